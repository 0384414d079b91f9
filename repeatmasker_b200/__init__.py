"""repeatmasker_b200: B200-native search engine for RepeatMasker's crossmatch path."""
from .engine import GpuSearchEngine, Session, RmgpuError, make_params, search_files, measure_dpx_peak, lib  # noqa: F401
