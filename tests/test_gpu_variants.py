"""The fallback paths of the origin pass must give the same hits as the default (windowed, checkpointed) path:
RMGPU_K2_FULL=1 sweeps every passing triple with origins over the whole fragment; RMGPU_NO_CHECKPOINTS=1 uses the
score-only / origin mode mask from column 0; RMGPU_QUERY_CHUNK=2 forces the memory-bounded chunking of query records;
RMGPU_NO_WIDE=1 sends the triples whose int16 scores wrapped through the one-sweep path instead of the int32 score pass
(K2w) + block-parallel origin pass; RMGPU_K1_NO_IMMEDIATES=1 runs the generic instantiation of the score pass (gap
penalties in registers) where a recipe's penalties would select one with immediates; RMGPU_K1_SEGMENTS=3 cuts every
ticket of the score pass into three column segments (by default only launches with several waves of tickets are cut, in
four), RMGPU_K1_SEGMENTS=1 never cuts.  The switches are read once per process, hence the subprocesses."""
import os
import subprocess
import sys

import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("var,val", [("RMGPU_K2_FULL", "1"), ("RMGPU_NO_CHECKPOINTS", "1"), ("RMGPU_QUERY_CHUNK", "2"),
                                     ("RMGPU_NO_WIDE", "1"), ("RMGPU_K1_NO_IMMEDIATES", "1"), ("RMGPU_K1_SEGMENTS", "3"),
                                     ("RMGPU_K1_SEGMENTS", "1")])
def test_origin_pass_fallbacks_match_oracle(var, val):
    env = dict(os.environ, **{var: val})
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu",
                          os.path.join(util.ROOT, "tests", "test_gpu_parity.py"), "-k",
                          "synthetic_genome_full_parity or long_subject_chains or kat_seq1_2 or alphabet" +
                          (" or int16_overflow" if var == "RMGPU_NO_WIDE" else "") +
                          (" or score_pass or int16_overflow or wrapped or packing" if var == "RMGPU_K1_SEGMENTS" else "")],
                         env=env, capture_output=True, text=True, cwd=util.ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
