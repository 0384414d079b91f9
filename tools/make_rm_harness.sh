#!/bin/bash
# Builds baseline/_ref/ : a writable copy of the reference RepeatMasker tree (git-ignored, travels to the GPU box
# with gpurun) so that the UNMODIFIED driver can be run with the oracle as `cross_match` (-engine crossmatch) and,
# after the two documented edits, with librmgpu (-engine gpu).  Run in the build container only:
# it reads /root/reference.  Nothing of the reference is committed to this repository.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${REF:-/root/reference}
DST="$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
if [ -d "$DST" ]; then mv "$DST" "$DST.old.$$"; fi
mkdir -p "$DST"
cp -r "$REF" "$DST/RepeatMasker"
chmod -R u+w "$DST/RepeatMasker"
RM="$DST/RepeatMasker"
# stand-ins for blobs that are not part of the reference checkout (.MISSING_LARGE_BLOBS)
mkdir -p "$DST/libdir/famdb" "$DST/libdir/general" "$DST/bin"
cat > "$DST/libdir/RepeatAnnotationData.pm" <<'PM'
package RepeatAnnotationData;
# stand-in: the real file is a large generated data blob absent from the reference checkout
our %repeatDB = ();
our %preProcData = ();
1;
PM
printf '#!/bin/sh\nexit 0\n' > "$DST/bin/trf"
chmod +x "$DST/bin/trf"
# the oracle, under the name the driver looks for: CROSSMATCH_DIR points at oracle/_build itself (never a stale copy)
make -C "$ROOT/oracle" >/dev/null
# ---- the two driver edits that add "-engine gpu"  ----
python3 - "$RM/RepeatMasker" <<'PY'
import sys
p = sys.argv[1]
s = open(p).read()
a = "/^(ncbi|rmblast|wublast|abblast|crossmatch|hmmer|nhmmer)$/i"
assert s.count(a) == 1
s = s.replace(a, "/^(ncbi|rmblast|wublast|abblast|crossmatch|hmmer|nhmmer|gpu)$/i")
b = 'elsif ( $engine eq "ncbi" ) {\n  $searchEngine = NCBIBlastSearchEngine->new('
assert s.count(b) == 1
s = s.replace(b, '''elsif ( $engine eq "gpu" ) {
  require GpuSearchEngine;
  $searchEngine = GpuSearchEngine->new( pathToEngine => $ENV{'RMGPU_LIB'},
                                        DEBUG        => $engineDEBUG );
  print "Search Engine: rmgpu [ " . $searchEngine->getVersion() . " ]\\n";
}
''' + b)
open(p, "w").write(s)
PY
cat > "$DST/env.sh" <<ENV
# source this: configuration of the harness through the driver's own environment overrides.
# HARNESS_ROOT is resolved at source time so the tree can live at another path on the GPU box.
HARNESS_ROOT="\$(cd "\$(dirname "\${BASH_SOURCE[0]}")/../.." && pwd)"
export CROSSMATCH_DIR="\$HARNESS_ROOT/oracle/_build"
export TRF_PRGM="\$HARNESS_ROOT/baseline/_ref/bin/trf"
export LIBDIR="\$HARNESS_ROOT/baseline/_ref/libdir"
export PERL5LIB="\$HARNESS_ROOT/baseline/_ref/libdir:\$HARNESS_ROOT/repeatmasker_b200/perl:\$PERL5LIB"
export RMGPU_LIB="\$HARNESS_ROOT/repeatmasker_b200/librmgpu.so"
export RM_HOME="\$HARNESS_ROOT/baseline/_ref/RepeatMasker"
ENV
echo "harness ready in $DST"
