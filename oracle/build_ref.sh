#!/usr/bin/env bash
# Install the UNMODIFIED reference (johnveitch/cpnest) into oracle/_ref/ so that it can be
# imported as `cpnest` by the test-fixture generator (tests/golden/gen_golden.py) and by
# `bench.py --impl reference`.  oracle/_ref/ is git-ignored (it is a build artefact, not
# product source) but it is NOT gpurun-ignored, so it travels to the GPU box.
#
# This is what `pip install --target` would do (setup.py itself cannot run here:
# setuptools_scm is absent and there is no network): the pure-Python modules are installed
# as-is and the one native component, cpnest/parameter.pyx, is compiled with the flags of
# the reference's setup.py:23-28 (-O3 -ffast-math).
#
# TEST INFRASTRUCTURE ONLY: nothing under cpnest_b200/ may import from here.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/cpnest" ]; then
    echo "build_ref: $REF/cpnest not present (GPU box?) - keeping prebuilt $OUT" >&2
    exit 0
fi
PY=${PYTHON:-python}
mkdir -p "$OUT/cpnest" "$OUT/build"
install -m 0644 "$REF"/cpnest/*.py "$OUT/cpnest/"
NPINC=$($PY -c 'import numpy; print(numpy.get_include())')
PYINC=$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')
EXT=$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')
$PY -m cython -3 "$REF/cpnest/parameter.pyx" -o "$OUT/build/parameter.c"
gcc -shared -fPIC -O3 -ffast-math -w -I"$NPINC" -I"$PYINC" \
    "$OUT/build/parameter.c" -o "$OUT/cpnest/parameter$EXT"
rm -rf "$OUT/build"
$PY - <<PYEOF
import sys; sys.path.insert(0, "$OUT")
import cpnest, cpnest.parameter
print("build_ref: reference cpnest", cpnest.__version__, "installed in $OUT")
PYEOF
