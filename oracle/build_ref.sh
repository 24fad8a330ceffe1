#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference integral engine
# (/root/reference/TUNA/tuna_integrals/tuna_integral.pyx, Cython -> C -> .so) into oracle/_ref/.
# Nothing from the reference is copied into the repository history: the generated C file and the
# shared object live only under oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
# Flags follow the reference's own setup.py:56-83 (-O3 -fopenmp, bounds checks off).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${TUNA_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
PYX="$REF/TUNA/tuna_integrals/tuna_integral.pyx"
if [ ! -f "$PYX" ]; then
    echo "build_ref.sh: $PYX not present; keeping any prebuilt $OUT/*.so" >&2
    exit 0
fi
mkdir -p "$OUT"
PY="${PYTHON:-python}"
INC_PY=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
INC_NP=$($PY -c "import numpy; print(numpy.get_include())")
EXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
$PY -m cython -3 -X boundscheck=False -X wraparound=False -X cdivision=True -X nonecheck=False \
    -X initializedcheck=False "$PYX" -o "$OUT/tuna_integral.c"
/usr/bin/gcc -O3 -fopenmp -fPIC -shared -w -I"$INC_PY" -I"$INC_NP" \
    "$OUT/tuna_integral.c" -o "$OUT/tuna_integral$EXT" -lm
rm -f "$OUT/tuna_integral.c"
echo "built $OUT/tuna_integral$EXT"
# Stage the reference's own Python modules (UNMODIFIED, byte for byte) next to the engine so that the `-m gpu` tests can run the real
# reference driver (tuna_energy.evaluate_molecular_energy) on the GPU box, where /root/reference does not exist.  Like the .so this
# copy lives only under the git-ignored oracle/_ref/ and never enters the repository history.
mkdir -p "$OUT/TUNA"
cp -f "$REF"/TUNA/*.py "$OUT/TUNA/"
echo "staged $(ls "$OUT/TUNA" | wc -l) reference modules under $OUT/TUNA"
