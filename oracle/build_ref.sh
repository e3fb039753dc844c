#!/bin/bash
# Builds the UNMODIFIED reference hot path (src/python_bindings.cpp + headers, where they lie under
# /root/reference) into oracle/_ref/TRACS<ext>.so, against the two-header Boost stand-in in
# oracle/standin/. Output only under oracle/_ref/ (git-ignored, travels to the GPU box).
# TEST INFRASTRUCTURE: used to pin the oracle restatement and as the CPU baseline in bench.py.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${TRACS_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -f "$REF/src/python_bindings.cpp" ]; then
  echo "build_ref: $REF not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
EXT=$(python3 -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
INC=$(python3 -m pybind11 --includes)
# shipped flags (setup.py:46). Two builds: "native" (-march=native of THIS container, as setup.py ships) and
# "v3" (-march=x86-64-v3) in case the GPU box's host CPU lacks this container's ISA extensions;
# oracle/refmod.py picks native when it runs, else v3.
for V in native v3; do
  if [ $V = native ]; then M="-march=native"; else M="-march=x86-64-v3"; fi
  mkdir -p "$OUT/$V"
  g++ -std=c++17 -O3 -ffast-math $M -fopenmp -shared -fPIC -w -include "$HERE/standin/nosig.h" \
      -I"$HERE/standin" $INC "$REF/src/python_bindings.cpp" -o "$OUT/$V/TRACS$EXT" -lz &
done
wait
ls "$OUT"/native/TRACS$EXT "$OUT"/v3/TRACS$EXT
# The reference's Python callers of the native module and its own tests of the hot path, staged UNCHANGED under the
# git-ignored oracle/_ref/py/ so that the GPU box (which has no /root/reference) can run them on top of the CUDA
# drop-in (tests/test_gpu_reference_callers.py). pyfastx is imported by tracs/utils.py:8 at module top and never used
# on this path: an empty stub module stands in for it.
PY="$OUT/py"
rm -rf "$PY"
mkdir -p "$PY/tracs" "$PY/ref_tests"
for f in __init__ distance transcluster cluster utils threshold; do cp "$REF/tracs/$f.py" "$PY/tracs/$f.py"; done
for f in conftest test_llk test_pairsnp test_trans_distance; do cp "$REF/tests/$f.py" "$PY/ref_tests/$f.py"; done
printf '"""stub: tracs/utils.py imports pyfastx at module top; nothing on the distance / cluster path uses it"""\n' > "$PY/pyfastx.py"
ls "$PY/tracs" "$PY/ref_tests"
