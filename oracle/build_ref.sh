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
