#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference CPU driver (src/selection.cpp,
# OpenMP) from the sources where they lie under /root/reference into oracle/_ref/.
# No reference source is copied into this repo; only the binary lands in oracle/_ref/
# (git-ignored, but it travels to the GPU box with gpurun).
#
# We do not run the reference's own Makefile: it uses -march=native, which fails on
# AVX-512 hosts at sketch/include/sketch/bbmh.h:1499. Flags otherwise
# follow the reference Makefile:32-33 (CXXFLAGS) and :39 (LDFLAGS_CPU).
set -euo pipefail
REF="${REFERENCE_ROOT:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/selection.cpp" ]; then
  echo "build_ref: $REF not present; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
# The reference's own CUDA kernels + launchers (src/selection_kernels.cu), compiled in place for
# sm_100a: the checker for our link-level shims and the "existing GPU kernel" speed bar.
if [ ! "$OUT/libref_kernels.so" -nt "$REF/src/selection_kernels.cu" ] || [ "${FORCE:-0}" = "1" ]; then
  (cd "$REF" && /usr/local/cuda/bin/nvcc -O3 --std=c++17 -gencode arch=compute_100a,code=sm_100a -ccbin g++ \
      -DNDEBUG -I. -Iinclude -shared -Xcompiler -fPIC src/selection_kernels.cu -o "$OUT/libref_kernels.so") \
    && echo "build_ref: built $OUT/libref_kernels.so"
fi
# The reference's sketch builder (src/build_sketch.cpp), checker for our GPU build_sketch.
if [ ! "$OUT/build_sketch" -nt "$REF/src/build_sketch.cpp" ] || [ "${FORCE:-0}" = "1" ]; then
  (cd "$REF" && g++ -O3 -std=c++17 -march=x86-64-v3 -fopenmp -DSEQAN_HAS_ZLIB=1 -DNDEBUG -w \
      -I. -Isketch -Isketch/include -Isketch/include/blaze -Iseqan-library-2.4.0/include -Iinclude \
      src/build_sketch.cpp -lz -pthread -o "$OUT/build_sketch") && echo "build_ref: built $OUT/build_sketch"
fi
if [ "$OUT/selection" -nt "$REF/src/selection.cpp" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "build_ref: $OUT/selection up to date"
  exit 0
fi
cd "$REF"
g++ -O3 -std=c++17 -march=x86-64-v3 -fopenmp -DSEQAN_HAS_ZLIB=1 -DNDEBUG -w \
    -I. -Isketch -Isketch/include -Isketch/include/blaze \
    -Iseqan-library-2.4.0/include -Iinclude \
    src/selection.cpp -lz -pthread -o "$OUT/selection"
echo "build_ref: built $OUT/selection"
