#!/usr/bin/env bash
# Builds the UNMODIFIED reference CUDA extension (models/csrc, pybind module `vren`) for sm_100a into
# oracle/_ref/vren.so, straight from the sources where they lie under /root/reference.  Nothing is copied
# into the repo; oracle/_ref/ is git-ignored but travels to the GPU box with the gpurun snapshot.
# The reference's own setup.py is NOT run; flags mirror it (-O2, default -fmad=true, no fast-math).
# TEST INFRASTRUCTURE ONLY: the parity tests (-m gpu) import it to pin our kernels and the C oracle.
set -euo pipefail
REF=${ARN_REFERENCE_DIR:-/root/reference}/models/csrc
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "reference not present at $REF; keeping any prebuilt $OUT/vren.so"; exit 0; }
mkdir -p "$OUT/obj"
PY=${PYTHON:-python}
TORCH_DIR=$($PY -c "import torch,os;print(os.path.dirname(torch.__file__))" 2>/dev/null | tail -1)
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
EXT=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
COMMON=(-O2 -std=c++17 -DTORCH_EXTENSION_NAME=vren -DTORCH_API_INCLUDE_EXTENSION_H
        -I"$TORCH_DIR/include" -I"$TORCH_DIR/include/torch/csrc/api/include" -I"$PYINC" -I"$REF/include")
NVCC=(nvcc "${COMMON[@]}" -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC
      -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__
      -include "$HERE/ref_shim.h")
pids=()
for f in raymarching volumerendering intersection losses; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/$f.cu" -nt "$OUT/obj/$f.o" ]; then
    "${NVCC[@]}" -c "$REF/$f.cu" -o "$OUT/obj/$f.o" 2> "$OUT/obj/$f.log" & pids+=($!)
  fi
done
if [ ! -f "$OUT/obj/binding.o" ] || [ "$REF/binding.cpp" -nt "$OUT/obj/binding.o" ]; then
  g++ "${COMMON[@]}" -fPIC -c "$REF/binding.cpp" -o "$OUT/obj/binding.o" 2> "$OUT/obj/binding.log" & pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
g++ -shared -o "$OUT/vren$EXT" "$OUT"/obj/*.o -L"$TORCH_DIR/lib" -L/usr/local/cuda/lib64 \
    -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python -lcudart \
    -Wl,-rpath,"$TORCH_DIR/lib"
echo "built $OUT/vren$EXT"
