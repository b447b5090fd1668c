#!/usr/bin/env bash
# Build libonebit.so (sm_100a only) next to the Python package.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../libonebit.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -t 4 -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
  -Xcompiler -fPIC -shared -I"$ROOT/include" -I"$HERE" "$@" \
  "$HERE/ob_api.cu" "$HERE/ob_quant.cu" "$HERE/ob_gemm.cu" "$HERE/ob_gemv.cu" "$HERE/ob_decode.cu" "$HERE/ob_norm.cu" "$HERE/ob_attn.cu" "$HERE/ob_gemm_f32.cu" "$HERE/ob_conv.cu" "$HERE/ob_frontend.cu" "$HERE/ob_ctc.cu" -o "$OUT"
echo "built $OUT"
