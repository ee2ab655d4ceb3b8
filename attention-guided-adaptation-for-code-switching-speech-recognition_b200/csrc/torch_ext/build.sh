#!/bin/bash
# Builds aga_torch.so (TORCH_LIBRARY operators over the C ABI of libaga_b200.so) in-tree, next to the package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
PKG="$(cd "$HERE/../.." && pwd)"
PY="${PYTHON:-python}"
TORCH_DIR="$($PY -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
ABI="$($PY -c 'import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')"
OUT="$PKG/aga_torch.so"
if [ "$OUT" -nt "$HERE/aga_torch.cpp" ] && [ "$OUT" -nt "$PKG/../include/aga_b200.h" ]; then exit 0; fi
g++ -O2 -std=c++17 -fPIC -shared "$HERE/aga_torch.cpp" -o "$OUT" \
  -D_GLIBCXX_USE_CXX11_ABI=$ABI -DTORCH_API_INCLUDE_EXTENSION_H \
  -I"$PKG/../include" -I"$TORCH_DIR/include" -I"$TORCH_DIR/include/torch/csrc/api/include" -I/usr/local/cuda/include \
  -L"$PKG" -laga_b200 -L"$TORCH_DIR/lib" -ltorch -ltorch_cpu -lc10 -ltorch_cuda -lc10_cuda \
  -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$TORCH_DIR/lib" -Wl,--no-as-needed
