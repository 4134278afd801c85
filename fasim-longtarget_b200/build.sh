#!/bin/sh
# Builds libfasim_b200.so (CUDA kernels + C ABI + host pipeline) and the `fasim` CLI for sm_100a.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
"$NVCC" $ARCH -O3 -std=c++17 -lineinfo -Xptxas -v --fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off,-Wall \
    -ccbin /usr/bin/g++ -shared -o "$HERE/libfasim_b200.so" "$HERE/csrc/engine.cu" 2> "$HERE/build.log" || { cat "$HERE/build.log"; exit 1; }
/usr/bin/g++ -O2 -std=c++17 -o "$HERE/fasim" "$HERE/host/fasim_cli.cpp" -L"$HERE" -lfasim_b200 -Wl,-rpath,'$ORIGIN'
