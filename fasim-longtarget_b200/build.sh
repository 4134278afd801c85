#!/bin/sh
# Builds libfasim_b200.so (CUDA kernels + C ABI + host pipeline) and the `fasim` CLI for sm_100a.
#   build.sh                      the product
#   build.sh <tag> "<-D flags>"   a tuning variant variants/libfasim_b200_<tag>.so (selected with FASIM_B200_LIB)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
OUT="$HERE/libfasim_b200.so"; LOG="$HERE/build.log"
if [ -n "$1" ]; then mkdir -p "$HERE/variants"; OUT="$HERE/variants/libfasim_b200_$1.so"; LOG="$HERE/variants/build_$1.log"; fi
"$NVCC" $ARCH -O3 -std=c++17 -lineinfo -Xptxas -v --fmad=false $2 -Xcompiler -fPIC,-O2,-ffp-contract=off,-Wall,-pthread \
    -ccbin /usr/bin/g++ -shared -o "$OUT" "$HERE/csrc/engine.cu" -lz 2> "$LOG" || { cat "$LOG"; exit 1; }
if [ -z "$1" ]; then
    /usr/bin/g++ -O2 -std=c++17 -o "$HERE/fasim" "$HERE/host/fasim_cli.cpp" -L"$HERE" -lfasim_b200 -Wl,-rpath,'$ORIGIN'
fi
