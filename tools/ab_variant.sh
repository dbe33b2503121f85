#!/bin/bash
# Builds a variant of _lwop.so with extra -D flags on ONE translation unit (compile-time A/B of a kernel), next to the
# default library: usage ab_variant.sh <name> <file.cu> <-D flags...>; then LWOP_LIB_PATH=build/_lwop_<name>.so <tool>.
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p build/obj
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
    "$@" -c lightweight_openpose_b200/csrc/$src -o build/obj/${src%.cu}_$name.o
objs=""
for f in lwop_direct lwop_dw lwop_gemm lwop_sep lwop_heads lwop_decode lwop_draw lwop_net; do
    if [ "$f.cu" == "$src" ]; then objs="$objs build/obj/${f}_$name.o"; else objs="$objs build/obj/$f.o"; fi
done
/usr/local/cuda/bin/nvcc -shared -o build/_lwop_$name.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static
echo built build/_lwop_$name.so
