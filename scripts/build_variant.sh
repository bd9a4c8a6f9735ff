#!/bin/bash
# usage: build_variant.sh NAME -DMVHMR_WARPS=.. ...   -> gpurun_out/variants/NAME.so (only the softmax fp32/bf16 V=4 paths matter)
set -e
name=$1; shift
mkdir -p /root/repo/variants
cd /root/repo/multiviewhmr_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static "$@" -o /root/repo/variants/$name.so abi.cu geometry.cu unproject.cu softargmax.cu backward.cu
