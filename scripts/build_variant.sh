#!/bin/bash
# usage: build_variant.sh NAME -DMVHMR_WARPS=.. ...   -> variants/NAME.so (a tuning build of the whole library; select it with MVHMR_LIB)
set -e
name=$1; shift
mkdir -p /root/repo/variants
cd /root/repo/multiviewhmr_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static "$@" -o /root/repo/variants/$name.so abi.cu geometry.cu unproject.cu unproject_out0.cu unproject_out1.cu unproject_out2.cu unproject_out3.cu unproject_staged.cu unproject_tex.cu softargmax.cu backward.cu
