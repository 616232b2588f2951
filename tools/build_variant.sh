#!/bin/bash
# Build a variant of libmapf_b200.so for kernel A/B experiments: tools/build_variant.sh NAME [extra nvcc flags...]
# -> scratch/variants/libmapf_NAME.so (use with MAPF_B200_LIB=...).  -DMAPF_DEV_MINIMAL keeps only sensor range 2.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scratch/variants
nvcc -std=c++17 -O3 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -pthread \
  -DMAPF_DEV_MINIMAL "$@" -o scratch/variants/libmapf_$name.so \
  dl_reference_models_b200/csrc/mapf_b200.cu dl_reference_models_b200/csrc/mapf_host_unpack.cpp
echo scratch/variants/libmapf_$name.so
