#!/bin/bash
# tools/build_variant.sh NAME -DMACRO=V ... : crumble_b200/lib/variants/libcrumble_gpu_NAME.so with extra nvcc flags
# (A/B of kernel tunables on the GPU box: CRUMBLE_GPU_LIB selects the library, see tools/ab.sh)
set -e
cd "$(dirname "$0")/../crumble_b200/csrc"
name=$1; shift
mkdir -p ../lib/variants build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-O2 -Xptxas -v "$@" -c cg_device.cu -o build/cg_device_$name.o 2> build/ptxas_$name.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libcrumble_gpu_$name.so build/cg_device_$name.o build/cg_host.o build/transcode_gpu.o build/cg_multi.o build/crumble_main.o build/crumble_opts.o build/cg_params.o build/crumble_bed.o build/sam.o build/sam_hdr.o -lz -lm -lpthread -cudart static
grep -A2 "k_column" build/ptxas_$name.log | grep -E "spill|Used" | tr '\n' ' '; echo
