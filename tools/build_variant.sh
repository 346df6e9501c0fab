#!/bin/bash
# Build a compile-time variant of libsolb.so next to the shipped one for A/B runs on the GPU box:
#   bash tools/build_variant.sh r72 -DSOLB_WF_MIN_CTAS=7          -> sol_rs_b200/libsolb_r72.so
#   bash tools/build_variant.sh prmt -DSOLB_Q2M_PRMT               -> the PRMT dequantisation of sessions 1-2
# then, on the box:  SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_r72.so python bench.py --steps 4 --no-cpu-baseline
# (tools/sweep.sh takes the same VAR=value words).  Variant libraries are git-ignored (*.so) and travel with gpurun.
set -e
name=$1; shift
cd "$(dirname "$0")/../sol_rs_b200/csrc"
FL="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ccbin /usr/bin/g++ --expt-relaxed-constexpr"
mkdir -p build
nvcc $FL "$@" -c trace.cu -o build/trace_$name.o &
nvcc $FL "$@" -Xptxas -dlcm=cg -c build.cu -o build/build_$name.o &
nvcc $FL "$@" -c solb_api.cu -o build/api_$name.o &
nvcc $FL "$@" -c comm.cu -o build/comm_$name.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsolb_$name.so build/trace_$name.o build/build_$name.o build/api_$name.o build/comm_$name.o -lcudart_static -ldl -ccbin /usr/bin/g++ -Xcompiler -fPIC
ls -la ../libsolb_$name.so
