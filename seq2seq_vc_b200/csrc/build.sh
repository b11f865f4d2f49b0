#!/bin/bash
# Builds libs2svc_b200.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libs2svc_b200.so
SRCS="api.cu gemm_simt.cu gemm_tc.cu gemm_split.cu ops_norm.cu ops_attn.cu attn_fused.cu attn_tc.cu ops_misc.cu ops_conformer.cu ops_align.cu decode.cu mas.cu logmel.cu ops_sdp.cu ops_lr.cu griffinlim.cu"
mkdir -p _obj
pids=()
for s in $SRCS; do
  o=_obj/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ tc_common.cuh -nt "$o" ] || [ ../../include/s2svc_b200.h -nt "$o" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -Wno-deprecated-gpu-targets -shared -o $OUT _obj/*.o -lcuda
echo "built $OUT"
