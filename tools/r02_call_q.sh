#!/bin/bash
# producer with in-order probe loops: 2 / 4 / 8 producer warps
mkdir -p gpurun_out
run() { echo "== $1 flags $2"; CRABML_CUDA_LIB=$3 CRABML_MEGA_FLAGS=$2 timeout 200 python tools/mega_profile.py Q8_0 2>&1 | grep -E "token total|gate/up|qkv|k=10K|producer|rror" | cut -c1-150; }
{
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1)
run pw4 0x064d ""
run pw8 0x064d $PWD/crabml_b200/lib/libcrabml_cuda_pw8.so
run pw2 0x064d $PWD/crabml_b200/lib/libcrabml_cuda_pw2.so
run pw4 0x064d ""
run pw8 0x064d $PWD/crabml_b200/lib/libcrabml_cuda_pw8.so
(CRABML_CUDA_LIB=$PWD/crabml_b200/lib/libcrabml_cuda_pw8.so timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1)
} > gpurun_out/r02q_producer_inorder.txt 2>&1
cat gpurun_out/r02q_producer_inorder.txt
