#!/bin/bash
# KV prefetch into L2 by the producer warps (flag 0x800): A/B at KV ~60 and ~200, then bit identity + targeted tests with the flag on
mkdir -p gpurun_out
run() { echo "== flags $1 KV ~$2"; CRABML_MEGA_FLAGS=$1 MEGA_PROFILE_WARM=$2 timeout 200 python tools/mega_profile.py Q8_0 2>&1 | grep -E "token total|attn|qkv|rror" | cut -c1-120; }
{
run 0x064d 40
run 0x0e4d 40
run 0x064d 200
run 0x0e4d 200
run 0x064d 40
run 0x0e4d 40
(CRABML_MEGA_FLAGS=0x0e4d timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1)
(CRABML_MEGA_FLAGS=0x0e4d timeout 600 python -m pytest tests/test_gpu_runner.py tests/test_gpu_llama.py tests/test_gpu_phase_taps.py -q -x 2>&1 | tail -2)
} > gpurun_out/r02z_kv_prefetch.txt 2>&1
cat gpurun_out/r02z_kv_prefetch.txt
