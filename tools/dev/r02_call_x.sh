#!/bin/bash
# attention chunk size at the KV lengths of the default bench (positions 100-250)
mkdir -p gpurun_out
for cfg in "32 200" "64 200" "32 200" "64 200" "48 200" "64 40"; do
    set -- $cfg
    echo "== attention chunk $1, KV length ~$2"
    CRABML_RING_ATCH=$1 MEGA_PROFILE_WARM=$2 timeout 200 python tools/mega_profile.py Q8_0 2>&1 | grep -E "token total|attn|gate/up|rror" | cut -c1-130
done > gpurun_out/r02x_atch_long.txt 2>&1
cat gpurun_out/r02x_atch_long.txt
