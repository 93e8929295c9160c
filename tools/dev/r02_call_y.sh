#!/bin/bash
mkdir -p gpurun_out
for w in 168 165 168 360 357; do
  echo -n "KV capacity $((w+88)) rows: "
  MEGA_PROFILE_WARM=$w timeout 200 python tools/mega_profile.py Q8_0 2>&1 | grep -E "tokens back" | cut -c1-60
  MEGA_PROFILE_WARM=$w timeout 200 python tools/mega_profile.py Q8_0 2>&1 | grep -E "attn" | cut -c1-110
done > gpurun_out/r02y_kv_capacity.txt 2>&1
cat gpurun_out/r02y_kv_capacity.txt
