#!/bin/bash
mkdir -p gpurun_out
for f in 0x24d 0x64d 0x24d 0x64d 0x4d; do
    echo "== flags $f (Q8_0)"
    CRABML_MEGA_FLAGS=$f timeout 200 python tools/mega_profile.py Q8_0 2>&1 | grep -E "tokens back|token total|n= |rror|producer" | grep -v "normq\|rows  " | head -14
done > gpurun_out/r02n_profile6.txt 2>&1
cat gpurun_out/r02n_profile6.txt
