#!/bin/bash
# ring kernel: slot count / attention chunk A/B, then the bench line and the ncu captures of the new default
mkdir -p gpurun_out
for cfg in "0 32" "34 32" "28 32" "0 64" "0 16" "0 32"; do
    set -- $cfg
    echo "== ring slots $1, attention chunk $2 (Q8_0)"
    CRABML_RING_SLOTS=$1 CRABML_RING_ATCH=$2 timeout 200 python tools/mega_profile.py Q8_0 2>&1 | grep -E "tokens back|token total|n= |rror" | grep -v "normq\|rows  \|mv k=4K" | cut -c1-130 | head -8
done > gpurun_out/r02o_ring_slots_atch.txt 2>&1
cat gpurun_out/r02o_ring_slots_atch.txt
(timeout 600 python bench.py --steps 32 --warmup 5 2>gpurun_out/r02o_bench.err | tail -1) > gpurun_out/r02o_bench_n1.json; cat gpurun_out/r02o_bench_n1.json
