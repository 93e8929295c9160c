#!/bin/bash
# tail dealing: the last 2 n units of a phase one row at a time (flags bits 12-15 = n)
mkdir -p gpurun_out
run() { echo "== tail $1 $3 flags $2"; CRABML_MEGA_FLAGS=$2 timeout 200 python tools/mega_profile.py $3 2>&1 | grep -E "token total|gate/up|qkv|k=10K|mv\+res k=4K|rror" | cut -c1-150; }
{
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1)
run 0 0x064d Q8_0
run 8 0x864d Q8_0
run 4 0x464d Q8_0
run 15 0xf64d Q8_0
run 0 0x064d Q8_0
run 8 0x864d Q8_0
run 0 0x064d Q4_0
run 8 0x864d Q4_0
} > gpurun_out/r02q_tail_dealing.txt 2>&1
cat gpurun_out/r02q_tail_dealing.txt
