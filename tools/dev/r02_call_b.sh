#!/bin/bash
mkdir -p gpurun_out
(CRABML_MEGA_FLAGS=0x183d timeout 300 python -m pytest tests/test_gpu_runner.py -x -q -m gpu -k "lazy_7b or modes_bit or lazy_fused" 2>&1 | tail -5) > gpurun_out/r02b_tests.log
(tools/r02_flags_ab.sh "0x5 0xd 0x25 0x2d 0x1815 0x1835 0x183d 0x3035 0x0c35 0x5" Q8_0 2>&1) > gpurun_out/r02b_ab.log
tail -3 gpurun_out/r02b_tests.log; grep -E "flags|token total" gpurun_out/r02b_ab.log
