#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r02c_tests.log
(tools/r02_flags_ab.sh "0x5 0x2d 0x6d 0x4d 0x6d 0x2d" Q8_0 2>&1) > gpurun_out/r02c_ab.log
tail -5 gpurun_out/r02c_tests.log; grep -E "flags|token total|rror" gpurun_out/r02c_ab.log
