#!/bin/bash
# GPU-box job: full GPU test suite, membench (L2 / bulk-prefetch microbenchmarks), megakernel flag A/B.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r02a_tests.log
(timeout 120 tools/build/membench 512 2>&1 | tail -70) > gpurun_out/r02a_membench.log
(tools/r02_flags_ab.sh "0x1 0x5 0x9 0x1005 0x1805 0x2005 0x3005 0x180d" Q8_0 2>&1) > gpurun_out/r02a_ab.log
(timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -8) > gpurun_out/r02a_smoke.log
tail -4 gpurun_out/r02a_tests.log; tail -4 gpurun_out/r02a_smoke.log
