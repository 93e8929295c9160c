#!/bin/bash
# generic (K-quant) phases as a called function in the ring kernel: mixed model profile, the new two-kernel test, K-quant tests
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_runner.py -q -k "both_persistent or lazy_7b" 2>&1 | tail -4) > gpurun_out/r02s_pytest.log; cat gpurun_out/r02s_pytest.log
for w in Q4_0-Q6K Q4_0; do echo "== $w"; timeout 200 python tools/mega_profile.py $w 2>&1 | grep -E "tokens back|token total|n= |rror" | cut -c1-150; done > gpurun_out/r02s_kquant_phases2.txt 2>&1; cat gpurun_out/r02s_kquant_phases2.txt
