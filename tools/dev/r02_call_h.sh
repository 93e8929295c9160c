#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/r02h_tests.log
tail -4 gpurun_out/r02h_tests.log
(timeout 300 python bench.py --steps 3 --warmup 1 --workload mistral-7b-q8_0-prefill 2>>gpurun_out/r02h_bench.err | tail -1) > gpurun_out/r02h_bench_prefill.json
(CRABML_PREFILL_MT1=1 timeout 300 python bench.py --steps 3 --warmup 1 --workload mistral-7b-q8_0-prefill 2>>gpurun_out/r02h_bench.err | tail -1) > gpurun_out/r02h_bench_prefill_mt1.json
timeout 120 python tools/prof_prefill.py Q8_0 4096 4096 4096 5 > gpurun_out/r02h_prefill_timing.txt 2>&1
timeout 120 python tools/prof_prefill.py Q8_0 14336 4096 4096 5 >> gpurun_out/r02h_prefill_timing.txt 2>&1
timeout 300 ncu --clock-control none --set full --import-source on -k regex:umma_gemm -s 2 -c 1 -o gpurun_out/r02h_gemm_4096_mt2 python tools/prof_prefill.py Q8_0 4096 4096 4096 3 > gpurun_out/r02h_ncu_gemm.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02h_bench_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split("bench_")[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "TF", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300])
PY
cat gpurun_out/r02h_prefill_timing.txt; tail -3 gpurun_out/r02h_bench.err
