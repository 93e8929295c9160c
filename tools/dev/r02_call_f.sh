#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r02f_tests.log
tail -4 gpurun_out/r02f_tests.log
(timeout 300 python bench.py --steps 32 --warmup 5 2>gpurun_out/r02f_bench.err | tail -1) > gpurun_out/r02f_bench_default.json
(timeout 300 python bench.py --steps 3 --warmup 1 --workload mistral-7b-q8_0-prefill 2>>gpurun_out/r02f_bench.err | tail -1) > gpurun_out/r02f_bench_prefill.json
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -s 690 -c 60 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r02f_bench_under_ncu.log 2>&1
timeout 200 $NCU --set full -k regex:matvec_stream_kernel -s 2 -c 1 -o gpurun_out/r02f_mvs_q8_0_32000x4096 python tools/prof_matvec.py Q8_0 32000 4096 1 > gpurun_out/r02f_ncu_mvs.log 2>&1
timeout 120 python tools/prof_matvec.py Q8_0 > gpurun_out/r02f_matvec_timing.txt 2>&1
timeout 120 python tools/prof_matvec.py Q4_0 >> gpurun_out/r02f_matvec_timing.txt 2>&1
timeout 120 python tools/prof_prefill.py Q8_0 4096 4096 4096 5 > gpurun_out/r02f_prefill_timing.txt 2>&1
timeout 120 python tools/prof_prefill.py Q8_0 14336 4096 4096 5 >> gpurun_out/r02f_prefill_timing.txt 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02f_bench_default.json").read())
print("default", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "sync", round(d["e2e"]["synchronous_variant"]["value"],1), "frac", round(d["roofline"]["frac"],3), {k:(round(v.get("value",0),1), round(v.get("e2e",{}).get("value",0),1)) for k,v in d.get("also",{}).items()})
d=json.loads(open("gpurun_out/r02f_bench_prefill.json").read())
print("prefill", round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "TF", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3))
PY
cat gpurun_out/r02f_prefill_timing.txt gpurun_out/r02f_matvec_timing.txt; tail -3 gpurun_out/r02f_bench.err
