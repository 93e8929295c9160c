#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r02d_tests.log
tail -4 gpurun_out/r02d_tests.log
(timeout 300 python bench.py --steps 32 --warmup 5 2>gpurun_out/r02d_bench.err | tail -1) > gpurun_out/r02d_bench_default.json
for wl in llama2-7b-q4_0-q6k llama2-7b-q4_k tinyllamas-15m-q8_0; do
  (timeout 300 python bench.py --steps 32 --warmup 5 --workload $wl --no-cpu-baseline 2>>gpurun_out/r02d_bench.err | tail -1) > gpurun_out/r02d_bench_$wl.json
done
(timeout 300 python bench.py --steps 3 --warmup 1 --workload mistral-7b-q8_0-prefill 2>>gpurun_out/r02d_bench.err | tail -1) > gpurun_out/r02d_bench_prefill.json
(CRABML_MEGA_PROF=1 timeout 120 python tools/mega_profile.py Q8_0 2>&1 | tail -12) > gpurun_out/r02d_profile_q8_0.txt
(CRABML_MEGA_PROF=1 timeout 120 python tools/mega_profile.py Q4_0 2>&1 | tail -12) > gpurun_out/r02d_profile_q4_0.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d_bench_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split("bench_")[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "launches", d.get("gpu_launches_device_resident", d.get("gpu_launches")), {k:(round(v["value"],1), round(v["roofline"]["frac"],3)) if "value" in v else v for k,v in d.get("also",{}).items()})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300])
PY
tail -5 gpurun_out/r02d_bench.err
