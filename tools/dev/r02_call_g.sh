#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/r02g_tests.log
tail -4 gpurun_out/r02g_tests.log
for wl in llama2-7b-q4_k llama2-7b-q4_0-q6k; do
  (timeout 300 python bench.py --steps 32 --warmup 5 --workload $wl --no-cpu-baseline 2>>gpurun_out/r02g_bench.err | tail -1) > gpurun_out/r02g_bench_$wl.json
done
(timeout 300 python bench.py --steps 3 --warmup 1 --workload mistral-7b-q8_0-prefill 2>>gpurun_out/r02g_bench.err | tail -1) > gpurun_out/r02g_bench_prefill.json
(CRABML_MEGA_PROF=1 timeout 120 python tools/mega_profile.py Q4_K 2>&1 | tail -12) > gpurun_out/r02g_profile_q4_k.txt
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -k "regex:mega_kernel|argmax_kernel|dequant_rows|matvec|quantize|normq|attn_decode|exchange|binary|rms_norm|softmax|silu|scale_kernel|strided_copy|rope|bmm" -c 200 --csv --log-file gpurun_out/r02g_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r02g_bench_under_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g_bench_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split("bench_")[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300])
PY
cat gpurun_out/r02g_profile_q4_k.txt | head -8; tail -3 gpurun_out/r02g_bench.err
