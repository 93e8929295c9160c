#!/bin/bash
# evidence of the ring default: bench lines (all decode workloads + prefill), phase profiles, launch list, ncu --set full of mega_ring_kernel
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 32 --warmup 5 2>gpurun_out/r02p_bench.err | tail -1) > gpurun_out/r02p_bench_n1.json
for wl in llama2-7b-q4_k llama2-7b-q4_0-q6k tinyllamas-15m-q8_0; do
  (timeout 300 python bench.py --steps 32 --warmup 5 --workload $wl --no-cpu-baseline --no-also 2>>gpurun_out/r02p_bench.err | tail -1) > gpurun_out/r02p_bench_$wl.json
done
(timeout 300 python bench.py --steps 3 --warmup 1 --workload mistral-7b-q8_0-prefill --no-cpu-baseline 2>>gpurun_out/r02p_bench.err | tail -1) > gpurun_out/r02p_bench_prefill.json
for wl in Q8_0 Q4_0; do
  echo "== $wl (default flags: ring + pairs)"
  timeout 200 python tools/mega_profile.py $wl 2>&1 | grep -E "tokens back|token total|n= |rror|producer"
done > gpurun_out/r02p_mega_phases.txt 2>&1
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/r02p_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r02p_bench_under_ncu.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:mega_ring_kernel -s 8 -c 1 -o gpurun_out/r02p_mega_ring_q8_0 python tools/mega_profile.py Q8_0 > gpurun_out/r02p_ncu_mega.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02p_bench_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split("bench_")[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:40], {k:(round(v.get("value",0),1)) for k,v in d.get("also",{}).items()})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300])
PY
cat gpurun_out/r02p_mega_phases.txt; tail -3 gpurun_out/r02p_bench.err; ls -la gpurun_out | grep r02p
