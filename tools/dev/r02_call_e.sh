#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r02e_tests.log
tail -4 gpurun_out/r02e_tests.log
(timeout 300 python bench.py --steps 32 --warmup 5 2>gpurun_out/r02e_bench.err | tail -1) > gpurun_out/r02e_bench_default.json
(CRABML_MEGA_PROF=1 timeout 120 python tools/mega_profile.py Q4_K 2>&1 | tail -12) > gpurun_out/r02e_profile_q4_k.txt
NCU="ncu --clock-control none"
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02e_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r02e_bench_under_ncu.log 2>&1
# the dominant kernel
timeout 400 $NCU --set full --import-source on -k regex:mega_kernel -s 8 -c 1 -o gpurun_out/r02e_mega_q8_0 python tools/mega_profile.py Q8_0 > gpurun_out/r02e_ncu_mega.log 2>&1
# the four Llama-2-7B matvec shapes of the eager streaming kernel (+ Q4_0 for the ffn shape)
for shp in "4096 4096" "11008 4096" "4096 11008" "32000 4096"; do
  set -- $shp
  timeout 200 $NCU --set full -k regex:matvec_stream_kernel -s 6 -c 1 -o gpurun_out/r02e_mvs_q8_0_$1x$2 python tools/prof_matvec.py Q8_0 $1 $2 1 >> gpurun_out/r02e_ncu_mvs.log 2>&1
done
timeout 200 $NCU --set full -k regex:matvec_stream_kernel -s 6 -c 1 -o gpurun_out/r02e_mvs_q4_0_11008x4096 python tools/prof_matvec.py Q4_0 11008 4096 1 >> gpurun_out/r02e_ncu_mvs.log 2>&1
# the prefill GEMM
timeout 300 $NCU --set full --import-source on -k regex:umma_gemm -s 2 -c 1 -o gpurun_out/r02e_gemm_4096 python tools/prof_prefill.py Q8_0 4096 4096 4096 3 > gpurun_out/r02e_ncu_gemm.log 2>&1
timeout 120 python tools/prof_prefill.py Q8_0 4096 4096 4096 5 > gpurun_out/r02e_prefill_timing.txt 2>&1
timeout 120 python tools/prof_prefill.py Q8_0 14336 4096 4096 5 >> gpurun_out/r02e_prefill_timing.txt 2>&1
ls -la gpurun_out | grep r02e
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02e_bench_default.json").read())
print("default", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "sync", round(d["e2e"]["synchronous_variant"]["value"],1), "frac", round(d["roofline"]["frac"],3), {k:(round(v.get("value",0),1), round(v.get("e2e",{}).get("value",0),1)) for k,v in d.get("also",{}).items()})
PY
cat gpurun_out/r02e_prefill_timing.txt; tail -3 gpurun_out/r02e_bench.err
