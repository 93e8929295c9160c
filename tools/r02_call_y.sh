#!/bin/bash
mkdir -p gpurun_out
for k in 32 64 128 64; do
  echo -n "steps $k: "
  timeout 300 python bench.py --steps $k --warmup 8 --no-cpu-baseline --no-also 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'issue', round(d['host_ms_per_step']['issue_total'],3), d['config']['kv_positions'])"
done > gpurun_out/r02y_bench_steps.txt 2>&1
MEGA_PROFILE_WARM=40 timeout 200 python tools/mega_profile.py Q8_0 2>&1 | grep -E "tokens back|token total" >> gpurun_out/r02y_bench_steps.txt
cat gpurun_out/r02y_bench_steps.txt
