#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_runner.py -x -q -m gpu 2>&1 | tail -5) > gpurun_out/r02k_tests.log
tail -3 gpurun_out/r02k_tests.log
(timeout 300 $TR --master-port 29531 tools/mega_profile_sharded.py 2>&1 | grep -E "world|n= |activation") > gpurun_out/r02k_profile_sharded.txt
(timeout 300 python bench.py --steps 32 --warmup 5 --no-cpu-baseline --no-also 2>gpurun_out/r02k_bench.err | tail -1) > gpurun_out/r02k_bench_n1.json
(timeout 400 $TR --master-port 29533 bench.py --gpus 2 --steps 32 --warmup 5 --no-cpu-baseline 2>>gpurun_out/r02k_bench.err | tail -1) > gpurun_out/r02k_bench_n2.json
cat gpurun_out/r02k_profile_sharded.txt
python - <<'PY'
import json
for n in ("n1","n2"):
    d=json.loads(open(f"gpurun_out/r02k_bench_{n}.json").read()); print(n, round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
PY
