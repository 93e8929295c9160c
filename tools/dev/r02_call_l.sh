#!/bin/bash
# 8-GPU box: world-8 parity + soak, then the scaling table N = 8, 4, 2, 1 on the same box, NCCL baseline at 8, phase profile at 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 $TR --nproc-per-node 8 --master-port 29541 tests/sharded_worker.py --mode gpu --transports p2p 2>&1 | tail -12) > gpurun_out/r02l_worker_n8.log
tail -4 gpurun_out/r02l_worker_n8.log
for n in 8 4 2; do
  (timeout 400 $TR --nproc-per-node $n --master-port $((29550+n)) bench.py --gpus $n --steps 32 --warmup 5 --no-cpu-baseline 2>>gpurun_out/r02l_bench.err | tail -1) > gpurun_out/r02l_bench_n$n.json
done
(timeout 300 python bench.py --steps 32 --warmup 5 --no-cpu-baseline --no-also 2>>gpurun_out/r02l_bench.err | tail -1) > gpurun_out/r02l_bench_n1.json
(timeout 400 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 32 --warmup 5 --no-cpu-baseline --comm nccl 2>>gpurun_out/r02l_bench.err | tail -1) > gpurun_out/r02l_bench_n8_nccl.json
(timeout 300 $TR --nproc-per-node 8 --master-port 29563 tools/mega_profile_sharded.py 2>&1 | grep -E "world|n= |activation") > gpurun_out/r02l_profile_sharded_n8.txt
cat gpurun_out/r02l_profile_sharded_n8.txt
python - <<'PY'
import json
for n in ("n1","n2","n4","n8","n8_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/r02l_bench_{n}.json").read()); print(n, round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(n, "ERR", e)
PY
