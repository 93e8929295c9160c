#!/bin/bash
# 2-GPU box: world-2 parity (every mode and transport), the sharded bench with the ring kernel against the register-pipe kernel, phase profile
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 $TR --nproc-per-node 2 --master-port 29541 tests/sharded_worker.py --mode gpu --transports p2p,nccl 2>&1 | tail -12) > gpurun_out/r02u_worker_n2.log
tail -5 gpurun_out/r02u_worker_n2.log
(timeout 400 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 32 --warmup 5 --no-cpu-baseline 2>>gpurun_out/r02u_bench.err | tail -1) > gpurun_out/r02u_bench_n2.json
(CRABML_MEGA_FLAGS=0x4d timeout 400 $TR --nproc-per-node 2 --master-port 29553 bench.py --gpus 2 --steps 32 --warmup 5 --no-cpu-baseline 2>>gpurun_out/r02u_bench.err | tail -1) > gpurun_out/r02u_bench_n2_regpipe.json
(timeout 300 python bench.py --steps 32 --warmup 5 --no-cpu-baseline --no-also 2>>gpurun_out/r02u_bench.err | tail -1) > gpurun_out/r02u_bench_n1.json
(timeout 300 $TR --nproc-per-node 2 --master-port 29563 tools/mega_profile_sharded.py 2>&1 | grep -E "world|n= |activation|token") > gpurun_out/r02u_profile_sharded_n2.txt
cat gpurun_out/r02u_profile_sharded_n2.txt
python - <<'PY'
import json
for n in ("n1","n2","n2_regpipe"):
    try:
        d=json.loads(open(f"gpurun_out/r02u_bench_{n}.json").read()); print(n, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["roofline"]["kernel"][:20])
    except Exception as e: print(n, "ERR", e)
PY
tail -3 gpurun_out/r02u_bench.err
