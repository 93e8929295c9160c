#!/bin/bash
# 2-GPU box: full GPU test suite (incl. the 2-GPU sharded test), scaling bench at N = 1, 2 (fused p2p exchange) and the NCCL baseline
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02i_gpus.txt
(timeout 1800 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -vE "^\s*$" | tail -40) > gpurun_out/r02i_tests.log
tail -6 gpurun_out/r02i_tests.log
(timeout 300 python bench.py --steps 32 --warmup 5 --no-cpu-baseline --no-also 2>gpurun_out/r02i_bench.err | tail -1) > gpurun_out/r02i_bench_n1.json
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 5 --no-cpu-baseline 2>>gpurun_out/r02i_bench.err | tail -1) > gpurun_out/r02i_bench_n2.json
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 5 --no-cpu-baseline --comm nccl 2>>gpurun_out/r02i_bench.err | tail -1) > gpurun_out/r02i_bench_n2_nccl.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02i_bench_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split("bench_")[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300])
PY
tail -5 gpurun_out/r02i_bench.err
