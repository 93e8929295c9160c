#!/bin/bash
# 2-GPU box: architecture tests, sharded phase profile with / without the per-CTA system fence, correctness of the no-fence variant
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_architectures.py tests/test_gpu_prefill.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r02j_tests.log
tail -3 gpurun_out/r02j_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29521 tools/mega_profile_sharded.py 2>&1 | grep -E "world|n= |activation") > gpurun_out/r02j_profile_sharded_default.txt
(CRABML_MEGA_FLAGS=0x14d timeout 300 $TR --master-port 29522 tools/mega_profile_sharded.py 2>&1 | grep -E "world|n= |activation") > gpurun_out/r02j_profile_sharded_nofence.txt
(CRABML_MEGA_FLAGS=0x14d timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -5) > gpurun_out/r02j_tests_nofence.log
(CRABML_MEGA_FLAGS=0x14d timeout 400 $TR --master-port 29523 bench.py --gpus 2 --steps 32 --warmup 5 --no-cpu-baseline 2>>gpurun_out/r02j_bench.err | tail -1) > gpurun_out/r02j_bench_n2_nofence.json
cat gpurun_out/r02j_profile_sharded_default.txt; cat gpurun_out/r02j_profile_sharded_nofence.txt; tail -3 gpurun_out/r02j_tests_nofence.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02j_bench_n2_nofence.json").read()); print("n2 nofence", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
PY
