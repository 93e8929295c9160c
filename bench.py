#!/usr/bin/env python
"""bench.py -- decode tokens/s of the crabml CUDA backend on synthetic Llama-2-7B shapes (BASELINE.json metric).

One "step" = one decoded token = one pass of the hot path (225 quantized matvecs + the small ops around them).
  value : device-resident throughput (weights + KV in HBM, no host<->device traffic inside the timed region)
  e2e   : the same loop through the reference-facing runner API with HOST buffers: per step the token id goes
          host->device and the 128 KB logits come device->host (Llama2Runner::forward -> export, llama2.rs:209),
          then the host samples (argmax) and feeds the next token.
  roofline : the dominant kernel (the ffn_gate/ffn_up-shaped matvec, 11008x4096) timed with CUDA events
  cpu_baseline : the CPU restatement of the reference's AVX2 path (oracle/, all host threads) on a bounded sample
`--impl reference` runs ONLY that CPU path and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x5EED
WORKLOADS = {
    # name: (config name, body weight type, classifier weight type)
    "llama2-7b-q8_0": ("LLAMA2_7B", "Q8_0", "Q8_0"),
    "llama2-7b-q4_0": ("LLAMA2_7B", "Q4_0", "Q4_0"),
    "llama2-7b-q4_0-q6k": ("LLAMA2_7B", "Q4_0", "Q6_K"),
    "llama2-7b-q4_k": ("LLAMA2_7B", "Q4_K", "Q6_K"),
    "tinyllamas-15m-q8_0": ("TINYLLAMAS_15M", "Q8_0", "Q8_0"),
}
# dram bytes per megakernel launch from the committed ncu --set full capture (profiles/); None until captured
TRAFFIC = {("llama2-7b-q8_0", 1): 7059924000 + 11929600}     # profiles/r01f_megakernel_ncu.md
TYPE_ID = {"Q4_0": 2, "Q4_1": 3, "Q5_0": 6, "Q5_1": 7, "Q8_0": 8, "Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}


def usable_cpus():
    """CPUs this process may actually use: affinity mask capped by the cgroup CPU quota (spin-waiting workers beyond the
    quota get throttled by CFS and make the CPU arm erratic)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return n


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path restated (oracle/), AVX2 order, all host threads, bounded sample
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_tokens_per_s(workload: str, budget_tokens: int = 3):
    from crabml_b200 import runner as R          # config table + byte accounting only (no GPU use)
    from oracle import oracle as oc
    from oracle.llama_replay import Llama2Runner, LlamaConfig, LlamaWeights
    from oracle.synth import synth_weight
    from oracle.tensor_ref import OracleDevice, OracleTensor

    cname, wt_name, ct_name = WORKLOADS[workload]
    conf = getattr(R, cname)
    wt, ct = TYPE_ID[wt_name], TYPE_ID[ct_name]
    dim, hid, kv = conf.embedding_dim, conf.hidden_dim, conf.head_size() * conf.n_kv_heads
    # thread count: the reference takes it from the command line; use the count that is fastest for the dominant matvec on
    # this host (ascending, stop when it gets slower -- oversubscribed spin-waiting workers are pathological on big hosts)
    probe_w = synth_weight(wt, 2048, dim, SEED, 99, R.synth_scale(wt, dim))
    probe_x = np.random.default_rng(0).standard_normal(dim).astype(np.float32)
    threads, best = 1, float("inf")
    cpus = min(usable_cpus(), oc.hw_threads())
    for cand in sorted({c for c in (1, 2, 4, 8, 12, 16, 24, 32, 48, 64, 96, 128, cpus) if c <= cpus}):
        oc.gemv(wt, probe_w, 2048, dim, probe_x, threads=cand, flags=oc.ORDER_AVX2)
        dt = float("inf")
        for _ in range(3):                       # best of 3 batches of 5
            t0 = time.perf_counter()
            for _ in range(5):
                oc.gemv(wt, probe_w, 2048, dim, probe_x, threads=cand, flags=oc.ORDER_AVX2)
            dt = min(dt, time.perf_counter() - t0)
        if dt < best * 0.97:
            threads, best = cand, dt
        elif dt > best * 1.3:
            break
    dev = OracleDevice(thread_num=threads, flags=oc.ORDER_AVX2)
    n_sample_layers = min(2, conf.n_layers)

    def syn(rows, cols, t, tid):
        return OracleTensor.from_cpu(synth_weight(t, rows, cols, SEED, tid, R.synth_scale(t, cols)), [rows, cols], t, dev)
    rng = np.random.default_rng(SEED)

    def norm():
        return OracleTensor.from_cpu((1.0 + 0.05 * rng.standard_normal(dim)).astype(np.float32), [dim], oc.F32, dev)
    L = conf.n_layers
    w = dict(wq=[], wk=[], wv=[], wo=[], gate=[], up=[], down=[], ra=[], rf=[])
    for l in range(n_sample_layers):            # same tensor ids as crabml_b200.runner.synthetic_weights
        b = 7 * l
        w["wq"].append(syn(dim, dim, wt, b + 1)); w["wk"].append(syn(kv, dim, wt, b + 2)); w["wv"].append(syn(kv, dim, wt, b + 3))
        w["wo"].append(syn(dim, dim, wt, b + 4)); w["gate"].append(syn(hid, dim, wt, b + 5)); w["up"].append(syn(hid, dim, wt, b + 6))
        w["down"].append(syn(dim, hid, wt, b + 7)); w["ra"].append(norm()); w["rf"].append(norm())
    tok_embed = syn(conf.vocab_size, dim, wt, 7 * L + 1)
    out_w = syn(conf.vocab_size, dim, ct, 7 * L + 2)
    times = {}
    for nl in sorted({1, n_sample_layers}):
        c = LlamaConfig(conf.n_heads, conf.n_kv_heads, nl, dim, hid, conf.seq_len, conf.vocab_size, conf.rms_norm_eps, conf.rope_dim or None)
        lw = LlamaWeights(tok_embed, w["wq"][:nl], w["wk"][:nl], w["wv"][:nl], w["wo"][:nl], w["gate"][:nl], w["down"][:nl], w["up"][:nl],
                          w["ra"][:nl], w["rf"][:nl], norm(), out_w)
        r = Llama2Runner(OracleTensor, c, lw, dev, 16)
        r.forward([1], 0)                        # warm-up token
        best_t = float("inf")
        for i in range(budget_tokens):
            t0 = time.perf_counter()
            r.forward([2 + i], 1 + i)
            best_t = min(best_t, time.perf_counter() - t0)
        times[nl] = best_t
    if n_sample_layers > 1:
        per_layer = (times[n_sample_layers] - times[1]) / (n_sample_layers - 1)
        if per_layer <= 0:                       # timer noise: split the 2-layer time by streamed bytes instead
            lb = 4 * R.weight_bytes(wt, dim, dim) + 3 * R.weight_bytes(wt, hid, dim)
            per_layer = times[n_sample_layers] * lb / (n_sample_layers * lb + R.weight_bytes(ct, conf.vocab_size, dim))
        rest = times[1] - per_layer
    else:
        per_layer, rest = times[1], 0.0
    per_token = per_layer * conf.n_layers + max(rest, 0.0)
    sample = (f"{n_sample_layers} of {conf.n_layers} layers + classifier of {workload} (same synthetic weights, seed {SEED:#x}), "
              f"best of {budget_tokens} tokens after 1 warm-up, per-layer time x {conf.n_layers} + classifier; AVX2-order restatement, "
              f"{threads} threads (fastest count of an ascending probe; {cpus} usable CPUs = affinity capped by the cgroup quota, {oc.hw_threads()} hardware threads)")
    return 1.0 / per_token, threads, sample, per_token


def run_reference(args, rank, world):
    if rank != 0:
        return
    tps, threads, sample, per_token = cpu_reference_tokens_per_s(args.workload)
    line = {
        "impl": "reference", "metric": "decode_tokens_per_s", "value": tps, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_token * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8", "data": "synthetic",
        "config": {"workload": f"{args.workload}-decode-synthetic", "note": "reference CPU path restated in C (oracle/): the Rust reference cannot be built here"},
        "cpu_baseline": {"value": tps, "unit": "tok/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    from crabml_b200 import CudaTensor, CudaTensorDevice
    from crabml_b200 import runner as R

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    sharded = world > 1 and args.multi == "sharded"

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cname, wt_name, ct_name = WORKLOADS[args.workload]
    conf = getattr(R, cname)
    wt, ct = TYPE_ID[wt_name], TYPE_ID[ct_name]
    transport = args.comm
    lazy = args.lazy if not (sharded and transport == "nccl" and args.lazy == 2) else 1     # NCCL cannot run inside the megakernel
    dev = CudaTensorDevice(local_rank, lazy=lazy)
    plan = None
    if sharded:
        from crabml_b200 import sharding

        def exchange(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out
        plan = sharding.make_plan(conf.n_heads, conf.n_kv_heads, conf.embedding_dim, conf.hidden_dim, conf.vocab_size, wt, rank, world)
        dev.init_comm(rank, world, exchange, transport)
    args.lazy = lazy
    weights = R.synthetic_weights(dev, conf, wt, ct, seed=SEED, plan=plan)
    K, W = args.steps, args.warmup
    kv_len = min(conf.seq_len, args.start_pos + 2 * (W + K) + 8)
    runner = R.LlamaRunner(dev, conf, weights, kv_len, plan=plan)
    bytes_per_token = runner.weight_bytes_per_token()

    # context: fill the KV cache up to start_pos with untimed steps so that decode runs at a realistic position
    pos, tok = 0, 1
    for _ in range(args.start_pos):
        runner.forward([tok], pos, export=False); pos += 1; tok = (tok * 7 + 3) % conf.vocab_size

    def step_e2e(t, p):
        lg = runner.forward([t], p, export=True)      # token id H2D (8 B) + logits D2H (vocab*4 B)
        return int(np.flatnonzero(lg == lg.max())[-1])   # sampler.rs:109-116 argmax (last max)

    for _ in range(W):
        tok = step_e2e(tok, pos); pos += 1
    # ---- timed region 1: e2e through the runner API with host buffers ---------------------------------------
    sampler = ClockSampler(local_rank); sampler.start()
    dev.synchronize(); barrier()
    launches0 = dev.launch_count()
    dev.timer_begin(); t0 = time.perf_counter()
    for _ in range(K):
        tok = step_e2e(tok, pos); pos += 1
    e2e_ms_dev = dev.timer_end(); e2e_wall = time.perf_counter() - t0
    launches_e2e = dev.launch_count() - launches0
    barrier()
    e2e_ms = max_over_ranks(max(e2e_ms_dev, e2e_wall * 1e3))      # host work (sampling) is part of e2e
    # ---- timed region 2: device-resident (no per-step host<->device traffic) -----------------------------------
    toks = [(tok * 31 + 7 * i) % conf.vocab_size for i in range(K)]
    dev.synchronize(); barrier()
    launches1 = dev.launch_count()
    st0 = dev.lazy_stats() if args.lazy else None
    dev.timer_begin(); th0 = time.perf_counter()
    for i in range(K):
        runner.forward([toks[i]], pos, export=False); pos += 1
    host_issue_ms = (time.perf_counter() - th0) * 1e3          # host time to record+submit K tokens (GPU runs behind)
    val_ms = max_over_ranks(dev.timer_end())
    st1 = dev.lazy_stats() if args.lazy else None
    launches_val = dev.launch_count() - launches1
    clocks = sampler.stop()
    barrier()

    # ---- roofline of the dominant kernel: ffn_gate/ffn_up-shaped matvec over all layers' weights (>> L2) ----------
    # (timed through an eager-mode device handle on the same GPU so that each matmul_vec is its own launch pair)
    m, k = weights["ffn_gate"][0].shape()
    dev.synchronize()
    edev = CudaTensorDevice(local_rank, lazy=False) if args.lazy else dev
    x = CudaTensor.new(np.random.default_rng(1).standard_normal(k).astype(np.float32), [k], edev)
    mats = [CudaTensor(wm.buf, wm.strider(), edev) for wm in weights["ffn_gate"] + weights["ffn_up"]]
    for wmat in mats[:4]:
        wmat.matmul_vec(x)
    reps = 3
    edev.synchronize()
    l0 = edev.launch_count()
    edev.timer_begin()
    for _ in range(reps):
        for wmat in mats:
            wmat.matmul_vec(x)
    mv_ms = edev.timer_end()
    n_mv = reps * len(mats)
    launches_per_mv = (edev.launch_count() - l0) / n_mv
    mv_bytes = R.weight_bytes(wt, m, k)
    mv_gbs = mv_bytes / (mv_ms / n_mv * 1e-3) / 1e9
    peaks, peak_src = measured_peaks()

    if rank == 0:
        streams = 1 if sharded or world == 1 else world       # sharded: ONE token stream over N GPUs; replicas: N streams
        value = streams * K / (val_ms * 1e-3)
        e2e = streams * K / (e2e_ms * 1e-3)
        if lazy == 2 and launches_val == K:
            # the dominant kernel IS the step: one mega_kernel launch per token streams every weight byte of this rank once
            mega_gbs = bytes_per_token / (val_ms / K * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "mega_kernel (mega.cu): one persistent launch per decoded token, all matvec/attention/norm phases",
                        "achieved": mega_gbs, "peak": peaks["hbm_gbs"], "peak_source": peak_src, "unit": "GB/s", "frac": mega_gbs / peaks["hbm_gbs"],
                        "traffic": TRAFFIC.get((args.workload, world)), "traffic_source": "profiles/ (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" if TRAFFIC.get((args.workload, world)) else None,
                        "algorithmic_bytes_per_launch": bytes_per_token, "us_per_launch": val_ms / K * 1e3, "frac_of_8TBs_nominal": mega_gbs / 8000.0}
        else:
            roofline = {"bound": "hbm", "kernel": f"matvec_stream_kernel<{wt_name}> {m}x{k} (+ activation quantize: {launches_per_mv:.0f} launches per matmul_vec)",
                        "achieved": mv_gbs, "peak": peaks["hbm_gbs"], "peak_source": peak_src, "unit": "GB/s",
                        "frac": mv_gbs / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes_per_launch": mv_bytes,
                        "us_per_launch": mv_ms / n_mv * 1e3, "frac_of_8TBs_nominal": mv_gbs / 8000.0}
        line = {
            "metric": "decode_tokens_per_s", "value": value, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": val_ms / K, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}-decode-synthetic", "weights": wt_name, "classifier": ct_name, "kv_cache": "f32",
                       "start_pos": args.start_pos, "mode": "lazy: plan not megakernel-eligible (matvec types outside Q8_0/Q4_0 use the warp-per-row kernels): fused kernels, CUDA-graph replay" if (lazy == 2 and launches_val != K) else {0: "eager (one launch per trait call)", 1: "lazy: fused kernels, CUDA-graph replay", 2: "lazy: one persistent megakernel per token, CUDA-graph replay"}[args.lazy],
                       "multi_gpu": ("single GPU" if world == 1 else
                                     f"one token stream sharded over {world} GPUs: rows of wq/wk/wv/gate/up/classifier, block columns of wo/down; "
                                     f"exchange = {'one-shot NVLink peer stores fused into the megakernel' if transport == 'p2p' and lazy == 2 else 'one-shot NVLink peer-store kernel' if transport == 'p2p' else 'ncclAllReduce/ncclAllGather'} "
                                     f"(2 allreduce of [dim] f32 per layer + 1 allgather of logits)" if sharded else f"{world} independent replicas"),
                       "weight_bytes_per_token_per_gpu": bytes_per_token,
                       "l2_policy": f"weights streamed once per token ({bytes_per_token / 1e9:.2f} GB >> 126 MB L2): inputs larger than L2",
                       "weight_bytes_per_token": bytes_per_token,
                       "hbm_frac_whole_step": bytes_per_token / (val_ms / K * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "e2e": {"value": e2e, "unit": "tok/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": conf.vocab_size * 4,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches_e2e),
            "gpu_launches_device_resident": int(launches_val),
            "roofline": roofline,
            "roofline_matvec_eager": {"kernel": f"matvec_stream_kernel<{wt_name}> {m}x{k} + activation quantize ({launches_per_mv:.0f} launches per matmul_vec, eager handle, not in the timed region)",
                                      "achieved": mv_gbs, "unit": "GB/s", "frac": mv_gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": mv_bytes,
                                      "us_per_launch": mv_ms / n_mv * 1e3},
            "clocks": clocks,
            "lazy_stats": dev.lazy_stats() if args.lazy else None,
            "host_ms_per_step": {"issue_total": host_issue_ms / K,
                                 **({k: (st1[k] - st0[k]) / 1e3 / K for k in ("host_us_record", "host_us_fuse", "host_us_submit")} if args.lazy else {})},
        }
        if world == 1 and not args.no_cpu_baseline:
            tps, threads, sample, _ = cpu_reference_tokens_per_s(args.workload)
            line["cpu_baseline"] = {"value": tps, "unit": "tok/s", "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line))
    runner.close()
    dev.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="llama2-7b-q8_0", choices=sorted(WORKLOADS))
    ap.add_argument("--start-pos", type=int, default=32, help="KV-cache length before the timed decode steps")
    ap.add_argument("--lazy", type=int, default=2, help="2 = record+fuse, one persistent megakernel per token (default); 1 = fused kernels in a CUDA graph; 0 = one launch per trait call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "replicas"], help="N > 1: shard one token stream (strong scaling, default) or run N independent replicas")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="exchange transport of the sharded path: one-shot NVLink peer stores (default) or the NCCL baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
