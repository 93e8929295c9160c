"""Sharded decode path on real GPUs.  The one-process tests run the exchange protocol with a world of 1 (the
peer stores, flag handshake and slot reduction all execute, against this GPU's own window), so they also run on a
single-GPU box; the N-process test needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tests.gpu_common import make_device

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("lazy", [0, 1, 2])
def test_exchange_ops_world_of_one(lazy):
    """matvec -> all_reduce -> + residual and matvec -> all_gather through every execution mode: with one rank the
    exchange is the identity, so the result must equal the plain ops bit for bit."""
    from crabml_b200 import CudaTensor, capi
    from crabml_b200.runner import synth_scale
    ref_dev = make_device(lazy=0)
    dev = make_device(lazy=lazy)
    try:
        dev.init_comm(0, 1)
        rng = np.random.default_rng(3)
        k, m = 4096, 1024
        xs = [rng.standard_normal(k).astype(np.float32) for _ in range(4)]
        res = rng.standard_normal(m).astype(np.float32)
        outs = {}
        for name, d in (("ref", ref_dev), ("dut", dev)):
            w = CudaTensor.synth([m, k], capi.Q8_0, d, 5, 1, synth_scale(capi.Q8_0, k))
            got = []
            for x in xs:           # several rounds: sequence numbers / slot parity advance, graphs are replayed
                xt = CudaTensor.new(x, [k], d)
                r = CudaTensor.new(res, [m], d)
                y = w.matmul_vec(xt)
                if name == "dut":
                    y = y.all_reduce_sum_inplace()
                y = y.add_inplace(r)
                z = w.matmul_vec(xt)
                if name == "dut":
                    z = CudaTensor.alloc([m], capi.F32, d).all_gather_from(z)
                got.append(np.concatenate([y.export(), z.export()]))
            outs[name] = np.stack(got)
        ref, dut = outs["ref"], outs["dut"]
        tol = 1e-6 * float(np.abs(ref).max()) * 64      # eager ref uses the warp-per-row kernel, lazy the streaming one
        np.testing.assert_allclose(dut, ref, rtol=0, atol=tol)
        if lazy:
            st = dev.lazy_stats()
            assert st["uncached"] == 0, st
    finally:
        dev.close()
        ref_dev.close()


def test_exchange_requires_a_communicator():
    from crabml_b200 import CudaTensor, TensorError
    dev = make_device()
    try:
        with pytest.raises(TensorError):
            CudaTensor.new(np.zeros(32, np.float32), [32], dev).all_reduce_sum_inplace()
    finally:
        dev.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_gpus_sharded_llama_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "sharded_worker.py"), "--mode", "gpu"]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]


def _run_ranks(fns):
    """one host thread per rank (ctypes releases the GIL inside the library): the ranks' kernels must be in flight together"""
    import threading
    out, err = [None] * len(fns), [None] * len(fns)

    def wrap(i):
        try:
            out[i] = fns[i]()
        except BaseException as e:      # noqa: BLE001
            err[i] = e
    ts = [threading.Thread(target=wrap, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    for e in err:
        if e is not None:
            raise e
    assert all(not t.is_alive() for t in ts), "a rank is stuck"
    return out


def _world_of_two_on_one_gpu(lazy):
    import torch
    devs = [make_device(lazy=lazy) for _ in range(2)]
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    for r, d in enumerate(devs):
        d.set_sm_limit(sm // 2)             # two persistent kernels side by side on one GPU: half of the SMs each
        d.create_comm_local(r, 2)
    for d in devs:
        d.connect_comm_local(devs)
    return devs


@pytest.mark.parametrize("lazy", [0])
def test_world_of_two_on_one_gpu_exchange_is_the_rank_ordered_sum(lazy):
    """The REAL two-rank protocol on a single GPU (two devices of this process, half of the SMs each, windows wired in process):
    column-split matvec -> all_reduce (+ residual) and row-split matvec -> all_gather in every execution mode, including the
    exchange fused into the megakernel (coalesced peer stores from the matvec epilogue, handshake on the grid barrier, reduction in
    the consumer's prologue).  Expected = the partial rows of the eager single-device kernels added in rank order, bit for bit, and
    both ranks must hold identical bits.
    In-process ranks share one CUDA context, where graph instantiation / first-use allocations of one rank can wait for the other
    rank's spinning exchange kernel: this in-process form runs the eager mode only; the CUDA-graph and megakernel modes of the
    two-rank protocol run as two PROCESSES on the one GPU (test_two_processes_on_one_gpu_sharded_llama below)."""
    from crabml_b200 import CudaTensor, capi
    from crabml_b200.runner import synth_scale
    k, m = 4096, 4096
    rng = np.random.default_rng(11)
    xs = [rng.standard_normal(k).astype(np.float32) for _ in range(5)]
    res = rng.standard_normal(m).astype(np.float32)
    nw = (1.0 + 0.05 * rng.standard_normal(m)).astype(np.float32)
    # reference partials on a plain eager device
    ref = make_device(lazy=0)
    try:
        ws = [CudaTensor.synth([m, k], capi.Q8_0, ref, 5, r + 1, synth_scale(capi.Q8_0, k)) for r in range(2)]
        w2 = CudaTensor.synth([64, m], capi.Q8_0, ref, 5, 9, synth_scale(capi.Q8_0, m))
        want = []
        for x in xs:
            xt = CudaTensor.new(x, [k], ref)
            p = [w.matmul_vec(xt).export().copy() for w in ws]
            y = (p[0] + p[1]) + res                                   # rank order, then the residual (comm.cu / mega.cu prologue)
            yt = CudaTensor.new(y, [m], ref)
            z = w2.matmul_vec(yt.dup().rms_norm_inplace(1e-5).mul_inplace(CudaTensor.new(nw, [m], ref))).export().copy()
            want.append((y, np.concatenate(p), z))
    finally:
        ref.close()
    devs = _world_of_two_on_one_gpu(lazy)
    try:
        # inputs are created up front: the two ranks share ONE CUDA context here, and freeing a non-pooled buffer (cudaFree) waits for
        # every kernel of the context -- including the other rank's exchange kernel that is spinning for this rank (separate processes,
        # the real deployment, do not have that coupling)
        inputs = []
        for r, d in enumerate(devs):
            inputs.append(dict(w=CudaTensor.synth([m, k], capi.Q8_0, d, 5, r + 1, synth_scale(capi.Q8_0, k)),
                               w2=CudaTensor.synth([64, m], capi.Q8_0, d, 5, 9, synth_scale(capi.Q8_0, m)),
                               nw=CudaTensor.new(nw, [m], d), res=CudaTensor.new(res, [m], d), xs=[CudaTensor.new(x, [k], d) for x in xs]))
            d.synchronize()

        def rank_fn(r):
            def run():
                d = devs[r]
                w, w2r, nwt, rest = inputs[r]["w"], inputs[r]["w2"], inputs[r]["nw"], inputs[r]["res"]
                got = []
                for xt in inputs[r]["xs"]:
                    y = w.matmul_vec(xt).all_reduce_sum_inplace().add_inplace(rest)
                    # a consumer of the reduced row in the decode layer's own shape (llama2.rs:227-232: dup, rms_norm, mul, matmul_vec):
                    # the megakernel folds the rank-ordered reduction into that consumer's fused norm + quantise prologue
                    yo = y.dup()
                    z = w2r.matmul_vec(y.rms_norm_inplace(1e-5).mul_inplace(nwt))
                    g = CudaTensor.alloc([2 * m], capi.F32, d).all_gather_from(w.matmul_vec(xt))
                    got.append((yo.export().copy(), g.export().copy(), z.export().copy()))
                return got
            return run
        got = _run_ranks([rank_fn(0), rank_fn(1)])
        for i in range(len(xs)):
            for r in range(2):
                for a, b, what in zip(got[r][i], want[i], ("allreduce+residual", "allgather", "consumer of the reduced row")):
                    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=f"lazy={lazy} round {i} rank {r}: {what}")
        if lazy:
            for d in devs:
                assert d.lazy_stats()["uncached"] == 0
    finally:
        for d in devs:
            d.close()


def test_two_processes_on_one_gpu_sharded_llama():
    """World of 2 on ONE GPU, one process per rank (time-sliced contexts, peers' windows mapped through CUDA IPC exactly as across
    GPUs): the exchange ops bit-exact against the rank-ordered numpy sum, and the 2-layer Llama-2-7B-shaped sharded model in the
    eager, CUDA-graph and megakernel modes -- ranks bit-identical, modes bit-identical, close to the unsharded logits.
    This is the multi-GPU path's parity test on a single-GPU lease (tests/sharded_worker.py --one-gpu)."""
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "sharded_worker.py"), "--mode", "gpu", "--one-gpu"]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert "sharded parity vs single GPU" in p.stdout
