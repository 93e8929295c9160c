// lazy.cu -- lazy mode: record the trait calls of one forward(), fuse, replay as a CUDA graph.
//
// The reference's Tensor API is eager: ~31 calls per layer, ~1000 per token, each a separate launch in
// eager mode.  profiles/r01c shows that per-launch overhead, not kernel bodies, is what caps decode.  In lazy mode the
// SAME C-ABI calls only append to a queue (after the same argument checks); the queue is executed at the first call that
// needs a result on the host (export / debug tap / synchronize -- the reference synchronises only there too, llama2.rs:209):
//   1. fuse: runs of ops that match the Llama decode layer (llama2.rs:226-269, 527-638) are replaced by the kernels of
//      fused.cu / matvec_stream.cu; anything unrecognised falls back to its eager kernel, in order;
//   2. the resulting launch list is hashed (kernel ids, pointers, static sizes).  Values that change every token
//      (position, KV length, token ids, RoPE table) live in a small device buffer `dyn` that the kernels read;
//   3. first time a hash is seen the launches are stream-captured into a CUDA graph (programmatic-dependent-launch edges
//      included); afterwards a token = one 1-2 KB H2D copy of `dyn` + one cudaGraphLaunch.
// The activation pool hands out the same pointers for the same call sequence (lowest free address first), which is what makes the
// hash stable from token to token.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <functional>

#include "common.cuh"

enum LKind { L_COPY_ROWS, L_DUP, L_RMS_NORM, L_MUL, L_ADD, L_SCALE, L_MATVEC, L_ROPE, L_CONCAT, L_CONTIGUOUS, L_BMM, L_SOFTMAX, L_SILU, L_GELU, L_ALLREDUCE, L_ALLGATHER, L_ARGMAX };

struct LView { cc_buf* buf = nullptr; int ndim = 0; int64_t shape[CC_MAX_DIMS] = {0, 0, 0, 0}, strides[CC_MAX_DIMS] = {0, 0, 0, 0}; };

struct LOp {
    int kind;
    LView a, b;              // a: self / lhs / dst ; b: rhs / src
    cc_buf* out = nullptr;   // freshly allocated result (MATVEC, BMM, DUP, CONTIGUOUS)
    float f = 0.0f;
    int64_t i0 = 0, i1 = 0, i2 = 0;
    std::vector<int64_t> rows;
    bool done = false;
};

// one captured token graph.  The cache key is a 64-bit fold of the launch signature; `sig` is the signature itself and is compared
// on every hit (a colliding key must re-capture, never replay another plan's baked pointers); `last_use` drives the LRU bound.
struct GraphEntry { cudaGraphExec_t exec = nullptr; size_t dyn_bytes = 0; uint64_t launches = 0; MkPhase* phases_dev = nullptr;
                    std::vector<uint64_t> sig; uint64_t last_use = 0; int mega_variant = 0; };
#define LZ_MAX_GRAPHS 64                             // cached token graphs per device (a decode loop needs 2-4)

struct LazyState {
    std::vector<LOp> q;
    std::unordered_map<cc_buf*, int> qrefs;
#define LZ_DYN_SLOTS 8                               // tokens the host may run ahead of the GPU (absorbs host scheduling hiccups)
    uint8_t* dyn_host[LZ_DYN_SLOTS] = {nullptr};   // pinned, rotated so the host can build tokens t+1.. while t runs
    cudaEvent_t dyn_ev[LZ_DYN_SLOTS] = {nullptr};
    int dyn_slot = 0;
    uint8_t* dyn_dev = nullptr;
    size_t dyn_cap = 1 << 16;
    void* act[2] = {nullptr, nullptr};
    unsigned* bar_dev = nullptr;     // megakernel grid barrier {count, generation}
    unsigned long long* prof_dev = nullptr;   // CRABML_MEGA_PROF=1: per-phase start timestamps of the last megakernel run
    std::vector<int> prof_types;
    size_t act_cap = 0;
    std::unordered_map<uint64_t, GraphEntry> cache;
    uint64_t flushes = 0, graph_hits = 0, captures = 0, uncached = 0, evictions = 0, collisions = 0;
    uint64_t ns_record = 0, ns_fuse = 0, ns_submit = 0, n_ops = 0;      // host-side cost accounting
    int mega_variant = 0;            // persistent kernel of the last megakernel flush: 1 mega_kernel, 2 mega_ring_kernel
};

static LView mkview(const cc_view* v) {
    LView l;
    if (!v) return l;
    l.buf = v->buf; l.ndim = v->ndim;
    for (int i = 0; i < v->ndim; i++) { l.shape[i] = v->shape[i]; l.strides[i] = v->strides[i]; }
    return l;
}
static int64_t vlen(const LView& v) { int64_t n = 1; for (int i = 0; i < v.ndim; i++) n *= v.shape[i]; return n; }
static bool vcontig(const LView& v) {
    if (v.ndim == 0) return true;
    if (v.strides[v.ndim - 1] != 1) return false;
    int64_t last = 1;
    for (int i = v.ndim - 1; i >= 0; i--) { if (last != v.strides[i]) return false; last *= v.shape[i]; }
    return true;
}

static void graph_entry_free(GraphEntry& ge) {
    if (ge.exec) cudaGraphExecDestroy(ge.exec);
    if (ge.phases_dev) cudaFree(ge.phases_dev);
    ge.exec = nullptr; ge.phases_dev = nullptr;
}
// drop every cached graph (their baked pointers are stale after the scratch buffers moved); the stream must be idle
static void graph_cache_clear(LazyState* lz) {
    for (auto& kv : lz->cache) graph_entry_free(kv.second);
    lz->cache.clear();
}

LazyState* cc_lazy_create(cc_device* dev) {
    LazyState* lz = new LazyState();
    for (int i = 0; i < LZ_DYN_SLOTS; i++)
        if (cudaMallocHost(&lz->dyn_host[i], lz->dyn_cap) != cudaSuccess || cudaEventCreateWithFlags(&lz->dyn_ev[i], cudaEventDisableTiming) != cudaSuccess) { delete lz; return nullptr; }
    if (cudaMalloc(&lz->dyn_dev, lz->dyn_cap) != cudaSuccess) { delete lz; return nullptr; }
    if (getenv("CRABML_MEGA_PROF")) cudaMalloc(&lz->prof_dev, 8 * 8 * 4097);
    if (cudaMalloc(&lz->bar_dev, 4096) != cudaSuccess || cudaMemset(lz->bar_dev, 0, 4096) != cudaSuccess) { delete lz; return nullptr; }
    return lz;
}
void cc_lazy_destroy(cc_device* dev) {
    LazyState* lz = dev->lz;
    if (!lz) return;
    for (auto& op : lz->q) { if (op.a.buf) cc_tensor_release(op.a.buf); if (op.b.buf) cc_tensor_release(op.b.buf); if (op.out) cc_tensor_release(op.out); }
    for (auto& kv : lz->cache) graph_entry_free(kv.second);
    lz->cache.clear();
    if (lz->bar_dev) cudaFree(lz->bar_dev);
    for (int i = 0; i < LZ_DYN_SLOTS; i++) { if (lz->dyn_host[i]) cudaFreeHost(lz->dyn_host[i]); if (lz->dyn_ev[i]) cudaEventDestroy(lz->dyn_ev[i]); }
    if (lz->dyn_dev) cudaFree(lz->dyn_dev);
    for (int i = 0; i < 2; i++) if (lz->act[i]) cudaFree(lz->act[i]);
    delete lz;
    dev->lz = nullptr;
}

// ---- recording ---------------------------------------------------------------------------------------------------
static void hold(LazyState* lz, cc_buf* b) { if (b) { cc_tensor_retain(b); lz->qrefs[b]++; } }

int cc_lazy_record(cc_device* dev, int kind, const cc_view* a, const cc_view* b, cc_buf* out, float f, int64_t i0, int64_t i1, int64_t i2,
                   const int64_t* rows, int n_rows) {
    LazyState* lz = dev->lz;
    auto t0 = std::chrono::steady_clock::now();
    LOp op;
    op.kind = kind; op.a = mkview(a); op.b = mkview(b); op.out = out; op.f = f; op.i0 = i0; op.i1 = i1; op.i2 = i2;
    if (rows) op.rows.assign(rows, rows + n_rows);
    hold(lz, op.a.buf); hold(lz, op.b.buf); hold(lz, op.out);
    lz->q.push_back(std::move(op));
    lz->n_ops++;
    lz->ns_record += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    return CC_OK;
}

// ---- plan building ---------------------------------------------------------------------------------------------------
struct Plan {
    std::vector<uint64_t> sig;
    std::vector<uint8_t> dyn;
    std::vector<std::function<int(uint8_t* dyn_dev)>> steps;
    bool cacheable = true;
    std::vector<MkPhase> phases;     // megakernel form of the same plan (valid while mega_ok)
    bool mega_ok = true;
    size_t mega_smem = 1024, mega_wstage = 0;
    bool mega_ring = false;          // weights through the shared-memory ring (mega_ring.cu)
    int ring_slot = 0, ring_at_ch = 64;
    bool mega_generic = false;       // some MATVEC phase is generic (K-quant weights): launch the instantiation that carries that code
    void S(uint64_t v) { sig.push_back(v); }
    void SP(const void* p) { sig.push_back((uint64_t)(uintptr_t)p); }
    size_t dyn_put(const void* p, size_t n) {
        size_t off = (dyn.size() + 15) & ~(size_t)15;
        dyn.resize(off + n);
        memcpy(dyn.data() + off, p, n);
        return off;
    }
};

static uint64_t hash_sig(const std::vector<uint64_t>& s) {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t v : s) { h ^= v; h *= 1099511628211ull; h ^= h >> 29; }
    return h;
}

struct Fuser {
    cc_device* dev;
    LazyState* lz;
    std::vector<LOp>& q;
    Plan& P;
    size_t rope_off = (size_t)-1; int64_t rope_pos = -1; int rope_hd = 0, rope_dim = 0;

    bool is(size_t i, int kind) const { return i < q.size() && !q[i].done && q[i].kind == kind; }
    std::unordered_map<cc_buf*, size_t> last_use;      // buffer -> index of the last queued op that touches it (built once per flush)
    void index_uses() {
        last_use.reserve(q.size() * 2);
        for (size_t j = 0; j < q.size(); j++) {
            if (q[j].a.buf) last_use[q[j].a.buf] = j;
            if (q[j].b.buf) last_use[q[j].b.buf] = j;
            if (q[j].out) last_use[q[j].out] = j;
        }
    }
    // no op at or after `from` touches buf, and nobody outside the queue holds it
    bool dead_after(cc_buf* b, size_t from) const {
        auto lu = last_use.find(b);
        if (lu != last_use.end() && lu->second >= from) return false;
        auto it = lz->qrefs.find(b);
        int held = it == lz->qrefs.end() ? 0 : it->second;
        return b->refs.load() == held;
    }
    static bool same_dense_1xN(const LView& v, int64_t n) { return vcontig(v) && vlen(v) == n; }

    // ---- eager fallback for one op -----------------------------------------------------------------------------
    bool covered_by_phase = false;     // set while try_generic emits the eager steps of ops its megakernel phase covers
    void fallback(size_t i) {
        if (!covered_by_phase) P.mega_ok = false;
        LOp op = q[i];       // copy: lambdas outlive the queue only until flush ends, but keep them self-contained
        cc_device* d = dev;
        P.S(0x1000 + op.kind); P.SP(op.a.buf ? op.a.buf->plane[0] : nullptr); P.SP(op.b.buf ? op.b.buf->plane[0] : nullptr);
        P.SP(op.out ? op.out->plane[0] : nullptr);
        for (int k = 0; k < CC_MAX_DIMS; k++) { P.S(op.a.shape[k]); P.S(op.a.strides[k]); P.S(op.b.shape[k]); P.S(op.b.strides[k]); }
        P.S(op.i0); P.S(op.i1); P.S(op.i2); uint32_t fb; memcpy(&fb, &op.f, 4); P.S(fb);
        switch (op.kind) {
        case L_ROPE: case L_CONCAT: case L_BMM: case L_SOFTMAX: P.cacheable = false; break;   // position / length dependent
        case L_COPY_ROWS: P.cacheable = false; break;   // (normally taken by try_copy_rows) host row list would be baked into a graph
        case L_MATVEC: if (op.b.ndim > 1 && op.b.shape[0] > 1) P.cacheable = false; break;   // batched (prefill) matmul: its scratch may be (re)allocated, never captured
        default: break;
        }
        for (int64_t r : op.rows) P.S((uint64_t)r);
        P.steps.push_back([d, op](uint8_t*) -> int {
            const LView &a = op.a, &b = op.b;
            switch (op.kind) {
            case L_DUP: {
                int64_t n = vlen(a);
                if (n && cudaMemcpyAsync(op.out->base, a.buf->plane[0], (size_t)n * 4, cudaMemcpyDeviceToDevice, d->stream) != cudaSuccess) return cc_fail(d, CC_ERR_CUDA, "dup copy failed");
                return CC_OK;
            }
            case L_RMS_NORM: return cc_launch_rms_norm(d, (float*)a.buf->plane[0], a.ndim == 1 ? 1 : a.shape[0], a.shape[a.ndim - 1], op.f);
            case L_MUL: return cc_launch_binary(d, (float*)a.buf->plane[0], op.i0, (const float*)b.buf->plane[0], op.i1, 1);
            case L_ADD: return cc_launch_binary(d, (float*)a.buf->plane[0], op.i0, (const float*)b.buf->plane[0], op.i1, 0);
            case L_SCALE: return cc_launch_scale(d, (float*)a.buf->plane[0], vlen(a), op.f);
            case L_SILU: return cc_launch_silu(d, (float*)a.buf->plane[0], vlen(a));
            case L_GELU: return cc_launch_gelu(d, (float*)a.buf->plane[0], vlen(a));
            case L_ALLREDUCE: return cc_launch_all_reduce(d, (float*)a.buf->plane[0], op.i0, nullptr);
            case L_ALLGATHER: return cc_launch_all_gather(d, (const float*)b.buf->plane[0], op.i0, (float*)a.buf->plane[0]);
            case L_ARGMAX: return cc_launch_argmax(d, (const float*)a.buf->plane[0], vlen(a), d->slots + op.i0, d->history, nullptr, op.i1);
            case L_SOFTMAX: { int64_t cols = a.shape[a.ndim - 1]; return cc_launch_softmax(d, (float*)a.buf->plane[0], cols ? vlen(a) / cols : 0, cols); }
            case L_ROPE: return cc_launch_rope_exact(d, (float*)a.buf->plane[0], op.i1, op.i2, a.shape[a.ndim - 1], (int)op.f, op.i0, op.rows[0]);
            case L_CONCAT:
                return cc_launch_strided_copy(d, b.buf->plane[0], b.buf->dtype, b.shape, b.strides, a.buf->plane[0], a.buf->dtype, a.strides,
                                              a.shape[op.i0] * a.strides[op.i0], a.ndim);
            case L_CONTIGUOUS: {
                int64_t dstr[CC_MAX_DIMS]; int64_t s = 1;
                for (int k = a.ndim - 1; k >= 0; k--) { dstr[k] = s; s *= a.shape[k]; }
                return cc_launch_strided_copy(d, a.buf->plane[0], a.buf->dtype, a.shape, a.strides, op.out->base, a.buf->dtype, dstr, 0, a.ndim);
            }
            case L_BMM:
                return cc_launch_batch_matmul(d, (const float*)a.buf->plane[0], b.buf->plane[0], b.buf->dtype, (float*)op.out->base, a.shape[0], b.shape[0],
                                              a.shape[1], a.shape[2], b.shape[2], b.strides[0], b.strides[1], b.strides[2]);
            case L_COPY_ROWS: {
                int n = (int)op.rows.size();
                int rc = cc_ensure_dev_idx(d, (size_t)n * 8);
                if (rc) return rc;
                if (cudaMemcpyAsync(d->dev_idx, op.rows.data(), (size_t)n * 8, cudaMemcpyHostToDevice, d->stream) != cudaSuccess) return cc_fail(d, CC_ERR_CUDA, "row index upload failed");
                return cc_launch_dequant_rows(d, b.buf, (const int64_t*)d->dev_idx, n, a.shape[a.ndim - 1], a.buf->plane[0], a.buf->dtype);
            }
            case L_MATVEC: {
                const int64_t m = a.shape[0], k = a.shape[1], bb = b.ndim == 1 ? 1 : b.shape[0];
                const int wt = a.buf->dtype, at = cc_partner_type(wt);
                const float* xf = (const float*)b.buf->plane[0];
                int rc = CC_OK;
                const bool dense = !(bb == 1 && cc_stream_supported(wt, k)) && cc_prefill_supported(wt, m, k, bb);
                if (at != CC_F32 && !(dense && at == CC_Q8_0)) rc = cc_launch_quantize(d, xf, bb * k, at, d->act_scratch);
                if (rc) return rc;
                if (bb == 1 && cc_stream_supported(wt, k)) return cc_launch_matvec_stream_plain(d, a.buf, d->act_scratch, (float*)op.out->base, m, k);
                if (dense) return cc_launch_prefill_matmul(d, a.buf, d->act_scratch, at == CC_Q8_0 ? xf : nullptr, (float*)op.out->base, m, k, bb);
                return cc_launch_matvec(d, a.buf, d->act_scratch, xf, (float*)op.out->base, m, k, bb);
            }
            }
            return cc_fail(d, CC_ERR_UNSUPPORTED, "lazy: unknown op kind %d", op.kind);
        });
        q[i].done = true;
    }

    // ---- pattern: [DUP] RMS_NORM MUL -> normq ; returns number of ops consumed (0 = no match) ---------------------------------
    // On success *act_sel receives the scratch index that now holds quantize(x).
    size_t try_normq(size_t i, int act_sel, cc_buf** xbuf) {
        size_t j = i;
        cc_buf* orig = nullptr;
        if (is(j, L_DUP) && is(j + 1, L_RMS_NORM) && q[j + 1].a.buf == q[j].a.buf) { orig = q[j].out; j++; }
        if (!(is(j, L_RMS_NORM) && is(j + 1, L_MUL))) return 0;
        const LOp &rn = q[j], &mu = q[j + 1];
        if (mu.a.buf != rn.a.buf) return 0;
        const int64_t n = vlen(rn.a);
        if (rn.a.ndim > 2 || (rn.a.ndim == 2 && rn.a.shape[0] != 1) || !vcontig(rn.a) || n % 32) return 0;
        if (vlen(mu.b) != n || mu.b.buf->dtype != CC_F32 || !vcontig(mu.b)) return 0;
        if (n > 65536) return 0;
        float* x = (float*)rn.a.buf->plane[0];
        float* og = orig ? (float*)orig->base : nullptr;
        const float* w = (const float*)mu.b.buf->plane[0];
        float eps = rn.f;
        void* act = lz->act[act_sel];
        cc_device* d = dev;
        // the normalised f32 row only has to be materialised if something other than the following matvecs reads it
        // (only matvecs the streaming kernel will take consume the quantised scratch; any other reader -- a K-quant or batched
        // matvec falling back to its eager kernel -- needs the f32 row)
        size_t end = j + 2;
        while (is(end, L_MATVEC) && q[end].b.buf == rn.a.buf && cc_stream_supported(q[end].a.buf->dtype, q[end].a.shape[1]) &&
               vlen(q[end].b) == q[end].a.shape[1] && vcontig(q[end].b)) end++;
        const bool write_back = !dead_after(rn.a.buf, end);
        P.S(0x2001); P.SP(x); P.SP(og); P.SP(w); P.SP(act); P.S((uint64_t)n); uint32_t eb; memcpy(&eb, &eps, 4); P.S(eb); P.S(write_back);
        P.steps.push_back([=](uint8_t*) { return cc_launch_normq(d, x, og, w, eps, n, act, write_back); });
        { MkPhase ph = {}; ph.type = MK_NORMQ; ph.write_back = write_back; ph.x = x; ph.orig = og; ph.norm_w = w; ph.eps = eps; ph.n = (int)n; ph.act = cc_act_q8_0(act, n); P.phases.push_back(ph); }
        *xbuf = rn.a.buf;
        size_t used = (j + 2) - i;
        for (size_t t = i; t < i + used; t++) q[t].done = true;
        return used;
    }

    // ---- pattern: 1-3 MATVECs on the same activation through the streaming kernel (+ optional epilogues) ------------------------------
    // `act_sel` >= 0: scratch already holds quantize(x) ; < 0: emit a plain quantise first.
    size_t try_stream(size_t i, cc_buf* xbuf, int act_sel) {
        if (!is(i, L_MATVEC)) return 0;
        const LOp& m0 = q[i];
        const int wt = m0.a.buf->dtype;
        const int64_t k = m0.a.shape[1];
        if (!cc_stream_supported(wt, k) || m0.b.buf != xbuf || vlen(m0.b) != k || !vcontig(m0.b)) return 0;
        // count consecutive matvecs sharing x / type / k
        size_t n = 1;
        while (n < 3 && is(i + n, L_MATVEC) && q[i + n].b.buf == xbuf && q[i + n].a.buf->dtype == wt && q[i + n].a.shape[1] == k && vlen(q[i + n].b) == k) n++;
        StreamArgs A = {};
        A.k = (int)k;
        A.exp_lut = dev->exp_lut;
        size_t used = n;
        // gate/up + silu + mul (llama2.rs:620-630)
        if (n >= 2 && is(i + 2, L_SILU) && is(i + 3, L_MUL) && q[i + 2].a.buf == q[i].out && q[i + 3].a.buf == q[i].out && q[i + 3].b.buf == q[i + 1].out &&
            q[i].a.shape[0] == q[i + 1].a.shape[0] && vlen(q[i + 3].b) == q[i].a.shape[0] && dead_after(q[i + 1].out, i + 4)) {
            n = 2;
            A.epilogue = 2;
            used = 4;
        } else if (n == 1 && is(i + 1, L_ADD) && q[i + 1].a.buf == m0.out && vlen(q[i + 1].b) == m0.a.shape[0] && q[i + 1].b.buf->dtype == CC_F32 && vcontig(q[i + 1].b)) {
            A.epilogue = 1;                    // x = matvec + residual (llama2.rs:266,636)
            A.residual = (const float*)q[i + 1].b.buf->plane[0];
            used = 2;
        } else if (n > 1 && is(i + n, L_SILU)) {
            n = 1; used = 1;                   // do not swallow a gate/up pair we could not fuse as a pair
        }
        // sharded path: column-split matvec -> allreduce [-> + residual]  /  row-split classifier -> allgather (comm.cu)
        int xchg = 0; float* xdst = nullptr; const float* xres = nullptr;
        if (A.epilogue == 0 && is(i + 1, L_ALLREDUCE) && q[i + 1].a.buf == m0.out && q[i + 1].i0 == m0.a.shape[0]) {
            n = 1; xchg = 1; used = 2; xdst = (float*)m0.out->base;
            if (is(i + 2, L_ADD) && q[i + 2].a.buf == m0.out && vlen(q[i + 2].b) == m0.a.shape[0] && q[i + 2].b.buf->dtype == CC_F32 && vcontig(q[i + 2].b)) {
                xres = (const float*)q[i + 2].b.buf->plane[0];
                used = 3;
            }
        } else if (A.epilogue == 0 && is(i + 1, L_ALLGATHER) && q[i + 1].b.buf == m0.out && q[i + 1].i0 == m0.a.shape[0]) {
            n = 1; xchg = 2; used = 2; xdst = (float*)q[i + 1].a.buf->plane[0];
        }
        A.mats.n = (int)n;
        for (size_t t = 0; t < n; t++) {
            A.mats.qs[t] = q[i + t].a.buf->plane[0];
            A.mats.d[t] = (const uint16_t*)q[i + t].a.buf->plane[1];
            A.mats.out[t] = (float*)q[i + t].out->base;
            A.mats.m[t] = (int)q[i + t].a.shape[0];
        }
        cc_device* d = dev;
        void* act = lz->act[act_sel >= 0 ? act_sel : 1];
        A.act = act;
        if (act_sel < 0) {                     // plain quantise of x (matmul_vec.rs:37-40)
            float* x = (float*)m0.b.buf->plane[0];
            P.S(0x2002); P.SP(x); P.SP(act); P.S((uint64_t)k);
            P.steps.push_back([=](uint8_t*) { return cc_launch_normq(d, x, nullptr, nullptr, 0.0f, k, act, false); });
            { MkPhase ph = {}; ph.type = MK_NORMQ; ph.x = x; ph.n = (int)k; ph.act = cc_act_q8_0(act, k); P.phases.push_back(ph); }
        }
        P.S(0x2003); P.S(wt); P.S(k); P.S(A.epilogue); P.SP(A.residual); P.SP(act);
        for (size_t t = 0; t < n; t++) { P.SP(A.mats.qs[t]); P.SP(A.mats.out[t]); P.S(A.mats.m[t]); }
        P.steps.push_back([=](uint8_t*) { return cc_launch_matvec_stream(d, wt, A); });
        { MkPhase ph = {}; ph.type = MK_MATVEC; ph.wtype = wt; ph.mv = A; if (xchg) { ph.mv.epilogue = 3; ph.xgpu = 1; }
          P.phases.push_back(ph); }
        if (xchg) {
            const int64_t mrows = m0.a.shape[0];
            float* part = A.mats.out[0];
            P.S(0x2006); P.S(xchg); P.SP(xdst); P.SP(xres); P.S(mrows);
            if (xchg == 1) P.steps.push_back([=](uint8_t*) { return cc_launch_all_reduce(d, part, mrows, xres); });
            else P.steps.push_back([=](uint8_t*) { return cc_launch_all_gather(d, part, mrows, xdst); });
            MkPhase ph = {}; ph.type = xchg == 1 ? MK_REDUCE : MK_GATHER; ph.red_n = (int)mrows; ph.red_dst = xdst; ph.red_res = xres; P.phases.push_back(ph);
            if (!cc_comm_dev(dev) || cc_comm_is_nccl(dev)) P.mega_ok = false;      // NCCL baseline: graph of kernels + NCCL nodes (lazy mode 1)
            if ((((mrows + dev->sm_count - 1) / dev->sm_count + 3) & ~(int64_t)3) > 512) P.mega_ok = false;   // one CTA's row block must fit the exchange stage (mega.cu MK_XSTAGE_ROWS)
        }
        for (size_t t = i; t < i + used; t++) q[t].done = true;
        return used;
    }

    // ---- pattern: rope q,k + kv append + attention (llama2.rs:252-256, 541-590), n_batch == 1 -----------------------------------------------
    size_t try_attention(size_t i, cc_buf** obuf, bool* quantized) {
        if (!(is(i, L_ROPE) && is(i + 1, L_ROPE) && is(i + 2, L_CONCAT) && is(i + 3, L_CONCAT) && is(i + 4, L_CONTIGUOUS) && is(i + 5, L_SCALE) &&
              is(i + 6, L_BMM) && is(i + 7, L_SOFTMAX) && is(i + 8, L_BMM))) return 0;
        const LOp &rq = q[i], &rk = q[i + 1], &ck = q[i + 2], &cv = q[i + 3], &ct = q[i + 4], &sc = q[i + 5], &b1 = q[i + 6], &sm = q[i + 7], &b2 = q[i + 8];
        if (rq.a.ndim != 3 || rk.a.ndim != 3 || rq.a.shape[0] != 1 || rk.a.shape[0] != 1) return 0;
        const int64_t n_heads = rq.a.shape[1], hd = rq.a.shape[2], n_kv = rk.a.shape[1];
        if (rk.a.shape[2] != hd || (int)rq.f != CC_ROPE_LLAMA || (int)rk.f != CC_ROPE_LLAMA || rq.i0 != rk.i0 || rq.rows[0] != rk.rows[0]) return 0;
        if (hd > 256 || n_heads % n_kv) return 0;
        *quantized = hd % 32 == 0;
        cc_buf *qb = rq.a.buf, *kb = rk.a.buf, *kc = ck.a.buf, *vc = cv.a.buf, *vb = cv.b.buf;
        if (ck.b.buf != kb || ck.i0 != 1 || cv.i0 != 1 || ck.a.ndim != 3 || cv.a.ndim != 3) return 0;
        if (ck.a.shape[0] != n_kv || ck.a.shape[2] != hd || ck.a.strides[2] != 1 || ck.a.strides[1] != hd) return 0;
        if (cv.a.shape[0] != n_kv || cv.a.shape[2] != hd || cv.a.strides[2] != 1 || cv.a.strides[1] != hd || cv.a.strides[0] != ck.a.strides[0]) return 0;
        if (ck.a.shape[1] != cv.a.shape[1] || kc->dtype != vc->dtype) return 0;
        const int64_t kv_len = ck.a.shape[1], seq_stride = ck.a.strides[0];
        // rhs of the concats: [n_kv, 1, hd] views of the raw k / v rows
        if (ck.b.ndim != 3 || ck.b.shape[0] != n_kv || ck.b.shape[1] != 1 || ck.b.shape[2] != hd || ck.b.strides[0] != hd || ck.b.strides[2] != 1) return 0;
        if (cv.b.ndim != 3 || cv.b.shape[0] != n_kv || cv.b.shape[1] != 1 || cv.b.shape[2] != hd || cv.b.strides[0] != hd || cv.b.strides[2] != 1) return 0;
        if (vb->dtype != CC_F32 || kb->dtype != CC_F32 || qb->dtype != CC_F32) return 0;
        if (ct.a.buf != qb || sc.a.buf != ct.out || b1.a.buf != ct.out || b1.b.buf != kc || sm.a.buf != b1.out || b2.a.buf != b1.out || b2.b.buf != vc) return 0;
        if (b1.b.shape[2] != kv_len + 1 || b1.b.strides[1] != 1 || b2.b.shape[1] != kv_len + 1 || b2.b.strides[2] != 1) return 0;
        if (kv_len + 1 > seq_stride / hd) return 0;
        // intermediates must not be observable afterwards
        if (!dead_after(ct.out, i + 9) || !dead_after(b1.out, i + 9) || !dead_after(qb, i + 9) || !dead_after(kb, i + 9) || !dead_after(vb, i + 9)) return 0;
        const int64_t pos = rq.i0, rope_dim = rq.rows[0];
        if (rope_dim > hd || rope_dim % 2) return 0;
        // RoPE table: host libm, identical calls to the reference (rope.rs:47-63); shared by all layers of this flush
        if (rope_off == (size_t)-1 || rope_pos != pos || rope_hd != hd || this->rope_dim != rope_dim) {
            std::vector<float> tab((size_t)rope_dim);
            const int pairs = (int)rope_dim / 2;
            float theta_scale = powf(10000.0f, -2.0f / (float)hd), theta = (float)pos;
            for (int j = 0; j < pairs; j++) { tab[j] = cosf(theta); tab[pairs + j] = sinf(theta); theta *= theta_scale; }
            rope_off = P.dyn_put(tab.data(), tab.size() * 4);
            rope_pos = pos; rope_hd = (int)hd; this->rope_dim = (int)rope_dim;
        }
        int64_t dynv[2] = {pos, kv_len};
        size_t dyn_off = P.dyn_put(dynv, sizeof(dynv));
        AttnArgs A = {};
        A.q = (const float*)qb->plane[0]; A.k = (const float*)kb->plane[0]; A.v = (const float*)vb->plane[0];
        A.kcache = kc->plane[0]; A.vcache = vc->plane[0];
        A.out = (float*)b2.out->base;
        if (*quantized && is(i + 9, L_MATVEC) && !cc_stream_supported(q[i + 9].a.buf->dtype, q[i + 9].a.shape[1])) *quantized = false;   // K-quant wo quantises (Q8_K) itself
        A.act_scratch = *quantized ? lz->act[1] : nullptr;
        A.n_heads = (int)n_heads; A.n_kv = (int)n_kv; A.hd = (int)hd; A.rope_dim = (int)rope_dim;
        A.max_len = (int)(seq_stride / hd); A.kv_f16 = kc->dtype == CC_F16;
        A.seq_stride = seq_stride; A.scale = sc.f;
        cc_device* d = dev;
        size_t roff = rope_off;
        P.S(0x2004); P.SP(A.q); P.SP(A.k); P.SP(A.v); P.SP(A.kcache); P.SP(A.vcache); P.SP(A.out); P.SP(A.act_scratch);
        P.S(n_heads); P.S(n_kv); P.S(hd); P.S(rope_dim); P.S(seq_stride); P.S(A.kv_f16); uint32_t sb; memcpy(&sb, &A.scale, 4); P.S(sb); P.S(dyn_off); P.S(roff);
        P.steps.push_back([=](uint8_t* dyn_dev) {
            AttnArgs B = A;
            B.dyn = (const int64_t*)(dyn_dev + dyn_off);
            B.rope_tab = (const float*)(dyn_dev + roff);
            return cc_launch_attn_decode(d, B);
        });
        { MkPhase ph = {}; ph.type = MK_ATTN; ph.at = A; ph.dyn_off = dyn_off; ph.rope_off = roff; ph.act = cc_act_q8_0(lz->act[1], n_heads * hd);
          P.phases.push_back(ph); }
        *obuf = b2.out;
        for (size_t t = i; t < i + 9; t++) q[t].done = true;
        return 9;
    }

    // ---- pattern: embedding / row pick: COPY_ROWS with the row indices in dyn ---------------------------------------------------------------------
    size_t try_copy_rows(size_t i) {
        if (!is(i, L_COPY_ROWS)) return 0;
        const LOp& op = q[i];
        cc_device* d = dev;
        const cc_buf* src = op.b.buf;
        void* dst = op.a.buf->plane[0];
        int dt = op.a.buf->dtype;
        int64_t cols = op.a.shape[op.a.ndim - 1];
        const int slot = (int)op.i2 - 1;                     // >= 0: the single row index lives in a device slot (cc_copy_rows_from_slot)
        int n = slot >= 0 ? 1 : (int)op.rows.size();
        size_t off = slot >= 0 ? 0 : P.dyn_put(op.rows.data(), op.rows.size() * 8);
        const int64_t* rows_dev = slot >= 0 ? dev->slots + slot : nullptr;
        P.S(0x2005); P.SP(src->plane[0]); P.SP(dst); P.S(dt); P.S(n); P.S(cols); P.S(off); P.SP(rows_dev);
        P.steps.push_back([=](uint8_t* dyn_dev) { return cc_launch_dequant_rows(d, src, rows_dev ? rows_dev : (const int64_t*)(dyn_dev + off), n, cols, dst, dt); });
        { MkPhase ph = {}; ph.type = MK_ROWS; ph.dyn_off = off; for (int t = 0; t < CC_MAX_PLANES; t++) ph.planes.p[t] = src->plane[t];
          ph.planes.cols = src->cols > 0 ? src->cols : cols; ph.src_dtype = src->dtype; ph.dst_dtype = dt; ph.n_rows = n; ph.cols = cols; ph.dst = dst;
          ph.rows_dev = (const long long*)rows_dev; P.phases.push_back(ph); }
        q[i].done = true;
        return 1;
    }

    // ---- pattern: greedy sampling on the device (cc_argmax_to_slot): the history index travels in dyn ------------------------------------------------
    size_t try_argmax(size_t i) {
        if (!is(i, L_ARGMAX)) return 0;
        const LOp& op = q[i];
        cc_device* d = dev;
        const float* x = (const float*)op.a.buf->plane[0];
        const int64_t n = vlen(op.a);
        int64_t* slot = dev->slots + op.i0;
        int64_t* hist = dev->history;
        const int64_t hidx = op.i1;
        size_t off = P.dyn_put(&hidx, 8);
        P.S(0x2007); P.SP(x); P.S((uint64_t)n); P.SP(slot); P.S(off);
        P.steps.push_back([=](uint8_t* dyn_dev) { return cc_launch_argmax(d, x, n, slot, hist, (const int64_t*)(dyn_dev + off), -1); });
        { MkPhase ph = {}; ph.type = MK_ARGMAX; ph.x = (float*)x; ph.n = (int)n; ph.dyn_off = off; ph.slot_dev = (long long*)slot; ph.hist_dev = (long long*)hist; P.phases.push_back(ph); }
        q[i].done = true;
        return 1;
    }

    // ---- pattern: K-quant matvecs (generic MATVEC phase of the megakernel) ----------------------------------------------------------------
    //   [[DUP] RMS_NORM MUL]  MATVEC{1..3, same K-quant type, same f32 row x, b = 1}  [SILU MUL | ADD]
    // In the CUDA-graph mode (lazy = 1) these ops keep running as their eager kernels, in order (fallback steps); for the megakernel
    // the whole group is ONE phase: fused prologue (norm + Q8_K quantisation of x) + T::row_dot rows + epilogue -- the same
    // arithmetic as the eager kernels, so the two modes agree bit for bit.
    size_t try_generic(size_t i) {
        size_t j = i;
        cc_buf* orig = nullptr; cc_buf* xb = nullptr;
        const float* norm_w = nullptr; float eps = 0.0f;
        if (is(j, L_DUP) && is(j + 1, L_RMS_NORM) && q[j + 1].a.buf == q[j].a.buf) { orig = q[j].out; j++; }
        if (is(j, L_RMS_NORM) && is(j + 1, L_MUL) && q[j + 1].a.buf == q[j].a.buf) {
            const LOp &rn = q[j], &mu = q[j + 1];
            const int64_t n = vlen(rn.a);
            if (rn.a.ndim > 2 || (rn.a.ndim == 2 && rn.a.shape[0] != 1) || !vcontig(rn.a)) return 0;
            if (vlen(mu.b) != n || mu.b.buf->dtype != CC_F32 || !vcontig(mu.b)) return 0;
            xb = rn.a.buf; norm_w = (const float*)mu.b.buf->plane[0]; eps = rn.f;
            j += 2;
        } else if (orig) return 0;
        if (!is(j, L_MATVEC)) return 0;
        const LOp& m0 = q[j];
        const int wt = m0.a.buf->dtype;
        const int64_t k = m0.a.shape[1];
        if (!cc_mega_generic_supported(wt, k)) return 0;
        if (!xb) xb = m0.b.buf;
        if (m0.b.buf != xb || xb->dtype != CC_F32 || vlen(m0.b) != k || !vcontig(m0.b) || (m0.b.ndim == 2 && m0.b.shape[0] != 1) || m0.b.ndim > 2) return 0;
        size_t n = 1;
        while (n < 3 && is(j + n, L_MATVEC) && q[j + n].b.buf == xb && q[j + n].a.buf->dtype == wt && q[j + n].a.shape[1] == k && vlen(q[j + n].b) == k) n++;
        StreamArgs A = {};
        A.k = (int)k; A.exp_lut = dev->exp_lut;
        size_t used_mv = n;
        if (n >= 2 && is(j + 2, L_SILU) && is(j + 3, L_MUL) && q[j + 2].a.buf == q[j].out && q[j + 3].a.buf == q[j].out && q[j + 3].b.buf == q[j + 1].out &&
            q[j].a.shape[0] == q[j + 1].a.shape[0] && vlen(q[j + 3].b) == q[j].a.shape[0] && dead_after(q[j + 1].out, j + 4)) {
            n = 2; A.epilogue = 2; used_mv = 4;
        } else if (n == 1 && is(j + 1, L_ADD) && q[j + 1].a.buf == m0.out && vlen(q[j + 1].b) == m0.a.shape[0] && q[j + 1].b.buf->dtype == CC_F32 && vcontig(q[j + 1].b) &&
                   q[j + 1].b.buf != xb) {
            A.epilogue = 1; A.residual = (const float*)q[j + 1].b.buf->plane[0]; used_mv = 2;
        } else if (n > 1 && is(j + n, L_SILU)) { n = 1; used_mv = 1; }
        if (is(j + used_mv, L_ALLREDUCE) || is(j + used_mv, L_ALLGATHER)) return 0;      // sharded K-quant models run in the CUDA-graph mode
        const size_t end = j + used_mv;
        // the normalised row is overwritten in place by the eager ops; the fused prologue never materialises it: nobody else may read it
        if (norm_w && !dead_after(xb, end)) return 0;
        A.mats.n = (int)n;
        for (size_t t = 0; t < n; t++) {
            const cc_buf* w = q[j + t].a.buf;
            if (w->cols != k) return 0;
            A.mats.qs[t] = w->plane[0]; A.mats.d[t] = (const uint16_t*)w->plane[1]; A.mats.p2[t] = w->plane[2]; A.mats.p3[t] = w->plane[3];
            A.mats.out[t] = (float*)q[j + t].out->base;
            A.mats.m[t] = (int)q[j + t].a.shape[0];
        }
        MkPhase ph = {};
        ph.type = MK_MATVEC; ph.wtype = wt; ph.act_type = CC_Q8_K; ph.mv = A;
        ph.x = (float*)xb->plane[0]; ph.orig = orig ? (float*)orig->base : nullptr; ph.norm_w = norm_w; ph.eps = eps; ph.n = (int)k;
        // eager steps for the CUDA-graph mode, op by op, without disqualifying the megakernel form
        covered_by_phase = true;
        for (size_t t = i; t < end; t++) fallback(t);
        covered_by_phase = false;
        P.phases.push_back(ph);
        P.mega_generic = true;
        return end - i;
    }

    // megakernel only: phases[at] = NORMQ (not write-back) directly followed by its single MATVEC consumer -> one MATVEC
    // phase with a fused prologue (saves a grid barrier per merge; see mega.cu phase_matvec)
    void merge_prologue(size_t at) {
        // [at] NORMQ, [at + 1] MATVEC, optionally [at + 2] the REDUCE / GATHER half of the matvec's exchange (sharded path)
        const bool tail = at + 3 == P.phases.size() && (P.phases[at + 2].type == MK_REDUCE || P.phases[at + 2].type == MK_GATHER);
        if (at + 2 != P.phases.size() && !tail) return;
        MkPhase& nq = P.phases[at];
        MkPhase& mv = P.phases[at + 1];
        if (nq.type != MK_NORMQ || mv.type != MK_MATVEC || nq.write_back || nq.n != mv.mv.k) return;
        mv.x = nq.x; mv.orig = nq.orig; mv.norm_w = nq.norm_w; mv.eps = nq.eps; mv.n = nq.n;
        P.phases.erase(P.phases.begin() + at);
        // sharded path: the REDUCE phase that produced x folds into the same prologue (x itself is dead after this group:
        // !write_back), so an exchange costs no phase of its own
        if (at >= 1 && P.phases[at - 1].type == MK_REDUCE && P.phases[at - 1].red_dst == P.phases[at].x && P.phases[at - 1].red_n == P.phases[at].n) {
            MkPhase& m2 = P.phases[at];
            m2.red_n = P.phases[at - 1].red_n; m2.red_res = P.phases[at - 1].red_res;
            P.phases.erase(P.phases.begin() + (at - 1));
        }
    }

    void run() {
        index_uses();
        size_t i = 0;
        while (i < q.size()) {
            if (q[i].done) { i++; continue; }
            size_t used;
            cc_buf* xb = nullptr;
            if ((used = try_copy_rows(i))) { i += used; continue; }
            if ((used = try_argmax(i))) { i += used; continue; }
            if ((used = try_generic(i))) { i += used; continue; }
            if ((used = try_normq(i, 0, &xb))) {
                i += used;
                // every following group of matvecs on the normalised x reuses scratch 0
                const size_t ph0 = P.phases.size();
                int groups = 0;
                while (size_t u2 = try_stream(i, xb, 0)) { i += u2; groups++; }
                if (groups == 1) merge_prologue(ph0 - 1);
                continue;
            }
            bool quantized = false;
            if ((used = try_attention(i, &xb, &quantized))) {
                i += used;
                if (size_t u2 = try_stream(i, xb, quantized ? 1 : -1)) i += u2;      // wo consumes the (quantised) attention output
                continue;
            }
            if (is(i, L_MATVEC) && q[i].b.buf->dtype == CC_F32 && (q[i].b.ndim == 1 || q[i].b.shape[0] == 1)) {
                const size_t ph0 = P.phases.size();
                if ((used = try_stream(i, q[i].b.buf, -1))) { i += used; merge_prologue(ph0); continue; }
            }
            fallback(i);
            i++;
        }
    }
};

int cc_lazy_flush(cc_device* dev) {
    LazyState* lz = dev->lz;
    if (!lz || lz->q.empty()) return CC_OK;
    lz->flushes++;
    // scratch for the quantised activations (largest k in the queue)
    int64_t max_k = 0;
    for (auto& op : lz->q) if (op.kind == L_MATVEC) max_k = std::max<int64_t>(max_k, std::max<int64_t>(op.a.shape[1], 0));
    // the fused attention quantises its OUTPUT row (n_heads * head_dim = the PV product's a_batch * n) into act[1]; QK^T products
    // ([heads, 1, kv_len + 1]) are never quantised, so the scratch does not grow with the context
    for (auto& op : lz->q) if (op.kind == L_BMM && op.b.ndim == 3 && op.b.strides[2] == 1) max_k = std::max<int64_t>(max_k, op.a.shape[0] * op.b.shape[2]);
    size_t need = cc_act_bytes(CC_Q8_0, (max_k + 255) / 256 * 256) + 256;
    int rc = CC_OK;
    if (need > lz->act_cap) {
        cudaStreamSynchronize(dev->stream);
        graph_cache_clear(lz);                       // cached graphs hold the old scratch pointers
        for (int i = 0; i < 2; i++) { if (lz->act[i]) cudaFree(lz->act[i]); lz->act[i] = nullptr; }
        size_t cap = (size_t)1 << 20; while (cap < need) cap <<= 1;      // generous from the start: growing means cudaFree (context-wide wait)
        for (int i = 0; i < 2; i++) if (cudaMalloc(&lz->act[i], cap) != cudaSuccess) return cc_fail(dev, CC_ERR_CUDA, "lazy: scratch alloc failed");
        lz->act_cap = cap;
    }
    // eager matvec fallbacks use dev->act_scratch: make sure it is large enough BEFORE any capture
    for (auto& op : lz->q) if (op.kind == L_MATVEC) {
        int at = cc_partner_type(op.a.buf->dtype);
        int64_t bb = op.b.ndim == 1 ? 1 : op.b.shape[0];
        if (at != CC_F32 && (rc = cc_ensure_act_scratch(dev, cc_act_bytes(at, bb * op.a.shape[1])))) return rc;
    }
    auto t_f0 = std::chrono::steady_clock::now();
    Plan P;
    Fuser F{dev, lz, lz->q, P};
    F.run();
    auto t_f1 = std::chrono::steady_clock::now();
    lz->ns_fuse += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_f1 - t_f0).count();
    if (P.dyn.size() > lz->dyn_cap) P.cacheable = false;

    auto run_steps = [&](uint8_t* dyn_dev) -> int {
        for (auto& st : P.steps) { int r = st(dyn_dev); if (r) return r; }
        return CC_OK;
    };
    // per-token values: pinned slot -> device block, as an ordinary stream copy in front of the launches / the graph
    if (P.dyn.size() > lz->dyn_cap) rc = cc_fail(dev, CC_ERR_UNSUPPORTED, "lazy: dynamic argument block too large");
    if (!rc && !P.dyn.empty()) {
        const int s = lz->dyn_slot;
        lz->dyn_slot = (s + 1) % LZ_DYN_SLOTS;
        cudaEventSynchronize(lz->dyn_ev[s]);          // the copy that last read this slot (LZ_DYN_SLOTS flushes ago) is long done
        memcpy(lz->dyn_host[s], P.dyn.data(), P.dyn.size());
        if (cudaMemcpyAsync(lz->dyn_dev, lz->dyn_host[s], P.dyn.size(), cudaMemcpyHostToDevice, dev->stream) != cudaSuccess) rc = cc_fail(dev, CC_ERR_CUDA, "lazy: dyn upload failed");
        cudaEventRecord(lz->dyn_ev[s], dev->stream);
    }
    if (rc) {
    } else if (!P.cacheable) {
        lz->uncached++;
        rc = run_steps(lz->dyn_dev);
    } else {
        P.S(P.dyn.size());
        uint64_t key = hash_sig(P.sig);
        auto it = lz->cache.find(key);
        if (it != lz->cache.end() && it->second.sig != P.sig) {          // 64-bit key collision: never replay the other plan's graph
            lz->collisions++;
            cudaStreamSynchronize(dev->stream);
            graph_entry_free(it->second);
            lz->cache.erase(it);
            it = lz->cache.end();
        }
        if (it == lz->cache.end() && lz->cache.size() >= LZ_MAX_GRAPHS) {   // LRU bound: evict the entry replayed longest ago
            auto victim = lz->cache.begin();
            for (auto c = lz->cache.begin(); c != lz->cache.end(); ++c) if (c->second.last_use < victim->second.last_use) victim = c;
            cudaStreamSynchronize(dev->stream);          // its last launch may still be running
            graph_entry_free(victim->second);
            lz->cache.erase(victim);
            lz->evictions++;
        }
        if (it == lz->cache.end()) {
            lz->captures++;
            uint64_t l0 = dev->launches;
            cudaGraph_t graph = nullptr;
            GraphEntry ge;
            bool use_mega = dev->mega && P.mega_ok && !P.phases.empty();
            if (use_mega) {       // a phase whose working area cannot fit beside anything (e.g. the score row of a 32 K-token context): CUDA-graph mode
                size_t work = 0, wst = 0;
                for (auto& ph : P.phases) {
                    work = std::max(work, cc_mega_smem_for_phase(ph));
                    if (ph.type == MK_MATVEC && ph.x && ph.norm_w) wst = std::max(wst, (size_t)ph.n * 4);
                }
                if (work + wst + 4096 > 227 * 1024) use_mega = false;
            }
            if (use_mega) {       // phase table lives in device memory for the lifetime of the graph
                int nxt = -1, nxt2 = -1;
                for (int t = (int)P.phases.size() - 1; t >= 0; t--) {
                    P.phases[t].next_matvec = nxt; P.phases[t].next_matvec2 = nxt2;
                    if (P.phases[t].type == MK_MATVEC) { nxt2 = nxt; nxt = t; }
                }
                // weights through the shared-memory ring (mega_ring.cu) when every streaming MATVEC phase can be fed by bulk copies
                bool ring = cc_mega_ring_enabled(), any_stream = false;
                for (auto& ph : P.phases) {
                    if (ph.type == MK_MATVEC && ph.act_type != CC_Q8_K) any_stream = true;
                    if (!cc_mega_ring_phase_ok(ph)) ring = false;
                }
                P.mega_ring = ring && any_stream;
                if (P.mega_ring) {              // is there room for a useful ring beside the working area?  else: the register-pipe kernel
                    size_t work = 0, wst = 0; int slot = 0;
                    for (auto& ph : P.phases) {
                        work = std::max(work, cc_mega_ring_smem_for_phase(ph));
                        if (ph.type == MK_MATVEC && ph.act_type == CC_Q8_K && ph.x && ph.norm_w) wst = std::max(wst, (size_t)ph.n * 4);
                        if (ph.type == MK_MATVEC && ph.act_type != CC_Q8_K) slot = std::max(slot, ph.wtype == CC_Q8_0 ? 4352 : 2304);
                    }
                    if (!cc_mega_ring_fits(work, wst, slot, P.mega_generic)) P.mega_ring = false;
                }
                for (auto& ph : P.phases) {
                    if (P.mega_ring) {
                        P.mega_smem = std::max(P.mega_smem, cc_mega_ring_smem_for_phase(ph));
                        if (ph.type == MK_MATVEC && ph.act_type == CC_Q8_K && ph.x && ph.norm_w) P.mega_wstage = std::max(P.mega_wstage, (size_t)ph.n * 4);
                        if (ph.type == MK_MATVEC && ph.act_type != CC_Q8_K) P.ring_slot = std::max(P.ring_slot, ph.wtype == CC_Q8_0 ? 4352 : 2304);
                        if (ph.type == MK_ATTN) P.ring_at_ch = cc_mega_ring_at_ch(ph);
                    } else {
                        P.mega_smem = std::max(P.mega_smem, cc_mega_smem_for_phase(ph));
                        if (ph.type == MK_MATVEC && ph.x && ph.norm_w) P.mega_wstage = std::max(P.mega_wstage, (size_t)ph.n * 4);
                    }
                }
                if (cudaMalloc(&ge.phases_dev, P.phases.size() * sizeof(MkPhase)) != cudaSuccess ||
                    cudaMemcpy(ge.phases_dev, P.phases.data(), P.phases.size() * sizeof(MkPhase), cudaMemcpyHostToDevice) != cudaSuccess)
                    rc = cc_fail(dev, CC_ERR_CUDA, "lazy: phase table upload failed");
            }
            cudaError_t e = rc ? cudaSuccess : cudaStreamBeginCapture(dev->stream, cudaStreamCaptureModeRelaxed);
            if (!rc && e != cudaSuccess) rc = cc_fail(dev, CC_ERR_CUDA, "lazy: begin capture: %s", cudaGetErrorString(e));
            if (!rc) {
                if (use_mega && lz->prof_dev && P.phases.size() < 4000) { lz->prof_types.clear(); for (auto& ph : P.phases) lz->prof_types.push_back(ph.type * 16 + (ph.type == MK_MATVEC ? ph.mv.mats.n + 4 * ph.mv.epilogue + 1024 * (ph.mv.k >> 10) : 0)); }
                unsigned long long* prof = P.phases.size() < 4000 ? lz->prof_dev : nullptr;
                if (!use_mega) rc = run_steps(lz->dyn_dev);
                else if (P.mega_ring) rc = cc_launch_mega_ring(dev, ge.phases_dev, (int)P.phases.size(), lz->dyn_dev, lz->bar_dev, P.mega_smem, P.mega_wstage, prof, cc_comm_dev(dev),
                                                               P.mega_generic, P.ring_slot, P.ring_at_ch, cc_mega_flags());
                else rc = cc_launch_mega(dev, ge.phases_dev, (int)P.phases.size(), lz->dyn_dev, lz->bar_dev, P.mega_smem, P.mega_wstage, prof, cc_comm_dev(dev), P.mega_generic);
                e = cudaStreamEndCapture(dev->stream, &graph);
                if (!rc && e != cudaSuccess) rc = cc_fail(dev, CC_ERR_CUDA, "lazy: end capture: %s", cudaGetErrorString(e));
            }
            if (!rc) {
                e = cudaGraphInstantiate(&ge.exec, graph, 0);
                if (e != cudaSuccess) rc = cc_fail(dev, CC_ERR_CUDA, "lazy: graph instantiate: %s", cudaGetErrorString(e));
            }
            if (graph) cudaGraphDestroy(graph);
            if (!rc) {
                ge.mega_variant = use_mega ? (P.mega_ring ? 2 : 1) : 0;
                ge.dyn_bytes = P.dyn.size();
                ge.sig = P.sig;
                ge.launches = dev->launches - l0;
                dev->launches = l0;            // counted when the graph is launched
                it = lz->cache.emplace(key, ge).first;
            }
        } else {
            lz->graph_hits++;
        }
        if (!rc) {
            it->second.last_use = lz->flushes;
            if (it->second.mega_variant) lz->mega_variant = it->second.mega_variant;
            cudaError_t e = cudaGraphLaunch(it->second.exec, dev->stream);
            if (e != cudaSuccess) rc = cc_fail(dev, CC_ERR_CUDA, "lazy: graph launch: %s", cudaGetErrorString(e));
            else dev->launches += it->second.launches;
        }
    }
    lz->ns_submit += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_f1).count();
    // drop the queue's references, last op first (keeps pool pointer assignment identical from token to token)
    std::vector<LOp> old;
    old.swap(lz->q);
    lz->qrefs.clear();
    for (size_t t = old.size(); t-- > 0;) {
        LOp& op = old[t];
        if (op.out) cc_tensor_release(op.out);
        if (op.b.buf) cc_tensor_release(op.b.buf);
        if (op.a.buf) cc_tensor_release(op.a.buf);
    }
    return rc;
}

// Called after a stream synchronize.  A persistent kernel that gave up on a barrier (MkSpin in mega.cu) has raised the host-mapped
// error word: report it, and reset the barrier words (their monotonic counters are inconsistent after a drained launch).
int cc_check_async_error(cc_device* dev) {
    if (!dev || !dev->err_host) return CC_OK;
    const unsigned code = *(volatile unsigned*)dev->err_host;
    if (!code) return CC_OK;
    *(volatile unsigned*)dev->err_host = 0u;
    if (dev->err_dev) cudaMemset(dev->err_dev, 0, 256);
    if (dev->lz && dev->lz->bar_dev) cudaMemset(dev->lz->bar_dev, 0, 4096);
    return cc_fail(dev, CC_ERR_CUDA, "%s timeout (%s): the grid was not co-resident or a peer GPU stopped responding",
                   code == 3u ? "exchange kernel" : "megakernel barrier", code == 1u ? "grid barrier" : code == 4u ? "weight ring" : "cross-GPU handshake");
}

// developer profiling: per-phase start timestamps (ns) of the last megakernel run + phase type codes
extern "C" CC_API int cc_lazy_mega_profile(cc_device* dev, unsigned long long* ts, int* types, int cap, int* n_out) {
    if (!dev || !dev->lz || !dev->lz->prof_dev || !ts || !types || !n_out) return CC_ERR_ARG;
    cudaStreamSynchronize(dev->stream);
    int n = (int)dev->lz->prof_types.size();
    if ((n + 1) * 8 > cap) return CC_ERR_ARG;          // 8 stamps per phase (mega.cu MK_PROF_SLOTS)
    if (cudaMemcpy(ts, dev->lz->prof_dev, (size_t)(n + 1) * 8 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return CC_ERR_CUDA;
    for (int i = 0; i < n; i++) types[i] = dev->lz->prof_types[i];
    *n_out = n;
    return CC_OK;
}

extern "C" CC_API int cc_lazy_mega_variant(cc_device* dev) { return dev && dev->lz ? dev->lz->mega_variant : 0; }

extern "C" CC_API int cc_lazy_stats(cc_device* dev, uint64_t* out4) {
    if (!dev || !dev->lz || !out4) return CC_ERR_ARG;
    out4[0] = dev->lz->flushes; out4[1] = dev->lz->graph_hits; out4[2] = dev->lz->captures; out4[3] = dev->lz->uncached;
    out4[4] = dev->lz->ns_record; out4[5] = dev->lz->ns_fuse; out4[6] = dev->lz->ns_submit; out4[7] = dev->lz->n_ops;
    return CC_OK;
}
