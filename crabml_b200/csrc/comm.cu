// comm.cu -- the exchange step of the row/column-sharded decode path: one allreduce(sum) of a [dim] f32
// row after `wo` and after `ffn_down`, one allgather of the classifier's logit slices.  One process per GPU.
//
// The reference has no multi-device path; what shards is its thread-pool split of output rows (matmul_vec.rs:41-76).
//
// Two transports behind the same two entry points (cc_all_reduce_sum_inplace / cc_all_gather):
//   * p2p (default): ONE-SHOT exchange over NVLink peer memory.  Every rank owns a cudaMalloc'd window that all peers map
//     with CUDA IPC.  A rank STORES its 16 KB partial straight into slot[rank] of every peer's window, publishes a sequence
//     number in each peer's flag word (st.release.sys), polls its own flag words (ld.acquire.sys) and sums the slots in rank
//     order 0..N-1 -- every rank adds in the same order, so the replicated activation stays bit-identical on all ranks.
//     Data slots are double buffered by sequence parity: a peer can only be one exchange ahead (it needs everybody's flag
//     of exchange e before it starts e+1), so parity (e+1)&1 is never still being read.
//     The same protocol runs INSIDE the megakernel (mega.cu): the matvec epilogue stores partial rows directly into the
//     peers' slots and the cross-GPU flag handshake rides on the phase's grid barrier -- compute and collective in one kernel.
//   * nccl (baseline, CRABML_COMM=nccl): ncclAllReduce / ncclAllGather on the device's stream, libnccl.so.2 resolved with
//     dlopen so that libcrabml_cuda has no link-time dependency on it.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

// ---- minimal NCCL surface (nccl.h 2.x ABI; we only need five entry points) ------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclFloat32 = 7, ncclSumOp = 0 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static bool nccl_load(cc_device* dev) {
    if (g_nccl.lib) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy torch already mapped, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { cc_fail(dev, CC_ERR_UNSUPPORTED, "libnccl.so.2 not found: %s", dlerror()); return false; }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather) {
        cc_fail(dev, CC_ERR_UNSUPPORTED, "libnccl.so.2 lacks the expected entry points");
        return false;
    }
    g_nccl.lib = h;
    return true;
}

// ---- window layout -------------------------------------------------------------------------------------------------------------
//   [0, 4096)        flags: word [src_rank * 32] = last sequence number published by src_rank (one 128 B line per source)
//   [4096, 8192)     seq:   word 0 = sequence number of the last finished exchange of THIS rank
//   [8192, ...)      data:  [parity 2][src_rank 8][CC_COMM_MAX_ELEMS] f32
static const size_t COMM_DATA_OFF = 8192;
static const size_t COMM_BYTES = COMM_DATA_OFF + (size_t)2 * CC_COMM_MAX_RANKS * CC_COMM_MAX_ELEMS * 4;

struct cc_comm {
    int rank = 0, world = 1;
    uint8_t* local = nullptr;
    uint8_t* peer[CC_COMM_MAX_RANKS] = {nullptr};
    bool connected = false;
    bool local_peers = false;      // peers are devices of this process (cc_comm_connect_local): nothing to unmap
    ncclComm_t nccl = nullptr;
    CommDev cd = {};
};

const CommDev* cc_comm_dev(cc_device* dev) { return dev->comm && dev->comm->connected ? &dev->comm->cd : nullptr; }
bool cc_comm_is_nccl(cc_device* dev) { return dev->comm && dev->comm->nccl; }
int cc_comm_world(cc_device* dev) { return dev->comm ? dev->comm->world : 1; }

extern "C" CC_API int cc_comm_create(cc_device* dev, int32_t rank, int32_t world, uint8_t* handle_out /* 64 bytes */) {
    if (!dev || !handle_out) return cc_fail(dev, CC_ERR_ARG, "cc_comm_create: bad argument");
    CC_ENTER(dev);
    CC_REQUIRE(dev, world >= 1 && world <= CC_COMM_MAX_RANKS && rank >= 0 && rank < world, "comm: rank %d of %d unsupported (max %d ranks)", rank, world, CC_COMM_MAX_RANKS);
    CC_REQUIRE(dev, !dev->comm, "comm: already created on this device");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cc_comm* c = new cc_comm();
    c->rank = rank; c->world = world;
    cudaError_t e = cudaMalloc(&c->local, COMM_BYTES);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, COMM_BYTES);
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) { if (c->local) cudaFree(c->local); delete c; return cc_fail(dev, CC_ERR_CUDA, "comm window: %s", cudaGetErrorString(e)); }
    memcpy(handle_out, &h, 64);
    dev->comm = c;
    return CC_OK;
}

// handles: world x 64 bytes, handles[rank] ignored.  After this call the window of every peer is mapped.
extern "C" CC_API int cc_comm_connect(cc_device* dev, const uint8_t* handles) {
    if (!dev || !dev->comm || !handles) return cc_fail(dev, CC_ERR_ARG, "cc_comm_connect: create the communicator first");
    CC_ENTER(dev);
    cc_comm* c = dev->comm;
    for (int p = 0; p < c->world; p++) {
        if (p == c->rank) { c->peer[p] = c->local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)p * 64, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cc_fail(dev, CC_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
        c->peer[p] = (uint8_t*)ptr;
    }
    c->cd.rank = c->rank; c->cd.world = c->world;
    for (int p = 0; p < c->world; p++) { c->cd.flag[p] = (unsigned*)c->peer[p]; c->cd.data[p] = (float*)(c->peer[p] + COMM_DATA_OFF); }
    c->cd.seq = (unsigned*)(c->local + 4096);
    c->connected = true;
    return CC_OK;
}

extern "C" CC_API int cc_comm_nccl_unique_id(cc_device* dev, uint8_t* id_out /* 128 bytes */) {
    if (!dev || !id_out) return cc_fail(dev, CC_ERR_ARG, "cc_comm_nccl_unique_id: bad argument");
    if (!nccl_load(dev)) return CC_ERR_UNSUPPORTED;
    ncclUniqueId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc) return cc_fail(dev, CC_ERR_CUDA, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    memcpy(id_out, &id, 128);
    return CC_OK;
}

// The same wiring for ranks that live in ONE process (one cc_device per rank; on different GPUs with peer access, or on the same
// GPU for the single-GPU world-of-2 test): the peers' windows are ordinary device pointers, no IPC handle is needed.
extern "C" CC_API int cc_comm_connect_local(cc_device* dev, cc_device* const* peers) {
    if (!dev || !dev->comm || !peers) return cc_fail(dev, CC_ERR_ARG, "cc_comm_connect_local: create the communicator first");
    CC_ENTER(dev);
    cc_comm* c = dev->comm;
    for (int p = 0; p < c->world; p++) {
        CC_REQUIRE(dev, peers[p] && peers[p]->comm && peers[p]->comm->world == c->world && peers[p]->comm->rank == p, "cc_comm_connect_local: peer %d has no matching communicator", p);
        if (peers[p]->ordinal != dev->ordinal) {
            cudaError_t e = cudaDeviceEnablePeerAccess(peers[p]->ordinal, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cc_fail(dev, CC_ERR_CUDA, "peer access to device %d: %s", peers[p]->ordinal, cudaGetErrorString(e));
            cudaGetLastError();
        }
        c->peer[p] = peers[p]->comm->local;
    }
    c->cd.rank = c->rank; c->cd.world = c->world;
    for (int p = 0; p < c->world; p++) { c->cd.flag[p] = (unsigned*)c->peer[p]; c->cd.data[p] = (float*)(c->peer[p] + COMM_DATA_OFF); }
    c->cd.seq = (unsigned*)(c->local + 4096);
    c->connected = true;
    c->local_peers = true;
    return CC_OK;
}

// NCCL baseline transport: every rank passes rank 0's unique id (collective call).
extern "C" CC_API int cc_comm_init_nccl(cc_device* dev, const uint8_t* id) {
    if (!dev || !dev->comm || !id) return cc_fail(dev, CC_ERR_ARG, "cc_comm_init_nccl: create the communicator first");
    if (!nccl_load(dev)) return CC_ERR_UNSUPPORTED;
    CC_ENTER(dev);
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    int rc = g_nccl.CommInitRank(&dev->comm->nccl, dev->comm->world, uid, dev->comm->rank);
    if (rc) { dev->comm->nccl = nullptr; return cc_fail(dev, CC_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); }
    return CC_OK;
}

void cc_comm_destroy(cc_device* dev) {
    cc_comm* c = dev->comm;
    if (!c) return;
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
    if (!c->local_peers) for (int p = 0; p < c->world; p++) if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
    if (c->local) cudaFree(c->local);
    delete c;
    dev->comm = nullptr;
}
extern "C" CC_API int32_t cc_comm_rank(cc_device* dev) { return dev && dev->comm ? dev->comm->rank : 0; }
extern "C" CC_API int32_t cc_comm_world_size(cc_device* dev) { return dev && dev->comm ? dev->comm->world : 1; }

// ---- one-shot exchange kernel (eager mode and lazy mode 1; the megakernel carries its own copy of the protocol) ----------------
// mode 0: x[i] = sum_p part_p[i] (+ residual[i])      (n elements, in place)
// mode 1: dst[p * n + i] = src_p[i]                   (allgather of n-element slices)
#define XC_THREADS 1024
__global__ void __launch_bounds__(XC_THREADS) exchange_kernel(CommDev c, float* x, const float* residual, float* dst, int n, int mode, unsigned* err_dev, unsigned* err_host) {
    __shared__ unsigned s_seq;
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = 0;
    if (threadIdx.x == 0) s_seq = *c.seq + 1u;
    __syncthreads();
    const unsigned seq = s_seq;
    const size_t slot = (size_t)(seq & 1u) * CC_COMM_MAX_RANKS;
    const int n4 = n >> 2;
    const float4* x4 = (const float4*)x;
    for (int i = threadIdx.x; i < n4; i += XC_THREADS) {
        const float4 v = x4[i];
        for (int p = 0; p < c.world; p++) ((float4*)(c.data[p] + (slot + c.rank) * CC_COMM_MAX_ELEMS))[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < c.world) {
        cc_st_release_sys(c.flag[threadIdx.x] + c.rank * 32, seq);
        const unsigned* mine = c.flag[c.rank] + threadIdx.x * 32;
        CcSpin sp;
        while ((int)(cc_ld_acquire_sys(mine) - seq) < 0) if (sp.expired(err_dev, err_host, 3u)) { s_abort = 1; break; }      // a peer never arrived
    }
    __syncthreads();
    if (s_abort) return;
    const float* base = c.data[c.rank] + slot * CC_COMM_MAX_ELEMS;
    if (mode == 0) {
        for (int i = threadIdx.x; i < n4; i += XC_THREADS) {
            float4 a = __ldcg((const float4*)base + i);
            for (int p = 1; p < c.world; p++) {
                const float4 b = __ldcg((const float4*)(base + (size_t)p * CC_COMM_MAX_ELEMS) + i);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            if (residual) { const float4 r = ((const float4*)residual)[i]; a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w; }
            ((float4*)x)[i] = a;
        }
    } else {
        for (int p = 0; p < c.world; p++)
            for (int i = threadIdx.x; i < n4; i += XC_THREADS) ((float4*)(dst + (size_t)p * n))[i] = __ldcg((const float4*)(base + (size_t)p * CC_COMM_MAX_ELEMS) + i);
    }
    if (threadIdx.x == 0) *c.seq = seq;
}

// x += residual after an NCCL allreduce (the p2p kernel folds this in)
__global__ void add_residual_kernel(float* x, const float* r, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] + r[i];
}

int cc_launch_all_reduce(cc_device* dev, float* x, int64_t n, const float* residual) {
    cc_comm* c = dev->comm;
    if (!c) return cc_fail(dev, CC_ERR_UNSUPPORTED, "all_reduce: no communicator on this device (cc_comm_create / cc_comm_connect)");
    CC_REQUIRE(dev, n % 4 == 0 && n <= CC_COMM_MAX_ELEMS, "all_reduce: %lld elements unsupported (multiple of 4, at most %d)", (long long)n, CC_COMM_MAX_ELEMS);
    if (c->nccl) {
        int rc = g_nccl.AllReduce(x, x, (size_t)n, ncclFloat32, ncclSumOp, c->nccl, dev->stream);
        if (rc) return cc_fail(dev, CC_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        dev->launches++;
        if (residual) { add_residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, dev->stream>>>(x, residual, (int)n); CC_LAUNCH_CHECK(dev); }
        return CC_OK;
    }
    CC_REQUIRE(dev, c->connected, "all_reduce: communicator not connected");
    exchange_kernel<<<1, XC_THREADS, 0, dev->stream>>>(c->cd, x, residual, nullptr, (int)n, 0, dev->err_dev, dev->err_host);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}

int cc_launch_all_gather(cc_device* dev, const float* src, int64_t n, float* dst) {
    cc_comm* c = dev->comm;
    if (!c) return cc_fail(dev, CC_ERR_UNSUPPORTED, "all_gather: no communicator on this device");
    CC_REQUIRE(dev, n % 4 == 0 && n <= CC_COMM_MAX_ELEMS, "all_gather: %lld elements per rank unsupported (multiple of 4, at most %d)", (long long)n, CC_COMM_MAX_ELEMS);
    if (c->nccl) {
        int rc = g_nccl.AllGather(src, dst, (size_t)n, ncclFloat32, c->nccl, dev->stream);
        if (rc) return cc_fail(dev, CC_ERR_CUDA, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        dev->launches++;
        return CC_OK;
    }
    CC_REQUIRE(dev, c->connected, "all_gather: communicator not connected");
    exchange_kernel<<<1, XC_THREADS, 0, dev->stream>>>(c->cd, (float*)src, nullptr, dst, (int)n, 1, dev->err_dev, dev->err_host);
    CC_LAUNCH_CHECK(dev);
    return CC_OK;
}
