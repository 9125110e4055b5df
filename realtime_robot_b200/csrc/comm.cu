// comm.cu — the one collective of the path: an ncclAllGather of fixed 128-byte pose records.
//
// The path shards by candidate model cloud (rank r registers its own models against the replicated scan) and by RANSAC
// hypothesis range; nothing else crosses GPUs.  After a rank's batch, every rank needs every record (best model per scan,
// arg-min over hypothesis shards), which is ONE all-gather of n_local x 128 bytes per rank over NVLink — latency, not
// bandwidth.  The communicator belongs to the context and works on its stream, from a pre-allocated device buffer and the
// context's pinned area: no allocation, no Python, no extra synchronisation on the way.
//
// Peer-memory exchange (round 2).  The records are 1 KB per rank and the exchange is pure latency, so after the communicator
// exists every rank maps every peer's "mailbox" (cudaIpc handles, all-gathered once through NCCL) and a batch's exchange is ONE
// kernel of `world` CTAs on the context's stream: CTA d stores this rank's records straight into rank d's mailbox over NVLink
// (plain 16-byte stores, __threadfence_system, then the sequence number as a release flag), waits for rank d's flag in its OWN
// mailbox and copies rank d's records into the gathered buffer.  No proxy thread, no protocol buffers, no second launch.
// Two mailbox slots alternate by sequence parity (a rank cannot be two exchanges ahead of a peer: it needs the peer's records of
// the exchange in between); every spin is bounded (2 s) and ends in an error status at the C ABI, never in a hung GPU.
// Opt-in (RTR_COMM_P2P=1; peers that cannot be mapped keep NCCL): correct on 2 and 8 GPUs (tools/check_comm.py checks every
// gathered byte on every rank), and NOT faster — a step of the bench ends 0.02 - 0.09 ms after its slowest rank with either
// exchange (2.667 against 2.669 ms per step at 8 GPUs): what a step waits for is the slowest GPU of the box, not the collective.
//
// NCCL is bound at run time (dlopen): the process usually has one already (torch ships libnccl.so.2 and loads it), and
// librtr.so must keep loading on hosts without NCCL — the single-GPU path does not need it.  The C++ host reaches multi-GPU
// through the same three calls (rtr_comm_unique_id on rank 0, any out-of-band exchange of the 128-byte id, rtr_comm_init).
#include "common.cuh"
#include <dlfcn.h>
#include <cstring>
#include <mutex>

// the slice of nccl.h this file uses (ABI-stable across NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { RTR_NCCL_UINT8 = 1 };       // ncclUint8

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_bind() {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.ok) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);         // the copy the process already uses (torch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return rtr_fail("comm", "libnccl.so.2 not found (multi-GPU needs NCCL; the single-GPU path does not)", RTR_ERR_NOT_READY);
    g_nccl.GetUniqueId = (ncclResult_t(*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (ncclResult_t(*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(ncclResult_t))dlsym(h, "ncclGetErrorString");
    g_nccl.GetVersion = (ncclResult_t(*)(int*))dlsym(h, "ncclGetVersion");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather)
        return rtr_fail("comm", "libnccl.so.2 lacks a required symbol", RTR_ERR_NOT_READY);
    g_nccl.ok = true;
    return 0;
}

#define RTR_NCCL(call, tag)                                                                                 \
    do {                                                                                                    \
        ncclResult_t r__ = (call);                                                                          \
        if (r__ != 0) {                                                                                     \
            fprintf(stderr, "rtr[%s] %s failed: %s\n", tag, #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "NCCL error"); \
            return 1000 + (int)r__;                                                                         \
        }                                                                                                   \
    } while (0)

// records every rank may contribute per gather (device staging: world x this x 128 B)
#define RTR_COMM_MAX_LOCAL 64

// ---- peer-memory exchange ------------------------------------------------------------------------------------------------
#define RTR_P2P_MAX_WORLD 16
#define RTR_P2P_SLOT_BYTES ((size_t)RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result))          // one rank's records of one exchange
struct P2pView {
    int world, rank;
    char* box[RTR_P2P_MAX_WORLD];          // every rank's mailbox as this process sees it (box[rank]: the local allocation)
};
struct P2pState {
    P2pView v;
    char* local = nullptr;                 // [2 parities][world][slot] records, then [2][world] flags, 128 bytes apart
    unsigned seq = 0;                      // exchanges queued so far
    int* d_error = nullptr;                // set by a kernel whose wait timed out
};
static inline size_t p2p_records_bytes(int world) { return 2 * (size_t)world * RTR_P2P_SLOT_BYTES; }
static inline size_t p2p_box_bytes(int world) { return p2p_records_bytes(world) + 2 * (size_t)world * 128; }
__device__ __forceinline__ char* p2p_slot(char* box, int world, unsigned parity, int src) { return box + ((size_t)parity * world + src) * RTR_P2P_SLOT_BYTES; }
__device__ __forceinline__ unsigned* p2p_flag(char* box, int world, unsigned parity, int src) {
    return (unsigned*)(box + 2 * (size_t)world * RTR_P2P_SLOT_BYTES + ((size_t)parity * world + src) * 128);
}
// CTA d: my records -> rank d's mailbox, flag; then rank d's records (from my mailbox) -> out_all[d * n_local ...]
__global__ void __launch_bounds__(128) k_p2p_exchange(const __grid_constant__ P2pView v, const rtr_pose_result* __restrict__ local, int n_local, unsigned seq,
                                                     rtr_pose_result* __restrict__ out_all, int* __restrict__ error) {
    const int d = blockIdx.x;
    const unsigned parity = seq & 1u;
    const int n16 = n_local * (int)(sizeof(rtr_pose_result) / 16);
    {
        uint4* dst = (uint4*)p2p_slot(v.box[d], v.world, parity, v.rank);
        const uint4* src = (const uint4*)local;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned* f = p2p_flag(v.box[d], v.world, parity, v.rank);
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
        }
    }
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const unsigned* f = p2p_flag(v.box[v.rank], v.world, parity, d);
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        int good = 0;
        for (;;) {
            unsigned val;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(val) : "l"(f) : "memory");
            if (val == seq) { good = 1; break; }
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) break;                    // 2 s: a peer is gone
            __nanosleep(100);
        }
        if (!good) atomicExch(error, 1);
        ok = good;
    }
    __syncthreads();
    if (!ok) return;
    const uint4* src = (const uint4*)p2p_slot(v.box[v.rank], v.world, parity, d);
    uint4* dst = (uint4*)(out_all + (size_t)d * n_local);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) {
        uint4 x;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(src + i) : "memory");
        dst[i] = x;
    }
}

static inline int* comm_pinned_error(rtr_context* ctx) {
    return (int*)((char*)ctx->comm_pinned + ((size_t)ctx->comm_world + 1) * RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result));
}
// one exchange on the context's stream: d_local (n_local records) -> d_all (world x n_local, rank order); the error word
// travels to the pinned area behind it (checked by whoever hands the records out)
static int p2p_exchange(rtr_context* ctx, const rtr_pose_result* d_local, int n_local, rtr_pose_result* d_all) {
    P2pState* st = (P2pState*)ctx->comm_p2p;
    st->seq++;
    k_p2p_exchange<<<st->v.world, 128, 0, ctx->stream>>>(st->v, d_local, n_local, st->seq, d_all, st->d_error);
    RTR_LAUNCH_CHECK(ctx, "comm.p2p_exchange");
    RTR_CHECK(cudaMemcpyAsync(comm_pinned_error(ctx), st->d_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "comm.p2p");
    return 0;
}
static void p2p_teardown(rtr_context* ctx) {
    P2pState* st = (P2pState*)ctx->comm_p2p;
    if (!st) return;
    for (int r = 0; r < st->v.world; ++r) if (r != st->v.rank && st->v.box[r]) cudaIpcCloseMemHandle(st->v.box[r]);
    if (st->local) cudaFree(st->local);
    if (st->d_error) cudaFree(st->d_error);
    delete st;
    ctx->comm_p2p = nullptr;
}
// after the communicator exists: allocate the mailbox, all-gather the IPC handles through NCCL, map the peers.  Any failure
// leaves comm_p2p == nullptr (NCCL serves the exchange) — but the decision must be the same on every rank, so the outcome is
// all-reduced by a second tiny all-gather.
static int p2p_setup(rtr_context* ctx, int world, int rank);

// queue the all-gather of n_local device records on the context's stream (register_many_enqueue); d_local must stay valid
// until the stream has passed it
int rtr_comm_allgather_dev(rtr_context* ctx, const rtr_pose_result* d_local, int n_local) {
    if (!ctx->comm) return 0;
    if (n_local > RTR_COMM_MAX_LOCAL) return rtr_fail("allgather", "at most 64 records per rank and batch", RTR_ERR_CAPACITY);
    const size_t lb = sizeof(rtr_pose_result) * (size_t)n_local;
    if (ctx->comm_p2p) {
        if (int e = p2p_exchange(ctx, d_local, n_local, (rtr_pose_result*)ctx->comm_dev)) return e;
    } else {
        RTR_NCCL(g_nccl.AllGather(d_local, ctx->comm_dev, lb, RTR_NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream), "allgather");
        RTR_MARK(ctx, "comm.allgather");
    }
    RTR_CHECK(cudaMemcpyAsync(ctx->comm_pinned, ctx->comm_dev, lb * (size_t)ctx->comm_world, cudaMemcpyDeviceToHost, ctx->stream), "allgather");
    return 0;
}

static int p2p_setup(rtr_context* ctx, int world, int rank) {
    if (world <= 1) return 0;
    // opt-in (RTR_COMM_P2P=1): measured on 2 and 8 B200s of one box, a step that ends in this exchange takes 2.657 / 2.667 ms against
    // 2.637 / 2.669 ms with the in-stream ncclAllGather — the exchange is not what a step waits for, the slowest rank is
    const char* env = getenv("RTR_COMM_P2P");
    if (!(env && env[0] == '1')) return 0;
    int mine_ok = world <= RTR_P2P_MAX_WORLD && world > 1;
    P2pState* st = new P2pState();
    st->v.world = world; st->v.rank = rank;
    for (int r = 0; r < RTR_P2P_MAX_WORLD; ++r) st->v.box[r] = nullptr;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (mine_ok) {
        if (cudaMalloc(&st->local, p2p_box_bytes(world)) != cudaSuccess || cudaMalloc(&st->d_error, sizeof(int)) != cudaSuccess) mine_ok = 0;
        else {
            cudaMemsetAsync(st->local, 0, p2p_box_bytes(world), ctx->stream);
            cudaMemsetAsync(st->d_error, 0, sizeof(int), ctx->stream);
            if (cudaIpcGetMemHandle(&mine, st->local) != cudaSuccess) mine_ok = 0;
        }
        cudaGetLastError();
    }
    // handles of all ranks (64 bytes each + this rank's verdict so far) through the communicator
    struct Msg { cudaIpcMemHandle_t h; int ok; int pad[15]; };
    static_assert(sizeof(Msg) == 128, "one record-sized message per rank");
    Msg* pin = (Msg*)ctx->comm_pinned;
    char* dev_all = (char*)ctx->comm_dev;
    char* dev_local = dev_all + (size_t)world * RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result);
    Msg m;
    memset(&m, 0, sizeof(m));
    m.h = mine; m.ok = mine_ok;
    for (int round = 0; round < 2; ++round) {
        // round 0: handles + "I could allocate"; round 1: "I could map every peer"
        pin[world] = m;
        RTR_CHECK(cudaMemcpyAsync(dev_local, &pin[world], sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream), "comm.p2p");
        RTR_NCCL(g_nccl.AllGather(dev_local, dev_all, sizeof(Msg), RTR_NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream), "comm.p2p");
        RTR_CHECK(cudaMemcpyAsync(pin, dev_all, sizeof(Msg) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream), "comm.p2p");
        RTR_CHECK(cudaStreamSynchronize(ctx->stream), "comm.p2p");
        int all_ok = 1;
        for (int r = 0; r < world; ++r) all_ok &= pin[r].ok;
        if (!all_ok) {
            if (rank == 0) fprintf(stderr, "rtr[comm] peer-memory exchange not available on every rank: NCCL serves the all-gather\n");
            ctx->comm_p2p = st;
            p2p_teardown(ctx);
            return 0;
        }
        if (round == 0) {
            for (int r = 0; r < world; ++r) {
                if (r == rank) { st->v.box[r] = st->local; continue; }
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, pin[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); m.ok = 0; ptr = nullptr; }
                st->v.box[r] = (char*)ptr;
            }
        }
    }
    ctx->comm_p2p = st;
    if (rank == 0 && getenv("RTR_COMM_VERBOSE")) fprintf(stderr, "rtr[comm] peer-memory exchange enabled: %d mailboxes mapped over NVLink\n", world);
    return 0;
}

// a peer that never arrived (2 s) leaves the device error flag set: report it instead of handing out stale records
static int rtr_comm_check_peers(rtr_context* ctx) {
    if (!ctx->comm_p2p) return 0;
    if (*comm_pinned_error(ctx)) return rtr_fail("comm", "a peer did not deliver its records within 2 s (peer-memory exchange)", RTR_ERR_NOT_READY);
    return 0;
}

extern "C" {

int rtr_comm_unique_id(char* id128) {
    if (!id128) return rtr_fail("comm", "bad argument", RTR_ERR_INVALID);
    if (int e = nccl_bind()) return e;
    ncclUniqueId id;
    RTR_NCCL(g_nccl.GetUniqueId(&id), "comm");
    memcpy(id128, id.internal, 128);
    return 0;
}

int rtr_comm_init(rtr_context* ctx, int world, int rank, const char* id128) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return rtr_fail("comm", "bad argument", RTR_ERR_INVALID);
    if (ctx->comm) return rtr_fail("comm", "this context already has a communicator", RTR_ERR_INVALID);
    if (int e = nccl_bind()) return e;
    RTR_CHECK(cudaSetDevice(ctx->device), "comm");
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t comm = nullptr;
    RTR_NCCL(g_nccl.CommInitRank(&comm, world, id, rank), "comm");
    ctx->comm = (void*)comm; ctx->comm_world = world; ctx->comm_rank = rank;
    const size_t bytes = (size_t)world * RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result);
    RTR_CHECK(cudaMalloc(&ctx->comm_dev, bytes + (size_t)RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result)), "comm");
    RTR_CHECK(cudaMallocHost(&ctx->comm_pinned, bytes + (size_t)RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result) + 128), "comm");     // + the exchange's error word
    *comm_pinned_error(ctx) = 0;
    return p2p_setup(ctx, world, rank);
}

int rtr_comm_destroy(rtr_context* ctx) {
    if (!ctx || !ctx->comm) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    p2p_teardown(ctx);
    if (g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr; ctx->comm_world = 0; ctx->comm_rank = 0;
    if (ctx->comm_dev) cudaFree(ctx->comm_dev);
    if (ctx->comm_pinned) cudaFreeHost(ctx->comm_pinned);
    ctx->comm_dev = nullptr; ctx->comm_pinned = nullptr;
    return 0;
}

int rtr_comm_world(rtr_context* ctx, int* world, int* rank) {
    if (!ctx) return rtr_fail("comm", "bad argument", RTR_ERR_INVALID);
    if (world) *world = ctx->comm ? ctx->comm_world : 1;
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    return 0;
}

// Every rank contributes n_local records (the same n_local everywhere, <= 64); host_all receives world x n_local records
// in rank order.  Without a communicator (single GPU) it is a copy.  One H2D of the local records, ONE ncclAllGather, one
// D2H, one stream synchronisation.
int rtr_allgather_results(rtr_context* ctx, const rtr_pose_result* host_local, int n_local, rtr_pose_result* host_all) {
    if (!ctx || n_local < 0 || (n_local > 0 && (!host_local || !host_all))) return rtr_fail("allgather", "bad argument", RTR_ERR_INVALID);
    if (n_local == 0) return 0;
    if (!ctx->comm) { memmove(host_all, host_local, sizeof(rtr_pose_result) * (size_t)n_local); return 0; }
    if (n_local > RTR_COMM_MAX_LOCAL) return rtr_fail("allgather", "at most 64 records per rank and call", RTR_ERR_CAPACITY);
    RTR_CHECK(cudaSetDevice(ctx->device), "allgather");
    const size_t lb = sizeof(rtr_pose_result) * (size_t)n_local, ab = lb * (size_t)ctx->comm_world;
    char* dev_all = (char*)ctx->comm_dev;
    char* dev_local = dev_all + (size_t)ctx->comm_world * RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result);
    char* pin_all = (char*)ctx->comm_pinned;
    char* pin_local = pin_all + (size_t)ctx->comm_world * RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result);
    memcpy(pin_local, host_local, lb);
    RTR_CHECK(cudaMemcpyAsync(dev_local, pin_local, lb, cudaMemcpyHostToDevice, ctx->stream), "allgather");
    if (ctx->comm_p2p) {
        if (int e = p2p_exchange(ctx, (const rtr_pose_result*)dev_local, n_local, (rtr_pose_result*)dev_all)) return e;
    } else {
        RTR_NCCL(g_nccl.AllGather(dev_local, dev_all, lb, RTR_NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream), "allgather");
        RTR_MARK(ctx, "comm.allgather");
    }
    RTR_CHECK(cudaMemcpyAsync(pin_all, dev_all, ab, cudaMemcpyDeviceToHost, ctx->stream), "allgather");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "allgather");
    if (int e = rtr_comm_check_peers(ctx)) return e;
    memcpy(host_all, pin_all, ab);
    return 0;
}

// In-stream form for batches: with rtr_comm_gather_batches(ctx, 1, base) every rtr_register_many* on this context ends with the
// all-gather of its records queued on the context's stream right behind the batch — device buffer to device buffer, then one
// D2H next to the local records — so the exchange costs no extra host round trip or synchronisation; rtr_gathered_results
// hands out the world x n_models records after rtr_register_many_end.  model_id of local record k is base + k.  Every rank
// must run batches of the same size.
int rtr_comm_gather_batches(rtr_context* ctx, int on, int model_id_base) {
    if (!ctx) return rtr_fail("comm", "bad argument", RTR_ERR_INVALID);
    ctx->gather_batches = on ? 1 : 0;
    ctx->model_id_base = model_id_base;
    return 0;
}
int rtr_gathered_results(rtr_context* ctx, rtr_pose_result* host_all, int capacity, int* n_records) {
    if (!ctx || !host_all || !n_records) return rtr_fail("comm", "bad argument", RTR_ERR_INVALID);
    const int n = ctx->gathered_records;
    *n_records = n;
    if (n <= 0) return rtr_fail("comm", "no gathered batch on this context (rtr_comm_gather_batches + rtr_register_many_end first)", RTR_ERR_NOT_READY);
    if (capacity < n) return rtr_fail("comm", "result buffer too small", RTR_ERR_CAPACITY);
    if (ctx->comm) if (int e = rtr_comm_check_peers(ctx)) return e;
    memcpy(host_all, ctx->comm ? ctx->comm_pinned : ctx->pinned, sizeof(rtr_pose_result) * (size_t)n);
    return 0;
}

// Hypothesis-sharded RANSAC: every rank evaluated its own hypothesis range of the SAME registration; the
// winner is the arg-min over (fitness, hypothesis id) among the accepted shards — the sequential rule "error < lowest_error"
// keeps the first lowest — so the answer is identical on every rank and for every world size.  `evaluated` is summed.
int rtr_select_best_hypothesis(const rtr_pose_result* records, int n, rtr_pose_result* best) {
    if (!records || !best || n < 1) return rtr_fail("select", "bad argument", RTR_ERR_INVALID);
    int w = -1;
    long long evaluated = 0;
    for (int i = 0; i < n; ++i) {
        evaluated += records[i].evaluated;
        if (records[i].hypothesis < 0 || !records[i].converged) continue;
        if (w < 0 || records[i].fitness < records[w].fitness || (records[i].fitness == records[w].fitness && records[i].hypothesis < records[w].hypothesis)) w = i;
    }
    *best = records[w < 0 ? 0 : w];
    if (w < 0) { best->hypothesis = -1; best->converged = 0; best->inliers = 0; best->fitness = FLT_MAX; for (int i = 0; i < 16; ++i) best->pose[i] = (i % 5 == 0) ? 1.f : 0.f; }
    best->evaluated = evaluated;
    return 0;
}

}  // extern "C"
