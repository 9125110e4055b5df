// comm.cu — the one collective of the path: an ncclAllGather of fixed 128-byte pose records.
//
// The path shards by candidate model cloud (rank r registers its own models against the replicated scan) and by RANSAC
// hypothesis range; nothing else crosses GPUs.  After a rank's batch, every rank needs every record (best model per scan,
// arg-min over hypothesis shards), which is ONE all-gather of n_local x 128 bytes per rank over NVLink — latency, not
// bandwidth.  The communicator belongs to the context and works on its stream, from a pre-allocated device buffer and the
// context's pinned area: no allocation, no Python, no extra synchronisation on the way.
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

// queue the all-gather of n_local device records on the context's stream (register_many_enqueue); d_local must stay valid
// until the stream has passed it
int rtr_comm_allgather_dev(rtr_context* ctx, const rtr_pose_result* d_local, int n_local) {
    if (!ctx->comm) return 0;
    if (n_local > RTR_COMM_MAX_LOCAL) return rtr_fail("allgather", "at most 64 records per rank and batch", RTR_ERR_CAPACITY);
    const size_t lb = sizeof(rtr_pose_result) * (size_t)n_local;
    RTR_NCCL(g_nccl.AllGather(d_local, ctx->comm_dev, lb, RTR_NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream), "allgather");
    RTR_MARK(ctx, "comm.allgather");
    RTR_CHECK(cudaMemcpyAsync(ctx->comm_pinned, ctx->comm_dev, lb * (size_t)ctx->comm_world, cudaMemcpyDeviceToHost, ctx->stream), "allgather");
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
    RTR_CHECK(cudaMallocHost(&ctx->comm_pinned, bytes + (size_t)RTR_COMM_MAX_LOCAL * sizeof(rtr_pose_result)), "comm");
    return 0;
}

int rtr_comm_destroy(rtr_context* ctx) {
    if (!ctx || !ctx->comm) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
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
    RTR_NCCL(g_nccl.AllGather(dev_local, dev_all, lb, RTR_NCCL_UINT8, (ncclComm_t)ctx->comm, ctx->stream), "allgather");
    RTR_MARK(ctx, "comm.allgather");
    RTR_CHECK(cudaMemcpyAsync(pin_all, dev_all, ab, cudaMemcpyDeviceToHost, ctx->stream), "allgather");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "allgather");
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
