// common.cuh — shared host/device definitions of librtr.so (sm_100a only).
//
// Arithmetic contract (DESIGN.md "Parity contract"): neighbour membership is decided by the float expression
// (dx*dx + dy*dy) + dz*dz with no FMA (FLANN L2_Simple); reductions accumulate in fp64 and are
// stored as fp32; small solves are cyclic Jacobi in fp64 using + - * / sqrt only.  The library is compiled with
// -fmad=false so that no contraction changes those operation sequences.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cfloat>
#include <map>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "../../include/rtr.h"

// NVTX range around the host-side enqueue of a stage (SURVEY section 5 "tracing"): shows up as rtr.<stage> rows in Nsight
// Systems / ncu --nvtx; header-only NVTX v3, a pointer test when no tool is attached.
struct RtrRange {
    explicit RtrRange(const char* name) { nvtxRangePushA(name); }
    ~RtrRange() { nvtxRangePop(); }
    RtrRange(const RtrRange&) = delete;
    RtrRange& operator=(const RtrRange&) = delete;
};

#define RTR_NUM_EVENTS 16
// clouds up to this size use one WARP per point in the gather kernels (one thread per point cannot fill 148 SMs)
#define RTR_WARP_PER_POINT_MAX 262144

// ----------------------------------------------------------------------------- host-side objects
struct DevGrid {
    float h = 0.f, inv_h = 0.f;
    float mnx = 0.f, mny = 0.f, mnz = 0.f;
    int dx = 1, dy = 1, dz = 1;
    int ncells = 1;
    int n = 0;
    int* cell_begin = nullptr;       // ncells + 1
    float4* sorted = nullptr;        // points in cell order; .w carries the ORIGINAL index (int bits)
    float4* sorted_normals = nullptr;  // optional: normals4 permuted into this grid's order
    int normals_version = -1;
    // Segmented grid of a model set (see rtr_cloud::seg_begin): one independent uniform grid per segment.  cell_begin is
    // the concatenation of the segments' cell tables (segment k at cell_off, ncells_k entries, values are GLOBAL positions
    // into `sorted`, so the table of segment k+1 closes segment k), sorted[.].w carries the GLOBAL original index.
    std::vector<struct SegHdr> segs;
};

// one segment of a segmented grid
struct SegHdr {
    float inv_h, h;
    float mnx, mny, mnz;
    int dx, dy, dz;
    int cell_off;      // first entry of this segment's cell table inside the concatenated cell_begin
};

// what the kernels see (passed by value)
struct GridView {
    float inv_h, h;
    float mnx, mny, mnz;
    int dx, dy, dz;
    int n;
    const int* __restrict__ cell_begin;
    const float4* __restrict__ sorted;
};

// Grid source of the per-point kernels: the grid a work item (a position in cell order) belongs to.
//   OneGrid   — an ordinary cloud: every position is in the one grid (compiles to the plain GridView kernels).
//   ManyGrids — a model set: up to RTR_MAX_SEGMENTS clouds concatenated, each with its own grid; positions are global
//               (segment k owns [begin[k], begin[k+1])), so ONE launch serves every cloud of the set and a point only
//               ever sees neighbours of its own cloud.
#define RTR_MAX_SEGMENTS 32
struct OneGrid {
    GridView g;
    __host__ __device__ int total() const { return g.n; }
#ifdef __CUDACC__
    __device__ __forceinline__ GridView at(int) const { return g; }
    __device__ __forceinline__ GridView view(int) const { return g; }
    __device__ __forceinline__ int segment_of(int) const { return 0; }
    __device__ __forceinline__ int seg_start(int) const { return 0; }
    __device__ __forceinline__ int seg_size(int) const { return g.n; }
#endif
};
struct ManyGrids {
    int nseg;
    int begin[RTR_MAX_SEGMENTS + 1];        // unused tail = total
    SegHdr seg[RTR_MAX_SEGMENTS];
    const int* __restrict__ cell_begin;
    const float4* __restrict__ sorted;
    __host__ __device__ int total() const { return begin[RTR_MAX_SEGMENTS]; }
#ifdef __CUDACC__
    // binary search over the 32 segment starts (positions past the end fall into the last slot)
    __device__ __forceinline__ int segment_of(int s) const {
        int k = 0;
#pragma unroll
        for (int step = RTR_MAX_SEGMENTS / 2; step > 0; step >>= 1) if (s >= begin[k + step]) k += step;
        return k;
    }
    __device__ __forceinline__ GridView view(int k) const {
        GridView g;
        const SegHdr& h = seg[k];
        g.inv_h = h.inv_h; g.h = h.h; g.mnx = h.mnx; g.mny = h.mny; g.mnz = h.mnz; g.dx = h.dx; g.dy = h.dy; g.dz = h.dz;
        g.n = begin[k + 1] - begin[k];
        g.cell_begin = cell_begin + h.cell_off; g.sorted = sorted;
        return g;
    }
    __device__ __forceinline__ GridView at(int s) const { return view(segment_of(s)); }
    __device__ __forceinline__ int seg_start(int k) const { return begin[k]; }
    __device__ __forceinline__ int seg_size(int k) const { return begin[k + 1] - begin[k]; }
#endif
};

struct ProfMark { const char* tag; cudaEvent_t ev; };
struct ArenaSlab { char* base; size_t cap; };

struct rtr_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    // second stream of a registration: the scene's stages (grid, normals, Harris, FPFH) run here while the model's run on
    // `stream`; forked and joined with the two events below inside rtr_register_begin, idle otherwise
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t fork_event = nullptr, join_event = nullptr;
    long long launches = 0;
    bool profile = false;                 // when set, every stream operation is followed by an event mark
    std::vector<ProfMark> marks;
    std::vector<cudaEvent_t> event_pool;
    int sm_count = 148;
    int match_stats[3] = {-1, 0, 0};
    // bump arena for temporaries of one API call: all work of a context is ordered on one stream, so a temporary's memory
    // can be handed out again as soon as the call that used it has been ENQUEUED.  Replaces ~100 cudaMallocAsync /
    // cudaFreeAsync pairs per registration (8 concurrent registrations were partly bound by those driver calls).
    std::vector<ArenaSlab> slabs;
    size_t arena_top = 0;
    int arena_depth = 0;
    cudaEvent_t events[RTR_NUM_EVENTS] = {};
    // small pinned staging area for results / counters
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // this context's own stream-ordered memory pool: blocks freed on this stream are only ever reused on this stream, so the
    // allocator never makes one context's stream wait for another's (the device's default pool is shared by all streams)
    cudaMemPool_t pool = nullptr;
    // asynchronous registration (rtr_register_begin / _end): at most one in flight per context
    int register_pending = 0;
    struct rtr_cloud* pending_cloud[2] = {nullptr, nullptr};
    int many_models = 0;                  // > 0: the registration in flight is a batch of that many models (rtr_register_many_begin)
    std::vector<char> kp_preview;         // corner previews of the last batch (rtr_register_many_keypoints)
    int kp_members = 0;
    // multi-GPU (comm.cu): an NCCL communicator on this context's stream + pre-allocated staging for the record all-gather
    void* comm = nullptr; int comm_world = 0, comm_rank = 0;
    void* comm_dev = nullptr; void* comm_pinned = nullptr;
    void* comm_p2p = nullptr;             // peer-memory exchange over NVLink (comm.cu P2pState), nullptr: NCCL only
    int gather_batches = 0;               // rtr_comm_gather_batches: every batch ends with the in-stream all-gather of its records
    int model_id_base = 0;                // model_id of a batch's record k = base + k
    int gathered_records = 0;             // records of the last gathered batch (world x n_models; n_models without a communicator)
    // pinned staging of rtr_pcd_load (decoded points, grown on demand)
    void* io_pinned = nullptr;
    size_t io_pinned_cap = 0;
};

struct rtr_cloud {
    rtr_context* ctx = nullptr;
    int n = 0;
    float4* pts = nullptr;                    // original order
    float bb_min[3] = {0, 0, 0}, bb_max[3] = {0, 0, 0};
    bool bbox_valid = false;
    std::map<int, DevGrid> grids;             // keyed by float bits of the requested cell size
    float4* normals = nullptr;   float normals_radius = -1.f;  int normals_version = 0;  int normals_mode = 0;
    float*  response = nullptr;
    float*  fpfh = nullptr;      float fpfh_radius = -1.f;
    // the k-NN correspondences name their target by identity: the handle plus the generation of its features at that time
    int*    knn = nullptr;       float* knn_dist = nullptr;    int knn_k = 0;
    const rtr_cloud* knn_target = nullptr;   long long knn_target_gen = -1;
    long long feature_gen = 0;                // process-wide unique stamp, renewed whenever this cloud's FPFH rows are recomputed or dropped
    int n_keypoints = -1;
    // rtr_cloud_prepare (the offline phase, RealTimeRobot.cpp:124-165): normals, Harris corners and FPFH rows stay on the cloud
    // together with the stage parameters they were computed with; kp_xyz holds the first RTR_KP_PREVIEW corners, kp_count their number
    bool prepared = false;
    float prep_normal_radius = 0.f, prep_harris_radius = 0.f, prep_harris_threshold = 0.f, prep_fpfh_radius = 0.f;
    int prep_harris_nms = 0, prep_harris_refine = 0, prep_normals_version = -1;
    float4* kp_xyz = nullptr;    int* kp_count = nullptr;
    // A model set (rtr_register_many): the points of several clouds concatenated, seg_begin[k] .. seg_begin[k+1] is cloud k
    // (size nseg + 1; empty for an ordinary cloud).  Every per-point stage then runs on all member clouds in one launch,
    // through segmented grids; outputs (normals, response, FPFH) are concatenated in the same order.
    std::vector<int> seg_begin;
    std::vector<float> seg_bb;                // 6 floats per segment: min x, y, z, max x, y, z
    int nseg() const { return seg_begin.empty() ? 0 : (int)seg_begin.size() - 1; }
};

#define RTR_CHECK(call, tag)                                                                       \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            fprintf(stderr, "rtr[%s] %s failed: %s\n", tag, #call, cudaGetErrorString(e__));       \
            return (int)e__;                                                                       \
        }                                                                                          \
    } while (0)

void rtr_prof_mark(rtr_context* ctx, const char* tag);
// a non-kernel stream operation (memset / memcpy / CUB pass): only marked when profiling
#define RTR_MARK(ctx, tag) do { if ((ctx)->profile) rtr_prof_mark((ctx), tag); } while (0)

#define RTR_LAUNCH_CHECK(ctx, tag)                                                                 \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        if ((ctx)->profile) rtr_prof_mark((ctx), tag);                                             \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) {                                                                  \
            fprintf(stderr, "rtr[%s] kernel launch failed: %s\n", tag, cudaGetErrorString(e__));   \
            return (int)e__;                                                                       \
        }                                                                                          \
    } while (0)

static inline int rtr_fail(const char* tag, const char* msg, int code) {
    fprintf(stderr, "rtr[%s] %s\n", tag, msg);
    return code;
}

template <typename T>
static inline int dev_alloc(rtr_context* ctx, T** p, size_t count, const char* tag) {
    *p = nullptr;
    if (count == 0) count = 1;
    RTR_CHECK(cudaMallocFromPoolAsync((void**)p, count * sizeof(T), ctx->pool, ctx->stream), tag);
    return 0;
}
static inline bool in_arena(const rtr_context* ctx, const void* p) {
    for (const ArenaSlab& s : ctx->slabs) if ((const char*)p >= s.base && (const char*)p < s.base + s.cap) return true;
    return false;
}
template <typename T>
static inline void dev_free(rtr_context* ctx, T* p) {
    if (p && !in_arena(ctx, (const void*)p)) cudaFreeAsync((void*)p, ctx->stream);     // arena memory is released by its scope
}
// temporary that does not outlive the current API call (falls back to the pool when no scope is open)
template <typename T>
static inline int tmp_alloc(rtr_context* ctx, T** p, size_t count, const char* tag) {
    if (ctx->arena_depth == 0) return dev_alloc(ctx, p, count, tag);
    size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
    if (ctx->slabs.empty() || ctx->arena_top + bytes > ctx->slabs.back().cap) {
        size_t cap = ctx->slabs.empty() ? ((size_t)32 << 20) : ctx->slabs.back().cap * 2;
        if (cap < bytes * 2) cap = bytes * 2;
        char* base = nullptr;
        RTR_CHECK(cudaMallocFromPoolAsync((void**)&base, cap, ctx->pool, ctx->stream), tag);
        ctx->slabs.push_back({base, cap});      // older slabs stay alive until the outermost scope closes
        ctx->arena_top = 0;
    }
    *p = (T*)(ctx->slabs.back().base + ctx->arena_top);
    ctx->arena_top += bytes;
    return 0;
}
struct TmpScope {
    rtr_context* ctx; size_t mark; size_t nslabs;
    explicit TmpScope(rtr_context* c) : ctx(c), mark(c->arena_top), nslabs(c->slabs.size()) { ctx->arena_depth++; }
    ~TmpScope() {
        ctx->arena_depth--;
        if (ctx->slabs.size() == nslabs) { ctx->arena_top = mark; return; }
        if (ctx->arena_depth == 0) {            // the arena grew during this call: keep only the newest (largest) slab
            for (size_t i = 0; i + 1 < ctx->slabs.size(); ++i) cudaFreeAsync(ctx->slabs[i].base, ctx->stream);
            ArenaSlab last = ctx->slabs.back();
            ctx->slabs.clear(); ctx->slabs.push_back(last);
            ctx->arena_top = 0;
        }
    }
    TmpScope(const TmpScope&) = delete;
    TmpScope& operator=(const TmpScope&) = delete;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once per (device, kernel), under a mutex
// (contexts may live on any device and on any host thread; ComputeTDFWithCuda always uses device 0).
int  rtr_func_smem(const void* func, int device, int bytes);
template <typename K> static inline int rtr_kernel_smem(K kernel, const rtr_context* ctx, size_t bytes) {
    return rtr_func_smem((const void*)kernel, ctx->device, (int)bytes);
}
// cached cudaOccupancyMaxActiveBlocksPerMultiprocessor per (device, kernel, block, smem); >= 1
int  rtr_func_occupancy(const void* func, int device, int block, size_t smem, int* per_sm);

// internal API between translation units
int  rtr_get_grid(rtr_cloud* c, float cell, DevGrid** out);          // build-or-fetch the grid that serves radius `cell`
// a cached grid whose cell size h satisfies lo <= h <= hi (the one closest to `want`), or build one for `want`
int  rtr_get_grid_any(rtr_cloud* c, float want, float lo, float hi, DevGrid** out);
int  rtr_grid_normals(rtr_cloud* c, DevGrid* g);                     // make g->sorted_normals current
GridView rtr_view(const DevGrid* g);
OneGrid  rtr_one(const DevGrid* g);
ManyGrids rtr_many(const DevGrid* g, const rtr_cloud* c);
// one segment of a segmented grid as an ordinary view: sorted / cell_begin are offset so that positions are LOCAL to the
// segment (0 .. n_k), while sorted[.].w still carries the global original index
GridView rtr_segment_view(const DevGrid* g, const rtr_cloud* c, int k);
// build-or-fetch several grids of a model set in as few launches as possible (one per cell size)
int  rtr_get_grids(rtr_cloud* c, const float* cells, int n_cells, DevGrid** out);
int  rtr_ensure_bbox(rtr_cloud* c);
void rtr_invalidate(rtr_cloud* c);
float rtr_icp_cell(const rtr_cloud* c);
int  rtr_normals_dev(rtr_cloud* c, float radius);
int  rtr_normals_mode_dev(rtr_cloud* c, float radius, int mode);
int  rtr_harris_dev(rtr_cloud* c, float radius, float threshold, int nms, int refine, int** d_kp_idx, float4** d_kp_xyz,
                    int** d_count);
int  rtr_fpfh_dev(rtr_cloud* c, float radius);
int  rtr_match_dev(rtr_cloud* src, rtr_cloud* tgt, int k);
int  rtr_ransac_dev(rtr_cloud* src, rtr_cloud* tgt, const rtr_ransac_params* p, rtr_pose_result* d_result);
int  rtr_icp_dev(rtr_cloud* src, rtr_cloud* tgt, const rtr_icp_params* p, const float* d_init_pose16, int init_from_result,
                 rtr_pose_result* d_result);
int  rtr_validate_register_params(const rtr_register_params* p);
int  rtr_comm_allgather_dev(rtr_context* ctx, const rtr_pose_result* d_local, int n_local);
int  rtr_nearest_bvh_dev(rtr_cloud* tgt, const float4* d_q, int nq, int* d_idx, float* d_d2);
// model sets: member clouds (device handles / host buffers) concatenated into one segmented rtr_cloud; free with rtr_cloud_free
int  rtr_model_set_from_clouds(rtr_context* ctx, rtr_cloud* const* members, int nseg, rtr_cloud** out);
int  rtr_model_set_from_host(rtr_context* ctx, const float* const* host_xyz1, const int* ns, int nseg, rtr_cloud** out);
int  rtr_match_features_dev(rtr_context* ctx, const float* fa, int na, const float* fb, int nb, int k, int* out_idx, float* out_dist);
long long rtr_next_generation();          // process-wide, never repeats (a freed handle's address may be reused, its stamps are not)

// launch with programmatic stream serialization (see pdl_wait below); RTR_PDL=0 falls back to ordinary launches
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
    static const bool enabled = []() { const char* e = getenv("RTR_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be
// scheduled while its predecessor in the stream is still finishing; pdl_wait() blocks until the predecessor has completed
// and its writes are visible (no-op for an ordinary launch), pdl_launch_dependents() tells the scheduler this CTA no
// longer needs the SM slots the successor is waiting for.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// A kernel launched with the attribute must execute pdl_wait before it exits, or its own completion would no longer imply
// its predecessors'.  Used by the ICP iteration kernels: the next iteration's CTAs are scheduled while the last CTA of the
// current one folds the partials and solves (one registration: 0.58 -> 0.55 ms; 100 k -> 1 M ICP: 29.5 k -> 32.3 k
// iterations/s).  Extending it to every kernel of the registration chain was measured: no further latency gain and -1.2 %
// on the 8-registration step (early-resident CTAs hold SM slots other streams could use), so only ICP uses it.

__device__ __forceinline__ float dist2f(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// pcl::transformPointCloud semantics: ((m00*x + m01*y) + m02*z) + m03, column-major m, float, no FMA
__device__ __forceinline__ float4 xform(const float* __restrict__ m, float4 p) {
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], p.x), __fmul_rn(m[4], p.y)), __fmul_rn(m[8], p.z)), m[12]);
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[1], p.x), __fmul_rn(m[5], p.y)), __fmul_rn(m[9], p.z)), m[13]);
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[2], p.x), __fmul_rn(m[6], p.y)), __fmul_rn(m[10], p.z)), m[14]);
    o.w = 1.0f;
    return o;
}

// c = a * b (4x4 column-major float), k left to right, no FMA
__device__ __forceinline__ void matmul4(const float* a, const float* b, float* c) {
    float out[16];
#pragma unroll
    for (int col = 0; col < 4; ++col)
#pragma unroll
        for (int row = 0; row < 4; ++row) {
            float acc = __fmul_rn(a[row], b[col * 4]);
#pragma unroll
            for (int k = 1; k < 4; ++k) acc = __fadd_rn(acc, __fmul_rn(a[k * 4 + row], b[col * 4 + k]));
            out[col * 4 + row] = acc;
        }
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = out[i];
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t rand_below(uint64_t seed, uint64_t h, uint32_t draw, uint32_t n) {
    uint64_t r = mix64(mix64(seed + 0x9E3779B97F4A7C15ULL * (h + 1)) + 0x9E3779B97F4A7C15ULL * (uint64_t)(draw + 1));
    return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// unclamped cell coordinate of a float coordinate
__device__ __forceinline__ int cell_coord(float v, float mn, float inv_h) {
    float f = floorf(__fmul_rn(__fsub_rn(v, mn), inv_h));
    // keep the int conversion defined for far-away / non-finite queries
    if (!(f > -1.0e9f)) f = -1.0e9f;
    if (f > 1.0e9f) f = 1.0e9f;
    return (int)f;
}

__device__ __forceinline__ int cell_key(const GridView& g, int x, int y, int z) { return (z * g.dy + y) * g.dx + x; }

// Visit every point of the 3x3x3 cell block around q: f(point_with_index, d2).  The three x-adjacent cells of a row are
// one contiguous range of the cell-sorted array, so a query touches 9 ranges.  Returns false if q is farther than one
// cell outside the grid (no neighbour within h possible).
template <typename F>
__device__ __forceinline__ bool for_block27(const GridView& g, float qx, float qy, float qz, F&& f) {
    int cx = cell_coord(qx, g.mnx, g.inv_h), cy = cell_coord(qy, g.mny, g.inv_h), cz = cell_coord(qz, g.mnz, g.inv_h);
    if (cx < -1 || cy < -1 || cz < -1 || cx > g.dx || cy > g.dy || cz > g.dz) return false;
    cx = clampi(cx, 0, g.dx - 1); cy = clampi(cy, 0, g.dy - 1); cz = clampi(cz, 0, g.dz - 1);
    int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
    for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dz - 1); ++z)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dy - 1); ++y) {
            int s0 = __ldg(g.cell_begin + cell_key(g, x0, y, z));
            int s1 = __ldg(g.cell_begin + cell_key(g, x1, y, z) + 1);
            for (int s = s0; s < s1; ++s) {
                float4 p = __ldg(g.sorted + s);
                f(s, p, dist2f(qx, qy, qz, p.x, p.y, p.z));
            }
        }
    return true;
}

// Phase 2 of the exact search (see grid_nearest_ex): slab / row walk outward from the query cell, starting from the
// best candidate found so far (best < 0: none).
__device__ __forceinline__ void grid_nearest_far(const GridView& g, float qx, float qy, float qz, float prune2, int& best,
                                                 float& best_d2, float4& bp) {
    int cy = clampi(cell_coord(qy, g.mny, g.inv_h), 0, g.dy - 1);
    int cz = clampi(cell_coord(qz, g.mnz, g.inv_h), 0, g.dz - 1);
    float ext = g.h * (float)max(g.dx, max(g.dy, g.dz));
    float slack = g.h * 1e-3f + 2e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + fabsf(g.mnx) + fabsf(g.mny) + fabsf(g.mnz) + ext);
    for (int zdir = 0; zdir < 2; ++zdir) {
        for (int z = zdir == 0 ? cz : cz - 1; z >= 0 && z < g.dz; z += (zdir == 0 ? 1 : -1)) {
            float lo = g.mnz + (float)z * g.h, hi = lo + g.h;
            float dzl = qz < lo ? lo - qz : (qz > hi ? qz - hi : 0.f);
            dzl = fmaxf(dzl - slack, 0.f);
            float bound = fminf(best_d2, prune2);
            if (dzl * dzl > bound) break;
            for (int ydir = 0; ydir < 2; ++ydir) {
                for (int y = ydir == 0 ? cy : cy - 1; y >= 0 && y < g.dy; y += (ydir == 0 ? 1 : -1)) {
                    float lo2 = g.mny + (float)y * g.h, hi2 = lo2 + g.h;
                    float dyl = qy < lo2 ? lo2 - qy : (qy > hi2 ? qy - hi2 : 0.f);
                    dyl = fmaxf(dyl - slack, 0.f);
                    bound = fminf(best_d2, prune2);
                    float rem = bound - (dzl * dzl + dyl * dyl);
                    if (rem < 0.f) break;
                    float rx = sqrtf(rem) * 1.000001f + slack;
                    int xa = clampi(cell_coord(qx - rx, g.mnx, g.inv_h), 0, g.dx - 1);
                    int xb = clampi(cell_coord(qx + rx, g.mnx, g.inv_h), 0, g.dx - 1);
                    int s0 = __ldg(g.cell_begin + cell_key(g, xa, y, z));
                    int s1 = __ldg(g.cell_begin + cell_key(g, xb, y, z) + 1);
                    for (int s = s0; s < s1; ++s) {
                        float4 p = __ldg(g.sorted + s);
                        float d = dist2f(qx, qy, qz, p.x, p.y, p.z);
                        int id = __float_as_int(p.w);
                        if (d < best_d2 || (d == best_d2 && id < best)) { best_d2 = d; best = id; bp = p; }
                    }
                }
            }
        }
    }
}

// Exact nearest neighbour; ties -> lowest original index.  prune2: candidates farther than this (squared) are of no
// interest to the caller (ICP's max correspondence distance), FLT_MAX for none.
//   phase 1: the 3x3x3 cell block around the (clamped) query cell — 9 contiguous ranges.  Every point outside that
//            block is farther than one cell size along some axis, so best_d2 <= (0.999 h)^2 ends the search
//            (the common case in ICP).
//   phase 2: otherwise walk slabs (z) and rows (y) outward from the query cell, pruning each slab / row by its exact
//            box distance to the query and clipping the x range of a row to the current search sphere.  Cost is
//            proportional to the rows inside the sphere, also for queries far outside the grid.
// `slack` covers float cell assignment (a point may sit a hair outside its cell's nominal box).
__device__ __forceinline__ void grid_nearest_ex(const GridView& g, float qx, float qy, float qz, float prune2, int& best,
                                                float& best_d2, float4& bp) {
    best = -1; best_d2 = FLT_MAX; bp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.n == 0) return;
    int cx = clampi(cell_coord(qx, g.mnx, g.inv_h), 0, g.dx - 1);
    int cy = clampi(cell_coord(qy, g.mny, g.inv_h), 0, g.dy - 1);
    int cz = clampi(cell_coord(qz, g.mnz, g.inv_h), 0, g.dz - 1);
    {
        int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
        for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dz - 1); ++z)
            for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dy - 1); ++y) {
                int s0 = __ldg(g.cell_begin + cell_key(g, x0, y, z));
                int s1 = __ldg(g.cell_begin + cell_key(g, x1, y, z) + 1);
                for (int s = s0; s < s1; ++s) {
                    float4 p = __ldg(g.sorted + s);
                    float d = dist2f(qx, qy, qz, p.x, p.y, p.z);
                    int id = __float_as_int(p.w);
                    if (d < best_d2 || (d == best_d2 && id < best)) { best_d2 = d; best = id; bp = p; }
                }
            }
    }
    float hh = g.h * 0.999f;
    if (best >= 0 && best_d2 <= hh * hh) return;
    if (prune2 <= hh * hh) return;        // everything outside the block is farther than the caller cares about
    grid_nearest_far(g, qx, qy, qz, prune2, best, best_d2, bp);
}

// The 3x3x3 block of a query is 9 contiguous ranges; their 18 bounds are fetched by 18 lanes at once (one round trip
// instead of a chain of 18 dependent loads) and handed out with shuffles.  Range r = zi * 3 + yi over the clipped block.
struct BlockRanges { int z0, z1, y0, y1; int bound; };
__device__ __forceinline__ BlockRanges warp_block_ranges(const GridView& g, int cx, int cy, int cz, int lane) {
    BlockRanges br;
    br.z0 = max(cz - 1, 0); br.z1 = min(cz + 1, g.dz - 1); br.y0 = max(cy - 1, 0); br.y1 = min(cy + 1, g.dy - 1);
    int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
    int r = lane % 9, zi = r / 3, yi = r - zi * 3;
    int z = br.z0 + zi, y = br.y0 + yi;
    br.bound = 0;
    if (lane < 18 && z <= br.z1 && y <= br.y1)
        br.bound = (lane < 9) ? __ldg(g.cell_begin + cell_key(g, x0, y, z)) : __ldg(g.cell_begin + cell_key(g, x1, y, z) + 1);
    return br;
}

// The 9 ranges of a 27-cell block as ONE candidate list for a warp: candidate j of [0, total) lives at sorted position
// locate(j), in the same order as walking the ranges one after another.  Walking range by range costs a dependent memory
// round trip per range and leaves most lanes idle when a range holds 3-5 points (sparse clouds: ~30 neighbours over 9
// ranges); the flat list needs ceil(total / 32) steps with every lane busy and all loads of a step in flight together.
struct WarpCand { int total; int pre[9]; int first[9]; };
__device__ __forceinline__ WarpCand warp_candidates(const GridView& g, int cx, int cy, int cz, int lane) {
    BlockRanges br = warp_block_ranges(g, cx, cy, cz, lane);
    int end = __shfl_sync(0xffffffffu, br.bound, (lane + 9) & 31);
    int len = lane < 9 ? max(end - br.bound, 0) : 0;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    WarpCand w;
    w.total = __shfl_sync(0xffffffffu, incl, 8);
    int excl = incl - len;
#pragma unroll
    for (int k = 0; k < 9; ++k) { w.pre[k] = __shfl_sync(0xffffffffu, excl, k); w.first[k] = __shfl_sync(0xffffffffu, br.bound, k); }
    return w;
}
// The same list with its tables in shared memory (tab[0..8] = exclusive prefix, tab[9..17] = range begin), for kernels
// that cannot spare 18 registers; returns the total.  tab is this warp's own; the caller __syncwarp()s before reuse.
__device__ __forceinline__ int warp_candidates_smem(const GridView& g, int cx, int cy, int cz, int lane, int* tab) {
    BlockRanges br = warp_block_ranges(g, cx, cy, cz, lane);
    int end = __shfl_sync(0xffffffffu, br.bound, (lane + 9) & 31);
    int len = lane < 9 ? max(end - br.bound, 0) : 0;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane < 9) { tab[lane] = incl - len; tab[9 + lane] = br.bound; }
    __syncwarp();
    return __shfl_sync(0xffffffffu, incl, 8);
}
__device__ __forceinline__ int locate(const WarpCand& w, int j) {
    int start = 0, first = w.first[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) if (j >= w.pre[k]) { start = w.pre[k]; first = w.first[k]; }
    return first + (j - start);
}

// Warp-cooperative form of grid_nearest_ex for small query sets (one warp per query): the lanes stride over the points
// of every visited range and the (d2, index) minimum is folded with shuffles.  Same visiting rules, same result.
__device__ __forceinline__ void warp_argmin(float& d, int& id, float4& p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, d, o);
        int oi = __shfl_xor_sync(0xffffffffu, id, o);
        float ox = __shfl_xor_sync(0xffffffffu, p.x, o), oy = __shfl_xor_sync(0xffffffffu, p.y, o), oz = __shfl_xor_sync(0xffffffffu, p.z, o);
        bool take = (oi >= 0) && (id < 0 || od < d || (od == d && oi < id));
        if (take) { d = od; id = oi; p.x = ox; p.y = oy; p.z = oz; }
    }
}
__device__ __forceinline__ void grid_nearest_warp(const GridView& g, float qx, float qy, float qz, float prune2, int lane, int& best,
                                                  float& best_d2, float4& bp) {
    best = -1; best_d2 = FLT_MAX; bp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.n == 0) return;
    int cx = clampi(cell_coord(qx, g.mnx, g.inv_h), 0, g.dx - 1);
    int cy = clampi(cell_coord(qy, g.mny, g.inv_h), 0, g.dy - 1);
    int cz = clampi(cell_coord(qz, g.mnz, g.inv_h), 0, g.dz - 1);
    {
        WarpCand w = warp_candidates(g, cx, cy, cz, lane);
        for (int j = lane; j < w.total; j += 32) {
            float4 p = __ldg(g.sorted + locate(w, j));
            float d = dist2f(qx, qy, qz, p.x, p.y, p.z);
            int id = __float_as_int(p.w);
            if (best < 0 || d < best_d2 || (d == best_d2 && id < best)) { best_d2 = d; best = id; bp = p; }
        }
    }
    warp_argmin(best_d2, best, bp);
    float hh = g.h * 0.999f;
    if (best >= 0 && best_d2 <= hh * hh) return;
    if (prune2 <= hh * hh) return;        // everything outside the block is farther than the caller cares about
    float ext = g.h * (float)max(g.dx, max(g.dy, g.dz));
    float slack = g.h * 1e-3f + 2e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + fabsf(g.mnx) + fabsf(g.mny) + fabsf(g.mnz) + ext);
    // Phase 2: square rings of x-rows around (cy, cz), 32 rows per step — one row per lane: box-distance test, x range
    // clipped to the current sphere, the two range bounds.  Surfaces leave most rows empty, so the lanes mostly retire
    // empty rows in parallel; the few non-empty ones are then scanned by the whole warp.  Ring k lies at least
    // (k - 1) h away in y or z, which ends the walk.
    const int kmax = max(max(cy, g.dy - 1 - cy), max(cz, g.dz - 1 - cz));
    for (int k = 0; k <= kmax; ++k) {
        float bound = fminf(best >= 0 ? best_d2 : FLT_MAX, prune2);      // best is warp-uniform here
        float rm = fmaxf((float)(k - 1) * g.h - slack, 0.f);
        if (rm * rm > bound) break;
        const int nrows = k == 0 ? 1 : 8 * k;
        for (int i0 = 0; i0 < nrows; i0 += 32) {
            int i = i0 + lane, s0 = 0, s1 = 0;
            if (i < nrows) {
                int oy = 0, oz = 0;
                if (k > 0) {
                    int side = i / (2 * k), o = i - side * 2 * k;
                    if (side == 0) { oy = -k + o; oz = -k; }
                    else if (side == 1) { oy = k; oz = -k + o; }
                    else if (side == 2) { oy = k - o; oz = k; }
                    else { oy = -k; oz = k - o; }
                }
                int y = cy + oy, z = cz + oz;
                if (y >= 0 && y < g.dy && z >= 0 && z < g.dz) {
                    float lo = g.mnz + (float)z * g.h, hi = lo + g.h;
                    float dzl = qz < lo ? lo - qz : (qz > hi ? qz - hi : 0.f);
                    dzl = fmaxf(dzl - slack, 0.f);
                    float lo2 = g.mny + (float)y * g.h, hi2 = lo2 + g.h;
                    float dyl = qy < lo2 ? lo2 - qy : (qy > hi2 ? qy - hi2 : 0.f);
                    dyl = fmaxf(dyl - slack, 0.f);
                    float rem = bound - (dzl * dzl + dyl * dyl);
                    if (rem >= 0.f) {
                        float rx = sqrtf(rem) * 1.000001f + slack;
                        int xa = clampi(cell_coord(qx - rx, g.mnx, g.inv_h), 0, g.dx - 1);
                        int xb = clampi(cell_coord(qx + rx, g.mnx, g.inv_h), 0, g.dx - 1);
                        s0 = __ldg(g.cell_begin + cell_key(g, xa, y, z));
                        s1 = __ldg(g.cell_begin + cell_key(g, xb, y, z) + 1);
                    }
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, s1 > s0);
            if (m) {
                while (m) {
                    int j = __ffs(m) - 1;
                    m &= m - 1;
                    int ra = __shfl_sync(0xffffffffu, s0, j), rb = __shfl_sync(0xffffffffu, s1, j);
                    for (int s = ra + lane; s < rb; s += 32) {
                        float4 p = __ldg(g.sorted + s);
                        float d = dist2f(qx, qy, qz, p.x, p.y, p.z);
                        int id = __float_as_int(p.w);
                        if (best < 0 || d < best_d2 || (d == best_d2 && id < best)) { best_d2 = d; best = id; bp = p; }
                    }
                }
                warp_argmin(best_d2, best, bp);     // keep the pruning bound uniform across the warp
                bound = fminf(best >= 0 ? best_d2 : FLT_MAX, prune2);
            }
        }
    }
}

// Small targets (every scan the reference ships has ~2-3 k points): no search structure at all.  One warp answers TWO
// queries at once — the lanes stride over all target points (30 KB, L1 resident), each loaded point is tested against
// both queries — then two shuffle arg-mins.  ~n/32 x 14 instructions per query whatever the alignment of the clouds,
// against an unbounded ring walk for queries far from the target.  Same distance expression and tie rule: same result.
#define RTR_BRUTE_NN_MAX 4096
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, o);
        k = ok < k ? ok : k;
    }
    return k;
}
// (d2 bits, original index) packed in one 64-bit key: d2 >= +0, so its bit pattern orders like the value, and the
// unsigned minimum is "smallest distance, then lowest index" — the tie rule — in one compare.  NaN distances (bits above
// +inf) never count as a match.
__device__ __forceinline__ void brute_nearest_warp2(const GridView& g, float ax, float ay, float az, float bx, float by, float bz, int lane,
                                                    int& best_a, float& d2_a, int& best_b, float& d2_b) {
    unsigned long long ka = ~0ull, kb = ~0ull;
#pragma unroll 4
    for (int s = lane; s < g.n; s += 32) {
        float4 p = __ldg(g.sorted + s);
        unsigned id = (unsigned)__float_as_int(p.w);
        unsigned long long ca = ((unsigned long long)__float_as_uint(dist2f(ax, ay, az, p.x, p.y, p.z)) << 32) | id;
        unsigned long long cb = ((unsigned long long)__float_as_uint(dist2f(bx, by, bz, p.x, p.y, p.z)) << 32) | id;
        ka = ca < ka ? ca : ka;
        kb = cb < kb ? cb : kb;
    }
    ka = warp_min_u64(ka);
    kb = warp_min_u64(kb);
    unsigned ha = (unsigned)(ka >> 32), hb = (unsigned)(kb >> 32);
    best_a = ha <= 0x7f800000u ? (int)(unsigned)ka : -1; d2_a = __uint_as_float(ha);
    best_b = hb <= 0x7f800000u ? (int)(unsigned)kb : -1; d2_b = __uint_as_float(hb);
}

__device__ __forceinline__ void grid_nearest(const GridView& g, float qx, float qy, float qz, int& best, float& best_d2) {
    float4 bp;
    grid_nearest_ex(g, qx, qy, qz, FLT_MAX, best, best_d2, bp);
}

// cyclic Jacobi, symmetric NxN, fp64, + - * / sqrt only (same operation sequence as the oracle's restatement)
template <int N>
__device__ __forceinline__ void jacobi_eig(double (&a)[N][N], double (&v)[N][N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < N; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) off += fabs(a[p][q]);
        if (off == 0.0) break;
#pragma unroll
        for (int p = 0; p < N; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                double apq = a[p][q];
                if (apq != 0.0) {
                    double g = 100.0 * fabs(apq);
                    if (sweep > 3 && fabs(a[p][p]) + g == fabs(a[p][p]) && fabs(a[q][q]) + g == fabs(a[q][q])) {
                        a[p][q] = 0.0; a[q][p] = 0.0;
                    } else {
                        double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                        if (theta < 0) t = -t;
                        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                        a[p][p] = a[p][p] - t * apq;
                        a[q][q] = a[q][q] + t * apq;
                        a[p][q] = 0.0; a[q][p] = 0.0;
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            if (r != p && r != q) {
                                double arp = a[r][p], arq = a[r][q];
                                a[r][p] = c * arp - s * arq; a[p][r] = a[r][p];
                                a[r][q] = s * arp + c * arq; a[q][r] = a[r][q];
                            }
                            double vrp = v[r][p], vrq = v[r][q];
                            v[r][p] = c * vrp - s * vrq;
                            v[r][q] = s * vrp + c * vrq;
                        }
                    }
                }
            }
    }
}

// Largest eigenpair of the symmetric 4x4 Horn matrix without an iterative diagonalisation: the characteristic quartic's
// top root by Newton from an upper bound (monotone for real-rooted polynomials), then one column of adj(N - lambda I).
// + - * / sqrt only, fixed order.  ~40x shorter dependent chain than cyclic Jacobi (the ICP solve is one serial thread).
// Returns false when the top eigenvalue is not well separated (collinear / symmetric inputs): the caller falls back
// to jacobi_eig, which handles repeated eigenvalues gracefully.
__device__ __forceinline__ void adj4(double m00, double m01, double m02, double m03, double m10, double m11, double m12, double m13,
                                  double m20, double m21, double m22, double m23, double m30, double m31, double m32, double m33,
                                  double* A, double* det) {
    double s0 = m00 * m11 - m10 * m01, s1 = m00 * m12 - m10 * m02, s2 = m00 * m13 - m10 * m03;
    double s3 = m01 * m12 - m11 * m02, s4 = m01 * m13 - m11 * m03, s5 = m02 * m13 - m12 * m03;
    double c5 = m22 * m33 - m32 * m23, c4 = m21 * m33 - m31 * m23, c3 = m21 * m32 - m31 * m22;
    double c2 = m20 * m33 - m30 * m23, c1 = m20 * m32 - m30 * m22, c0 = m20 * m31 - m30 * m21;
    *det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
    A[0] = (m11 * c5 - m12 * c4) + m13 * c3;   A[1] = (m02 * c4 - m01 * c5) - m03 * c3;
    A[2] = (m31 * s5 - m32 * s4) + m33 * s3;   A[3] = (m22 * s4 - m21 * s5) - m23 * s3;
    A[4] = (m12 * c2 - m10 * c5) - m13 * c1;   A[5] = (m00 * c5 - m02 * c2) + m03 * c1;
    A[6] = (m32 * s2 - m30 * s5) - m33 * s1;   A[7] = (m20 * s5 - m22 * s2) + m23 * s1;
    A[8] = (m10 * c4 - m11 * c2) + m13 * c0;   A[9] = (m01 * c2 - m00 * c4) - m03 * c0;
    A[10] = (m30 * s4 - m31 * s2) + m33 * s0;  A[11] = (m21 * s2 - m20 * s4) - m23 * s0;
    A[12] = (m11 * c1 - m10 * c3) - m12 * c0;  A[13] = (m00 * c3 - m01 * c1) + m02 * c0;
    A[14] = (m31 * s1 - m30 * s3) - m32 * s0;  A[15] = (m20 * s3 - m21 * s1) + m22 * s0;
}
__device__ __forceinline__ bool horn_top_eigvec(const double (&N)[4][4], double (&q)[4]) {
    const double a = N[0][0], b = N[0][1], c = N[0][2], d = N[0][3], e = N[1][1], f = N[1][2], g = N[1][3], h = N[2][2], i = N[2][3], j = N[3][3];
    const double F2 = (((a * a + e * e) + h * h) + j * j) + 2.0 * (((((b * b + c * c) + d * d) + f * f) + g * g) + i * i);
    if (!(F2 > 0.0) || !(F2 < 1e200)) return false;
    // det(x I - N) = x^4 + k3 x^3 + k2 x^2 + k1 x + k0 : k3 = -trace, k2 = sum of principal 2x2 minors,
    // k1 = -(sum of principal 3x3 minors) = -(trace of adj N), k0 = det N
    double A[16], det;
    adj4(a, b, c, d, b, e, f, g, c, f, h, i, d, g, i, j, A, &det);
    const double k3 = -(((a + e) + h) + j);
    const double k2 = (((((a * e - b * b) + (a * h - c * c)) + (a * j - d * d)) + (e * h - f * f)) + (e * j - g * g)) + (h * j - i * i);
    const double k1 = -(((A[0] + A[5]) + A[10]) + A[15]);
    const double k0 = det;
    // Newton from above the largest root (all roots real, sum ~ 0  =>  max <= sqrt(3/4 sum x^2)): monotone descent
    double lam = sqrt(0.75 * F2) * 1.000000001;
    for (int it = 0; it < 64; ++it) {
        double p = (((lam + k3) * lam + k2) * lam + k1) * lam + k0;
        double dp = ((4.0 * lam + 3.0 * k3) * lam + 2.0 * k2) * lam + k1;
        if (!(dp > 0.0)) return false;
        double ln = lam - p / dp;
        if (it > 0 && !(ln < lam)) break;
        lam = ln;
    }
    // eigenvector: adj(N - lam I) = (product of the gaps) v v^T ; take the column with the largest diagonal entry
    adj4(a - lam, b, c, d, b, e - lam, f, g, c, f, h - lam, i, d, g, i, j - lam, A, &det);
    int col = 0;
    double best = fabs(A[0]);
    if (fabs(A[5]) > best) { best = fabs(A[5]); col = 1; }
    if (fabs(A[10]) > best) { best = fabs(A[10]); col = 2; }
    if (fabs(A[15]) > best) { best = fabs(A[15]); col = 3; }
    q[0] = A[0 * 4 + col]; q[1] = A[1 * 4 + col]; q[2] = A[2 * 4 + col]; q[3] = A[3 * 4 + col];
    const double nq2 = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
    // well separated top eigenvalue only (product of gaps > ~1e-4 |N|^3); otherwise the caller runs the Jacobi solve
    return nq2 > 1e-9 * ((F2 * F2) * F2);
}

// Horn's unit-quaternion rigid fit from raw fp64 sums (ss = sum s, st = sum t, m[a][b] = sum s_a t_b, n pairs);
// writes a column-major float 4x4 (source -> target).
__device__ __forceinline__ void horn_pose(const double* ss, const double* st, const double* m9, double n, float* pose) {
    double cs[3], ct[3], S[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { cs[a] = ss[a] / n; ct[a] = st[a] / n; }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) S[a][b] = m9[a * 3 + b] - n * cs[a] * ct[b];
    double N[4][4], V[4][4];
    N[0][0] = S[0][0] + S[1][1] + S[2][2];
    N[0][1] = S[1][2] - S[2][1]; N[0][2] = S[2][0] - S[0][2]; N[0][3] = S[0][1] - S[1][0];
    N[1][1] = S[0][0] - S[1][1] - S[2][2];
    N[1][2] = S[0][1] + S[1][0]; N[1][3] = S[2][0] + S[0][2];
    N[2][2] = -S[0][0] + S[1][1] - S[2][2];
    N[2][3] = S[1][2] + S[2][1];
    N[3][3] = -S[0][0] - S[1][1] + S[2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < i) N[i][j] = N[j][i];
    double q0, q1, q2, q3;
    double qv[4];
    if (horn_top_eigvec(N, qv)) { q0 = qv[0]; q1 = qv[1]; q2 = qv[2]; q3 = qv[3]; }
    else {
        jacobi_eig<4>(N, V);
        double best = N[0][0];
        q0 = V[0][0]; q1 = V[1][0]; q2 = V[2][0]; q3 = V[3][0];
#pragma unroll
        for (int i = 1; i < 4; ++i) if (N[i][i] > best) { best = N[i][i]; q0 = V[0][i]; q1 = V[1][i]; q2 = V[2][i]; q3 = V[3][i]; }
    }
    double nq = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 /= nq; q1 /= nq; q2 /= nq; q3 /= nq;
    double R[3][3];
    R[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3; R[0][1] = 2.0 * (q1 * q2 - q0 * q3); R[0][2] = 2.0 * (q1 * q3 + q0 * q2);
    R[1][0] = 2.0 * (q1 * q2 + q0 * q3); R[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3; R[1][2] = 2.0 * (q2 * q3 - q0 * q1);
    R[2][0] = 2.0 * (q1 * q3 - q0 * q2); R[2][1] = 2.0 * (q2 * q3 + q0 * q1); R[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double t = ct[r] - ((R[r][0] * cs[0] + R[r][1] * cs[1]) + R[r][2] * cs[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) pose[c * 4 + r] = (float)R[r][c];
        pose[12 + r] = (float)t;
    }
    pose[3] = 0.f; pose[7] = 0.f; pose[11] = 0.f; pose[15] = 1.f;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ bool finite3(float4 n) { return isfinite(n.x) && isfinite(n.y) && isfinite(n.z); }

#endif  // __CUDACC__
