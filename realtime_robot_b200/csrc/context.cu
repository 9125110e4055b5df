// context.cu — contexts, device-resident clouds, the uniform-grid neighbour index, and the two index-level entry
// points (radius sets, 1-NN).  The grid replaces pcl::search::KdTree / KdTreeFLANN, which the reference only reaches
// through HarrisKeypoint3D (model_point.h:127-136) and IterativeClosestPoint (function.h:112-117).
//
// Layout in HBM: a cloud is n x float4 (pcl::PointXYZ's own 16-byte layout).  A grid holds the same points permuted
// into cell order (key = (z*dy + y)*dx + x, stable, so in-cell order is ascending original index) with the original
// index in .w, plus cell_begin[ncells+1].  The three x-adjacent cells of a row are one contiguous range, so a
// 27-cell query is 9 coalesced range scans.
#include "common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>

// ----------------------------------------------------------------------------- kernels
__global__ void k_cell_keys(const float4* __restrict__ pts, int n, float mnx, float mny, float mnz, float inv_h, int dx,
                            int dy, int dz, int* __restrict__ keys, int* __restrict__ vals, int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = __ldg(pts + i);
    int cx = clampi(cell_coord(p.x, mnx, inv_h), 0, dx - 1);
    int cy = clampi(cell_coord(p.y, mny, inv_h), 0, dy - 1);
    int cz = clampi(cell_coord(p.z, mnz, inv_h), 0, dz - 1);
    int key = (cz * dy + cy) * dx + cx;
    keys[i] = key;
    vals[i] = i;
    atomicAdd(counts + key, 1);
}

__global__ void k_gather_sorted(const float4* __restrict__ pts, const int* __restrict__ order, int n, float4* __restrict__ sorted) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = order[s];
    float4 p = __ldg(pts + i);
    p.w = __int_as_float(i);
    sorted[s] = p;
}

// Grid build for repo-sized clouds in ONE launch: a thread-block CLUSTER of 8 CTAs (8192 threads on 8 SMs) walks the
// phases histogram -> exclusive scan -> scatter -> in-cell ranking by original index (the same stable order the
// radix-sort path produces) -> gather, with hardware cluster barriers between them.  Replaces memset + keys + radix
// sort + scan + gather (7-8 launches, each a few microseconds of pure latency at these sizes).
#define GB_THREADS 1024
#define GB_CLUSTER 8
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Tiny clouds (<= 4096 points, <= 8192 cells: chair1, chair2, Chair_025 and every scan the reference ships): the whole
// build in ONE CTA with keys, cell table and slots in shared memory — __syncthreads instead of six cluster barriers and no
// trips through L2 between the phases.  Same result as the kernels below: points ordered by (cell, original index).
//   counts -> inclusive scan (cell ends) -> scatter by atomicSub from the cell end (the table ends up holding the cell
//   BEGINS) -> rank inside the cell by original index -> write.
#define GB1_MAXN 4096
#define GB1_MAXC 8192
__global__ void __launch_bounds__(GB_THREADS)
k_grid_build_tiny(const float4* __restrict__ pts, int n, float mnx, float mny, float mnz, float inv_h, int dx, int dy, int dz, int ncells,
                  int* __restrict__ cell_begin, float4* __restrict__ sorted) {
    extern __shared__ int gb1_smem[];
    int* s_key = gb1_smem;                 // n
    int* s_slot = gb1_smem + n;            // n
    int* s_cell = gb1_smem + 2 * n;        // ncells + 1
    __shared__ int warp_tot[GB_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int c = t; c <= ncells; c += GB_THREADS) s_cell[c] = 0;
    __syncthreads();
    for (int i = t; i < n; i += GB_THREADS) {
        float4 p = __ldg(pts + i);
        int cx = clampi(cell_coord(p.x, mnx, inv_h), 0, dx - 1);
        int cy = clampi(cell_coord(p.y, mny, inv_h), 0, dy - 1);
        int cz = clampi(cell_coord(p.z, mnz, inv_h), 0, dz - 1);
        int key = (cz * dy + cy) * dx + cx;
        s_key[i] = key;
        atomicAdd(s_cell + key, 1);
    }
    __syncthreads();
    // inclusive scan of the counts: contiguous chunk per thread, shuffle scan of the chunk sums, warp totals through smem
    const int per = (ncells + GB_THREADS - 1) / GB_THREADS;
    const int c0 = min(t * per, ncells), c1 = min(c0 + per, ncells);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += s_cell[c];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
        warp_tot[lane] = w;
    }
    __syncthreads();
    int run = incl - sum + (warp > 0 ? warp_tot[warp - 1] : 0);
    for (int c = c0; c < c1; ++c) { run += s_cell[c]; s_cell[c] = run; }       // s_cell[c] = end of cell c
    if (t == 0) s_cell[ncells] = n;
    __syncthreads();
    for (int i = t; i < n; i += GB_THREADS) s_slot[atomicSub(s_cell + s_key[i], 1) - 1] = i;     // unordered inside a cell; s_cell -> begins
    __syncthreads();
    for (int c = t; c <= ncells; c += GB_THREADS) cell_begin[c] = s_cell[c];
    for (int pos = t; pos < n; pos += GB_THREADS) {
        const int i = s_slot[pos];
        const int key = s_key[i];
        const int b = s_cell[key], e = s_cell[key + 1];
        int rank = 0;
        for (int q = b; q < e; ++q) rank += (s_slot[q] < i) ? 1 : 0;
        float4 p = __ldg(pts + i);
        p.w = __int_as_float(i);
        sorted[b + rank] = p;
    }
}

__global__ void __cluster_dims__(GB_CLUSTER, 1, 1) __launch_bounds__(GB_THREADS)
k_grid_build_small(const float4* __restrict__ pts, int n, float mnx, float mny, float mnz, float inv_h, int dx, int dy, int dz, int ncells,
                   int* cell_begin, int* keys, int* cursor, int* slot, int* cta_total, float4* __restrict__ sorted) {
    __shared__ int part[GB_THREADS];
    const int t = threadIdx.x, cta = blockIdx.x;                 // one cluster per launch: blockIdx.x = rank in the cluster
    const int gt = cta * GB_THREADS + t, GT = GB_CLUSTER * GB_THREADS;
    for (int c = gt; c <= ncells; c += GT) cell_begin[c] = 0;
    cluster_sync_all();
    for (int i = gt; i < n; i += GT) {
        float4 p = __ldg(pts + i);
        int cx = clampi(cell_coord(p.x, mnx, inv_h), 0, dx - 1);
        int cy = clampi(cell_coord(p.y, mny, inv_h), 0, dy - 1);
        int cz = clampi(cell_coord(p.z, mnz, inv_h), 0, dz - 1);
        int key = (cz * dy + cy) * dx + cx;
        keys[i] = key;
        atomicAdd(cell_begin + key, 1);
    }
    cluster_sync_all();
    // exclusive scan of cell_begin[0 .. ncells] in place: contiguous chunk per thread, block scan of the chunk sums,
    // then the totals of the lower-ranked CTAs
    int per = (ncells + 1 + GT - 1) / GT;
    int c0 = min(gt * per, ncells + 1), c1 = min(c0 + per, ncells + 1);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += cell_begin[c];
    part[t] = sum;
    __syncthreads();
    for (int off = 1; off < GB_THREADS; off <<= 1) {
        int v = (t >= off) ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    if (t == GB_THREADS - 1) cta_total[cta] = part[t];
    cluster_sync_all();
    int run = part[t] - sum;
    for (int k = 0; k < cta; ++k) run += cta_total[k];
    for (int c = c0; c < c1; ++c) { int v = cell_begin[c]; cell_begin[c] = run; cursor[c] = run; run += v; }
    cluster_sync_all();
    for (int i = gt; i < n; i += GT) slot[atomicAdd(cursor + keys[i], 1)] = i;       // unordered inside a cell
    cluster_sync_all();
    // rank inside the cell = number of members with a smaller original index -> deterministic ascending order
    for (int pos = gt; pos < n; pos += GT) {
        int i = slot[pos];
        int key = keys[i];
        int b = cell_begin[key], e = cell_begin[key + 1];
        int rank = 0;
        for (int q = b; q < e; ++q) rank += (slot[q] < i) ? 1 : 0;
        float4 p = __ldg(pts + i);
        p.w = __int_as_float(i);
        sorted[b + rank] = p;
    }
}

// The same build for every cloud of a model set in ONE launch: cluster j builds segment j (blockIdx.x = 8 j + rank).
// Positions written to the cell table and the index in sorted[.].w are GLOBAL (segment start + local), so the segment
// tables concatenate into one array (see DevGrid::segs).  keys / slot are per point, cursor mirrors the cell table.
__global__ void __cluster_dims__(GB_CLUSTER, 1, 1) __launch_bounds__(GB_THREADS)
k_grid_build_multi(const __grid_constant__ ManyGrids mg, const float4* __restrict__ pts, int* cell_table, int* keys_all, int* cursor_all,
                   int* slot_all, int* cta_total_all, float4* __restrict__ sorted) {
    __shared__ int part[GB_THREADS];
    const int job = blockIdx.x / GB_CLUSTER, cta = blockIdx.x % GB_CLUSTER;
    const SegHdr hd = mg.seg[job];
    const int p0 = mg.begin[job], n = mg.begin[job + 1] - p0;
    const int ncells = hd.dx * hd.dy * hd.dz;
    int* cell_begin = cell_table + hd.cell_off;
    int* cursor = cursor_all + hd.cell_off;
    int* keys = keys_all + p0;
    int* cta_total = cta_total_all + job * GB_CLUSTER;
    const int t = threadIdx.x;
    const int gt = cta * GB_THREADS + t, GT = GB_CLUSTER * GB_THREADS;
    // local counts first; the last segment also owns the closing entry of the whole table
    const int last = (job == mg.nseg - 1) ? 1 : 0;
    for (int c = gt; c < ncells + last; c += GT) cell_begin[c] = 0;
    cluster_sync_all();
    for (int i = gt; i < n; i += GT) {
        float4 p = __ldg(pts + p0 + i);
        int cx = clampi(cell_coord(p.x, hd.mnx, hd.inv_h), 0, hd.dx - 1);
        int cy = clampi(cell_coord(p.y, hd.mny, hd.inv_h), 0, hd.dy - 1);
        int cz = clampi(cell_coord(p.z, hd.mnz, hd.inv_h), 0, hd.dz - 1);
        int key = (cz * hd.dy + cy) * hd.dx + cx;
        keys[i] = key;
        atomicAdd(cell_begin + key, 1);
    }
    cluster_sync_all();
    // exclusive scan of the ncells counts (+ the closing entry of the last segment), offset by the segment start
    const int nscan = ncells + last;
    int per = (nscan + GT - 1) / GT;
    int c0 = min(gt * per, nscan), c1 = min(c0 + per, nscan);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += cell_begin[c];
    part[t] = sum;
    __syncthreads();
    for (int off = 1; off < GB_THREADS; off <<= 1) {
        int v = (t >= off) ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    if (t == GB_THREADS - 1) cta_total[cta] = part[t];
    cluster_sync_all();
    int run = p0 + part[t] - sum;
    for (int k = 0; k < cta; ++k) run += cta_total[k];
    for (int c = c0; c < c1; ++c) { int v = cell_begin[c]; cell_begin[c] = run; if (c < ncells) cursor[c] = run; run += v; }
    cluster_sync_all();
    for (int i = gt; i < n; i += GT) slot_all[atomicAdd(cursor + keys[i], 1)] = i;       // unordered inside a cell; slot_all is indexed by GLOBAL position
    cluster_sync_all();
    // rank inside the cell = number of members with a smaller original index -> deterministic ascending order.
    // The end of a cell is the next table entry; for the last cell of a segment that is the next segment's first entry
    // (= this segment's end), written by another cluster: use the segment end directly there.
    for (int pos = gt; pos < n; pos += GT) {
        int i = slot_all[p0 + pos];
        int key = keys[i];
        int b = cell_begin[key], e = (key + 1 < ncells) ? cell_begin[key + 1] : p0 + n;
        int rank = 0;
        for (int q = b; q < e; ++q) rank += (slot_all[q] < i) ? 1 : 0;
        float4 p = __ldg(pts + p0 + i);
        p.w = __int_as_float(p0 + i);
        sorted[b + rank] = p;
    }
}

// model set from device clouds: one gather over a pointer table (RTR_MAX_SEGMENTS members at most)
struct ConcatTable { int nseg; int begin[RTR_MAX_SEGMENTS + 1]; const float4* src[RTR_MAX_SEGMENTS]; };
__global__ void k_concat_clouds(const __grid_constant__ ConcatTable t, float4* __restrict__ out) {
    const int total = t.begin[RTR_MAX_SEGMENTS];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int k = 0;
#pragma unroll
        for (int step = RTR_MAX_SEGMENTS / 2; step > 0; step >>= 1) if (i >= t.begin[k + step]) k += step;
        out[i] = __ldg(t.src[k] + (i - t.begin[k]));
    }
}

__global__ void k_permute4(const float4* __restrict__ src, const float4* __restrict__ sorted, int n, float4* __restrict__ dst) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    dst[s] = __ldg(src + __float_as_int(__ldg(sorted + s).w));
}

__device__ __forceinline__ unsigned enc_f(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ __forceinline__ float dec_f(unsigned u) {
    unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(v);
#else
    memcpy(&f, &v, 4);
#endif
    return f;
}

// bounding box of finite points: 6 ordered-uint atomics (min x,y,z, max x,y,z)
__global__ void k_bbox(const float4* __restrict__ pts, int n, unsigned* __restrict__ box) {
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = __ldg(pts + i);
        if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
            unsigned e[3] = {enc_f(p.x), enc_f(p.y), enc_f(p.z)};
#pragma unroll
            for (int a = 0; a < 3; ++a) { mn[a] = min(mn[a], e[a]); mx[a] = max(mx[a], e[a]); }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(box + a, mn[a]); atomicMax(box + 3 + a, mx[a]); }
    }
}

__global__ void k_transform_inplace(float4* __restrict__ pts, int n, const float* __restrict__ pose) {
    __shared__ float m[16];
    if (threadIdx.x < 16) m[threadIdx.x] = pose[threadIdx.x];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pts[i] = xform(m, pts[i]);
}

// radius sets of the cloud's own points: pass 0 counts, pass 1 fills (then sorts each list ascending in place)
__global__ void k_radius_count(GridView g, float r2, int* __restrict__ counts) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n) return;
    float4 q = __ldg(g.sorted + s);
    int c = 0;
    for_block27(g, q.x, q.y, q.z, [&](int, float4, float d2) { if (d2 < r2) ++c; });
    counts[__float_as_int(q.w)] = c;
}
__global__ void k_radius_fill(GridView g, float r2, const long long* __restrict__ offsets, int* __restrict__ indices) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n) return;
    float4 q = __ldg(g.sorted + s);
    int* out = indices + offsets[__float_as_int(q.w)];
    int c = 0;
    for_block27(g, q.x, q.y, q.z, [&](int, float4 p, float d2) {
        if (d2 < r2) {   // insertion keeps the list ascending
            int id = __float_as_int(p.w), j = c++;
            while (j > 0 && out[j - 1] > id) { out[j] = out[j - 1]; --j; }
            out[j] = id;
        }
    });
}

__global__ void k_nearest(GridView g, const float4* __restrict__ q, int nq, int* __restrict__ idx, float* __restrict__ d2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float4 p = __ldg(q + i);
    int b; float bd;
    grid_nearest(g, p.x, p.y, p.z, b, bd);
    idx[i] = b; d2[i] = bd;
}

// ----------------------------------------------------------------------------- host
static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

long long rtr_next_generation() { static std::atomic<long long> g{0}; return ++g; }

static std::mutex g_attr_mu;
static std::map<std::pair<int, const void*>, int> g_attr_smem;            // (device, kernel) -> bytes already granted
static std::map<std::pair<int, const void*>, int> g_attr_occ;             // (device, kernel) -> resident CTAs per SM
int rtr_func_smem(const void* func, int device, int bytes) {
    std::lock_guard<std::mutex> lock(g_attr_mu);
    auto key = std::make_pair(device, func);
    auto it = g_attr_smem.find(key);
    if (it != g_attr_smem.end() && it->second >= bytes) return 0;
    RTR_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "smem.attr");
    g_attr_smem[key] = bytes;
    return 0;
}
int rtr_func_occupancy(const void* func, int device, int block, size_t smem, int* per_sm) {
    std::lock_guard<std::mutex> lock(g_attr_mu);
    auto key = std::make_pair(device, func);
    auto it = g_attr_occ.find(key);
    if (it != g_attr_occ.end()) { *per_sm = it->second; return 0; }
    int v = 0;
    RTR_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, func, block, smem), "occupancy");
    v = std::max(v, 1);
    g_attr_occ[key] = v;
    *per_sm = v;
    return 0;
}

GridView rtr_view(const DevGrid* g) {
    GridView v;
    v.inv_h = g->inv_h; v.h = g->h; v.mnx = g->mnx; v.mny = g->mny; v.mnz = g->mnz;
    v.dx = g->dx; v.dy = g->dy; v.dz = g->dz; v.n = g->n; v.cell_begin = g->cell_begin; v.sorted = g->sorted;
    return v;
}

OneGrid rtr_one(const DevGrid* g) { OneGrid o; o.g = rtr_view(g); return o; }

ManyGrids rtr_many(const DevGrid* g, const rtr_cloud* c) {
    ManyGrids m;
    memset(&m, 0, sizeof(m));
    m.nseg = c->nseg();
    for (int k = 0; k <= RTR_MAX_SEGMENTS; ++k) m.begin[k] = k < m.nseg ? c->seg_begin[k] : c->n;
    for (int k = 0; k < m.nseg; ++k) m.seg[k] = g->segs[k];
    m.cell_begin = g->cell_begin; m.sorted = g->sorted;
    return m;
}

GridView rtr_segment_view(const DevGrid* g, const rtr_cloud* c, int k) {
    GridView v;
    const SegHdr& h = g->segs[k];
    v.inv_h = h.inv_h; v.h = h.h; v.mnx = h.mnx; v.mny = h.mny; v.mnz = h.mnz; v.dx = h.dx; v.dy = h.dy; v.dz = h.dz;
    v.n = c->seg_begin[k + 1] - c->seg_begin[k];
    v.cell_begin = g->cell_begin + h.cell_off; v.sorted = g->sorted;
    return v;
}

int rtr_ensure_bbox(rtr_cloud* c) {
    if (c->bbox_valid) return 0;
    rtr_context* ctx = c->ctx;
    unsigned* d_box = nullptr;
    if (int e = tmp_alloc(ctx, &d_box, 6, "bbox")) return e;
    unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    RTR_CHECK(cudaMemcpyAsync(d_box, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream), "bbox");
    if (c->n > 0) {
        k_bbox<<<std::min(nblk(c->n, 256), 4 * ctx->sm_count), 256, 0, ctx->stream>>>(c->pts, c->n, d_box);
        RTR_LAUNCH_CHECK(ctx, "bbox");
    }
    unsigned h_box[6];
    RTR_CHECK(cudaMemcpyAsync(h_box, d_box, sizeof(h_box), cudaMemcpyDeviceToHost, ctx->stream), "bbox");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "bbox");
    dev_free(ctx, d_box);
    for (int a = 0; a < 3; ++a) {
        if (h_box[a] == 0xffffffffu && h_box[3 + a] == 0u) { c->bb_min[a] = 0.f; c->bb_max[a] = 0.f; }
        else { c->bb_min[a] = dec_f(h_box[a]); c->bb_max[a] = dec_f(h_box[3 + a]); }
    }
    c->bbox_valid = true;
    return 0;
}

static void free_grid(rtr_context* ctx, DevGrid& g) {
    dev_free(ctx, g.cell_begin); dev_free(ctx, g.sorted); dev_free(ctx, g.sorted_normals);
    g.cell_begin = nullptr; g.sorted = nullptr; g.sorted_normals = nullptr;
}

void rtr_invalidate(rtr_cloud* c) {
    rtr_context* ctx = c->ctx;
    for (auto& kv : c->grids) free_grid(ctx, kv.second);
    c->grids.clear();
    dev_free(ctx, c->normals); c->normals = nullptr; c->normals_radius = -1.f; c->normals_version++;
    dev_free(ctx, c->response); c->response = nullptr;
    dev_free(ctx, c->fpfh); c->fpfh = nullptr; c->fpfh_radius = -1.f; c->feature_gen = rtr_next_generation();
    dev_free(ctx, c->knn); c->knn = nullptr; dev_free(ctx, c->knn_dist); c->knn_dist = nullptr; c->knn_k = 0; c->knn_target = nullptr;
    c->n_keypoints = -1;
    dev_free(ctx, c->kp_xyz); c->kp_xyz = nullptr; dev_free(ctx, c->kp_count); c->kp_count = nullptr; c->prepared = false;
}

float rtr_icp_cell(const rtr_cloud* c) {
    // ~2x the mean sample spacing, estimated from the bounding-box surface (clouds here are surface samples)
    double ex = (double)c->bb_max[0] - c->bb_min[0], ey = (double)c->bb_max[1] - c->bb_min[1], ez = (double)c->bb_max[2] - c->bb_min[2];
    double area = 2.0 * (ex * ey + ey * ez + ex * ez);
    static const double factor = []() { const char* e = getenv("RTR_ICP_CELL_FACTOR"); return e ? atof(e) : 2.0; }();
    double cell = factor * std::sqrt(std::max(area, 1e-12) / (double)std::max(c->n, 1));
    return (float)std::max(cell, 1e-4);
}

// cell size and dimensions of the grid that serves radius `cell` over a bounding box.  The cell is 0.1 % wider than the
// radius so that float cell assignment can never put two points that are within the radius more than one cell apart
// (DESIGN.md "grid exactness"); dims are capped (2048 per axis, 2^26 cells) by widening the cell.
static void grid_dims(const float* bb_min, const float* bb_max, float cell, double max_cells, float* h_out, int* d) {
    double h = (double)cell * 1.001;
    for (;;) {
        double cells = 1;
        bool ok = true;
        for (int a = 0; a < 3; ++a) {
            double ext = (double)bb_max[a] - (double)bb_min[a];
            double da = std::floor(ext / h) + 2;     // +1 for the floor, +1 slack for float rounding at the upper face
            if (da > 2048) ok = false;
            d[a] = (int)std::min(da, 4096.0);
            cells *= d[a];
        }
        if (ok && cells <= max_cells) break;
        h *= 1.25;
    }
    *h_out = (float)h;
}

// Model sets: one launch builds the grids of every member cloud for one cell size (clusters of 8 CTAs, one per segment).
int rtr_get_grids(rtr_cloud* c, const float* cells, int n_cells, DevGrid** out) {
    RtrRange nvtx_range("rtr.grid_build.many");
    rtr_context* ctx = c->ctx;
    const int nseg = c->nseg();
    if (nseg <= 0 || nseg > RTR_MAX_SEGMENTS) return rtr_fail("grid", "rtr_get_grids needs a model set of 1..32 clouds", RTR_ERR_INVALID);
    for (int j = 0; j < n_cells; ++j) {
        const float cell = cells[j];
        if (!(cell > 0.f) || !std::isfinite(cell)) return rtr_fail("grid", "cell size (search radius) must be finite and > 0", RTR_ERR_INVALID);
        int keybits;
        memcpy(&keybits, &cell, 4);
        auto it = c->grids.find(keybits);
        if (it != c->grids.end()) { out[j] = &it->second; continue; }
        DevGrid g;
        g.n = c->n;
        g.segs.resize(nseg);
        long long cell_total = 0;
        for (int k = 0; k < nseg; ++k) {
            SegHdr& hd = g.segs[k];
            int d[3];
            const float* bb = &c->seg_bb[6 * k];
            grid_dims(bb, bb + 3, cell, 4194304.0, &hd.h, d);
            hd.inv_h = 1.0f / hd.h; hd.mnx = bb[0]; hd.mny = bb[1]; hd.mnz = bb[2];
            hd.dx = d[0]; hd.dy = d[1]; hd.dz = d[2];
            hd.cell_off = (int)cell_total;
            cell_total += (long long)d[0] * d[1] * d[2];
        }
        if (cell_total > (1LL << 28)) return rtr_fail("grid", "model set needs too many grid cells", RTR_ERR_INVALID);
        g.h = g.segs[0].h; g.inv_h = g.segs[0].inv_h; g.ncells = (int)cell_total;
        int *keys = nullptr, *cursor = nullptr, *slot = nullptr, *cta_total = nullptr;
        if (int e = tmp_alloc(ctx, &cta_total, (size_t)GB_CLUSTER * nseg, "grid")) return e;
        if (int e = tmp_alloc(ctx, &keys, c->n, "grid")) return e;
        if (int e = tmp_alloc(ctx, &cursor, (size_t)cell_total + 1, "grid")) return e;
        if (int e = tmp_alloc(ctx, &slot, c->n, "grid")) return e;
        if (int e = dev_alloc(ctx, &g.cell_begin, (size_t)cell_total + 1, "grid")) return e;
        if (int e = dev_alloc(ctx, &g.sorted, c->n, "grid")) return e;
        ManyGrids mg = rtr_many(&g, c);
        k_grid_build_multi<<<GB_CLUSTER * nseg, GB_THREADS, 0, ctx->stream>>>(mg, c->pts, g.cell_begin, keys, cursor, slot, cta_total, g.sorted);
        RTR_LAUNCH_CHECK(ctx, "grid.build_multi");
        dev_free(ctx, keys); dev_free(ctx, cursor); dev_free(ctx, slot); dev_free(ctx, cta_total);
        auto ins = c->grids.emplace(keybits, g);
        out[j] = &ins.first->second;
    }
    return 0;
}

int rtr_get_grid(rtr_cloud* c, float cell, DevGrid** out) {
    RtrRange nvtx_range("rtr.grid");
    if (c->nseg() > 0) return rtr_get_grids(c, &cell, 1, out);
    // a zero / negative / non-finite cell would never satisfy the dimension caps below (h *= 1.25 keeps 0 at 0)
    if (!(cell > 0.f) || !std::isfinite(cell)) return rtr_fail("grid", "cell size (search radius) must be finite and > 0", RTR_ERR_INVALID);
    int keybits;
    memcpy(&keybits, &cell, 4);
    auto it = c->grids.find(keybits);
    if (it != c->grids.end()) { *out = &it->second; return 0; }
    rtr_context* ctx = c->ctx;
    if (int e = rtr_ensure_bbox(c)) return e;
    DevGrid g;
    g.n = c->n;
    {
        int d[3];
        grid_dims(c->bb_min, c->bb_max, cell, 67108864.0, &g.h, d);
        g.dx = d[0]; g.dy = d[1]; g.dz = d[2];
    }
    g.inv_h = 1.0f / g.h;
    g.mnx = c->bb_min[0]; g.mny = c->bb_min[1]; g.mnz = c->bb_min[2];
    g.ncells = g.dx * g.dy * g.dz;
    int n = c->n;
    if (n > 0 && n <= GB1_MAXN && g.ncells <= GB1_MAXC) {
        if (int e = rtr_kernel_smem(k_grid_build_tiny, ctx, (2 * GB1_MAXN + GB1_MAXC + 1) * sizeof(int))) return e;
        if (int e = dev_alloc(ctx, &g.cell_begin, (size_t)g.ncells + 1, "grid")) return e;
        if (int e = dev_alloc(ctx, &g.sorted, n, "grid")) return e;
        k_grid_build_tiny<<<1, GB_THREADS, (2 * (size_t)n + g.ncells + 1) * sizeof(int), ctx->stream>>>(c->pts, n, g.mnx, g.mny, g.mnz, g.inv_h, g.dx, g.dy, g.dz,
                                                                                                     g.ncells, g.cell_begin, g.sorted);
        RTR_LAUNCH_CHECK(ctx, "grid.build_small");
        auto ins = c->grids.emplace(keybits, g);
        *out = &ins.first->second;
        return 0;
    }
    if (n > 0 && n <= 65536 && g.ncells <= (1 << 22)) {
        int *keys = nullptr, *cursor = nullptr, *slot = nullptr, *cta_total = nullptr;
        if (int e = tmp_alloc(ctx, &cta_total, GB_CLUSTER, "grid")) return e;
        if (int e = tmp_alloc(ctx, &keys, n, "grid")) return e;
        if (int e = tmp_alloc(ctx, &cursor, (size_t)g.ncells + 1, "grid")) return e;
        if (int e = tmp_alloc(ctx, &slot, n, "grid")) return e;
        if (int e = dev_alloc(ctx, &g.cell_begin, (size_t)g.ncells + 1, "grid")) return e;
        if (int e = dev_alloc(ctx, &g.sorted, n, "grid")) return e;
        k_grid_build_small<<<GB_CLUSTER, GB_THREADS, 0, ctx->stream>>>(c->pts, n, g.mnx, g.mny, g.mnz, g.inv_h, g.dx, g.dy, g.dz, g.ncells,
                                                                        g.cell_begin, keys, cursor, slot, cta_total, g.sorted);
        RTR_LAUNCH_CHECK(ctx, "grid.build_small");
        dev_free(ctx, keys); dev_free(ctx, cursor); dev_free(ctx, slot); dev_free(ctx, cta_total);
        auto ins = c->grids.emplace(keybits, g);
        *out = &ins.first->second;
        return 0;
    }
    int *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr, *counts = nullptr;
    if (int e = tmp_alloc(ctx, &keys, n, "grid")) return e;
    if (int e = tmp_alloc(ctx, &vals, n, "grid")) return e;
    if (int e = tmp_alloc(ctx, &keys2, n, "grid")) return e;
    if (int e = tmp_alloc(ctx, &vals2, n, "grid")) return e;
    if (int e = tmp_alloc(ctx, &counts, (size_t)g.ncells + 1, "grid")) return e;
    if (int e = dev_alloc(ctx, &g.cell_begin, (size_t)g.ncells + 1, "grid")) return e;
    if (int e = dev_alloc(ctx, &g.sorted, n, "grid")) return e;
    RTR_CHECK(cudaMemsetAsync(counts, 0, ((size_t)g.ncells + 1) * sizeof(int), ctx->stream), "grid");
    RTR_MARK(ctx, "grid.memset");
    if (n > 0) {
        k_cell_keys<<<nblk(n, 256), 256, 0, ctx->stream>>>(c->pts, n, g.mnx, g.mny, g.mnz, g.inv_h, g.dx, g.dy, g.dz, keys, vals, counts);
        RTR_LAUNCH_CHECK(ctx, "grid.keys");
    }
    int end_bit = 1;
    while ((1LL << end_bit) < (long long)g.ncells) ++end_bit;
    size_t tb_sort = 0, tb_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, keys, keys2, vals, vals2, n, 0, end_bit, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tb_scan, counts, g.cell_begin, g.ncells + 1, ctx->stream);
    void* temp = nullptr;
    size_t tb = std::max(tb_sort, tb_scan);
    if (int e = tmp_alloc(ctx, (char**)&temp, tb, "grid")) return e;
    if (n > 0) RTR_CHECK(cub::DeviceRadixSort::SortPairs(temp, tb_sort, keys, keys2, vals, vals2, n, 0, end_bit, ctx->stream), "grid.sort");
    RTR_MARK(ctx, "grid.cub_sort");
    RTR_CHECK(cub::DeviceScan::ExclusiveSum(temp, tb_scan, counts, g.cell_begin, g.ncells + 1, ctx->stream), "grid.scan");
    RTR_MARK(ctx, "grid.cub_scan");
    if (n > 0) {
        k_gather_sorted<<<nblk(n, 256), 256, 0, ctx->stream>>>(c->pts, vals2, n, g.sorted);
        RTR_LAUNCH_CHECK(ctx, "grid.gather");
    }
    dev_free(ctx, (char*)temp); dev_free(ctx, keys); dev_free(ctx, vals); dev_free(ctx, keys2); dev_free(ctx, vals2); dev_free(ctx, counts);
    auto ins = c->grids.emplace(keybits, g);
    *out = &ins.first->second;
    return 0;
}

// ---- model sets (rtr_register_many): several clouds concatenated into one rtr_cloud with a segment table ----
static int set_new(rtr_context* ctx, const int* ns, int nseg, rtr_cloud** out) {
    if (!ctx || !out || nseg < 1 || nseg > RTR_MAX_SEGMENTS) return rtr_fail("model_set", "1..32 member clouds", RTR_ERR_INVALID);
    long long total = 0;
    for (int k = 0; k < nseg; ++k) { if (ns[k] < 0) return rtr_fail("model_set", "negative size", RTR_ERR_INVALID); total += ns[k]; }
    if (total > 0x7fffffffLL) return rtr_fail("model_set", "too many points", RTR_ERR_INVALID);
    rtr_cloud* c = new rtr_cloud();
    c->ctx = ctx; c->n = (int)total;
    c->seg_begin.resize(nseg + 1);
    c->seg_begin[0] = 0;
    for (int k = 0; k < nseg; ++k) c->seg_begin[k + 1] = c->seg_begin[k] + ns[k];
    c->seg_bb.assign((size_t)6 * nseg, 0.f);
    if (int e = dev_alloc(ctx, &c->pts, (size_t)total, "model_set")) { delete c; return e; }
    *out = c;
    return 0;
}
static void set_finish_bbox(rtr_cloud* c) {
    const int nseg = c->nseg();
    for (int a = 0; a < 3; ++a) { c->bb_min[a] = FLT_MAX; c->bb_max[a] = -FLT_MAX; }
    for (int k = 0; k < nseg; ++k)
        for (int a = 0; a < 3; ++a) { c->bb_min[a] = std::min(c->bb_min[a], c->seg_bb[6 * k + a]); c->bb_max[a] = std::max(c->bb_max[a], c->seg_bb[6 * k + 3 + a]); }
    c->bbox_valid = true;
}

int rtr_model_set_from_clouds(rtr_context* ctx, rtr_cloud* const* members, int nseg, rtr_cloud** out) {
    if (!members || nseg < 1 || nseg > RTR_MAX_SEGMENTS) return rtr_fail("model_set", "1..32 member clouds", RTR_ERR_INVALID);
    int ns[RTR_MAX_SEGMENTS];
    for (int k = 0; k < nseg; ++k) {
        if (!members[k] || members[k]->ctx != ctx || members[k]->nseg() > 0) return rtr_fail("model_set", "members must be ordinary clouds of this context", RTR_ERR_INVALID);
        ns[k] = members[k]->n;
        if (int e = rtr_ensure_bbox(members[k])) return e;
    }
    if (int e = set_new(ctx, ns, nseg, out)) return e;
    rtr_cloud* c = *out;
    ConcatTable t;
    memset(&t, 0, sizeof(t));
    t.nseg = nseg;
    for (int k = 0; k <= RTR_MAX_SEGMENTS; ++k) t.begin[k] = k < nseg ? c->seg_begin[k] : c->n;
    for (int k = 0; k < nseg; ++k) {
        t.src[k] = members[k]->pts;
        for (int a = 0; a < 3; ++a) { c->seg_bb[6 * k + a] = members[k]->bb_min[a]; c->seg_bb[6 * k + 3 + a] = members[k]->bb_max[a]; }
    }
    if (c->n > 0) {
        k_concat_clouds<<<std::min(nblk(c->n, 256), 8 * ctx->sm_count), 256, 0, ctx->stream>>>(t, c->pts);
        RTR_LAUNCH_CHECK(ctx, "set.concat");
    }
    set_finish_bbox(c);
    return 0;
}

// Bounding box of the finite points of a host cloud (n x float4).  It sits on the critical path of every host-buffer entry
// point — the grid dimensions come from it, so no kernel can be queued before it is known — hence SSE: one point per
// 128-bit min / max (a NaN coordinate never wins: MINPS / MAXPS return the second operand then).  Points with an infinite
// coordinate must be skipped as a whole, which the fast loop cannot do: it only notes that one exists and the scalar loop
// redoes the cloud.
#include <emmintrin.h>
static void host_bbox(const float* xyz1, int n, float* mn3, float* mx3) {
    bool any = false, odd = false;
    float mn[4] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX}, mx[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (n > 0 && ((uintptr_t)xyz1 & 15) == 0) {
        __m128 vmn = _mm_set1_ps(FLT_MAX), vmx = _mm_set1_ps(-FLT_MAX);
        const __m128 absmask = _mm_castsi128_ps(_mm_set1_epi32(0x7fffffff)), big = _mm_set1_ps(FLT_MAX);
        __m128 bad = _mm_setzero_ps();
        for (int i = 0; i < n; ++i) {
            const __m128 p = _mm_load_ps(xyz1 + 4 * (size_t)i);
            // not (|p| <= FLT_MAX): NaN or infinity in some lane
            bad = _mm_or_ps(bad, _mm_cmpnle_ps(_mm_and_ps(p, absmask), big));
            vmn = _mm_min_ps(p, vmn);
            vmx = _mm_max_ps(p, vmx);
        }
        odd = (_mm_movemask_ps(bad) & 7) != 0;         // lane 3 is the pad
        _mm_storeu_ps(mn, vmn); _mm_storeu_ps(mx, vmx);
        any = true;
    }
    if (odd || !any) {
        any = false;
        for (int a = 0; a < 3; ++a) { mn[a] = FLT_MAX; mx[a] = -FLT_MAX; }
        for (int i = 0; i < n; ++i) {
            const float* p = xyz1 + 4 * (size_t)i;
            if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) {
                any = true;
                for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
            }
        }
    }
    for (int a = 0; a < 3; ++a) { mn3[a] = any ? mn[a] : 0.f; mx3[a] = any ? mx[a] : 0.f; }
}

int rtr_model_set_from_host(rtr_context* ctx, const float* const* host_xyz1, const int* ns, int nseg, rtr_cloud** out) {
    if (!host_xyz1 || !ns) return rtr_fail("model_set", "bad argument", RTR_ERR_INVALID);
    for (int k = 0; k < nseg; ++k) if (ns[k] > 0 && !host_xyz1[k]) return rtr_fail("model_set", "null points", RTR_ERR_INVALID);
    if (int e = set_new(ctx, ns, nseg, out)) return e;
    rtr_cloud* c = *out;
    for (int k = 0; k < nseg; ++k) {
        const int n = ns[k];
        if (n > 0) RTR_CHECK(cudaMemcpyAsync(c->pts + c->seg_begin[k], host_xyz1[k], (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream), "set.h2d");
        RTR_MARK(ctx, "cloud.h2d");
        // the caller's buffer is host memory: take the bounding box here, so no device round trip is needed later
        host_bbox(host_xyz1[k], n, &c->seg_bb[6 * k], &c->seg_bb[6 * k + 3]);
    }
    set_finish_bbox(c);
    return 0;
}

int rtr_get_grid_any(rtr_cloud* c, float want, float lo, float hi, DevGrid** out) {
    DevGrid* best = nullptr;
    for (auto& kv : c->grids) {
        DevGrid& g = kv.second;
        if (g.h >= lo && g.h <= hi && (!best || std::fabs(g.h - want) < std::fabs(best->h - want))) best = &g;
    }
    if (best) { *out = best; return 0; }
    return rtr_get_grid(c, want, out);
}

int rtr_grid_normals(rtr_cloud* c, DevGrid* g) {
    if (!c->normals) return rtr_fail("grid.normals", "normals have not been computed on this cloud", RTR_ERR_NOT_READY);
    if (g->sorted_normals && g->normals_version == c->normals_version) return 0;
    rtr_context* ctx = c->ctx;
    if (!g->sorted_normals) if (int e = dev_alloc(ctx, &g->sorted_normals, c->n, "grid.normals")) return e;
    if (c->n > 0) {
        k_permute4<<<nblk(c->n, 256), 256, 0, ctx->stream>>>(c->normals, g->sorted, c->n, g->sorted_normals);
        RTR_LAUNCH_CHECK(ctx, "grid.normals");
    }
    g->normals_version = c->normals_version;
    return 0;
}

void rtr_prof_mark(rtr_context* ctx, const char* tag) {
    cudaEvent_t ev;
    if (!ctx->event_pool.empty()) { ev = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
    else if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, ctx->stream);
    ctx->marks.push_back({tag, ev});
}

extern "C" {

// Per-operation device timing for bench.py's roofline: between begin and end every kernel / CUB pass / memset on the
// context's stream is followed by a CUDA event; the interval since the previous mark is attributed to that operation.
int rtr_profile_begin(rtr_context* ctx) {
    if (!ctx) return rtr_fail("profile", "null context", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(ctx->device), "profile");
    for (auto& m : ctx->marks) ctx->event_pool.push_back(m.ev);
    ctx->marks.clear();
    ctx->profile = true;
    rtr_prof_mark(ctx, "(begin)");
    return 0;
}

// Writes "tag count total_ms" lines into buf; returns 0, or RTR_ERR_CAPACITY if buf is too small.
int rtr_profile_end(rtr_context* ctx, char* buf, int capacity) {
    if (!ctx || !buf || capacity < 1) return rtr_fail("profile", "bad argument", RTR_ERR_INVALID);
    ctx->profile = false;
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "profile");
    std::map<std::string, std::pair<long long, double>> agg;
    for (size_t i = 1; i < ctx->marks.size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->marks[i - 1].ev, ctx->marks[i].ev) != cudaSuccess) continue;
        auto& a = agg[ctx->marks[i].tag];
        a.first++; a.second += ms;
    }
    for (auto& m : ctx->marks) ctx->event_pool.push_back(m.ev);
    ctx->marks.clear();
    std::string out;
    char line[256];
    for (auto& kv : agg) { snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second); out += line; }
    if ((int)out.size() + 1 > capacity) return RTR_ERR_CAPACITY;
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

// Drop every cached stage of a cloud (grids, normals, features, correspondences) but keep the points and the bounding
// box: the next stage call recomputes from the points.  bench.py calls this so no step reuses earlier results.
int rtr_cloud_reset(rtr_cloud* c) {
    if (!c) return rtr_fail("reset", "null cloud", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(c->ctx->device), "reset");
    rtr_invalidate(c);
    return 0;
}

void rtr_default_register_params(rtr_register_params* p) {
    memset(p, 0, sizeof(*p));
    p->normal_radius = 0.05f; p->harris_radius = 0.05f; p->harris_threshold = 0.01f; p->harris_nms = 1; p->harris_refine = 1;
    p->fpfh_radius = 0.10f; p->run_icp = 1;
    p->ransac.max_iterations = 50000; p->ransac.seed = 20170427ULL; p->ransac.correspondence_k = 5;
    p->ransac.similarity_threshold = 0.9f; p->ransac.max_correspondence_distance = 0.0365f; p->ransac.inlier_fraction = 0.25f;
    p->icp.max_iterations = 10; p->icp.mse_threshold_absolute = 1e-12;
}

int rtr_context_create(int device, rtr_context** out) { return rtr_context_create_prio(device, 0, out); }

int rtr_context_create_prio(int device, int urgency, rtr_context** out) {
    if (!out) return rtr_fail("context", "null output pointer", RTR_ERR_INVALID);
    *out = nullptr;
    RTR_CHECK(cudaSetDevice(device), "context");
    rtr_context* ctx = new rtr_context();
    ctx->device = device;
    // urgency 0 = default; larger = the device scheduler serves this context's stream first when several contexts compete
    // (longest job first: give the context that registers the largest cloud the highest urgency)
    int least = 0, greatest = 0;
    RTR_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest), "context");      // numerically lower = more urgent
    int prio = std::max(greatest, std::min(least, least - std::max(urgency, 0)));
    RTR_CHECK(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio), "context");
    RTR_CHECK(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio), "context");
    RTR_CHECK(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming), "context");
    RTR_CHECK(cudaEventCreateWithFlags(&ctx->join_event, cudaEventDisableTiming), "context");
    cudaDeviceProp prop;
    RTR_CHECK(cudaGetDeviceProperties(&prop, device), "context");
    ctx->sm_count = prop.multiProcessorCount;
    // keep freed blocks in the context's stream-ordered pool: no cudaMalloc/cudaFree on the hot path (cf. kernel.cu:50-56,101-102)
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    RTR_CHECK(cudaMemPoolCreate(&ctx->pool, &props), "context");
    unsigned long long thresh = ~0ULL;
    RTR_CHECK(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thresh), "context");
    for (int i = 0; i < RTR_NUM_EVENTS; ++i) RTR_CHECK(cudaEventCreate(&ctx->events[i]), "context");
    ctx->pinned_bytes = 1 << 16;
    RTR_CHECK(cudaMallocHost(&ctx->pinned, ctx->pinned_bytes), "context");
    *out = ctx;
    return 0;
}

int rtr_comm_destroy(rtr_context* ctx);

int rtr_context_destroy(rtr_context* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rtr_comm_destroy(ctx);
    for (int i = 0; i < RTR_NUM_EVENTS; ++i) if (ctx->events[i]) cudaEventDestroy(ctx->events[i]);
    for (auto& sl : ctx->slabs) cudaFree(sl.base);
    for (auto& m : ctx->marks) cudaEventDestroy(m.ev);
    for (auto& e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->io_pinned) cudaFreeHost(ctx->io_pinned);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    if (ctx->join_event) cudaEventDestroy(ctx->join_event);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    cudaStreamDestroy(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
    return 0;
}

int rtr_context_sync(rtr_context* ctx) {
    if (!ctx) return rtr_fail("sync", "null context", RTR_ERR_INVALID);
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "sync");
    return 0;
}

void* rtr_context_stream(rtr_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
long long rtr_context_launches(rtr_context* ctx) { return ctx ? ctx->launches : 0; }

int rtr_event_record(rtr_context* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= RTR_NUM_EVENTS) return rtr_fail("event", "bad slot", RTR_ERR_INVALID);
    RTR_CHECK(cudaEventRecord(ctx->events[slot], ctx->stream), "event");
    return 0;
}
int rtr_event_elapsed_ms(rtr_context* ctx, int a, int b, float* ms) {
    if (!ctx || !ms || a < 0 || b < 0 || a >= RTR_NUM_EVENTS || b >= RTR_NUM_EVENTS) return rtr_fail("event", "bad slot", RTR_ERR_INVALID);
    RTR_CHECK(cudaEventSynchronize(ctx->events[b]), "event");
    RTR_CHECK(cudaEventElapsedTime(ms, ctx->events[a], ctx->events[b]), "event");
    return 0;
}

static int cloud_new(rtr_context* ctx, int n, rtr_cloud** out) {
    if (!ctx || !out || n < 0) return rtr_fail("cloud", "bad argument", RTR_ERR_INVALID);
    *out = nullptr;
    RTR_CHECK(cudaSetDevice(ctx->device), "cloud");
    rtr_cloud* c = new rtr_cloud();
    c->ctx = ctx; c->n = n;
    if (int e = dev_alloc(ctx, &c->pts, n, "cloud")) { delete c; return e; }
    *out = c;
    return 0;
}

int rtr_cloud_upload(rtr_context* ctx, const float* host_xyz1, int n, rtr_cloud** out) {
    if (n > 0 && !host_xyz1) return rtr_fail("cloud", "null points", RTR_ERR_INVALID);
    if (int e = cloud_new(ctx, n, out)) return e;
    rtr_cloud* c = *out;
    if (n > 0) RTR_CHECK(cudaMemcpyAsync(c->pts, host_xyz1, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream), "cloud.h2d");
    RTR_MARK(ctx, "cloud.h2d");
    // the caller's buffer is host memory: take the bounding box here, so no device round trip is needed later
    host_bbox(host_xyz1, n, c->bb_min, c->bb_max);
    c->bbox_valid = true;
    return 0;
}

int rtr_cloud_from_device(rtr_context* ctx, const float* dev_xyz1, int n, rtr_cloud** out) {
    if (n > 0 && !dev_xyz1) return rtr_fail("cloud", "null points", RTR_ERR_INVALID);
    if (int e = cloud_new(ctx, n, out)) return e;
    if (n > 0) RTR_CHECK(cudaMemcpyAsync((*out)->pts, dev_xyz1, (size_t)n * 16, cudaMemcpyDeviceToDevice, ctx->stream), "cloud.d2d");
    return 0;
}

int rtr_cloud_free(rtr_cloud* c) {
    if (!c) return 0;
    cudaSetDevice(c->ctx->device);
    rtr_invalidate(c);
    dev_free(c->ctx, c->pts);
    delete c;
    return 0;
}

int rtr_cloud_size(const rtr_cloud* c) { return c ? c->n : -1; }

int rtr_cloud_transform(rtr_cloud* c, const float* pose16) {
    if (!c || !pose16) return rtr_fail("transform", "bad argument", RTR_ERR_INVALID);
    rtr_context* ctx = c->ctx;
    RTR_CHECK(cudaSetDevice(ctx->device), "transform");
    float* d_pose = nullptr;
    if (int e = dev_alloc(ctx, &d_pose, 16, "transform")) return e;
    memcpy(ctx->pinned, pose16, 64);
    RTR_CHECK(cudaMemcpyAsync(d_pose, ctx->pinned, 64, cudaMemcpyHostToDevice, ctx->stream), "transform");
    if (c->n > 0) {
        k_transform_inplace<<<nblk(c->n, 256), 256, 0, ctx->stream>>>(c->pts, c->n, d_pose);
        RTR_LAUNCH_CHECK(ctx, "transform");
    }
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "transform");   // ctx->pinned is reused by the next call
    dev_free(ctx, d_pose);
    rtr_invalidate(c);
    c->bbox_valid = false;
    return 0;
}

int rtr_cloud_download(rtr_cloud* c, float* host_xyz1) {
    if (!c || !host_xyz1) return rtr_fail("download", "bad argument", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(c->ctx->device), "download");
    if (c->n > 0) RTR_CHECK(cudaMemcpyAsync(host_xyz1, c->pts, (size_t)c->n * 16, cudaMemcpyDeviceToHost, c->ctx->stream), "download");
    RTR_CHECK(cudaStreamSynchronize(c->ctx->stream), "download");
    return 0;
}

int rtr_radius_neighbors(rtr_cloud* c, float radius, int* host_counts, long long* host_offsets, int* host_indices,
                         long long capacity, long long* total) {
    if (!c || !host_counts || !(radius > 0.f)) return rtr_fail("radius", "bad argument", RTR_ERR_INVALID);
    rtr_context* ctx = c->ctx;
    RTR_CHECK(cudaSetDevice(ctx->device), "radius");
    DevGrid* g;
    if (int e = rtr_get_grid(c, radius, &g)) return e;
    int n = c->n;
    int* d_counts = nullptr;
    if (int e = dev_alloc(ctx, &d_counts, n, "radius")) return e;
    float r2 = radius * radius;
    if (n > 0) {
        k_radius_count<<<nblk(n, 128), 128, 0, ctx->stream>>>(rtr_view(g), r2, d_counts);
        RTR_LAUNCH_CHECK(ctx, "radius.count");
        RTR_CHECK(cudaMemcpyAsync(host_counts, d_counts, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream), "radius");
    }
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "radius");
    std::vector<long long> off((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + host_counts[i];
    if (host_offsets) memcpy(host_offsets, off.data(), sizeof(long long) * ((size_t)n + 1));
    if (total) *total = off[n];
    int rc = 0;
    if (host_indices) {
        if (capacity < off[n]) rc = rtr_fail("radius", "index buffer too small", RTR_ERR_CAPACITY);
        else if (off[n] > 0) {
            long long* d_off = nullptr; int* d_idx = nullptr;
            if (int e = dev_alloc(ctx, &d_off, (size_t)n + 1, "radius")) return e;
            if (int e = dev_alloc(ctx, &d_idx, (size_t)off[n], "radius")) return e;
            RTR_CHECK(cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream), "radius");
            k_radius_fill<<<nblk(n, 128), 128, 0, ctx->stream>>>(rtr_view(g), r2, d_off, d_idx);
            RTR_LAUNCH_CHECK(ctx, "radius.fill");
            RTR_CHECK(cudaMemcpyAsync(host_indices, d_idx, (size_t)off[n] * 4, cudaMemcpyDeviceToHost, ctx->stream), "radius");
            RTR_CHECK(cudaStreamSynchronize(ctx->stream), "radius");
            dev_free(ctx, d_off); dev_free(ctx, d_idx);
        }
    }
    dev_free(ctx, d_counts);
    return rc;
}

int rtr_nearest(rtr_cloud* target, const float* host_queries_xyz1, int nq, int* host_idx, float* host_d2) {
    if (!target || nq < 0 || (nq > 0 && (!host_queries_xyz1 || !host_idx || !host_d2))) return rtr_fail("nearest", "bad argument", RTR_ERR_INVALID);
    rtr_context* ctx = target->ctx;
    RTR_CHECK(cudaSetDevice(ctx->device), "nearest");
    if (nq == 0) return 0;
    if (int e = rtr_ensure_bbox(target)) return e;
    DevGrid* g;
    if (int e = rtr_get_grid(target, rtr_icp_cell(target), &g)) return e;
    float4* d_q = nullptr; int* d_idx = nullptr; float* d_d2 = nullptr;
    if (int e = dev_alloc(ctx, &d_q, nq, "nearest")) return e;
    if (int e = dev_alloc(ctx, &d_idx, nq, "nearest")) return e;
    if (int e = dev_alloc(ctx, &d_d2, nq, "nearest")) return e;
    RTR_CHECK(cudaMemcpyAsync(d_q, host_queries_xyz1, (size_t)nq * 16, cudaMemcpyHostToDevice, ctx->stream), "nearest");
    // tests: answer from the hierarchies of the ICP kernels (bvh.cuh) — 1: the two-level one for targets of <= 4096 points,
    // 2: the 32-ary one (any size)
    const char* use_bvh = getenv("RTR_NEAREST_BVH");
    if (use_bvh && ((use_bvh[0] == '1' && target->n <= 4096) || use_bvh[0] == '2')) {
        TmpScope tmp_scope(ctx);
        if (int e = rtr_nearest_bvh_dev(target, d_q, nq, d_idx, d_d2)) return e;
    } else {
        k_nearest<<<nblk(nq, 128), 128, 0, ctx->stream>>>(rtr_view(g), d_q, nq, d_idx, d_d2);
        RTR_LAUNCH_CHECK(ctx, "nearest");
    }
    RTR_CHECK(cudaMemcpyAsync(host_idx, d_idx, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream), "nearest");
    RTR_CHECK(cudaMemcpyAsync(host_d2, d_d2, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream), "nearest");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "nearest");
    dev_free(ctx, d_q); dev_free(ctx, d_idx); dev_free(ctx, d_d2);
    return 0;
}

}  // extern "C"
