// pose.cu — pose hypotheses (prerejective RANSAC), ICP refinement, and the whole-registration sequencer.
//
// Replaces, for the path BASELINE.json's north_star names: the reference's consensus step (function.h:35-109, an
// exhaustive loop over keypoint pairs) by pcl::SampleConsensusPrerejective semantics, and
// keyPointICP (function.h:111-123: pcl::IterativeClosestPoint, all defaults; App. A.6).  rtr_register sequences the
// stages the way main() does (RealTimeRobot.cpp:39-105) without leaving the device.
//
// Work decomposition
//   RANSAC: K1 one thread per hypothesis (counter-hash draws, polygon prerejection, warp-aggregated append);
//           K2 one thread per survivor (3-point Horn fit, fp64 Jacobi); K3 one CTA per survivor, persistent over the
//           survivor list: every thread transforms source points and scans 9 cell ranges of the target grid, inlier
//           count / error reduced by warp shuffles; K4 one CTA: lexicographic (error, hypothesis) arg-min.
//   ICP:    per iteration ONE kernel: apply the previous step, exact 1-NN, 17 (SVD estimator) or 29 (point-to-plane 6x6
//           A^T A / A^T b) fp64 sums per thread -> warp-shuffle tree -> per-CTA partials; the last CTA to finish folds the
//           partials in a fixed shape, solves (Horn / 6x6 elimination) and runs the convergence tests.  Iterations are
//           chained with programmatic dependent launch.
#include "common.cuh"
#include "bvh.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

// ============================================================================= RANSAC
struct RansacArgs {
    const float4* src; int ns;
    const float4* tgt; int nt;
    const int* knn; int knn_stride; int k;
    unsigned long long seed;
    long long h_base; int h_count;
    float simsq; float dmax2; float inlier_fraction;
};

__device__ __forceinline__ bool draw_hypothesis(const RansacArgs& a, unsigned long long h, int* s, int* c) {
    unsigned ns = (unsigned)a.ns;
    unsigned x = rand_below(a.seed, h, 0, ns);
    unsigned y = rand_below(a.seed, h, 1, ns - 1); if (y >= x) ++y;
    unsigned z = rand_below(a.seed, h, 2, ns - 2);
    unsigned lo = min(x, y), hi = max(x, y);
    if (z >= lo) ++z;
    if (z >= hi) ++z;
    s[0] = (int)x; s[1] = (int)y; s[2] = (int)z;
    bool ok = true;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        unsigned pick = (a.k > 1) ? rand_below(a.seed, h, 3 + t, (unsigned)a.k) : 0u;
        c[t] = __ldg(a.knn + (size_t)s[t] * a.knn_stride + pick);
        if (c[t] < 0) ok = false;
    }
    return ok;
}

// K1: sample + CorrespondenceRejectorPoly::thresholdPolygon (cardinality 3)
__global__ void __launch_bounds__(256) k_ransac_sample(RansacArgs a, int* __restrict__ survivors, int* __restrict__ count) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (t < a.h_count) {
        int s[3], c[3];
        ok = draw_hypothesis(a, (unsigned long long)(a.h_base + t), s, c);
        if (ok) {
            float4 ps[3], pt[3];
#pragma unroll
            for (int e = 0; e < 3; ++e) { ps[e] = __ldg(a.src + s[e]); pt[e] = __ldg(a.tgt + c[e]); }
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                int f = (e + 1) % 3;
                float ds = dist2f(ps[e].x, ps[e].y, ps[e].z, ps[f].x, ps[f].y, ps[f].z);
                float dt = dist2f(pt[e].x, pt[e].y, pt[e].z, pt[f].x, pt[f].y, pt[f].z);
                float sim = ds < dt ? __fdiv_rn(ds, dt) : __fdiv_rn(dt, ds);
                if (!(sim >= a.simsq)) ok = false;
            }
        }
    }
    unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (mask) {
        int lane = threadIdx.x & 31;
        int leader = __ffs(mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(count, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (ok) survivors[base + __popc(mask & ((1u << lane) - 1u))] = t;
    }
}

// K2: TransformationEstimationSVD on the three sampled pairs (Horn / fp64 Jacobi)
__global__ void __launch_bounds__(64) k_ransac_pose(RansacArgs a, const int* __restrict__ survivors, const int* __restrict__ count,
                                                    float* __restrict__ poses) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *count) return;
    int s[3], c[3];
    draw_hypothesis(a, (unsigned long long)(a.h_base + survivors[t]), s, c);
    double ss[3] = {0, 0, 0}, st[3] = {0, 0, 0}, m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        float4 p = __ldg(a.src + s[e]), q = __ldg(a.tgt + c[e]);
        double sv[3] = {p.x, p.y, p.z}, tv[3] = {q.x, q.y, q.z};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            ss[i] += sv[i]; st[i] += tv[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) m[i * 3 + j] += sv[i] * tv[j];
        }
    }
    float pose[16];
    horn_pose(ss, st, m, 3.0, pose);
#pragma unroll
    for (int i = 0; i < 16; ++i) poses[(size_t)t * 16 + i] = pose[i];
}

// K3: getFitness — one CTA per (surviving hypothesis, slice of the source cloud), persistent over that work list.
// With few survivors and a large source a CTA per hypothesis would leave most SMs idle, hence the slices; the per-slice
// (inlier count, sum d2) partials are folded in slice order by K4.
// Block lists of the target grid: for every cell, the points of its 3x3x3 block as ONE contiguous list.  The inlier test of a
// transformed source point is "minimum distance to the points of my cell's block", and walking the block as 9 row ranges
// (for_block27) costs a warp the MAXIMUM trip count over its lanes nine times over: 7 of 32 lanes were busy in the cell walk and
// it was 70 % of k_ransac_eval_many's 148 M warp instructions (ncu, round 2).  With the block flattened once per scan a query is
// one (begin, end) lookup and one loop.  27 x the target's points (825 KB for a 1909-point scan), built by two small kernels and
// one scan; the candidate SET of a query is for_block27's, so the minimum — the only thing used — is the same bit for bit.
struct BlockLists { const int* begin; const float4* pts; };      // begin: ncells + 1 entries; begin == nullptr: walk the grid
__global__ void k_block_count(GridView g, int ncells, int* __restrict__ counts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncells) return;
    int cnt = 0;
    if (c < ncells) {
        const int x = c % g.dx, y = (c / g.dx) % g.dy, z = c / (g.dx * g.dy);
        const int x0 = max(x - 1, 0), x1 = min(x + 1, g.dx - 1);
        for (int zz = max(z - 1, 0); zz <= min(z + 1, g.dz - 1); ++zz)
            for (int yy = max(y - 1, 0); yy <= min(y + 1, g.dy - 1); ++yy)
                cnt += __ldg(g.cell_begin + cell_key(g, x1, yy, zz) + 1) - __ldg(g.cell_begin + cell_key(g, x0, yy, zz));
    }
    counts[c] = cnt;             // counts[ncells] = 0: the exclusive scan's last entry is the total
}
__global__ void k_block_fill(GridView g, int ncells, const int* __restrict__ begin, float4* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int x = c % g.dx, y = (c / g.dx) % g.dy, z = c / (g.dx * g.dy);
    const int x0 = max(x - 1, 0), x1 = min(x + 1, g.dx - 1);
    int o = __ldg(begin + c);
    for (int zz = max(z - 1, 0); zz <= min(z + 1, g.dz - 1); ++zz)
        for (int yy = max(y - 1, 0); yy <= min(y + 1, g.dy - 1); ++yy) {
            const int s0 = __ldg(g.cell_begin + cell_key(g, x0, yy, zz)), s1 = __ldg(g.cell_begin + cell_key(g, x1, yy, zz) + 1);
            for (int s = s0; s < s1; ++s) out[o++] = __ldg(g.sorted + s);
        }
}
// minimum squared distance from q to the points of its cell's block (FLT_MAX: none) — the candidates of for_block27(g, q)
__device__ __forceinline__ float block_min_d2(const GridView& g, const BlockLists& bl, float qx, float qy, float qz) {
    int cx = cell_coord(qx, g.mnx, g.inv_h), cy = cell_coord(qy, g.mny, g.inv_h), cz = cell_coord(qz, g.mnz, g.inv_h);
    if (cx < -1 || cy < -1 || cz < -1 || cx > g.dx || cy > g.dy || cz > g.dz) return FLT_MAX;
    cx = clampi(cx, 0, g.dx - 1); cy = clampi(cy, 0, g.dy - 1); cz = clampi(cz, 0, g.dz - 1);
    const int key = cell_key(g, cx, cy, cz);
    const int s0 = __ldg(bl.begin + key), s1 = __ldg(bl.begin + key + 1);
    float best = FLT_MAX;
    for (int s = s0; s < s1; ++s) {
        const float4 p = __ldg(bl.pts + s);
        const float d2 = dist2f(qx, qy, qz, p.x, p.y, p.z);
        if (d2 < best) best = d2;
    }
    return best;
}
__device__ __forceinline__ float inlier_min_d2(const GridView& g, const BlockLists& bl, float qx, float qy, float qz) {
    if (bl.begin) return block_min_d2(g, bl, qx, qy, qz);
    float best = FLT_MAX;
    for_block27(g, qx, qy, qz, [&](int, float4, float d2) { if (d2 < best) best = d2; });
    return best;
}
// RTR_RANSAC_BLOCKLISTS=0 keeps the grid walk (A/B runs, tests).  Lists are built for grids of up to 4 M cells.
static int block_lists_build_dev(rtr_context* ctx, const GridView& v, BlockLists* out) {
    out->begin = nullptr; out->pts = nullptr;
    static const bool wanted = []() { const char* e = getenv("RTR_RANSAC_BLOCKLISTS"); return !(e && e[0] == '0'); }();
    const long long nc = (long long)v.dx * v.dy * v.dz;
    if (!wanted || v.n <= 0 || nc <= 0 || nc > (1 << 22) || v.n > (1 << 21)) return 0;
    const int ncells = (int)nc;
    int *counts = nullptr, *begin = nullptr; float4* pts = nullptr; char* temp = nullptr;
    if (int e = tmp_alloc(ctx, &counts, (size_t)ncells + 1, "ransac.blocks")) return e;
    if (int e = tmp_alloc(ctx, &begin, (size_t)ncells + 1, "ransac.blocks")) return e;
    if (int e = tmp_alloc(ctx, &pts, (size_t)27 * v.n, "ransac.blocks")) return e;
    k_block_count<<<nblk((long long)ncells + 1, 256), 256, 0, ctx->stream>>>(v, ncells, counts);
    RTR_LAUNCH_CHECK(ctx, "ransac.block_count");
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, counts, begin, ncells + 1, ctx->stream);
    if (int e = tmp_alloc(ctx, &temp, tb, "ransac.blocks")) return e;
    RTR_CHECK(cub::DeviceScan::ExclusiveSum(temp, tb, counts, begin, ncells + 1, ctx->stream), "ransac.blocks");
    RTR_MARK(ctx, "ransac.block_scan");
    k_block_fill<<<nblk(ncells, 256), 256, 0, ctx->stream>>>(v, ncells, begin, pts);
    RTR_LAUNCH_CHECK(ctx, "ransac.block_fill");
    out->begin = begin; out->pts = pts;
    return 0;
}

#define EVAL_THREADS 256
__global__ void __launch_bounds__(EVAL_THREADS) k_ransac_eval(RansacArgs a, GridView g, const int* __restrict__ count,
                                                              const float* __restrict__ poses, int split,
                                                              double* __restrict__ psum, int* __restrict__ pcnt, BlockLists bl) {
    __shared__ float m[16];
    __shared__ double wsum[EVAL_THREADS / 32];
    __shared__ int wcnt[EVAL_THREADS / 32];
    int n = *count;
    int chunk = (a.ns + split - 1) / split;
    for (int w = blockIdx.x; w < n * split; w += gridDim.x) {
        int t = w / split, part = w - t * split;
        int i0 = part * chunk, i1 = min(a.ns, i0 + chunk);
        __syncthreads();
        if (threadIdx.x < 16) m[threadIdx.x] = poses[(size_t)t * 16 + threadIdx.x];
        __syncthreads();
        int cnt = 0;
        double sum = 0;
        for (int i = i0 + threadIdx.x; i < i1; i += EVAL_THREADS) {
            float4 q = xform(m, __ldg(a.src + i));
            const float best = inlier_min_d2(g, bl, q.x, q.y, q.z);
            if (best < a.dmax2) { ++cnt; sum += (double)best; }
        }
        cnt = warp_sum(cnt);
        sum = warp_sum(sum);
        if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = sum; wcnt[threadIdx.x >> 5] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double S = 0; int C = 0;
#pragma unroll
            for (int ww = 0; ww < EVAL_THREADS / 32; ++ww) { S += wsum[ww]; C += wcnt[ww]; }
            psum[w] = S; pcnt[w] = C;
        }
    }
}

// K4: accept iff inlier fraction >= threshold and error < best; parallel form: arg-min over (error, hypothesis),
// folded into the running best of earlier chunks.
__global__ void __launch_bounds__(1024) k_ransac_select(RansacArgs a, const int* __restrict__ survivors, const int* __restrict__ count,
                                                        const float* __restrict__ poses, int split, const double* __restrict__ psum,
                                                        const int* __restrict__ pcnt, rtr_pose_result* __restrict__ res) {
    __shared__ float s_err[32];
    __shared__ long long s_h[32];
    __shared__ int s_t[32];
    int n = *count;
    float be = FLT_MAX; long long bh = -1; int bt = -1;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        double S = 0; int c = 0;
        for (int p = 0; p < split; ++p) { S += psum[(size_t)t * split + p]; c += pcnt[(size_t)t * split + p]; }
        float frac = __fdiv_rn((float)c, (float)a.ns);
        if (frac >= a.inlier_fraction) {
            float e = c > 0 ? (float)(S / (double)c) : FLT_MAX;
            long long h = a.h_base + survivors[t];
            if (e < be || (e == be && (bh < 0 || h < bh))) { be = e; bh = h; bt = t; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float oe = __shfl_xor_sync(0xffffffffu, be, o);
        long long oh = __shfl_xor_sync(0xffffffffu, bh, o);
        int ot = __shfl_xor_sync(0xffffffffu, bt, o);
        if (oh >= 0 && (bh < 0 || oe < be || (oe == be && oh < bh))) { be = oe; bh = oh; bt = ot; }
    }
    if ((threadIdx.x & 31) == 0) { s_err[threadIdx.x >> 5] = be; s_h[threadIdx.x >> 5] = bh; s_t[threadIdx.x >> 5] = bt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        be = FLT_MAX; bh = -1; bt = -1;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            float oe = s_err[w]; long long oh = s_h[w]; int ot = s_t[w];
            if (oh >= 0 && (bh < 0 || oe < be || (oe == be && oh < bh))) { be = oe; bh = oh; bt = ot; }
        }
        res->evaluated += n;
        // strict "error < lowest_error" with lowest_error starting at FLT_MAX, sequential in h == lexicographic min
        if (bh >= 0 && be < FLT_MAX && (res->hypothesis < 0 || be < res->fitness || (be == res->fitness && bh < res->hypothesis))) {
            int c = 0;
            for (int p = 0; p < split; ++p) c += pcnt[(size_t)bt * split + p];
            for (int i = 0; i < 16; ++i) res->pose[i] = poses[(size_t)bt * 16 + i];
            res->fitness = be; res->inliers = c; res->hypothesis = bh; res->converged = 1;
        }
    }
}

__global__ void k_result_init(rtr_pose_result* res) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) res->pose[i] = (i % 5 == 0) ? 1.f : 0.f;
        res->fitness = FLT_MAX; res->inliers = 0; res->hypothesis = -1; res->evaluated = 0; res->converged = 0;
        res->iterations = 0; res->model_id = 0; res->n_keypoints_src = 0; res->n_keypoints_tgt = 0;
        for (int i = 0; i < 5; ++i) res->pad_[i] = 0;
    }
}

int rtr_ransac_dev(rtr_cloud* src, rtr_cloud* tgt, const rtr_ransac_params* p, rtr_pose_result* d_result) {
    RtrRange nvtx_range("rtr.ransac_prerejective");
    rtr_context* ctx = src->ctx;
    k_result_init<<<1, 32, 0, ctx->stream>>>(d_result);
    RTR_LAUNCH_CHECK(ctx, "ransac.init");
    long long h0 = p->hypothesis_begin, h1 = (p->hypothesis_end > 0) ? p->hypothesis_end : p->max_iterations;
    if (src->n < 3 || tgt->n < 1 || h1 <= h0) return 0;
    if (!src->knn || src->knn_k < p->correspondence_k || src->knn_target != tgt || src->knn_target_gen != tgt->feature_gen)
        return rtr_fail("ransac", "rtr_match_features(source, target, k >= correspondence_k) must run first (on THIS target, after its last rtr_fpfh)", RTR_ERR_NOT_READY);
    if (!(p->max_correspondence_distance > 0.f)) return rtr_fail("ransac", "max_correspondence_distance must be > 0", RTR_ERR_INVALID);
    DevGrid* g;
    // any cached grid whose cells are at least d_max wide (and not much wider) serves the 27-cell inlier test; a large
    // sweep repays a grid of exactly d_max (the normals grid, 0.05 against d_max = 0.0365, holds 1.9x the candidates per
    // block: 1e7 hypotheses 8.9 -> 5.9 ms), a single registration's 50 000 hypotheses do not repay the build
    const float hi_mult = (h1 - h0 >= 200000) ? 1.05f : 2.0f;
    if (int e = rtr_get_grid_any(tgt, p->max_correspondence_distance, p->max_correspondence_distance * 1.001f, p->max_correspondence_distance * hi_mult, &g)) return e;
    const long long CHUNK = 1 << 20;
    int cap = (int)std::min<long long>(CHUNK, h1 - h0);
    // slices of the source per surviving hypothesis, fewer when the partial arrays would get large
    // (512-point slices: two source points per thread; 2048 -> 512 took the evaluation of a 50 000-hypothesis registration
    //  from 72 to 54 us on the repo clouds — the chain of dependent cell loads per thread is what it costs;
    //  a 1e7 sweep keeps every SM busy with whole-cloud CTAs and prefers 2048: 5.9 ms against 6.9 ms)
    const int slice = (h1 - h0 >= 200000) ? 2048 : 512;
    int split = std::max(1, std::min(16, (src->n + slice - 1) / slice));
    while (split > 1 && (long long)cap * split > (1LL << 22)) split /= 2;
    int *survivors = nullptr, *count = nullptr, *pcnt = nullptr; float* poses = nullptr; double* psum = nullptr;
    if (int e = tmp_alloc(ctx, &survivors, cap, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &count, 1, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &pcnt, (size_t)cap * split, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &psum, (size_t)cap * split, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &poses, (size_t)cap * 16, "ransac")) return e;
    RansacArgs a;
    a.src = src->pts; a.ns = src->n; a.tgt = tgt->pts; a.nt = tgt->n;
    a.knn = src->knn; a.knn_stride = src->knn_k; a.k = p->correspondence_k; a.seed = p->seed;
    a.simsq = p->similarity_threshold * p->similarity_threshold;
    a.dmax2 = p->max_correspondence_distance * p->max_correspondence_distance;
    a.inlier_fraction = p->inlier_fraction;
    GridView v = rtr_view(g);
    // large sweeps repay the flattened block lists (~20 us to build); a single registration's few hundred survivors do not
    BlockLists bl{nullptr, nullptr};
    if (h1 - h0 >= 200000) if (int e = block_lists_build_dev(ctx, v, &bl)) return e;
    for (long long base = h0; base < h1; base += CHUNK) {
        a.h_base = base; a.h_count = (int)std::min<long long>(CHUNK, h1 - base);
        RTR_CHECK(cudaMemsetAsync(count, 0, sizeof(int), ctx->stream), "ransac");
        RTR_MARK(ctx, "ransac.memset");
        k_ransac_sample<<<nblk(a.h_count, 256), 256, 0, ctx->stream>>>(a, survivors, count);
        RTR_LAUNCH_CHECK(ctx, "ransac.sample");
        k_ransac_pose<<<nblk(a.h_count, 64), 64, 0, ctx->stream>>>(a, survivors, count, poses);
        RTR_LAUNCH_CHECK(ctx, "ransac.pose");
        int grid = (int)std::min<long long>((long long)a.h_count * split, ctx->sm_count * 8);
        k_ransac_eval<<<grid, EVAL_THREADS, 0, ctx->stream>>>(a, v, count, poses, split, psum, pcnt, bl);
        RTR_LAUNCH_CHECK(ctx, "ransac.eval");
        k_ransac_select<<<1, 1024, 0, ctx->stream>>>(a, survivors, count, poses, split, psum, pcnt, d_result);
        RTR_LAUNCH_CHECK(ctx, "ransac.select");
    }
    dev_free(ctx, survivors); dev_free(ctx, count); dev_free(ctx, pcnt); dev_free(ctx, psum); dev_free(ctx, poses);
    return 0;
}

// ============================================================================= RANSAC over a model set
// Every member cloud's hypotheses in the same launches: slot t of [0, nseg * H) is hypothesis h_base + t % H of member
// t / H, evaluated against the one target (the scan).  Draws, prerejection, pose fit, the slice shape of the inlier sums and
// the acceptance rule are the single-cloud kernels', so each member's record equals what rtr_ransac_prerejective returns
// for it alone.
struct RansacMany {
    int nseg;
    int pt_begin[RTR_MAX_SEGMENTS + 1];       // member k = source points [pt_begin[k], pt_begin[k+1]) of src_all / knn_all
    int split[RTR_MAX_SEGMENTS];              // slices of member k's cloud per surviving hypothesis (as rtr_ransac_dev sizes them)
    const float4* src_all;
    const int* knn_all;
    int H;                                    // hypotheses per member in this pass
};
__device__ __forceinline__ RansacArgs ransac_member(const RansacArgs& common, const RansacMany& rm, int k) {
    RansacArgs a = common;
    a.src = rm.src_all + rm.pt_begin[k];
    a.ns = rm.pt_begin[k + 1] - rm.pt_begin[k];
    a.knn = rm.knn_all + (size_t)rm.pt_begin[k] * common.knn_stride;
    return a;
}

__global__ void __launch_bounds__(256) k_ransac_sample_many(const __grid_constant__ RansacMany rm, RansacArgs common, int* __restrict__ survivors,
                                                            int* __restrict__ count) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (t < (long long)rm.nseg * rm.H) {
        const int k = (int)(t / rm.H), hl = (int)(t - (long long)k * rm.H);
        const RansacArgs a = ransac_member(common, rm, k);
        if (a.ns >= 3 && a.nt >= 1) {
            int s[3], c[3];
            ok = draw_hypothesis(a, (unsigned long long)(a.h_base + hl), s, c);
            if (ok) {
                float4 ps[3], pt[3];
#pragma unroll
                for (int e = 0; e < 3; ++e) { ps[e] = __ldg(a.src + s[e]); pt[e] = __ldg(a.tgt + c[e]); }
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    int f = (e + 1) % 3;
                    float ds = dist2f(ps[e].x, ps[e].y, ps[e].z, ps[f].x, ps[f].y, ps[f].z);
                    float dt = dist2f(pt[e].x, pt[e].y, pt[e].z, pt[f].x, pt[f].y, pt[f].z);
                    float sim = ds < dt ? __fdiv_rn(ds, dt) : __fdiv_rn(dt, ds);
                    if (!(sim >= a.simsq)) ok = false;
                }
            }
        }
    }
    unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (mask) {
        int lane = threadIdx.x & 31;
        int leader = __ffs(mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(count, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (ok) survivors[base + __popc(mask & ((1u << lane) - 1u))] = (int)t;
    }
}

// pose of every survivor + its evaluation work items: split[k] consecutive entries of `items` ((survivor << 5) | slice)
__global__ void __launch_bounds__(64) k_ransac_pose_many(const __grid_constant__ RansacMany rm, RansacArgs common, const int* __restrict__ survivors,
                                                         const int* __restrict__ count, float* __restrict__ poses, int* __restrict__ item_first,
                                                         int* __restrict__ items, int* __restrict__ n_items) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *count) return;
    const int slot = survivors[t];
    const int k = slot / rm.H, hl = slot - k * rm.H;
    const RansacArgs a = ransac_member(common, rm, k);
    int s[3], c[3];
    draw_hypothesis(a, (unsigned long long)(a.h_base + hl), s, c);
    double ss[3] = {0, 0, 0}, st[3] = {0, 0, 0}, m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        float4 p = __ldg(a.src + s[e]), q = __ldg(a.tgt + c[e]);
        double sv[3] = {p.x, p.y, p.z}, tv[3] = {q.x, q.y, q.z};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            ss[i] += sv[i]; st[i] += tv[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) m[i * 3 + j] += sv[i] * tv[j];
        }
    }
    float pose[16];
    horn_pose(ss, st, m, 3.0, pose);
#pragma unroll
    for (int i = 0; i < 16; ++i) poses[(size_t)t * 16 + i] = pose[i];
    const int split = rm.split[k];
    const int first = atomicAdd(n_items, split);
    item_first[t] = first;
    for (int j = 0; j < split; ++j) items[first + j] = (t << 5) | j;
}

__global__ void __launch_bounds__(EVAL_THREADS) k_ransac_eval_many(const __grid_constant__ RansacMany rm, RansacArgs common, GridView g,
                                                                   const int* __restrict__ survivors, const float* __restrict__ poses,
                                                                   const int* __restrict__ items, const int* __restrict__ n_items,
                                                                   double* __restrict__ psum, int* __restrict__ pcnt, BlockLists bl) {
    __shared__ float m[16];
    __shared__ double wsum[EVAL_THREADS / 32];
    __shared__ int wcnt[EVAL_THREADS / 32];
    const int n = *n_items;
    for (int w = blockIdx.x; w < n; w += gridDim.x) {
        const int item = items[w];
        const int t = item >> 5, part = item & 31;
        const int k = survivors[t] / rm.H;
        const float4* src = rm.src_all + rm.pt_begin[k];
        const int ns = rm.pt_begin[k + 1] - rm.pt_begin[k];
        const int split = rm.split[k];
        const int chunk = (ns + split - 1) / split;
        const int i0 = part * chunk, i1 = min(ns, i0 + chunk);
        __syncthreads();
        if (threadIdx.x < 16) m[threadIdx.x] = poses[(size_t)t * 16 + threadIdx.x];
        __syncthreads();
        int cnt = 0;
        double sum = 0;
        for (int i = i0 + threadIdx.x; i < i1; i += EVAL_THREADS) {
            float4 q = xform(m, __ldg(src + i));
            const float best = inlier_min_d2(g, bl, q.x, q.y, q.z);
            if (best < common.dmax2) { ++cnt; sum += (double)best; }
        }
        cnt = warp_sum(cnt);
        sum = warp_sum(sum);
        if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = sum; wcnt[threadIdx.x >> 5] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double S = 0; int C = 0;
#pragma unroll
            for (int ww = 0; ww < EVAL_THREADS / 32; ++ww) { S += wsum[ww]; C += wcnt[ww]; }
            psum[w] = S; pcnt[w] = C;
        }
    }
}

// one CTA per member: its record (initialised here), the arg-min over its survivors
__global__ void __launch_bounds__(1024) k_ransac_select_many(const __grid_constant__ RansacMany rm, RansacArgs common, const int* __restrict__ survivors,
                                                             const int* __restrict__ count, const float* __restrict__ poses,
                                                             const int* __restrict__ item_first, const double* __restrict__ psum,
                                                             const int* __restrict__ pcnt, rtr_pose_result* __restrict__ res_all) {
    __shared__ float s_err[32];
    __shared__ long long s_h[32];
    __shared__ int s_t[32];
    __shared__ int s_n[32];
    const int k = blockIdx.x;
    const int ns = rm.pt_begin[k + 1] - rm.pt_begin[k];
    const int split = rm.split[k];
    rtr_pose_result* res = res_all + k;
    const int n = *count;
    float be = FLT_MAX; long long bh = -1; int bt = -1; int mine = 0;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int slot = survivors[t];
        if (slot / rm.H != k) continue;
        ++mine;
        const int first = item_first[t];
        double S = 0; int c = 0;
        for (int p = 0; p < split; ++p) { S += psum[first + p]; c += pcnt[first + p]; }
        float frac = __fdiv_rn((float)c, (float)ns);
        if (frac >= common.inlier_fraction) {
            float e = c > 0 ? (float)(S / (double)c) : FLT_MAX;
            long long h = common.h_base + (slot - k * rm.H);
            if (e < be || (e == be && (bh < 0 || h < bh))) { be = e; bh = h; bt = t; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float oe = __shfl_xor_sync(0xffffffffu, be, o);
        long long oh = __shfl_xor_sync(0xffffffffu, bh, o);
        int ot = __shfl_xor_sync(0xffffffffu, bt, o);
        mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (oh >= 0 && (bh < 0 || oe < be || (oe == be && oh < bh))) { be = oe; bh = oh; bt = ot; }
    }
    if ((threadIdx.x & 31) == 0) { s_err[threadIdx.x >> 5] = be; s_h[threadIdx.x >> 5] = bh; s_t[threadIdx.x >> 5] = bt; s_n[threadIdx.x >> 5] = mine; }
    __syncthreads();
    if (threadIdx.x == 0) {
        be = FLT_MAX; bh = -1; bt = -1; mine = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            float oe = s_err[w]; long long oh = s_h[w]; int ot = s_t[w];
            mine += s_n[w];
            if (oh >= 0 && (bh < 0 || oe < be || (oe == be && oh < bh))) { be = oe; bh = oh; bt = ot; }
        }
        for (int i = 0; i < 16; ++i) res->pose[i] = (i % 5 == 0) ? 1.f : 0.f;
        res->fitness = FLT_MAX; res->inliers = 0; res->hypothesis = -1; res->evaluated = mine; res->converged = 0;
        res->iterations = 0; res->model_id = k; res->n_keypoints_src = 0; res->n_keypoints_tgt = 0;
        for (int i = 0; i < 5; ++i) res->pad_[i] = 0;
        if (bh >= 0 && be < FLT_MAX) {
            const int first = item_first[bt];
            int c = 0;
            for (int p = 0; p < split; ++p) c += pcnt[first + p];
            for (int i = 0; i < 16; ++i) res->pose[i] = poses[(size_t)bt * 16 + i];
            res->fitness = be; res->inliers = c; res->hypothesis = bh; res->converged = 1;
        }
    }
}

__global__ void k_result_init_many(rtr_pose_result* res, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    rtr_pose_result* r = res + k;
    for (int i = 0; i < 16; ++i) r->pose[i] = (i % 5 == 0) ? 1.f : 0.f;
    r->fitness = FLT_MAX; r->inliers = 0; r->hypothesis = -1; r->evaluated = 0; r->converged = 0;
    r->iterations = 0; r->model_id = k; r->n_keypoints_src = 0; r->n_keypoints_tgt = 0;
    for (int i = 0; i < 5; ++i) r->pad_[i] = 0;
}

// set = model set whose members [0, n_models) are the sources; the target is member `tgt_seg` (the scan).  knn_all: the
// members' correspondences (target-local indices), n_source_points x k.  d_results: n_models records.
static int ransac_many_dev(rtr_cloud* set, int n_models, int tgt_seg, const int* knn_all, int knn_k, const rtr_ransac_params* p,
                           rtr_pose_result* d_results, const GridView* scan_grid, const BlockLists* scan_blocks) {
    RtrRange nvtx_range("rtr.ransac_prerejective.many");
    rtr_context* ctx = set->ctx;
    const long long h0 = p->hypothesis_begin, h1 = (p->hypothesis_end > 0) ? p->hypothesis_end : p->max_iterations;
    const int nt = set->seg_begin[tgt_seg + 1] - set->seg_begin[tgt_seg];
    const long long H = h1 - h0;
    if (H <= 0 || nt < 1) {
        k_result_init_many<<<1, 32, 0, ctx->stream>>>(d_results, n_models);
        RTR_LAUNCH_CHECK(ctx, "ransac.init");
        return 0;
    }
    if (!(p->max_correspondence_distance > 0.f)) return rtr_fail("ransac", "max_correspondence_distance must be > 0", RTR_ERR_INVALID);
    if (H > (1 << 20) || (long long)n_models * H > (1LL << 24)) return rtr_fail("ransac", "model-set RANSAC takes at most 2^20 hypotheses per model", RTR_ERR_INVALID);
    // target grid for the inlier test and its block lists: built by the caller (ScanStructures)
    const GridView v = *scan_grid;
    const BlockLists bl = *scan_blocks;
    RansacMany rm;
    memset(&rm, 0, sizeof(rm));
    rm.nseg = n_models;
    for (int k = 0; k <= RTR_MAX_SEGMENTS; ++k) rm.pt_begin[k] = set->seg_begin[std::min(k, n_models)];
    long long max_items = 0;
    for (int k = 0; k < n_models; ++k) {
        const int ns = set->seg_begin[k + 1] - set->seg_begin[k];
        rm.split[k] = std::max(1, std::min(16, (ns + 511) / 512));        // 512-point slices, as rtr_ransac_dev
        max_items += (long long)H * rm.split[k];
    }
    rm.src_all = set->pts; rm.knn_all = knn_all; rm.H = (int)H;
    RansacArgs a;
    a.src = nullptr; a.ns = 0; a.tgt = set->pts + set->seg_begin[tgt_seg]; a.nt = nt;
    a.knn = nullptr; a.knn_stride = knn_k; a.k = p->correspondence_k; a.seed = p->seed;
    a.simsq = p->similarity_threshold * p->similarity_threshold;
    a.dmax2 = p->max_correspondence_distance * p->max_correspondence_distance;
    a.inlier_fraction = p->inlier_fraction;
    a.h_base = h0; a.h_count = (int)H;
    const long long slots = (long long)n_models * H;
    int *survivors = nullptr, *counters = nullptr, *item_first = nullptr, *items = nullptr, *pcnt = nullptr; float* poses = nullptr; double* psum = nullptr;
    if (int e = tmp_alloc(ctx, &survivors, (size_t)slots, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &counters, 2, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &item_first, (size_t)slots, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &items, (size_t)max_items, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &pcnt, (size_t)max_items, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &psum, (size_t)max_items, "ransac")) return e;
    if (int e = tmp_alloc(ctx, &poses, (size_t)slots * 16, "ransac")) return e;
    RTR_CHECK(cudaMemsetAsync(counters, 0, 2 * sizeof(int), ctx->stream), "ransac");
    RTR_MARK(ctx, "ransac.memset");
    k_ransac_sample_many<<<nblk(slots, 256), 256, 0, ctx->stream>>>(rm, a, survivors, counters);
    RTR_LAUNCH_CHECK(ctx, "ransac.sample");
    k_ransac_pose_many<<<nblk(slots, 64), 64, 0, ctx->stream>>>(rm, a, survivors, counters, poses, item_first, items, counters + 1);
    RTR_LAUNCH_CHECK(ctx, "ransac.pose");
    k_ransac_eval_many<<<ctx->sm_count * 8, EVAL_THREADS, 0, ctx->stream>>>(rm, a, v, survivors, poses, items, counters + 1, psum, pcnt, bl);
    RTR_LAUNCH_CHECK(ctx, "ransac.eval");
    k_ransac_select_many<<<n_models, 1024, 0, ctx->stream>>>(rm, a, survivors, counters, poses, item_first, psum, pcnt, d_results);
    RTR_LAUNCH_CHECK(ctx, "ransac.select");
    return 0;
}

// ============================================================================= ICP
struct IcpState {
    float final_[16];
    float step[16];
    double prev_mse;
    int iterations;
    int done;
    int state;
    int corr;
    int have_step;
    int skipped;      // registration only: RANSAC accepted nothing, so there is no pose to refine
};

#define ICP_THREADS 256
// Running sums of one ICP iteration.  EST 0 (TransformationEstimationSVD): 3 (sum s) + 3 (sum t) + 9 (sum s t^T);
// EST 1 (TransformationEstimationPointToPlaneLLS): the 21 upper-triangle entries of the 6x6 A^T A (row-major) + 6 of
// A^T b, rows [s x n ; n] per correspondence (the linearised (R s + t - d) . n).  Both end with (sum d2, count).
template <int EST> struct IcpSums { static constexpr int N = EST ? 29 : 17; };
#define ICP_NSUM_MAX 29

template <int EST>
__device__ __forceinline__ void icp_accumulate(double* acc, float4 q, float4 t, float4 nrm, float d2) {
    constexpr int N = IcpSums<EST>::N;
    const double sx = q.x, sy = q.y, sz = q.z, tx = t.x, ty = t.y, tz = t.z;
    if (EST == 0) {
        acc[0] += sx; acc[1] += sy; acc[2] += sz; acc[3] += tx; acc[4] += ty; acc[5] += tz;
        acc[6] += sx * tx; acc[7] += sx * ty; acc[8] += sx * tz;
        acc[9] += sy * tx; acc[10] += sy * ty; acc[11] += sy * tz;
        acc[12] += sz * tx; acc[13] += sz * ty; acc[14] += sz * tz;
    } else {
        const double nx = nrm.x, ny = nrm.y, nz = nrm.z;
        const double v[6] = {nz * sy - ny * sz, nx * sz - nz * sx, ny * sx - nx * sy, nx, ny, nz};
        const double dd = ((nx * tx + ny * ty) + nz * tz) - ((nx * sx + ny * sy) + nz * sz);
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = r; c < 6; ++c) acc[k++] += v[r] * v[c];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[21 + r] += v[r] * dd;
    }
    acc[N - 2] += (double)d2; acc[N - 1] += 1.0;
}

// x = A^-1 b: Gaussian elimination with partial pivoting (first largest |pivot|), fp64, the oracle's operation order
// (stands in for Eigen's ATA.inverse() * ATb); then pcl's constructTransformationMatrix(alpha, beta, gamma, tx, ty, tz)
__device__ bool lls_pose(const double* sums, float* pose) {
    double A[6][6], b[6], x[6];
    int k = 0;
    for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { A[r][c] = sums[k]; A[c][r] = sums[k]; ++k; }
    for (int r = 0; r < 6; ++r) b[r] = sums[21 + r];
    for (int c = 0; c < 6; ++c) {
        int piv = c; double best = fabs(A[c][c]);
        for (int r = c + 1; r < 6; ++r) { double v = fabs(A[r][c]); if (v > best) { best = v; piv = r; } }
        if (!(best > 0.0)) return false;
        if (piv != c) {
            for (int kk = 0; kk < 6; ++kk) { double t = A[c][kk]; A[c][kk] = A[piv][kk]; A[piv][kk] = t; }
            double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        for (int r = c + 1; r < 6; ++r) {
            double f = A[r][c] / A[c][c];
            for (int kk = c; kk < 6; ++kk) A[r][kk] = A[r][kk] - f * A[c][kk];
            b[r] = b[r] - f * b[c];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double sacc = b[r];
        for (int kk = r + 1; kk < 6; ++kk) sacc = sacc - A[r][kk] * x[kk];
        x[r] = sacc / A[r][r];
    }
    const double sa = sin(x[0]), ca = cos(x[0]), sb = sin(x[1]), cb = cos(x[1]), sg = sin(x[2]), cg = cos(x[2]);
    for (int i = 0; i < 16; ++i) pose[i] = (i % 5 == 0) ? 1.f : 0.f;
    pose[0] = (float)(cg * cb);  pose[4] = (float)(-sg * ca + cg * sb * sa);  pose[8]  = (float)(sg * sa + cg * sb * ca);   pose[12] = (float)x[3];
    pose[1] = (float)(sg * cb);  pose[5] = (float)(cg * ca + sg * sb * sa);   pose[9]  = (float)(-cg * sa + sg * sb * ca);  pose[13] = (float)x[4];
    pose[2] = (float)(-sb);      pose[6] = (float)(cb * sa);                  pose[10] = (float)(cb * ca);                  pose[14] = (float)x[5];
    return true;
}

__global__ void k_icp_init(const float4* __restrict__ src, int n, const float* __restrict__ init_pose, const rtr_pose_result* __restrict__ init_res,
                           float4* __restrict__ cur, IcpState* __restrict__ st, unsigned* __restrict__ ticket) {
    __shared__ float m[16];
    if (threadIdx.x < 16) {
        float v = (threadIdx.x % 5 == 0) ? 1.f : 0.f;
        if (init_res) v = init_res->pose[threadIdx.x];
        else if (init_pose) v = init_pose[threadIdx.x];
        m[threadIdx.x] = v;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < 16) {
        st->final_[threadIdx.x] = m[threadIdx.x];
        st->step[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.f : 0.f;
        if (threadIdx.x == 0) {
            int skip = (init_res && init_res->converged == 0) ? 1 : 0;
            st->prev_mse = DBL_MAX; st->iterations = 0; st->done = skip; st->state = 0; st->corr = 0; st->have_step = 0; st->skipped = skip;
            *ticket = 0u;
        }
    }
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cur[i] = xform(m, __ldg(src + i));
}

// Large sources are processed in the TARGET grid's cell order: neighbouring threads then walk the same cell ranges
// (coalesced, L1-resident) instead of 32 unrelated ones.  The sums are order independent up to fp64 rounding.
__global__ void k_icp_cell_keys(GridView g, const float4* __restrict__ cur, int n, int* __restrict__ keys, int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = cur[i];
    int cx = clampi(cell_coord(p.x, g.mnx, g.inv_h), 0, g.dx - 1);
    int cy = clampi(cell_coord(p.y, g.mny, g.inv_h), 0, g.dy - 1);
    int cz = clampi(cell_coord(p.z, g.mnz, g.inv_h), 0, g.dz - 1);
    keys[i] = cell_key(g, cx, cy, cz);
    vals[i] = i;
}
__global__ void k_icp_permute(const float4* __restrict__ cur, const float4* __restrict__ src, const int* __restrict__ perm, int n,
                              float4* __restrict__ cur_out, float4* __restrict__ src_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j = perm[i];
    cur_out[i] = cur[j];
    src_out[i] = __ldg(src + j);
}

// TransformationEstimationSVD + final = step * final + DefaultConvergenceCriteria, run by the LAST CTA of the
// correspondence kernel to finish (ticket counter), so an iteration is one launch.  The reduction over the per-CTA
// partials has a fixed shape (lane-strided, then a shuffle tree) whatever CTA happens to be last.
struct IcpSolveArgs { int max_iterations, force; double mse_abs; };

template <int EST>
__device__ void icp_solve_thread0(const double* sums, IcpState* st, const IcpSolveArgs& sa) {
    constexpr int N = IcpSums<EST>::N;
    double cnt = sums[N - 1];
    st->corr = (int)cnt;
    if (cnt < 3.0) { st->done = 1; st->state = 0; return; }
    float step[16], fin[16];
    if (EST == 0) horn_pose(&sums[0], &sums[3], &sums[6], cnt, step);
    else if (!lls_pose(sums, step)) { st->done = 1; st->state = 0; return; }
    for (int i = 0; i < 16; ++i) fin[i] = st->final_[i];
    matmul4(step, fin, fin);
    for (int i = 0; i < 16; ++i) { st->final_[i] = fin[i]; st->step[i] = step[i]; }
    st->have_step = 1;
    int it = st->iterations + 1;
    st->iterations = it;
    if (it >= sa.max_iterations) { st->done = 1; st->state = 1; return; }
    if (!sa.force) {
        double cos_angle = 0.5 * ((double)step[0] + (double)step[5] + (double)step[10] - 1.0);
        double tsq = (double)step[12] * (double)step[12] + (double)step[13] * (double)step[13] + (double)step[14] * (double)step[14];
        if (cos_angle >= 1.0 && tsq <= 0.0) { st->done = 1; st->state = 2; return; }
        double mse = sums[N - 2] / cnt;
        if (fabs(mse - st->prev_mse) < sa.mse_abs) { st->done = 1; st->state = 3; return; }
        st->prev_mse = mse;
    }
}

// called by every thread of a CTA (256 threads) after its partials are written
template <int EST>
__device__ void icp_last_cta_solve(const double* partials, IcpState* st, unsigned* ticket, const IcpSolveArgs& sa, double* sums /* smem[N] */,
                                   int* is_last /* smem */, const int nparts /* CTAs that share this ticket */,
                                   unsigned* publish = nullptr, unsigned publish_value = 0u /* persistent kernel: iteration counter */) {
    constexpr int N = IcpSums<EST>::N;
    constexpr int G = 255 / N;          // row groups: 15 for 17 sums, 8 for 29
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(ticket, 1u);
        *is_last = (t == (unsigned)nparts - 1u);
    }
    __syncthreads();
    pdl_launch_dependents();    // this CTA's part is done: the next iteration's CTAs may take their places and wait while the last CTA solves
    if (!*is_last) return;
    __threadfence();
    // partials is an [nparts][N] matrix: thread t < G * N owns column t % N and the rows t / N, t / N + G, ... (adjacent
    // threads read adjacent addresses), then N threads fold the G row groups in order
    __shared__ double grp[G][N];
    const int t = threadIdx.x;
    if (t < G * N) {
        const int k = t % N, rg = t / N;
        const double* col = partials + k;
        double v = 0;
        // eight loads in flight per thread, missing rows read as +0.0 (x + 0.0 == x): no tail loop of single loads, each of
        // which would be a full round trip to L2 behind the previous add
        int b = rg;
        for (; b + G * 7 < nparts; b += G * 8) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = __ldcg(col + (size_t)(b + G * u) * N);
#pragma unroll
            for (int u = 0; u < 8; ++u) v += x[u];
        }
        {   // the last, partial batch: rows past the end are re-reads of the last valid row, masked to +0.0
            double x[8];
            const int last = nparts - 1;
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = __ldcg(col + (size_t)min(b + G * u, last) * N);
#pragma unroll
            for (int u = 0; u < 8; ++u) v += (b + G * u < nparts) ? x[u] : 0.0;
        }
        grp[rg][k] = v;
    }
    __syncthreads();
    if (t < N) {
        double v = 0;
#pragma unroll
        for (int rg = 0; rg < G; ++rg) v += grp[rg][t];
        sums[t] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0u;
        icp_solve_thread0<EST>(sums, st, sa);
        if (publish) { __threadfence(); atomicExch(publish, publish_value); }      // state, step and ticket are visible before the counter moves
    }
}

// Lower bound of dist2f(q, p) over every p inside the box, built from the same float operations (subtraction, product
// and sum are monotone under round-to-nearest): a query whose bound exceeds the correspondence cap has no correspondence.
struct Box6 { float mnx, mny, mnz, mxx, mxy, mxz; };
__device__ __forceinline__ float box_dist2f(const Box6& b, float qx, float qy, float qz) {
    float ex = qx > b.mxx ? __fsub_rn(qx, b.mxx) : (qx < b.mnx ? __fsub_rn(qx, b.mnx) : 0.f);
    float ey = qy > b.mxy ? __fsub_rn(qy, b.mxy) : (qy < b.mny ? __fsub_rn(qy, b.mny) : 0.f);
    float ez = qz > b.mxz ? __fsub_rn(qz, b.mxz) : (qz < b.mnz ? __fsub_rn(qz, b.mnz) : 0.f);
    return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}
// Thread-per-query kernels, far queries: the lanes that still want a neighbour are served one after the other by the whole
// warp through the 32-ary hierarchy (bvh.cuh).  key0: this lane's start — (d2 bits << 32 | index) of a known target point
// (warm start), or (bits of the search radius^2 << 32 | 0xffffffff) for "anything at or inside the radius".
// Every lane of the warp must call this (want == false: nothing to search).
__device__ __forceinline__ void wide_nearest_lanes(const WideBvh& W, bool want, float4 q, unsigned long long key0, int lane, int& b, float& d2) {
    unsigned todo = __ballot_sync(0xffffffffu, want);
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const float x = __shfl_sync(0xffffffffu, q.x, j), y = __shfl_sync(0xffffffffu, q.y, j), z = __shfl_sync(0xffffffffu, q.z, j);
        const unsigned long long k0 = __shfl_sync(0xffffffffu, key0, j);
        const unsigned long long key = wide_nearest_warp(W, x, y, z, lane, k0);
        if (lane == j && (unsigned)key != 0xffffffffu) { b = (int)(unsigned)key; d2 = __uint_as_float((unsigned)(key >> 32)); }
    }
}
// The neighbour of the previous search as the first candidate of this one: its distance bounds the search sphere, so the
// grid search skips the 27-cell block and walks only the rows the sphere touches (grid_nearest_far: typically one row and one
// or two cells once ICP is converging, against 9 rows / 27 cells), and the hierarchy opens only the boxes inside it.  The
// candidate is a target point like any other: the result is the exact nearest neighbour, ties -> lowest index, as without it.
struct WarmStart { int b; float d2; float4 p; };
__device__ __forceinline__ WarmStart icp_warm_start(const int* __restrict__ nn_prev, int i, bool use, const float4* __restrict__ tgt_pts, float4 q) {
    WarmStart w;
    w.b = -1; w.d2 = FLT_MAX; w.p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (use) {
        const int pb = nn_prev[i];
        if (pb >= 0) {
            const float4 tp = __ldg(tgt_pts + pb);
            const float d = dist2f(q.x, q.y, q.z, tp.x, tp.y, tp.z);
            if (d < FLT_MAX) { w.b = pb; w.d2 = d; w.p = tp; }       // not for a NaN / overflowing distance
        }
    }
    return w;
}
// correspondence estimation + accumulation.  Applies the previous iteration's step first (transformCloud in place).
// A query farther from the target's bounding box than the cap is dropped before any cell is touched.
template <int EST>
__global__ void __launch_bounds__(ICP_THREADS, 4) k_icp_corr(GridView g, GridView gc, Box6 bb, float far2, const float4* __restrict__ tgt_normals, float4* __restrict__ cur, int n,
                                                           IcpState* st, double dmax2, float prune2, double* partials, unsigned* ticket,
                                                           IcpSolveArgs sa, const __grid_constant__ WideBvh W, const float4* __restrict__ tgt_pts,
                                                           int* __restrict__ nn_prev) {
    constexpr int N = IcpSums<EST>::N;
    __shared__ float m[16];
    __shared__ double red[ICP_THREADS / 32][N];
    __shared__ double sums[N];
    __shared__ int is_last;
    pdl_wait();                 // the previous iteration (or k_icp_init) has completed: st, cur, ticket are current
    if (st->done) return;
    int have = st->have_step;
    if (threadIdx.x < 16) m[threadIdx.x] = st->step[threadIdx.x];
    __syncthreads();
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = 0.0;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    {
        const bool live = i < n;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            q = cur[i];
            if (have) { q = xform(m, q); cur[i] = q; }
        }
        int b = -1; float d2 = 0.f; float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        // the grid copy carries the matched point's coordinates: the original-order target is never touched
        const float bd2 = box_dist2f(bb, q.x, q.y, q.z);
        const bool search = live && !(bd2 > prune2);
        const bool far = bd2 > far2;                        // far from the target: coarse cells, or the 32-ary hierarchy (see k_icp_fitness)
        const WarmStart ws = icp_warm_start(nn_prev, i, search && have && nn_prev != nullptr, tgt_pts, q);
        if (search && !(far && W.n > 0)) {
            if (ws.b >= 0) { b = ws.b; d2 = ws.d2; t = ws.p; grid_nearest_far(far ? gc : g, q.x, q.y, q.z, prune2, b, d2, t); }
            else grid_nearest_ex(far ? gc : g, q.x, q.y, q.z, prune2, b, d2, t);
        }
        if (W.n > 0) {                                      // warp-uniform
            const bool mine = search && far;
            const unsigned long long key0 = (ws.b >= 0 && ws.d2 <= prune2) ? (((unsigned long long)__float_as_uint(ws.d2) << 32) | (unsigned)ws.b)
                                                                          : (((unsigned long long)__float_as_uint(prune2) << 32) | 0xffffffffull);
            wide_nearest_lanes(W, mine, q, key0, threadIdx.x & 31, b, d2);
            if (mine && b >= 0) t = __ldg(tgt_pts + b);
        }
        if (search && nn_prev) nn_prev[i] = b;              // (a query outside the cap's reach keeps its older entry: any target point is a valid start)
        if (b >= 0 && (double)d2 <= dmax2) {
            float4 nrm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EST == 1) nrm = __ldg(tgt_normals + b);
            if (EST == 0 || finite3(nrm)) icp_accumulate<EST>(acc, q, t, nrm, d2);
        }
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < ICP_THREADS / 32; ++w) v += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * N + threadIdx.x] = v;
    }
    icp_last_cta_solve<EST>(partials, st, ticket, sa, sums, &is_last, (int)gridDim.x);
}

// Small sources (repo clouds): one WARP per source point — the lanes share the candidate scan, so the per-iteration
// latency is set by ~9 range steps instead of ~50 dependent loads.  Each warp walks a strided list of queries and keeps
// the 17 sums in lane 0; warps are then folded through shared memory.
#define ICPW_WARPS 8
// bid / nblocks: this CTA's place among the CTAs that work on this source cloud (== blockIdx.x / gridDim.x for one cloud; a
// model set gives every member cloud its own range of CTAs, state, ticket and partials — see k_icp_corr_warp_many)
// Target search of the warp-per-query kernels.
//   Scans up to RTR_BRUTE_NN_MAX points (every scan the reference ships) are searched WITHOUT a structure: the lanes stride
//   over all target points, two queries per pass.  That pass is instruction-bound (~30 instructions per 64 point pairs: the
//   step's ICP executes 36.6 M warp instructions per iteration for 26 k queries x 1909 targets).  Measured and dropped in
//   round 2, all exact, none faster, because most of a model's points have NO scan point nearby (chair4: 26 % inliers):
//   (a) the uniform grid with the previous neighbour's distance as prune radius — the ring walks of the far queries cost as
//   much as the scan (43.7 -> 32.8 M instructions over the ten iterations against 36.1 M); (b) "neighbour provably unchanged"
//   from the second smallest distance and the distance moved (dist(q', p) >= dist(q, p) - delta) — for a far query the second
//   nearest scan point is within a fraction of a millimetre of the nearest, the bound never holds (711 us against 640).
//   Larger scans use the exact grid search (block, then rings of rows), warm-started with the previous neighbour's distance
//   as its prune radius (the nearest distance cannot exceed it).
struct IcpTarget {
    GridView g;                       // grid over the target (positions as stored in g.cell_begin)
    const float4* brute;              // the target's points in any order with the original index in .w (brute_n of them; 0: none)
    int brute_n;
    const float4* pts;                // target points by original index (the index .w / the search returns)
    const float4* normals;            // target normals by original index (estimator 1)
    WbvhView bvh;                     // small targets: two-level 32-wide hierarchy for the warp-per-query kernels (bvh.n == 0: plain scan)
};
template <int EST>
__device__ __forceinline__ void icp_corr_warp_body(const IcpTarget& T, float4* __restrict__ cur, int* __restrict__ nn_prev, int n, IcpState* st,
                                                   double dmax2, float prune2, double* partials, unsigned* ticket, const IcpSolveArgs& sa,
                                                   const int bid, const int nblocks, unsigned* publish = nullptr, unsigned publish_value = 0u) {
    constexpr int N = IcpSums<EST>::N;
    __shared__ float m[16];
    __shared__ double red[ICPW_WARPS][N];      // lane 0 of every warp accumulates its queries here, in query order
    __shared__ double sums[N];
    __shared__ int is_last;
    if (st->done) return;
    int have = st->have_step;
    if (threadIdx.x < 16) m[threadIdx.x] = st->step[threadIdx.x];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < N) red[warp][lane] = 0.0;
    __syncthreads();
    int nwarps = nblocks * ICPW_WARPS;
    double* acc = red[warp];
    const float4* __restrict__ tgt_pts = T.pts;
    auto add = [&](float4 q, int b, float d2, float4 t) {
        if (lane == 0 && b >= 0 && (double)d2 <= dmax2) {
            float4 nrm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (EST == 1) nrm = __ldg(T.normals + b);
            if (EST == 0 || finite3(nrm)) icp_accumulate<EST>(acc, q, t, nrm, d2);
        }
    };
    const bool warm = have && nn_prev != nullptr;
    if (T.bvh.n > 0) {
        // small target: two-level 32-wide hierarchy (bvh.cuh), started from the query's previous neighbour.  Same warp -> query
        // assignment and accumulation order as the plain scan below (i, i + nwarps, i + 2 nwarps, ...): same sums, same bits.
        for (int i = bid * ICPW_WARPS + warp; i < n; i += nwarps) {
            float4 q = cur[i];
            if (have) { q = xform(m, q); if (lane == 0) cur[i] = q; }
            unsigned long long key = ~0ull;
            if (warm) {
                const int pb = nn_prev[i];
                if (pb >= 0) {
                    const float4 tp = __ldg(tgt_pts + pb);
                    key = ((unsigned long long)__float_as_uint(dist2f(q.x, q.y, q.z, tp.x, tp.y, tp.z)) << 32) | (unsigned)pb;
                }
            }
            key = wbvh_nearest_warp(T.bvh, q.x, q.y, q.z, lane, key);
            const unsigned h = (unsigned)(key >> 32);
            const int b = h <= 0x7f800000u ? (int)(unsigned)key : -1;
            if (lane == 0) {
                if (nn_prev) nn_prev[i] = b;
                if (b >= 0) add(q, b, __uint_as_float(h), __ldg(tgt_pts + b));
            }
        }
    } else if (T.brute_n > 0) {
        GridView gb = T.g;
        gb.sorted = T.brute; gb.n = T.brute_n;
        // two of this warp's queries per pass (i, i + nwarps): same warp -> query assignment and accumulation order as below
        for (int i = bid * ICPW_WARPS + warp; i < n; i += 2 * nwarps) {
            int i2 = i + nwarps;
            bool two = i2 < n;
            float4 q = cur[i], q2 = two ? cur[i2] : q;
            if (have) {
                q = xform(m, q); if (lane == 0) cur[i] = q;
                if (two) { q2 = xform(m, q2); if (lane == 0) cur[i2] = q2; } else q2 = q;
            }
            int b, b2; float d2, d22;
            brute_nearest_warp2(gb, q.x, q.y, q.z, q2.x, q2.y, q2.z, lane, b, d2, b2, d22);
            if (lane == 0) {
                if (b >= 0) add(q, b, d2, __ldg(tgt_pts + b));
                if (two && b2 >= 0) add(q2, b2, d22, __ldg(tgt_pts + b2));
            }
        }
    } else {
        for (int i = bid * ICPW_WARPS + warp; i < n; i += nwarps) {
            float4 q = cur[i];
            if (have) { q = xform(m, q); if (lane == 0) cur[i] = q; }
            float pr2 = prune2;
            if (warm) {
                const int pb = nn_prev[i];
                if (pb >= 0) {
                    const float4 tp = __ldg(tgt_pts + pb);
                    // the previous neighbour is a candidate: the nearest distance cannot exceed this (a hair of slack keeps it inside)
                    pr2 = fminf(prune2, __fmul_rn(dist2f(q.x, q.y, q.z, tp.x, tp.y, tp.z), 1.000001f));
                }
            }
            int b; float d2; float4 t;
            grid_nearest_warp(T.g, q.x, q.y, q.z, pr2, lane, b, d2, t);
            if (lane == 0 && nn_prev) nn_prev[i] = b;
            add(q, b, d2, t);
        }
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < ICPW_WARPS; ++w) v += red[w][threadIdx.x];
        partials[(size_t)bid * N + threadIdx.x] = v;
    }
    icp_last_cta_solve<EST>(partials, st, ticket, sa, sums, &is_last, nblocks, publish, publish_value);
}

template <int EST>
__global__ void __launch_bounds__(ICPW_WARPS * 32, 4) k_icp_corr_warp(IcpTarget T, float4* __restrict__ cur, int* __restrict__ nn_prev, int n, IcpState* st,
                                                                   double dmax2, float prune2, double* partials, unsigned* ticket, IcpSolveArgs sa) {
    pdl_wait();                 // the previous iteration (or k_icp_init) has completed: st, cur, ticket are current
    icp_corr_warp_body<EST>(T, cur, nn_prev, n, st, dmax2, prune2, partials, ticket, sa, (int)blockIdx.x, (int)gridDim.x);
}

// Model set: the CTAs of every member cloud in one launch.  cta_begin[k] .. cta_begin[k+1] work on member k with the CTA
// count the single-cloud launch would use, so the summation shape — and with it every bit of the pose — is the same.
struct IcpMany {
    int nseg;
    int pt_begin[RTR_MAX_SEGMENTS + 1];
    int cta_begin[RTR_MAX_SEGMENTS + 1];      // unused tail = total CTAs
};
__device__ __forceinline__ int icp_many_segment(const IcpMany& im, int b) {
    int k = 0;
#pragma unroll
    for (int step = RTR_MAX_SEGMENTS / 2; step > 0; step >>= 1) if (b >= im.cta_begin[k + step]) k += step;
    return k;
}
template <int EST>
__global__ void __launch_bounds__(ICPW_WARPS * 32, 4) k_icp_corr_warp_many(const __grid_constant__ IcpMany im, IcpTarget T, float4* __restrict__ cur_all,
                                                                        int* __restrict__ nn_prev_all, IcpState* st_all, double dmax2, float prune2,
                                                                        double* partials_all, unsigned* ticket_all, IcpSolveArgs sa) {
    pdl_wait();
    const int k = icp_many_segment(im, (int)blockIdx.x);
    const int b0 = im.cta_begin[k];
    icp_corr_warp_body<EST>(T, cur_all + im.pt_begin[k], nn_prev_all + im.pt_begin[k], im.pt_begin[k + 1] - im.pt_begin[k], st_all + k, dmax2, prune2,
                            partials_all + (size_t)b0 * ICP_NSUM_MAX, ticket_all + k, sa, (int)blockIdx.x - b0, im.cta_begin[k + 1] - b0);
}

// last CTA of the fitness kernel: fold the (sum d2, count) partials in a fixed shape and write the result record
__device__ void icp_last_cta_finish(const double* partials, const IcpState* st, unsigned* ticket, rtr_pose_result* res, int keep_ransac_fields,
                                    double* sums /* smem[2] */, int* is_last /* smem */, const int nparts) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(ticket, 1u);
        *is_last = (t == (unsigned)nparts - 1u);
    }
    __syncthreads();
    if (!*is_last) return;
    __threadfence();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 2) {
        const double* col = partials + warp;
        double v = 0;
        int b = lane;
        for (; b + 32 * 7 < nparts; b += 32 * 8) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = __ldcg(col + (size_t)(b + 32 * u) * 2);
#pragma unroll
            for (int u = 0; u < 8; ++u) v += x[u];
        }
        for (; b < nparts; b += 32) v += __ldcg(col + (size_t)b * 2);
        v = warp_sum(v);
        if (lane == 0) sums[warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0u;
        for (int i = 0; i < 16; ++i) res->pose[i] = st->final_[i];
        res->fitness = sums[1] > 0 ? (float)(sums[0] / sums[1]) : FLT_MAX;
        res->iterations = st->iterations;
        if (keep_ransac_fields) {
            res->converged = res->converged ? st->state : 0;
        } else {
            res->converged = st->state; res->inliers = st->corr; res->hypothesis = -1; res->evaluated = 0;
            res->model_id = 0; res->n_keypoints_src = 0; res->n_keypoints_tgt = 0;
            for (int i = 0; i < 5; ++i) res->pad_[i] = 0;
        }
    }
}

// nn_prev: every query's neighbour at its last search (warm start of the grid search on large scans)
__device__ __forceinline__ void icp_fitness_warp_body(const IcpTarget& T, const float4* __restrict__ src, const float4* __restrict__ cur,
                                                      const int* __restrict__ nn_prev, int n,
                                                      const IcpState* __restrict__ st, double* partials, unsigned* ticket, rtr_pose_result* res,
                                                      int keep_ransac_fields, const int bid, const int nblocks) {
    __shared__ float m[16];
    __shared__ double red[ICPW_WARPS][2];
    __shared__ double sums[2];
    __shared__ int is_last;
    if (st->skipped) {     // pose / fitness stay RANSAC's (identity, FLT_MAX)
        if (bid == 0 && threadIdx.x == 0) { res->iterations = 0; res->converged = 0; }
        return;
    }
    if (threadIdx.x < 16) m[threadIdx.x] = st->final_[threadIdx.x];
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nwarps = nblocks * ICPW_WARPS;
    double s = 0, c = 0;
    // valid once at least one iteration has stored every query's neighbour
    const bool warm = nn_prev != nullptr && st->iterations > 0;
    if (T.bvh.n > 0) {
        for (int i = bid * ICPW_WARPS + warp; i < n; i += nwarps) {
            const float4 q = xform(m, __ldg(src + i));
            unsigned long long key = ~0ull;
            if (warm) {
                const int pb = nn_prev[i];
                if (pb >= 0) {
                    const float4 tp = __ldg(T.pts + pb);
                    key = ((unsigned long long)__float_as_uint(dist2f(q.x, q.y, q.z, tp.x, tp.y, tp.z)) << 32) | (unsigned)pb;
                }
            }
            key = wbvh_nearest_warp(T.bvh, q.x, q.y, q.z, lane, key);
            const unsigned h = (unsigned)(key >> 32);
            if (lane == 0 && h <= 0x7f800000u) { s += (double)__uint_as_float(h); c += 1.0; }
        }
    } else if (T.brute_n > 0) {
        GridView gb = T.g;
        gb.sorted = T.brute; gb.n = T.brute_n;
        for (int i = bid * ICPW_WARPS + warp; i < n; i += 2 * nwarps) {
            int i2 = i + nwarps;
            bool two = i2 < n;
            float4 q = xform(m, __ldg(src + i));
            float4 q2 = two ? xform(m, __ldg(src + i2)) : q;
            int b, b2; float d2, d22;
            brute_nearest_warp2(gb, q.x, q.y, q.z, q2.x, q2.y, q2.z, lane, b, d2, b2, d22);
            if (lane == 0 && b >= 0) { s += (double)d2; c += 1.0; }
            if (lane == 0 && two && b2 >= 0) { s += (double)d22; c += 1.0; }
        }
    } else {
        for (int i = bid * ICPW_WARPS + warp; i < n; i += nwarps) {
            float4 q = xform(m, __ldg(src + i));
            float pr2 = FLT_MAX;
            if (warm) {
                const int pb = nn_prev[i];
                if (pb >= 0) {
                    const float4 tp = __ldg(T.pts + pb);
                    pr2 = __fmul_rn(dist2f(q.x, q.y, q.z, tp.x, tp.y, tp.z), 1.000001f);
                }
            }
            int b; float d2; float4 t;
            grid_nearest_warp(T.g, q.x, q.y, q.z, pr2, lane, b, d2, t);
            if (lane == 0 && b >= 0) { s += (double)d2; c += 1.0; }
        }
    }
    if (lane == 0) { red[warp][0] = s; red[warp][1] = c; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < ICPW_WARPS; ++w) v += red[w][threadIdx.x];
        partials[(size_t)bid * 2 + threadIdx.x] = v;
    }
    icp_last_cta_finish(partials, st, ticket, res, keep_ransac_fields, sums, &is_last, nblocks);
}

__global__ void __launch_bounds__(ICPW_WARPS * 32) k_icp_fitness_warp(IcpTarget T, const float4* __restrict__ src, const float4* __restrict__ cur,
                                                                      const int* __restrict__ nn_prev, int n,
                                                                      const IcpState* __restrict__ st, double* partials, unsigned* ticket,
                                                                      rtr_pose_result* res, int keep_ransac_fields) {
    pdl_wait();
    icp_fitness_warp_body(T, src, cur, nn_prev, n, st, partials, ticket, res, keep_ransac_fields, (int)blockIdx.x, (int)gridDim.x);
}

__global__ void __launch_bounds__(ICPW_WARPS * 32) k_icp_fitness_warp_many(const __grid_constant__ IcpMany im, IcpTarget T, const float4* __restrict__ src_all,
                                                                           const float4* __restrict__ cur_all, const int* __restrict__ nn_prev_all,
                                                                           const IcpState* __restrict__ st_all,
                                                                           double* partials_all, unsigned* ticket_all, rtr_pose_result* res_all) {
    pdl_wait();
    const int k = icp_many_segment(im, (int)blockIdx.x);
    const int b0 = im.cta_begin[k], p0 = im.pt_begin[k];
    icp_fitness_warp_body(T, src_all + p0, cur_all + p0, nn_prev_all + p0, im.pt_begin[k + 1] - p0, st_all + k,
                          partials_all + (size_t)b0 * ICP_NSUM_MAX, ticket_all + k, res_all + k, 1, (int)blockIdx.x - b0, im.cta_begin[k + 1] - b0);
}

// ---- all iterations of every member cloud in ONE launch ----------------------------------------------------------------------
// With the hierarchy the search is a few hundred instructions per query, and an iteration launched on its own is bound by
// everything around it: ~6000 CTAs (10 waves of 4 per SM) that mostly do one or two queries each, the launch itself, the
// tail until the last CTA has solved (ncu / marks: 52 us per iteration for the bench batch, of which the searches need < 10).
// The persistent form keeps the same VIRTUAL CTAs — the same queries per warp, the same partial sums, the same fixed-shape
// fold by the member's last virtual CTA, hence the same bits — but runs them on a resident grid (cooperative launch: every
// CTA is on an SM, so waiting for another CTA cannot deadlock), iteration after iteration.  A member's next iteration starts
// as soon as ITS solve has been published (a per-member counter written behind a fence), independent of the other members;
// the fitness pass of a member follows its last iteration the same way.
__device__ __forceinline__ void icp_wait_member(const unsigned* gen, const IcpState* st, unsigned need) {
    if (threadIdx.x == 0) {
        const volatile unsigned* g = gen;
        const volatile int* done = &st->done;
        for (long long spin = 0; *g < need && !*done; ++spin) {
            if (spin > (1LL << 26)) __trap();          // a protocol bug must end in an error status, never in a hung GPU
            __nanosleep(40);
        }
        __threadfence();
    }
    __syncthreads();
}
template <int EST>
__global__ void __launch_bounds__(ICPW_WARPS * 32, 4) k_icp_persistent(const __grid_constant__ IcpMany im, IcpTarget T, const float4* __restrict__ src_all,
                                                                    float4* __restrict__ cur_all, int* __restrict__ nn_prev_all, IcpState* st_all,
                                                                    double dmax2, float prune2, double* partials_all, unsigned* ticket_all,
                                                                    unsigned* gen_all, IcpSolveArgs sa, int iterations, rtr_pose_result* res_all,
                                                                    int keep_ransac_fields) {
    const int total = im.cta_begin[RTR_MAX_SEGMENTS];
    for (int it = 0; it < iterations; ++it)
        for (int vb = blockIdx.x; vb < total; vb += gridDim.x) {
            const int k = icp_many_segment(im, vb);
            const int b0 = im.cta_begin[k], p0 = im.pt_begin[k];
            __syncthreads();                                                   // shared memory of the previous virtual CTA is free
            if (it > 0) icp_wait_member(gen_all + k, st_all + k, (unsigned)it);
            icp_corr_warp_body<EST>(T, cur_all + p0, nn_prev_all + p0, im.pt_begin[k + 1] - p0, st_all + k, dmax2, prune2,
                                    partials_all + (size_t)b0 * ICP_NSUM_MAX, ticket_all + k, sa, vb - b0, im.cta_begin[k + 1] - b0,
                                    gen_all + k, (unsigned)(it + 1));
        }
    for (int vb = blockIdx.x; vb < total; vb += gridDim.x) {
        const int k = icp_many_segment(im, vb);
        const int b0 = im.cta_begin[k], p0 = im.pt_begin[k];
        __syncthreads();
        if (iterations > 0) icp_wait_member(gen_all + k, st_all + k, (unsigned)iterations);
        icp_fitness_warp_body(T, src_all + p0, cur_all + p0, nn_prev_all + p0, im.pt_begin[k + 1] - p0, st_all + k,
                              partials_all + (size_t)b0 * ICP_NSUM_MAX, ticket_all + k, res_all + k, keep_ransac_fields, vb - b0, im.cta_begin[k + 1] - b0);
    }
}

// Build the two-level hierarchy of a small target (one launch, one CTA).  pts: the target's points by local index; idx_base:
// added to the reported indices.
static int wbvh_build_dev(rtr_context* ctx, const float4* pts, int n, int idx_base, const float* bb_min, const float* bb_max, WbvhView* out) {
    out->n = 0; out->nleaf = 0; out->boxes = nullptr; out->pts = nullptr;
    if (n <= 0 || n > WBVH_MAX_POINTS) return 0;
    const int nleaf = (n + WBVH_LEAF - 1) / WBVH_LEAF;
    float4 *boxes = nullptr, *mpts = nullptr;
    if (int e = tmp_alloc(ctx, &boxes, (size_t)2 * nleaf, "icp.bvh")) return e;
    if (int e = tmp_alloc(ctx, &mpts, n, "icp.bvh")) return e;
    k_wbvh_build<<<1, WBVH_BUILD_THREADS, 0, ctx->stream>>>(pts, n, idx_base, bb_min[0], bb_min[1], bb_min[2], bb_max[0], bb_max[1], bb_max[2], boxes, mpts);
    RTR_LAUNCH_CHECK(ctx, "icp.bvh_build");
    out->boxes = boxes; out->pts = mpts; out->nleaf = nleaf; out->n = n;
    return 0;
}
// Build the 32-ary hierarchy of a target of any size: Morton codes, one radix sort, one warp per leaf, one warp per upper node.
static int wide_build_dev(rtr_context* ctx, const float4* pts, int n, int idx_base, const float* bb_min, const float* bb_max, WideBvh* out) {
    memset(out, 0, sizeof(*out));
    if (n <= 0) return 0;
    int count[WIDE_MAX_LEVELS], offset[WIDE_MAX_LEVELS], depth = 0, total = 0;
    for (int c = nblk(n, WBVH_LEAF);; c = nblk(c, 32)) {
        count[depth] = c; offset[depth] = total; total += c; ++depth;
        if (c <= 32) break;
        if (depth == WIDE_MAX_LEVELS) return rtr_fail("icp.wide", "hierarchy deeper than WIDE_MAX_LEVELS", RTR_ERR_INVALID);
    }
    unsigned *keys = nullptr, *keys2 = nullptr; int *vals = nullptr, *vals2 = nullptr; char* temp = nullptr;
    float4 *boxes = nullptr, *mpts = nullptr;
    if (int e = tmp_alloc(ctx, &keys, n, "icp.wide")) return e;
    if (int e = tmp_alloc(ctx, &keys2, n, "icp.wide")) return e;
    if (int e = tmp_alloc(ctx, &vals, n, "icp.wide")) return e;
    if (int e = tmp_alloc(ctx, &vals2, n, "icp.wide")) return e;
    if (int e = tmp_alloc(ctx, &boxes, (size_t)2 * total, "icp.wide")) return e;
    if (int e = tmp_alloc(ctx, &mpts, n, "icp.wide")) return e;
    k_wide_morton<<<nblk(n, 256), 256, 0, ctx->stream>>>(pts, n, bb_min[0], bb_min[1], bb_min[2], bb_max[0], bb_max[1], bb_max[2], keys, vals);
    RTR_LAUNCH_CHECK(ctx, "icp.wide_keys");
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, vals, vals2, n, 0, 30, ctx->stream);
    if (int e = tmp_alloc(ctx, &temp, tb, "icp.wide")) return e;
    RTR_CHECK(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys2, vals, vals2, n, 0, 30, ctx->stream), "icp.wide_sort");
    RTR_MARK(ctx, "icp.wide_sort");
    k_wide_leaves<<<nblk((long long)count[0] * 32, 256), 256, 0, ctx->stream>>>(pts, vals2, n, idx_base, mpts, boxes);
    RTR_LAUNCH_CHECK(ctx, "icp.wide_leaves");
    for (int l = 1; l < depth; ++l) {
        k_wide_level<<<nblk((long long)count[l] * 32, 256), 256, 0, ctx->stream>>>(boxes + 2 * (size_t)offset[l - 1], count[l - 1], boxes + 2 * (size_t)offset[l]);
        RTR_LAUNCH_CHECK(ctx, "icp.wide_level");
    }
    dev_free(ctx, keys); dev_free(ctx, keys2); dev_free(ctx, vals); dev_free(ctx, vals2); dev_free(ctx, temp);
    out->pts = mpts; out->boxes = boxes; out->n = n; out->depth = depth;
    for (int l = 0; l < depth; ++l) { out->count[l] = count[l]; out->offset[l] = offset[l]; }
    return 0;
}
// RTR_ICP_WIDE=0: far queries of the thread-per-query kernels walk the coarse grid instead of the 32-ary hierarchy (A/B runs, tests)
static bool icp_wide_wanted() {
    const char* e = getenv("RTR_ICP_WIDE");
    return !(e && e[0] == '0');
}
__global__ void k_nearest_wide(const __grid_constant__ WideBvh B, const float4* __restrict__ q, int nq, int* __restrict__ idx, float* __restrict__ d2) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nq) return;
    const float4 p = __ldg(q + i);
    const unsigned long long key = wide_nearest_warp(B, p.x, p.y, p.z, lane, ~0ull);
    if (lane == 0) {
        const bool found = (unsigned)key != 0xffffffffu;
        idx[i] = found ? (int)(unsigned)key : -1;
        d2[i] = found ? __uint_as_float((unsigned)(key >> 32)) : __int_as_float(0x7f800000);
    }
}
// the hierarchy on its own (rtr_nearest with RTR_NEAREST_BVH=1: the tests compare it with the oracle's brute-force search)
__global__ void k_nearest_wbvh(WbvhView B, const float4* __restrict__ q, int nq, int* __restrict__ idx, float* __restrict__ d2) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nq) return;
    const float4 p = __ldg(q + i);
    const unsigned long long key = wbvh_nearest_warp(B, p.x, p.y, p.z, lane, ~0ull);
    const unsigned h = (unsigned)(key >> 32);
    if (lane == 0) { idx[i] = h <= 0x7f800000u ? (int)(unsigned)key : -1; d2[i] = __uint_as_float(h); }
}
int rtr_nearest_bvh_dev(rtr_cloud* tgt, const float4* d_q, int nq, int* d_idx, float* d_d2) {
    rtr_context* ctx = tgt->ctx;
    if (int e = rtr_ensure_bbox(tgt)) return e;
    const char* mode = getenv("RTR_NEAREST_BVH");
    const bool force_wide = mode && mode[0] == '2';
    if (tgt->n > WBVH_MAX_POINTS || force_wide) {       // any size: the 32-ary hierarchy
        WideBvh W;
        if (int e = wide_build_dev(ctx, tgt->pts, tgt->n, 0, tgt->bb_min, tgt->bb_max, &W)) return e;
        if (W.n == 0) return rtr_fail("nearest", "empty target", RTR_ERR_INVALID);
        k_nearest_wide<<<nblk((long long)nq * 32, 128), 128, 0, ctx->stream>>>(W, d_q, nq, d_idx, d_d2);
        RTR_LAUNCH_CHECK(ctx, "nearest.wide");
        return 0;
    }
    WbvhView B;
    if (int e = wbvh_build_dev(ctx, tgt->pts, tgt->n, 0, tgt->bb_min, tgt->bb_max, &B)) return e;
    if (B.n == 0) return rtr_fail("nearest", "empty target", RTR_ERR_INVALID);
    k_nearest_wbvh<<<nblk((long long)nq * 32, 128), 128, 0, ctx->stream>>>(B, d_q, nq, d_idx, d_d2);
    RTR_LAUNCH_CHECK(ctx, "nearest.bvh");
    return 0;
}
// RTR_ICP_BVH=0: keep the plain brute-force scan for small targets (A/B runs, tests)
static bool icp_bvh_wanted() {
    const char* e = getenv("RTR_ICP_BVH");
    return !(e && e[0] == '0');
}

// getFitnessScore(): mean squared NN distance of (final o source)
// Queries farther than sqrt(far2) from the target's bounding box search the COARSE grid gc: the walk beyond the 27-cell
// block visits every row inside the search sphere, and for a query metres away from a finely gridded target that is all
// of them; cells four times wider mean 16x fewer rows (either grid gives the exact nearest neighbour).
__global__ void __launch_bounds__(ICP_THREADS) k_icp_fitness(GridView g, GridView gc, Box6 bb, float far2, const float4* __restrict__ src, int n,
                                                             const IcpState* __restrict__ st, double* partials, unsigned* ticket,
                                                             rtr_pose_result* res, int keep_ransac_fields, const __grid_constant__ WideBvh W,
                                                             const float4* __restrict__ tgt_pts, const int* __restrict__ nn_prev) {
    __shared__ float m[16];
    __shared__ double red[ICP_THREADS / 32][2];
    __shared__ double sums[2];
    __shared__ int is_last;
    pdl_wait();
    if (st->skipped) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { res->iterations = 0; res->converged = 0; }
        return;
    }
    if (threadIdx.x < 16) m[threadIdx.x] = st->final_[threadIdx.x];
    __syncthreads();
    double s = 0, c = 0;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    {
        const bool live = i < n;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) q = xform(m, __ldg(src + i));
        int b = -1; float d2 = 0.f;
        const bool far = box_dist2f(bb, q.x, q.y, q.z) > far2;
        const WarmStart ws = icp_warm_start(nn_prev, i, live && nn_prev != nullptr, tgt_pts, q);
        if (live && !(far && W.n > 0)) {
            float4 t;
            if (ws.b >= 0) { b = ws.b; d2 = ws.d2; t = ws.p; grid_nearest_far(far ? gc : g, q.x, q.y, q.z, FLT_MAX, b, d2, t); }
            else grid_nearest(far ? gc : g, q.x, q.y, q.z, b, d2);
        }
        if (W.n > 0) {
            const unsigned long long key0 = ws.b >= 0 ? (((unsigned long long)__float_as_uint(ws.d2) << 32) | (unsigned)ws.b) : 0x7f800000ffffffffull;
            wide_nearest_lanes(W, live && far, q, key0, threadIdx.x & 31, b, d2);
        }
        if (b >= 0) { s = (double)d2; c = 1.0; }
    }
    s = warp_sum(s); c = warp_sum(c);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[warp][0] = s; red[warp][1] = c; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < ICP_THREADS / 32; ++w) v += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * 2 + threadIdx.x] = v;
    }
    icp_last_cta_finish(partials, st, ticket, res, keep_ransac_fields, sums, &is_last, (int)gridDim.x);
}

// CTAs of the warp-per-query kernels for a source of n points — a function of n alone, the same for a cloud on its own and
// as a member of a model set: it fixes the summation shape, hence the bits of the pose.  A warp takes 1 .. 4 queries one after
// the other (RTR_ICP_QPW caps it): small sources keep every warp they can get; for large ones an iteration is bound by the
// number of CTA waves (the bench batch: ~6000 CTAs = 10 waves of 4 per SM, 52 us per iteration at one query per warp, 44 us
// at four), not by the searches.
static int icp_warp_ctas(int n, int sm_count) {
    static const int qpw_max = []() { const char* e = getenv("RTR_ICP_QPW"); int v = e ? atoi(e) : 4; return v >= 1 && v <= 64 ? v : 4; }();
    const int qpw = std::max(1, std::min(qpw_max, n / (ICPW_WARPS * 2 * sm_count)));
    return std::max(1, std::min(nblk(n, ICPW_WARPS * qpw), sm_count * 8));
}
// The two-level hierarchy costs one 25 us build launch and pays per query and iteration: worth it from ~60 k query-iterations
// (a batch, a dense model); a single small model keeps the plain scan.  Either search is exact and the kernels add their
// correspondences in the same order, so the choice never changes a bit of the result.
static bool icp_hierarchy_pays(long long n_queries, int iterations) {
    return icp_bvh_wanted() && n_queries * (long long)std::max(iterations, 1) >= 60000;
}

// one cooperative launch for all iterations + the fitness pass (k_icp_persistent).  Built for SURVEY section 7 step 6, exact
// (bit-identical records, tests/test_gpu_parity.py), and MEASURED SLOWER than one launch per iteration on the bench batch:
// 715 us against 562 us for the ten iterations + fitness of the 8-model batch.  A resident CTA walks ~10 virtual CTAs per
// iteration one after the other (each a dependent chain of a few global round trips), where separate launches let the
// hardware scheduler run the same ~10 waves back to back with programmatic dependent launch hiding the hand-over; the
// per-member barrier adds its own round trips.  So it is opt-in: RTR_ICP_PERSISTENT=1.
static bool icp_persistent_wanted() {
    const char* e = getenv("RTR_ICP_PERSISTENT");
    return e && e[0] == '1';
}
static int icp_persistent_launch(rtr_context* ctx, const IcpMany& im, const IcpTarget& T, const float4* src_all, float4* cur, int* nn_prev, IcpState* st,
                                 double dmax2, float prune2, double* partials, unsigned* ticket, unsigned* gen, const IcpSolveArgs& sa, int iterations,
                                 rtr_pose_result* d_results, int keep_ransac_fields, bool plane) {
    const void* fn = plane ? (const void*)k_icp_persistent<1> : (const void*)k_icp_persistent<0>;
    int per_sm = 1;
    if (int e = rtr_func_occupancy(fn, ctx->device, ICPW_WARPS * 32, 0, &per_sm)) return e;
    const int total = im.cta_begin[RTR_MAX_SEGMENTS];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::max(1, std::min(total, ctx->sm_count * per_sm)));
    cfg.blockDim = dim3(ICPW_WARPS * 32); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t le = plane ? cudaLaunchKernelEx(&cfg, k_icp_persistent<1>, im, T, src_all, cur, nn_prev, st, dmax2, prune2, partials, ticket, gen, sa, iterations, d_results, keep_ransac_fields)
                           : cudaLaunchKernelEx(&cfg, k_icp_persistent<0>, im, T, src_all, cur, nn_prev, st, dmax2, prune2, partials, ticket, gen, sa, iterations, d_results, keep_ransac_fields);
    RTR_CHECK(le, "icp.persistent");
    return 0;
}

int rtr_icp_dev(rtr_cloud* src, rtr_cloud* tgt, const rtr_icp_params* p, const float* d_init_pose16, int init_from_result,
                rtr_pose_result* d_result) {
    RtrRange nvtx_range("rtr.icp");
    rtr_context* ctx = src->ctx;
    int n = src->n;
    if (p->estimator != 0 && p->estimator != 1) return rtr_fail("icp", "estimator must be 0 (SVD) or 1 (point-to-plane LLS)", RTR_ERR_INVALID);
    if (p->estimator == 1 && !tgt->normals) return rtr_fail("icp", "estimator 1 (point-to-plane) needs rtr_normals on the target", RTR_ERR_NOT_READY);
    if (int e = rtr_ensure_bbox(tgt)) return e;
    DevGrid* g;
    {   // the exact 1-NN search works on any grid: reuse a cached one of comparable cell size
        float want = rtr_icp_cell(tgt);
        // a cap of the order of the cell: make the cell the cap, so the 27-cell block holds every admissible correspondence
        // and no query ever needs the walk beyond it (grid_nearest_ex stops after the block when prune2 <= (0.999 h)^2)
        const float cap = p->max_correspondence_distance;
        if (cap > want && cap < 3.f * want) want = cap;
        if (n < 65536) {
            // warp-per-query kernels warm-start every search from the previous neighbour: prefer the finest cached grid in range
            g = nullptr;
            const float lo = want * (cap == want ? 1.0f : 0.4f), hi = want * 1.7f;
            for (auto& kv : tgt->grids) { DevGrid& c = kv.second; if (c.h >= lo && c.h <= hi && (!g || c.h < g->h)) g = &c; }
            if (!g) if (int e = rtr_get_grid(tgt, want, &g)) return e;
        } else if (int e = rtr_get_grid_any(tgt, want, want * (cap == want ? 1.0f : 0.6f), want * 1.7f, &g)) return e;
    }
    GridView v = rtr_view(g);
    // large sources: a second, coarser grid for the queries far from the target (k_icp_fitness).  1M scan points onto a 100k-point
    // model: fitness 17.8 -> 8.1 ms, an uncapped iteration 32.9 -> 15.4 ms with cells 4x wider (8x: 16.7 / 29.8 ms).
    // RTR_ICP_FAR_FACTOR <= 1 disables
    GridView vc = v;
    float far2 = FLT_MAX;
    WideBvh W;
    memset(&W, 0, sizeof(W));
    if (n >= 65536) {
        static const float far_factor = []() { const char* e = getenv("RTR_ICP_FAR_FACTOR"); return e ? (float)atof(e) : 4.0f; }();
        // worth building only if some of the source can be far from the target: judged on the untransformed bounding boxes
        // (a heuristic for speed only — either grid returns the exact neighbour)
        bool may_be_far = false;
        if (int e = rtr_ensure_bbox(src)) return e;
        for (int a = 0; a < 3; ++a)
            may_be_far |= src->bb_min[a] < tgt->bb_min[a] - 6.f * v.h || src->bb_max[a] > tgt->bb_max[a] + 6.f * v.h;
        if (may_be_far && icp_wide_wanted()) {
            // far queries: the 32-ary hierarchy (bvh.cuh) — a handful of 32-wide box tests per query instead of a ring walk
            if (int e = wide_build_dev(ctx, tgt->pts, tgt->n, 0, tgt->bb_min, tgt->bb_max, &W)) return e;
            far2 = (6.f * v.h) * (6.f * v.h);
        } else if (far_factor > 1.f && may_be_far) {
            DevGrid* gcoarse;
            if (int e = rtr_get_grid(tgt, g->h * far_factor, &gcoarse)) return e;
            vc = rtr_view(gcoarse);
            far2 = (6.f * v.h) * (6.f * v.h);
        }
    }
    float4* cur = nullptr; IcpState* st = nullptr; double* partials = nullptr;
    int nb = std::max(nblk(n, ICP_THREADS), 1);
    if (int e = tmp_alloc(ctx, &cur, n, "icp")) return e;
    if (int e = tmp_alloc(ctx, &st, 1, "icp")) return e;
    unsigned* ticket = nullptr;
    if (int e = tmp_alloc(ctx, &ticket, 1, "icp")) return e;
    const IcpSolveArgs sa{p->max_iterations, p->force_iterations, p->mse_threshold_absolute};
    const int nbw = icp_warp_ctas(n, ctx->sm_count);    // CTAs of the warp-per-query kernels
    if (int e = tmp_alloc(ctx, &partials, (size_t)std::max(nb, nbw) * ICP_NSUM_MAX, "icp")) return e;
    k_icp_init<<<nb, ICP_THREADS, 0, ctx->stream>>>(src->pts, n, d_init_pose16, init_from_result ? d_result : nullptr, cur, st, ticket);
    RTR_LAUNCH_CHECK(ctx, "icp.init");
    const float4* src_pts = src->pts;
    float4 *cur2 = nullptr, *src2 = nullptr;
    if (n >= 16384 && tgt->n >= 1) {
        int *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr; char* temp = nullptr;
        if (int e = tmp_alloc(ctx, &keys, n, "icp")) return e;
        if (int e = tmp_alloc(ctx, &vals, n, "icp")) return e;
        if (int e = tmp_alloc(ctx, &keys2, n, "icp")) return e;
        if (int e = tmp_alloc(ctx, &vals2, n, "icp")) return e;
        if (int e = tmp_alloc(ctx, &cur2, n, "icp")) return e;
        if (int e = tmp_alloc(ctx, &src2, n, "icp")) return e;
        k_icp_cell_keys<<<nb, ICP_THREADS, 0, ctx->stream>>>(v, cur, n, keys, vals);
        RTR_LAUNCH_CHECK(ctx, "icp.keys");
        int end_bit = 1;
        while ((1LL << end_bit) < (long long)g->ncells) ++end_bit;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, vals, vals2, n, 0, end_bit, ctx->stream);
        if (int e = tmp_alloc(ctx, &temp, tb, "icp")) return e;
        RTR_CHECK(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys2, vals, vals2, n, 0, end_bit, ctx->stream), "icp.sort");
        RTR_MARK(ctx, "icp.cub_sort");
        k_icp_permute<<<nb, ICP_THREADS, 0, ctx->stream>>>(cur, src->pts, vals2, n, cur2, src2);
        RTR_LAUNCH_CHECK(ctx, "icp.permute");
        dev_free(ctx, keys); dev_free(ctx, vals); dev_free(ctx, keys2); dev_free(ctx, vals2); dev_free(ctx, temp);
        dev_free(ctx, cur);
        cur = cur2; src_pts = src2;
    }
    double dmax2 = p->max_correspondence_distance > 0.f ? (double)p->max_correspondence_distance * (double)p->max_correspondence_distance : DBL_MAX;
    // search radius for pruning: the cap rounded UP in float so no candidate with d2 <= dmax2 is ever skipped
    float prune2 = FLT_MAX;
    if (p->max_correspondence_distance > 0.f) { prune2 = (float)dmax2; if ((double)prune2 < dmax2) prune2 = nextafterf(prune2, FLT_MAX); }
    // small sources: one warp per query (latency), large ones: one thread per query in cell order (throughput)
    const bool warp_per_query = n < 65536;
    const bool plane = p->estimator == 1;
    // warp-per-query kernels: the target as they see it, and each query's previous neighbour (warm start of the next search)
    IcpTarget T;
    T.g = v; T.brute = v.sorted; T.brute_n = (tgt->n <= RTR_BRUTE_NN_MAX) ? tgt->n : 0; T.pts = tgt->pts; T.normals = plane ? tgt->normals : nullptr;
    T.bvh.n = 0; T.bvh.nleaf = 0; T.bvh.boxes = nullptr; T.bvh.pts = nullptr;
    // every query's neighbour at its last search: the next search starts from it (RTR_ICP_WARM=0: thread-per-query kernels search cold)
    static const bool warm_large = []() { const char* e = getenv("RTR_ICP_WARM"); return !(e && e[0] == '0'); }();
    int* nn_prev = nullptr;
    if (warp_per_query || warm_large) if (int e = tmp_alloc(ctx, &nn_prev, std::max(n, 1), "icp")) return e;
    if (!warp_per_query && nn_prev) RTR_CHECK(cudaMemsetAsync(nn_prev, 0xFF, (size_t)std::max(n, 1) * sizeof(int), ctx->stream), "icp");      // -1: no neighbour yet
    // small targets (every scan the reference ships): the warp-per-query kernels search a two-level hierarchy of the target
    if (warp_per_query && tgt->n >= 1 && tgt->n <= WBVH_MAX_POINTS && icp_hierarchy_pays(n, p->max_iterations))
        if (int e = wbvh_build_dev(ctx, tgt->pts, tgt->n, 0, tgt->bb_min, tgt->bb_max, &T.bvh)) return e;
    const Box6 bb{tgt->bb_min[0], tgt->bb_min[1], tgt->bb_min[2], tgt->bb_max[0], tgt->bb_max[1], tgt->bb_max[2]};
    if (warp_per_query && n >= 1 && tgt->n >= 1 && icp_persistent_wanted()) {
        // all iterations and the fitness pass in one cooperative launch (k_icp_persistent), as a model set of one member
        IcpMany im;
        memset(&im, 0, sizeof(im));
        im.nseg = 1;
        for (int k = 1; k <= RTR_MAX_SEGMENTS; ++k) { im.pt_begin[k] = n; im.cta_begin[k] = nbw; }
        unsigned* gen = nullptr;
        if (int e = tmp_alloc(ctx, &gen, 1, "icp")) return e;
        RTR_CHECK(cudaMemsetAsync(gen, 0, sizeof(unsigned), ctx->stream), "icp");
        if (int e = icp_persistent_launch(ctx, im, T, src_pts, cur, nn_prev, st, dmax2, prune2, partials, ticket, gen, sa, p->max_iterations, d_result, init_from_result, plane)) return e;
        RTR_LAUNCH_CHECK(ctx, "icp.corr");
        dev_free(ctx, cur); dev_free(ctx, src2); dev_free(ctx, st); dev_free(ctx, partials); dev_free(ctx, ticket); dev_free(ctx, nn_prev); dev_free(ctx, gen);
        return 0;
    }
    if (n >= 1 && tgt->n >= 1) {
        for (int it = 0; it < p->max_iterations; ++it) {
            if (warp_per_query) {
                if (plane) launch_pdl(k_icp_corr_warp<1>, nbw, ICPW_WARPS * 32, 0, ctx->stream, T, cur, nn_prev, n, st, dmax2, prune2, partials, ticket, sa);
                else launch_pdl(k_icp_corr_warp<0>, nbw, ICPW_WARPS * 32, 0, ctx->stream, T, cur, nn_prev, n, st, dmax2, prune2, partials, ticket, sa);
            } else {
                if (plane) launch_pdl(k_icp_corr<1>, nb, ICP_THREADS, 0, ctx->stream, v, vc, bb, far2, (const float4*)tgt->normals, cur, n, st, dmax2, prune2, partials, ticket, sa, W, (const float4*)tgt->pts, nn_prev);
                else launch_pdl(k_icp_corr<0>, nb, ICP_THREADS, 0, ctx->stream, v, vc, bb, far2, (const float4*)nullptr, cur, n, st, dmax2, prune2, partials, ticket, sa, W, (const float4*)tgt->pts, nn_prev);
            }
            RTR_LAUNCH_CHECK(ctx, "icp.corr");
        }
    }
    if (warp_per_query) launch_pdl(k_icp_fitness_warp, nbw, ICPW_WARPS * 32, 0, ctx->stream, T, src_pts, (const float4*)cur, (const int*)nn_prev, n, (const IcpState*)st, partials, ticket, d_result, init_from_result);
    else launch_pdl(k_icp_fitness, nb, ICP_THREADS, 0, ctx->stream, v, vc, bb, far2, src_pts, n, (const IcpState*)st, partials, ticket, d_result, init_from_result, W, (const float4*)tgt->pts,
                    (const int*)nn_prev);
    RTR_LAUNCH_CHECK(ctx, "icp.fitness");
    dev_free(ctx, cur); dev_free(ctx, src2); dev_free(ctx, st); dev_free(ctx, partials); dev_free(ctx, ticket); dev_free(ctx, nn_prev);
    return 0;
}

// ---- ICP over a model set: every member's refinement in the same launches (one per iteration) --------------------------
__global__ void k_icp_init_many(const __grid_constant__ IcpMany im, const float4* __restrict__ src_all, const rtr_pose_result* __restrict__ res_all,
                                float4* __restrict__ cur_all, IcpState* __restrict__ st_all, unsigned* __restrict__ ticket_all) {
    const int total = im.pt_begin[RTR_MAX_SEGMENTS];
    if (blockIdx.x == 0 && threadIdx.x < im.nseg) {
        const int k = threadIdx.x;
        IcpState* st = st_all + k;
        for (int i = 0; i < 16; ++i) { st->final_[i] = res_all[k].pose[i]; st->step[i] = (i % 5 == 0) ? 1.f : 0.f; }
        const int skip = res_all[k].converged == 0 ? 1 : 0;
        st->prev_mse = DBL_MAX; st->iterations = 0; st->done = skip; st->state = 0; st->corr = 0; st->have_step = 0; st->skipped = skip;
        ticket_all[k] = 0u;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int k = 0;
#pragma unroll
        for (int step = RTR_MAX_SEGMENTS / 2; step > 0; step >>= 1) if (i >= im.pt_begin[k + step]) k += step;
        cur_all[i] = xform(res_all[k].pose, __ldg(src_all + i));
    }
}

// scan: nullptr — the target is member tgt_seg of the set, searched through the set's grids, neighbours named by set-wide index;
// or the scan as a cloud of its own (prepared path: the set carries no grids and no normals) — its own grids and normals,
// neighbours named by the scan's local index.  Lowest-index tie-breaking is the same in either index space: same records.
static int icp_many_dev(rtr_cloud* set, int n_models, int tgt_seg, const rtr_icp_params* p, rtr_pose_result* d_results, rtr_cloud* scan = nullptr,
                        const WbvhView* scan_bvh = nullptr /* built by the caller in the index space described below */) {
    RtrRange nvtx_range("rtr.icp.many");
    rtr_context* ctx = set->ctx;
    if (p->estimator != 0 && p->estimator != 1) return rtr_fail("icp", "estimator must be 0 (SVD) or 1 (point-to-plane LLS)", RTR_ERR_INVALID);
    if (p->estimator == 1 && !(scan ? scan->normals : set->normals)) return rtr_fail("icp", "estimator 1 (point-to-plane) needs normals on the target", RTR_ERR_NOT_READY);
    const int t0 = set->seg_begin[tgt_seg], nt = set->seg_begin[tgt_seg + 1] - t0;
    const int n_src = set->seg_begin[n_models];
    // target search structure: scans up to RTR_BRUTE_NN_MAX points are searched without one (the kernels only stride over the
    // target's points: any grid's copy serves); larger scans use the cached grid closest to the ICP cell, as rtr_icp_dev does.
    // Either way the neighbour is the exact one.
    GridView v;
    {
        rtr_cloud box;             // a bounding box + size only, for the cell heuristic
        box.n = nt;
        for (int a = 0; a < 3; ++a) { box.bb_min[a] = set->seg_bb[6 * tgt_seg + a]; box.bb_max[a] = set->seg_bb[6 * tgt_seg + 3 + a]; }
        float want = rtr_icp_cell(&box);
        const float cap = p->max_correspondence_distance;
        if (cap > want && cap < 3.f * want) want = cap;
        // warm-started searches are bounded by the previous neighbour's distance, usually far below the point spacing: the
        // FINEST cached grid in range keeps the 27-cell block small (the 0.05 normals grid holds ~80 candidates per block on
        // the repo scans against ~400 for the 0.10 FPFH grid, which made the warm start no faster than the brute-force pass)
        DevGrid* g = nullptr;
        const float lo = want * (cap == want ? 1.0f : 0.4f), hi = want * 1.7f;
        rtr_cloud* owner = scan ? scan : set;
        for (auto& kv : owner->grids) { DevGrid& c = kv.second; if (c.h >= lo && c.h <= hi && (!g || c.h < g->h)) g = &c; }
        if (!g) if (int e = rtr_get_grid(owner, want, &g)) return e;
        v = scan ? rtr_view(g) : rtr_segment_view(g, set, tgt_seg);
    }
    const int idx_base = scan ? 0 : t0;       // index space of the neighbours: the scan's own, or the set's
    IcpTarget T;
    T.g = v; T.brute = v.sorted + idx_base; T.brute_n = (nt <= RTR_BRUTE_NN_MAX) ? nt : 0;     // the scan's (slice of the) cell-ordered copy
    T.pts = set->pts + (t0 - idx_base); T.normals = (p->estimator == 1) ? (scan ? scan->normals : set->normals) : nullptr;
    T.bvh.n = 0; T.bvh.nleaf = 0; T.bvh.boxes = nullptr; T.bvh.pts = nullptr;
    // small scans: the warp-per-query kernels search a two-level hierarchy of the scan (indices reported in the same space)
    if (scan_bvh) T.bvh = *scan_bvh;
    else if (nt >= 1 && nt <= WBVH_MAX_POINTS && icp_hierarchy_pays(n_src, p->max_iterations))
        if (int e = wbvh_build_dev(ctx, set->pts + t0, nt, idx_base, &set->seg_bb[6 * tgt_seg], &set->seg_bb[6 * tgt_seg + 3], &T.bvh)) return e;
    IcpMany im;
    memset(&im, 0, sizeof(im));
    im.nseg = n_models;
    int ctas = 0;
    for (int k = 0; k <= RTR_MAX_SEGMENTS; ++k) {
        im.pt_begin[k] = set->seg_begin[std::min(k, n_models)];
        im.cta_begin[k] = ctas;
        // the CTA count the single-cloud launch of the same kernel would use (rtr_icp_dev): same summation shape, same bits
        if (k < n_models) ctas += icp_warp_ctas(set->seg_begin[k + 1] - set->seg_begin[k], ctx->sm_count);
    }
    float4* cur = nullptr; IcpState* st = nullptr; double* partials = nullptr; unsigned* ticket = nullptr; int* nn_prev = nullptr;
    if (int e = tmp_alloc(ctx, &cur, n_src, "icp")) return e;
    if (int e = tmp_alloc(ctx, &nn_prev, n_src, "icp")) return e;
    if (int e = tmp_alloc(ctx, &st, n_models, "icp")) return e;
    if (int e = tmp_alloc(ctx, &ticket, n_models, "icp")) return e;
    if (int e = tmp_alloc(ctx, &partials, (size_t)ctas * ICP_NSUM_MAX, "icp")) return e;
    const IcpSolveArgs sa{p->max_iterations, p->force_iterations, p->mse_threshold_absolute};
    k_icp_init_many<<<std::max(1, std::min(nblk(n_src, 256), 4 * ctx->sm_count)), 256, 0, ctx->stream>>>(im, set->pts, d_results, cur, st, ticket);
    RTR_LAUNCH_CHECK(ctx, "icp.init");
    double dmax2 = p->max_correspondence_distance > 0.f ? (double)p->max_correspondence_distance * (double)p->max_correspondence_distance : DBL_MAX;
    float prune2 = FLT_MAX;
    if (p->max_correspondence_distance > 0.f) { prune2 = (float)dmax2; if ((double)prune2 < dmax2) prune2 = nextafterf(prune2, FLT_MAX); }
    const bool plane = p->estimator == 1;
    if (icp_persistent_wanted() && nt >= 1) {
        unsigned* gen = nullptr;
        if (int e = tmp_alloc(ctx, &gen, n_models, "icp")) return e;
        RTR_CHECK(cudaMemsetAsync(gen, 0, sizeof(unsigned) * (size_t)n_models, ctx->stream), "icp");
        if (int e = icp_persistent_launch(ctx, im, T, set->pts, cur, nn_prev, st, dmax2, prune2, partials, ticket, gen, sa, p->max_iterations, d_results, 1, plane)) return e;
        RTR_LAUNCH_CHECK(ctx, "icp.corr");
        return 0;
    }
    if (nt >= 1) {
        for (int it = 0; it < p->max_iterations; ++it) {
            if (plane) launch_pdl(k_icp_corr_warp_many<1>, ctas, ICPW_WARPS * 32, 0, ctx->stream, im, T, cur, nn_prev, st, dmax2, prune2, partials, ticket, sa);
            else launch_pdl(k_icp_corr_warp_many<0>, ctas, ICPW_WARPS * 32, 0, ctx->stream, im, T, cur, nn_prev, st, dmax2, prune2, partials, ticket, sa);
            RTR_LAUNCH_CHECK(ctx, "icp.corr");
        }
    }
    launch_pdl(k_icp_fitness_warp_many, ctas, ICPW_WARPS * 32, 0, ctx->stream, im, T, (const float4*)set->pts, (const float4*)cur, (const int*)nn_prev, (const IcpState*)st, partials, ticket, d_results);
    RTR_LAUNCH_CHECK(ctx, "icp.fitness");
    return 0;
}

// ============================================================================= whole registration
// a zero-initialised or partly filled params struct must come back as RTR_ERR_INVALID, not hang the grid sizing
int rtr_validate_register_params(const rtr_register_params* p) {
    auto bad = [](float v) { return !(v > 0.f) || !std::isfinite(v); };
    if (bad(p->normal_radius) || bad(p->harris_radius) || bad(p->fpfh_radius))
        return rtr_fail("register", "normal_radius, harris_radius and fpfh_radius must be finite and > 0", RTR_ERR_INVALID);
    if (bad(p->ransac.max_correspondence_distance)) return rtr_fail("register", "ransac.max_correspondence_distance must be finite and > 0", RTR_ERR_INVALID);
    if (p->ransac.correspondence_k < 1 || p->ransac.correspondence_k > 8) return rtr_fail("register", "ransac.correspondence_k must be in [1, 8]", RTR_ERR_INVALID);
    if (p->run_icp && (p->icp.max_iterations < 0 || (p->icp.estimator != 0 && p->icp.estimator != 1)))
        return rtr_fail("register", "icp.max_iterations must be >= 0 and icp.estimator 0 or 1", RTR_ERR_INVALID);
    return 0;
}
__global__ void k_set_keypoints(rtr_pose_result* res, const int* n_src, const int* n_tgt) {
    if (threadIdx.x == 0) { res->n_keypoints_src = *n_src; res->n_keypoints_tgt = *n_tgt; }
}

static int fetch_result(rtr_context* ctx, rtr_pose_result* d_result, rtr_pose_result* host_result) {
    RTR_CHECK(cudaMemcpyAsync(ctx->pinned, d_result, sizeof(rtr_pose_result), cudaMemcpyDeviceToHost, ctx->stream), "result");
    RTR_MARK(ctx, "result.d2h");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "result");
    memcpy(host_result, ctx->pinned, sizeof(rtr_pose_result));
    return 0;
}

// ---- one scan against many database models ------------------------------------------------------------------------------
// The reference builds the scan side once (ScanPoint keypoints + descriptors, RealTimeRobot.cpp:45-60) and then loops over
// model keypoints (:62-102); its database holds many candidate models (README.md:10).  Here the scan and up to 31 models
// form one model set (the scan is the last member): every per-cloud stage is ONE launch over all of them, descriptor
// matching is one search of all model features against the scan's, and RANSAC / ICP run all models' hypotheses / iterations
// in shared launches.  About 35 launches for the whole batch instead of ~45 per registration, and the scan's stages run once.
#define RTR_KP_PREVIEW 64       // corners per cloud returned with the results (rtr_register_many_keypoints)
struct KpPreview { int count; int pad_[3]; float4 xyz[RTR_KP_PREVIEW]; };

__global__ void k_register_finish_many(const __grid_constant__ IcpMany im, rtr_pose_result* __restrict__ res, const int* __restrict__ kp_count,
                                       const float4* __restrict__ kp_xyz_all, int n_members, KpPreview* __restrict__ preview, int model_id_base) {
    const int k = blockIdx.x;       // member (models first, the scan last)
    const int n_models = n_members - 1;
    if (threadIdx.x == 0 && k < n_models) { res[k].n_keypoints_src = kp_count[k]; res[k].n_keypoints_tgt = kp_count[n_models]; res[k].model_id = model_id_base + k; }
    const int cnt = kp_count[k];
    if (threadIdx.x == 0) preview[k].count = cnt;
    // corner list of member k starts at its first point index (im.pt_begin covers the models; the scan follows them)
    const int first = k < n_models ? im.pt_begin[k] : im.pt_begin[RTR_MAX_SEGMENTS];
    for (int i = threadIdx.x; i < min(cnt, RTR_KP_PREVIEW); i += blockDim.x) preview[k].xyz[i] = kp_xyz_all[first + i];
}

// Second half of a batch, from the descriptor rows on: correspondences, prerejective RANSAC, ICP, the finishing kernel and the
// D2H of the records (and the corner previews) into the context's pinned area.  d_cnt / d_xyz: corner count per member and
// corner coordinates (member k's list starts at its first point index); join: waits for whatever produced them on the second
// stream.  scan: see icp_many_dev.
struct JoinGuard {
    rtr_context* ctx; bool fork; bool joined;
    void join() { if (fork && !joined) { cudaStreamWaitEvent(ctx->stream, ctx->join_event, 0); joined = true; } }
    ~JoinGuard() { join(); }
};
static int register_many_back(rtr_cloud* set, int n_models, const rtr_register_params* p, const int* d_cnt, const float4* d_xyz, JoinGuard& join_guard,
                              rtr_cloud* scan) {
    rtr_context* ctx = set->ctx;
    const int tgt = n_models, nseg = n_models + 1;
    const int n_src = set->seg_begin[n_models], nt = set->n - n_src;
    // What RANSAC and ICP need of the scan alone — its grid at exactly d_max with the flattened block lists (the inlier test: every
    // surviving hypothesis of every model reads it for every source point; 1.9x fewer candidates than the normals grid's cells of
    // 0.05 against d_max = 0.0365), and the two-level hierarchy of the ICP kernels — depends on nothing but the scan's points: it is
    // built on the context's second stream (behind the corner stage) while descriptor matching runs, ~60 us off the batch's chain.
    rtr_cloud scan_alias;                  // the scan's slice of the set as a cloud of its own: borrowed points, its own grid
    scan_alias.ctx = ctx; scan_alias.n = nt; scan_alias.pts = set->pts + n_src;
    for (int a = 0; a < 3; ++a) { scan_alias.bb_min[a] = set->seg_bb[6 * tgt + a]; scan_alias.bb_max[a] = set->seg_bb[6 * tgt + 3 + a]; }
    scan_alias.bbox_valid = true;
    struct AliasGuard {                    // released stream-ordered on the main stream — on every path only after the join
        rtr_cloud* c; JoinGuard* jg;
        ~AliasGuard() { jg->join(); c->pts = nullptr; rtr_invalidate(c); }
    } alias_guard{&scan_alias, &join_guard};
    GridView scan_grid;
    BlockLists scan_blocks{nullptr, nullptr};
    WbvhView scan_bvh;
    scan_bvh.n = 0; scan_bvh.nleaf = 0; scan_bvh.boxes = nullptr; scan_bvh.pts = nullptr;
    memset(&scan_grid, 0, sizeof(scan_grid));
    const long long H = ((p->ransac.hypothesis_end > 0) ? p->ransac.hypothesis_end : p->ransac.max_iterations) - p->ransac.hypothesis_begin;
    const bool want_grid = H > 0 && nt >= 1 && p->ransac.max_correspondence_distance > 0.f;
    const bool want_bvh = p->run_icp && nt >= 1 && nt <= WBVH_MAX_POINTS && icp_hierarchy_pays(n_src, p->icp.max_iterations);
    {
        const bool fork = !ctx->profile && ctx->aux_stream && (want_grid || want_bvh);
        cudaStream_t main_stream = ctx->stream;
        if (fork) {
            RTR_CHECK(cudaEventRecord(ctx->fork_event, main_stream), "register_many.fork");        // the scan's points are in place
            RTR_CHECK(cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_event, 0), "register_many.fork");
            ctx->stream = ctx->aux_stream;
        }
        int e = 0;
        if (want_grid) {
            DevGrid* g = nullptr;
            e = rtr_get_grid(&scan_alias, p->ransac.max_correspondence_distance, &g);
            if (!e) { scan_grid = rtr_view(g); e = block_lists_build_dev(ctx, scan_grid, &scan_blocks); }
        }
        // index space of the hierarchy: the scan's own when it comes as a cloud (prepared path), the set's otherwise (icp_many_dev)
        if (!e && want_bvh) e = wbvh_build_dev(ctx, set->pts + n_src, nt, scan ? 0 : n_src, &set->seg_bb[6 * tgt], &set->seg_bb[6 * tgt + 3], &scan_bvh);
        ctx->stream = main_stream;
        if (fork) {
            cudaEventRecord(ctx->join_event, ctx->aux_stream);      // behind everything the second stream was given so far
            join_guard.fork = true; join_guard.joined = false;
        }
        if (e) return e;
    }
    const int k = p->ransac.correspondence_k;
    int* knn = nullptr; float* knn_dist = nullptr;
    if (int e = tmp_alloc(ctx, &knn, (size_t)n_src * k, "register_many")) return e;
    if (int e = tmp_alloc(ctx, &knn_dist, (size_t)n_src * k, "register_many")) return e;
    if (int e = rtr_match_features_dev(ctx, set->fpfh, n_src, set->fpfh + (size_t)n_src * 33, nt, k, knn, knn_dist)) return e;
    rtr_pose_result* d_res = nullptr; KpPreview* d_prev = nullptr;
    if (int e = tmp_alloc(ctx, &d_res, n_models, "register_many")) return e;
    if (int e = tmp_alloc(ctx, &d_prev, nseg, "register_many")) return e;
    join_guard.join();                     // the scan's structures (and the corner stage before them) are complete
    if (int e = ransac_many_dev(set, n_models, tgt, knn, k, &p->ransac, d_res, &scan_grid, &scan_blocks)) return e;
    if (p->run_icp) if (int e = icp_many_dev(set, n_models, tgt, &p->icp, d_res, scan, want_bvh ? &scan_bvh : nullptr)) return e;
    IcpMany im;
    memset(&im, 0, sizeof(im));
    im.nseg = n_models;
    for (int kk = 0; kk <= RTR_MAX_SEGMENTS; ++kk) im.pt_begin[kk] = set->seg_begin[std::min(kk, n_models)];
    join_guard.join();
    k_register_finish_many<<<nseg, 64, 0, ctx->stream>>>(im, d_res, d_cnt, d_xyz, nseg, d_prev, ctx->gather_batches ? ctx->model_id_base : 0);
    RTR_LAUNCH_CHECK(ctx, "register.kp");
    const size_t res_bytes = sizeof(rtr_pose_result) * (size_t)n_models, prev_bytes = sizeof(KpPreview) * (size_t)nseg;
    if (res_bytes + prev_bytes > ctx->pinned_bytes) return rtr_fail("register_many", "pinned result area too small", RTR_ERR_CAPACITY);
    RTR_CHECK(cudaMemcpyAsync(ctx->pinned, d_res, res_bytes, cudaMemcpyDeviceToHost, ctx->stream), "result");
    RTR_CHECK(cudaMemcpyAsync((char*)ctx->pinned + res_bytes, d_prev, prev_bytes, cudaMemcpyDeviceToHost, ctx->stream), "result");
    RTR_MARK(ctx, "result.d2h");
    // multi-GPU: the one collective of the path, queued right behind the batch (rtr_comm_gather_batches)
    ctx->gathered_records = 0;
    if (ctx->gather_batches) {
        if (int e = rtr_comm_allgather_dev(ctx, d_res, n_models)) return e;
        ctx->gathered_records = n_models * (ctx->comm ? ctx->comm_world : 1);
    }
    return 0;
}

// set: n_models + 1 members, the scan last.  Queues everything on the context's stream.
static int register_many_enqueue(rtr_cloud* set, int n_models, const rtr_register_params* p) {
    RtrRange nvtx_range("rtr.register_many");
    rtr_context* ctx = set->ctx;
    const int nseg = n_models + 1;
    for (int k = 0; k < nseg; ++k)
        if (set->seg_begin[k + 1] - set->seg_begin[k] >= 65536) return rtr_fail("register_many", "member clouds of a model set hold fewer than 65536 points", RTR_ERR_INVALID);
    float cells[3] = {p->normal_radius, p->fpfh_radius, p->harris_radius};
    DevGrid* gs[3];
    if (int e = rtr_get_grids(set, cells, p->harris_radius == p->normal_radius ? 2 : 3, gs)) return e;
    if (int e = rtr_normals_dev(set, p->normal_radius)) return e;
    // Harris (response, NMS, corner lists, refinement) needs only the normals and feeds nothing but the records' corner
    // counts and previews: it runs on the context's second stream beside FPFH / matching / RANSAC / ICP and is joined before
    // the finishing kernel.  Temporaries come from the bump arena, which hands nothing out twice inside one entry point.
    int* d_idx = nullptr; float4* d_xyz = nullptr; int* d_cnt = nullptr;
    const bool fork = !ctx->profile && ctx->aux_stream;
    cudaStream_t main_stream = ctx->stream;
    if (fork) {
        RTR_CHECK(cudaEventRecord(ctx->fork_event, main_stream), "register_many.fork");
        RTR_CHECK(cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_event, 0), "register_many.fork");
        ctx->stream = ctx->aux_stream;
    }
    int eh = rtr_harris_dev(set, p->harris_radius, p->harris_threshold, p->harris_nms, p->harris_refine, &d_idx, &d_xyz, &d_cnt);
    ctx->stream = main_stream;
    if (fork) cudaEventRecord(ctx->join_event, ctx->aux_stream);
    // error paths below must not leave the second stream running behind a freed set: every return waits for the join
    JoinGuard join_guard{ctx, fork, false};
    if (eh) return eh;
    if (int e = rtr_fpfh_dev(set, p->fpfh_radius)) return e;
    return register_many_back(set, n_models, p, d_cnt, d_xyz, join_guard, nullptr);
}

// ---- prepared clouds: the reference's offline / online split (RealTimeRobot.cpp:124-165 builds every database model's
// keypoints and descriptors once; :45-104 does the per-scan work) --------------------------------------------------------
static bool prepared_with(const rtr_cloud* c, const rtr_register_params* p) {
    return c->prepared && c->normals && c->fpfh && c->kp_xyz && c->kp_count && c->prep_normal_radius == p->normal_radius &&
           c->prep_harris_radius == p->harris_radius && c->prep_harris_threshold == p->harris_threshold && c->prep_harris_nms == p->harris_nms &&
           c->prep_harris_refine == p->harris_refine && c->prep_fpfh_radius == p->fpfh_radius && c->normals_mode == 0 &&
           c->normals_radius == p->normal_radius && c->fpfh_radius == p->fpfh_radius && c->normals_version == c->prep_normals_version;
}
__global__ void k_store_corners(const int* __restrict__ cnt, const float4* __restrict__ xyz, int* __restrict__ cnt_out, float4* __restrict__ xyz_out) {
    const int c = *cnt;
    if (threadIdx.x == 0) *cnt_out = c;
    for (int i = threadIdx.x; i < min(c, RTR_KP_PREVIEW); i += blockDim.x) xyz_out[i] = xyz[i];
}
// Normals, Harris corners and FPFH rows of one cloud, kept on the cloud.  fork_harris: the corner stage runs on the context's
// second stream (the caller joins before it reads kp_count / kp_xyz).
static int cloud_prepare_dev(rtr_cloud* c, const rtr_register_params* p, bool fork_harris) {
    rtr_context* ctx = c->ctx;
    if (prepared_with(c, p)) return 0;
    if (c->nseg() > 0) return rtr_fail("prepare", "a model set cannot be prepared", RTR_ERR_INVALID);
    c->prepared = false;
    if (int e = rtr_ensure_bbox(c)) return e;
    if (int e = rtr_normals_dev(c, p->normal_radius)) return e;
    if (!c->kp_xyz) if (int e = dev_alloc(ctx, &c->kp_xyz, RTR_KP_PREVIEW, "prepare")) return e;
    if (!c->kp_count) if (int e = dev_alloc(ctx, &c->kp_count, 1, "prepare")) return e;
    cudaStream_t main_stream = ctx->stream;
    if (fork_harris) {
        RTR_CHECK(cudaEventRecord(ctx->fork_event, main_stream), "prepare.fork");
        RTR_CHECK(cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_event, 0), "prepare.fork");
        ctx->stream = ctx->aux_stream;
    }
    int* d_idx = nullptr; float4* d_xyz = nullptr; int* d_cnt = nullptr;
    int eh = rtr_harris_dev(c, p->harris_radius, p->harris_threshold, p->harris_nms, p->harris_refine, &d_idx, &d_xyz, &d_cnt);
    if (!eh) {
        k_store_corners<<<1, 64, 0, ctx->stream>>>(d_cnt, d_xyz, c->kp_count, c->kp_xyz);
        ctx->launches++;
        if (ctx->profile) rtr_prof_mark(ctx, "prepare.corners");
        if (cudaGetLastError() != cudaSuccess) eh = RTR_ERR_INVALID;
    }
    ctx->stream = main_stream;
    if (fork_harris) cudaEventRecord(ctx->join_event, ctx->aux_stream);
    if (eh) { if (fork_harris) cudaStreamWaitEvent(main_stream, ctx->join_event, 0); return eh; }
    if (int e = rtr_fpfh_dev(c, p->fpfh_radius)) { if (fork_harris) cudaStreamWaitEvent(main_stream, ctx->join_event, 0); return e; }
    c->prepared = true;
    c->prep_normal_radius = p->normal_radius; c->prep_harris_radius = p->harris_radius; c->prep_harris_threshold = p->harris_threshold;
    c->prep_harris_nms = p->harris_nms; c->prep_harris_refine = p->harris_refine; c->prep_fpfh_radius = p->fpfh_radius;
    c->prep_normals_version = c->normals_version;        // normals recomputed later (another radius or mode) end the preparation
    return 0;
}
struct PrepTable {
    int nseg; int begin[RTR_MAX_SEGMENTS + 1];
    const float* fpfh[RTR_MAX_SEGMENTS]; const float4* kp_xyz[RTR_MAX_SEGMENTS]; const int* kp_cnt[RTR_MAX_SEGMENTS];
};
// descriptor rows of all members, concatenated in member order
__global__ void k_concat_rows(const __grid_constant__ PrepTable t, float* __restrict__ fpfh) {
    const long long total = (long long)t.begin[RTR_MAX_SEGMENTS] * 33;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / 33);
        int k = 0;
#pragma unroll
        for (int step = RTR_MAX_SEGMENTS / 2; step > 0; step >>= 1) if (i >= t.begin[k + step]) k += step;
        fpfh[e] = __ldg(t.fpfh[k] + (e - (long long)t.begin[k] * 33));
    }
}
// corner counts and previews of all members in the layout k_register_finish_many reads (member k's list at its first point)
__global__ void k_gather_corners(const __grid_constant__ PrepTable t, int* __restrict__ cnt, float4* __restrict__ xyz_all) {
    const int k = blockIdx.x;
    const int c = *t.kp_cnt[k];
    if (threadIdx.x == 0) cnt[k] = c;
    for (int i = threadIdx.x; i < min(c, RTR_KP_PREVIEW); i += blockDim.x) xyz_all[t.begin[k] + i] = t.kp_xyz[k][i];
}
// The online phase: every model is prepared (same stage parameters); the scan's stages run here, once; then the second half of
// a batch.  set: the points of models + scan (rtr_model_set_from_clouds); it receives the concatenated descriptor rows.
static int register_prepared_enqueue(rtr_cloud* set, rtr_cloud* const* members, int n_models, const rtr_register_params* p) {
    RtrRange nvtx_range("rtr.register_prepared");
    rtr_context* ctx = set->ctx;
    const int nseg = n_models + 1;
    rtr_cloud* scan = members[n_models];
    const bool fork = !ctx->profile && ctx->aux_stream && !prepared_with(scan, p);
    int es = cloud_prepare_dev(scan, p, fork);
    JoinGuard join_guard{ctx, fork, false};
    if (es) return es;
    PrepTable t;
    memset(&t, 0, sizeof(t));
    t.nseg = nseg;
    for (int k = 0; k <= RTR_MAX_SEGMENTS; ++k) t.begin[k] = k < nseg ? set->seg_begin[k] : set->n;
    for (int k = 0; k < nseg; ++k) { t.fpfh[k] = members[k]->fpfh; t.kp_xyz[k] = members[k]->kp_xyz; t.kp_cnt[k] = members[k]->kp_count; }
    if (int e = dev_alloc(ctx, &set->fpfh, (size_t)set->n * 33, "register_prepared")) return e;
    set->fpfh_radius = p->fpfh_radius;
    if (set->n > 0) {
        k_concat_rows<<<std::max(1, std::min(nblk((long long)set->n * 33, 256), 8 * ctx->sm_count)), 256, 0, ctx->stream>>>(t, set->fpfh);
        RTR_LAUNCH_CHECK(ctx, "set.concat_rows");
    }
    int* d_cnt = nullptr; float4* d_xyz = nullptr;
    if (int e = tmp_alloc(ctx, &d_cnt, nseg, "register_prepared")) return e;
    if (int e = tmp_alloc(ctx, &d_xyz, std::max(set->n, 1), "register_prepared")) return e;
    // the corners are only read by the finishing kernel: gathered on the second stream behind the scan's corner stage when forked
    if (fork) {
        k_gather_corners<<<nseg, 64, 0, ctx->aux_stream>>>(t, d_cnt, d_xyz);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) return RTR_ERR_INVALID;
        cudaEventRecord(ctx->join_event, ctx->aux_stream);
    } else {
        k_gather_corners<<<nseg, 64, 0, ctx->stream>>>(t, d_cnt, d_xyz);
        RTR_LAUNCH_CHECK(ctx, "set.corners");
    }
    return register_many_back(set, n_models, p, d_cnt, d_xyz, join_guard, scan);
}

static bool many_shape_ok(const int* ns, int n_models, int n_scene, const rtr_register_params* p) {
    if (n_models < 1 || n_models > RTR_MAX_SEGMENTS - 1 || n_scene >= 65536) return false;
    for (int k = 0; k < n_models; ++k) if (ns[k] >= 65536) return false;
    const long long H = (p->ransac.hypothesis_end > 0 ? p->ransac.hypothesis_end : p->ransac.max_iterations) - p->ransac.hypothesis_begin;
    return H <= (1 << 20) && (long long)n_models * std::max(H, 0LL) <= (1LL << 24);
}

static int register_many_begin_set(rtr_context* ctx, rtr_cloud* set, int n_models, const rtr_register_params* p) {
    int e;
    {
        TmpScope tmp_scope(ctx);
        e = register_many_enqueue(set, n_models, p);
    }
    if (e) { cudaStreamSynchronize(ctx->stream); rtr_cloud_free(set); return e; }
    ctx->pending_cloud[0] = set;
    ctx->register_pending = n_models;
    ctx->many_models = n_models;
    return 0;
}

extern "C" {

int rtr_register_many_begin(rtr_cloud* const* models, int n_models, rtr_cloud* scene, const rtr_register_params* p) {
    if (!models || !scene || !p || n_models < 1) return rtr_fail("register_many", "bad argument", RTR_ERR_INVALID);
    if (int e = rtr_validate_register_params(p)) return e;
    rtr_context* ctx = scene->ctx;
    if (ctx->register_pending) return rtr_fail("register_many", "a registration is already in flight on this context", RTR_ERR_INVALID);
    if (n_models > RTR_MAX_SEGMENTS - 1) return rtr_fail("register_many", "at most 31 models per batch in the asynchronous form", RTR_ERR_INVALID);
    int ns[RTR_MAX_SEGMENTS];
    rtr_cloud* members[RTR_MAX_SEGMENTS];
    for (int k = 0; k < n_models; ++k) {
        if (!models[k] || models[k]->ctx != ctx) return rtr_fail("register_many", "models and scene must belong to one context", RTR_ERR_INVALID);
        ns[k] = models[k]->n; members[k] = models[k];
    }
    members[n_models] = scene;
    if (!many_shape_ok(ns, n_models, scene->n, p)) return rtr_fail("register_many", "batch shape not supported asynchronously (clouds >= 65536 points or > 2^20 hypotheses): use rtr_register_many", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(ctx->device), "register_many");
    rtr_cloud* set = nullptr;
    if (int e = rtr_model_set_from_clouds(ctx, members, n_models + 1, &set)) return e;
    return register_many_begin_set(ctx, set, n_models, p);
}

int rtr_register_many_host_begin(rtr_context* ctx, const float* const* host_models_xyz1, const int* n_points, int n_models,
                                 const float* host_scene_xyz1, int n_scene, const rtr_register_params* p) {
    if (!ctx || !host_models_xyz1 || !n_points || !p || n_models < 1 || n_scene < 0 || (n_scene > 0 && !host_scene_xyz1))
        return rtr_fail("register_many", "bad argument", RTR_ERR_INVALID);
    if (int e = rtr_validate_register_params(p)) return e;
    if (ctx->register_pending) return rtr_fail("register_many", "a registration is already in flight on this context", RTR_ERR_INVALID);
    if (n_models > RTR_MAX_SEGMENTS - 1 || !many_shape_ok(n_points, n_models, n_scene, p))
        return rtr_fail("register_many", "batch shape not supported asynchronously (> 31 models, clouds >= 65536 points or > 2^20 hypotheses): use rtr_register_many_host", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(ctx->device), "register_many");
    const float* ptrs[RTR_MAX_SEGMENTS]; int ns[RTR_MAX_SEGMENTS];
    for (int k = 0; k < n_models; ++k) { ptrs[k] = host_models_xyz1[k]; ns[k] = n_points[k]; }
    ptrs[n_models] = host_scene_xyz1; ns[n_models] = n_scene;
    rtr_cloud* set = nullptr;
    if (int e = rtr_model_set_from_host(ctx, ptrs, ns, n_models + 1, &set)) return e;
    return register_many_begin_set(ctx, set, n_models, p);
}

// ---- the offline / online split ------------------------------------------------------------------------------------------
int rtr_cloud_prepare(rtr_cloud* c, const rtr_register_params* p) {
    if (!c || !p) return rtr_fail("prepare", "bad argument", RTR_ERR_INVALID);
    if (int e = rtr_validate_register_params(p)) return e;
    rtr_context* ctx = c->ctx;
    if (ctx->register_pending) return rtr_fail("prepare", "a registration is in flight on this context", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(ctx->device), "prepare");
    TmpScope tmp_scope(ctx);
    return cloud_prepare_dev(c, p, false);
}

int rtr_register_prepared_begin(rtr_cloud* const* models, int n_models, rtr_cloud* scene, const rtr_register_params* p) {
    if (!models || !scene || !p || n_models < 1) return rtr_fail("register_prepared", "bad argument", RTR_ERR_INVALID);
    if (int e = rtr_validate_register_params(p)) return e;
    rtr_context* ctx = scene->ctx;
    if (ctx->register_pending) return rtr_fail("register_prepared", "a registration is already in flight on this context", RTR_ERR_INVALID);
    if (n_models > RTR_MAX_SEGMENTS - 1) return rtr_fail("register_prepared", "at most 31 models per batch", RTR_ERR_INVALID);
    int ns[RTR_MAX_SEGMENTS];
    rtr_cloud* members[RTR_MAX_SEGMENTS];
    for (int k = 0; k < n_models; ++k) {
        if (!models[k] || models[k]->ctx != ctx || models[k] == scene) return rtr_fail("register_prepared", "models and scene must be distinct clouds of one context", RTR_ERR_INVALID);
        if (!prepared_with(models[k], p)) return rtr_fail("register_prepared", "every model needs rtr_cloud_prepare with these stage parameters first", RTR_ERR_NOT_READY);
        ns[k] = models[k]->n; members[k] = models[k];
        if (ns[k] < 3) return rtr_fail("register_prepared", "models need at least 3 points", RTR_ERR_INVALID);
    }
    members[n_models] = scene;
    if (scene->n < 1 || !many_shape_ok(ns, n_models, scene->n, p))
        return rtr_fail("register_prepared", "batch shape not supported (empty scan, clouds >= 65536 points or > 2^20 hypotheses)", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(ctx->device), "register_prepared");
    rtr_cloud* set = nullptr;
    if (int e = rtr_model_set_from_clouds(ctx, members, n_models + 1, &set)) return e;
    int e;
    {
        TmpScope tmp_scope(ctx);
        e = register_prepared_enqueue(set, members, n_models, p);
    }
    if (e) { cudaStreamSynchronize(ctx->stream); rtr_cloud_free(set); return e; }
    ctx->pending_cloud[0] = set;
    ctx->register_pending = n_models;
    ctx->many_models = n_models;
    return 0;
}

int rtr_register_prepared(rtr_cloud* const* models, int n_models, rtr_cloud* scene, const rtr_register_params* p, rtr_pose_result* host_results) {
    if (!models || !scene || !p || !host_results || n_models < 1) return rtr_fail("register_prepared", "bad argument", RTR_ERR_INVALID);
    for (int m0 = 0; m0 < n_models; m0 += RTR_MAX_SEGMENTS - 1) {
        const int nb = std::min(n_models - m0, RTR_MAX_SEGMENTS - 1);
        if (int e = rtr_register_prepared_begin(models + m0, nb, scene, p)) return e;
        if (int e = rtr_register_many_end(scene->ctx, host_results + m0, nb)) return e;
        for (int k = 0; k < nb; ++k) host_results[m0 + k].model_id = m0 + k;
    }
    return 0;
}

int rtr_register_many_end(rtr_context* ctx, rtr_pose_result* host_results, int capacity) {
    if (!ctx || !host_results) return rtr_fail("register_many", "bad argument", RTR_ERR_INVALID);
    if (!ctx->register_pending || ctx->many_models <= 0) return rtr_fail("register_many", "no batch in flight on this context", RTR_ERR_INVALID);
    const int n = ctx->many_models;
    if (capacity < n) return rtr_fail("register_many", "result buffer too small", RTR_ERR_CAPACITY);
    ctx->register_pending = 0; ctx->many_models = 0;
    cudaError_t err = cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; ++i) if (ctx->pending_cloud[i]) { rtr_cloud_free(ctx->pending_cloud[i]); ctx->pending_cloud[i] = nullptr; }
    RTR_CHECK(err, "result");
    memcpy(host_results, ctx->pinned, sizeof(rtr_pose_result) * (size_t)n);
    // keep the corner previews for rtr_register_many_keypoints
    ctx->kp_preview.assign((const char*)ctx->pinned + sizeof(rtr_pose_result) * (size_t)n,
                           (const char*)ctx->pinned + sizeof(rtr_pose_result) * (size_t)n + sizeof(KpPreview) * (size_t)(n + 1));
    ctx->kp_members = n + 1;
    return 0;
}

int rtr_register_many_keypoints(rtr_context* ctx, int member, float* host_kp_xyz1, int capacity, int* n_keypoints) {
    if (!ctx || !n_keypoints || capacity < 0 || (capacity > 0 && !host_kp_xyz1)) return rtr_fail("register_many", "bad argument", RTR_ERR_INVALID);
    if (member < 0 || member >= ctx->kp_members) return rtr_fail("register_many", "no such member in the last batch", RTR_ERR_INVALID);
    const KpPreview* pv = reinterpret_cast<const KpPreview*>(ctx->kp_preview.data()) + member;
    *n_keypoints = pv->count;
    const int take = std::min(std::min(pv->count, RTR_KP_PREVIEW), capacity);
    if (take > 0) memcpy(host_kp_xyz1, pv->xyz, (size_t)take * 16);
    return (pv->count > take) ? RTR_ERR_CAPACITY : 0;
}

int rtr_register_many(rtr_cloud* const* models, int n_models, rtr_cloud* scene, const rtr_register_params* p, rtr_pose_result* host_results) {
    if (!models || !scene || !p || !host_results || n_models < 1) return rtr_fail("register_many", "bad argument", RTR_ERR_INVALID);
    if (int e = rtr_validate_register_params(p)) return e;
    // batches of up to 31 models; shapes the set path does not take (huge clouds, huge sweeps) go one model at a time
    for (int m0 = 0; m0 < n_models; m0 += RTR_MAX_SEGMENTS - 1) {
        const int nb = std::min(n_models - m0, RTR_MAX_SEGMENTS - 1);
        int ns[RTR_MAX_SEGMENTS];
        bool ok = true;
        for (int k = 0; k < nb; ++k) { if (!models[m0 + k]) return rtr_fail("register_many", "null model", RTR_ERR_INVALID); ns[k] = models[m0 + k]->n; ok &= ns[k] >= 3; }
        ok = ok && scene->n >= 1 && many_shape_ok(ns, nb, scene->n, p);
        if (ok) {
            if (int e = rtr_register_many_begin(models + m0, nb, scene, p)) return e;
            if (int e = rtr_register_many_end(scene->ctx, host_results + m0, nb)) return e;
        } else {
            for (int k = 0; k < nb; ++k) {
                if (int e = rtr_register(models[m0 + k], scene, p, host_results + m0 + k)) return e;
                host_results[m0 + k].model_id = k;
            }
        }
        for (int k = 0; k < nb; ++k) host_results[m0 + k].model_id = m0 + k;
    }
    return 0;
}

int rtr_register_many_host(rtr_context* ctx, const float* const* host_models_xyz1, const int* n_points, int n_models,
                           const float* host_scene_xyz1, int n_scene, const rtr_register_params* p, rtr_pose_result* host_results) {
    if (!ctx || !host_models_xyz1 || !n_points || !p || !host_results || n_models < 1) return rtr_fail("register_many", "bad argument", RTR_ERR_INVALID);
    if (int e = rtr_validate_register_params(p)) return e;
    for (int m0 = 0; m0 < n_models; m0 += RTR_MAX_SEGMENTS - 1) {
        const int nb = std::min(n_models - m0, RTR_MAX_SEGMENTS - 1);
        bool ok = n_scene >= 1 && many_shape_ok(n_points + m0, nb, n_scene, p);
        for (int k = 0; k < nb; ++k) ok &= n_points[m0 + k] >= 3;
        if (ok) {
            if (int e = rtr_register_many_host_begin(ctx, host_models_xyz1 + m0, n_points + m0, nb, host_scene_xyz1, n_scene, p)) return e;
            if (int e = rtr_register_many_end(ctx, host_results + m0, nb)) return e;
        } else {
            for (int k = 0; k < nb; ++k)
                if (int e = rtr_register_host(ctx, host_models_xyz1[m0 + k], n_points[m0 + k], host_scene_xyz1, n_scene, p, host_results + m0 + k)) return e;
        }
        for (int k = 0; k < nb; ++k) host_results[m0 + k].model_id = m0 + k;
    }
    return 0;
}

int rtr_ransac_prerejective(rtr_cloud* source, rtr_cloud* target, const rtr_ransac_params* p, rtr_pose_result* host_result) {
    if (!source || !target || !p || !host_result || source->ctx != target->ctx) return rtr_fail("ransac", "bad argument", RTR_ERR_INVALID);
    rtr_context* ctx = source->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "ransac");
    rtr_pose_result* d_res = nullptr;
    if (int e = tmp_alloc(ctx, &d_res, 1, "ransac")) return e;
    if (int e = rtr_ransac_dev(source, target, p, d_res)) return e;
    int rc = fetch_result(ctx, d_res, host_result);
    dev_free(ctx, d_res);
    return rc;
}

int rtr_icp(rtr_cloud* source, rtr_cloud* target, const rtr_icp_params* p, const float* init_pose16, rtr_pose_result* host_result) {
    if (!source || !target || !p || !host_result || source->ctx != target->ctx || p->max_iterations < 0) return rtr_fail("icp", "bad argument", RTR_ERR_INVALID);
    rtr_context* ctx = source->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "icp");
    rtr_pose_result* d_res = nullptr; float* d_init = nullptr;
    if (int e = tmp_alloc(ctx, &d_res, 1, "icp")) return e;
    if (init_pose16) {
        if (int e = tmp_alloc(ctx, &d_init, 16, "icp")) return e;
        RTR_CHECK(cudaMemcpyAsync(d_init, init_pose16, 64, cudaMemcpyHostToDevice, ctx->stream), "icp");
    }
    if (int e = rtr_icp_dev(source, target, p, d_init, 0, d_res)) return e;
    int rc = fetch_result(ctx, d_res, host_result);
    dev_free(ctx, d_res); dev_free(ctx, d_init);
    return rc;
}

// Enqueue a whole registration on the context's stream and return: no host synchronisation.  The 128-byte record is copied
// into the context's pinned result slot by the stream itself; rtr_register_end waits for it.  One registration may be in
// flight per context; a host thread can keep many contexts (streams) busy this way without one thread per stream.
// want_fork: the synchronous entry points (one registration, latency matters: 0.66 -> 0.58 ms for chair1) fork the scene's
// stages onto the second stream; rtr_register_begin (many registrations in flight on many contexts, throughput matters:
// 16 streams instead of 8 cost 6 % of the 8-registration step) does not.  RTR_REGISTER_FORK=0 / 1 forces either.
static int register_begin_impl(rtr_cloud* model, rtr_cloud* scene, const rtr_register_params* p, bool want_fork) {
    RtrRange nvtx_range("rtr.register");
    if (!model || !scene || !p || model->ctx != scene->ctx) return rtr_fail("register", "bad argument", RTR_ERR_INVALID);
    if (int e = rtr_validate_register_params(p)) return e;
    rtr_context* ctx = model->ctx;
    if (ctx->register_pending) return rtr_fail("register", "a registration is already in flight on this context", RTR_ERR_INVALID);
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "register");
    rtr_cloud* cl[2] = {model, scene};
    int* d_cnt[2] = {nullptr, nullptr};
    // The model's and the scene's stages are independent until the descriptor matching: the scene's are enqueued on the
    // context's second stream (fork after whatever the caller queued, e.g. the uploads; join before matching).  Every
    // stage function launches on ctx->stream, so the stream is swapped while the scene's stages are queued.  Temporaries
    // come from the bump arena, which hands nothing out twice inside one entry point, and the next entry point starts on
    // `stream` after the join: no buffer is shared by the two streams.  Profiling keeps one stream (its marks time a
    // single queue).
    static const int fork_env = []() { const char* e = getenv("RTR_REGISTER_FORK"); return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1; }();
    const bool fork = (fork_env < 0 ? want_fork : fork_env == 1) && !ctx->profile && ctx->aux_stream;
    cudaStream_t main_stream = ctx->stream;
    if (fork) {
        RTR_CHECK(cudaEventRecord(ctx->fork_event, main_stream), "register.fork");
        RTR_CHECK(cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_event, 0), "register.fork");
    }
    for (int i = 0; i < 2; ++i) {
        if (fork && i == 1) ctx->stream = ctx->aux_stream;
        int e = rtr_normals_dev(cl[i], p->normal_radius);
        int* d_idx = nullptr; float4* d_xyz = nullptr;
        if (!e) e = rtr_harris_dev(cl[i], p->harris_radius, p->harris_threshold, p->harris_nms, p->harris_refine, &d_idx, &d_xyz, &d_cnt[i]);
        if (!e) { dev_free(ctx, d_idx); dev_free(ctx, d_xyz); e = rtr_fpfh_dev(cl[i], p->fpfh_radius); }
        ctx->stream = main_stream;
        if (e) {
            if (fork) { cudaEventRecord(ctx->join_event, ctx->aux_stream); cudaStreamWaitEvent(main_stream, ctx->join_event, 0); }
            return e;
        }
    }
    if (fork) {
        RTR_CHECK(cudaEventRecord(ctx->join_event, ctx->aux_stream), "register.join");
        RTR_CHECK(cudaStreamWaitEvent(main_stream, ctx->join_event, 0), "register.join");
    }
    if (int e = rtr_match_dev(model, scene, p->ransac.correspondence_k)) return e;
    rtr_pose_result* d_res = nullptr;
    if (int e = tmp_alloc(ctx, &d_res, 1, "register")) return e;
    if (int e = rtr_ransac_dev(model, scene, &p->ransac, d_res)) return e;
    if (p->run_icp) if (int e = rtr_icp_dev(model, scene, &p->icp, nullptr, 1, d_res)) return e;
    k_set_keypoints<<<1, 32, 0, ctx->stream>>>(d_res, d_cnt[0], d_cnt[1]);
    RTR_LAUNCH_CHECK(ctx, "register.kp");
    RTR_CHECK(cudaMemcpyAsync(ctx->pinned, d_res, sizeof(rtr_pose_result), cudaMemcpyDeviceToHost, ctx->stream), "result");
    RTR_MARK(ctx, "result.d2h");
    dev_free(ctx, d_res); dev_free(ctx, d_cnt[0]); dev_free(ctx, d_cnt[1]);
    ctx->register_pending = 1;
    return 0;
}

int rtr_register_begin(rtr_cloud* model, rtr_cloud* scene, const rtr_register_params* p) { return register_begin_impl(model, scene, p, false); }

int rtr_register_end(rtr_context* ctx, rtr_pose_result* host_result) {
    if (!ctx || !host_result) return rtr_fail("register", "bad argument", RTR_ERR_INVALID);
    if (!ctx->register_pending) return rtr_fail("register", "no registration in flight on this context", RTR_ERR_INVALID);
    if (ctx->many_models > 0) return rtr_fail("register", "the registration in flight is a batch: call rtr_register_many_end", RTR_ERR_INVALID);
    ctx->register_pending = 0;
    cudaError_t err = cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; ++i) if (ctx->pending_cloud[i]) { rtr_cloud_free(ctx->pending_cloud[i]); ctx->pending_cloud[i] = nullptr; }
    RTR_CHECK(err, "result");
    memcpy(host_result, ctx->pinned, sizeof(rtr_pose_result));
    return 0;
}

int rtr_register(rtr_cloud* model, rtr_cloud* scene, const rtr_register_params* p, rtr_pose_result* host_result) {
    if (!host_result) return rtr_fail("register", "bad argument", RTR_ERR_INVALID);
    if (int e = register_begin_impl(model, scene, p, true)) return e;
    return rtr_register_end(model->ctx, host_result);
}

// host clouds in (pinned memory makes the two uploads asynchronous), record out; the device clouds live until _end
static int register_host_begin_impl(rtr_context* ctx, const float* host_model_xyz1, int n_model, const float* host_scene_xyz1, int n_scene,
                                    const rtr_register_params* p, bool want_fork) {
    if (!ctx) return rtr_fail("register", "bad argument", RTR_ERR_INVALID);
    if (ctx->register_pending) return rtr_fail("register", "a registration is already in flight on this context", RTR_ERR_INVALID);
    rtr_cloud *m = nullptr, *s = nullptr;
    if (int e = rtr_cloud_upload(ctx, host_model_xyz1, n_model, &m)) return e;
    if (int e = rtr_cloud_upload(ctx, host_scene_xyz1, n_scene, &s)) { rtr_cloud_free(m); return e; }
    if (int e = register_begin_impl(m, s, p, want_fork)) { rtr_cloud_free(m); rtr_cloud_free(s); return e; }
    ctx->pending_cloud[0] = m; ctx->pending_cloud[1] = s;
    return 0;
}
int rtr_register_host_begin(rtr_context* ctx, const float* host_model_xyz1, int n_model, const float* host_scene_xyz1, int n_scene,
                            const rtr_register_params* p) {
    return register_host_begin_impl(ctx, host_model_xyz1, n_model, host_scene_xyz1, n_scene, p, false);
}

int rtr_register_host(rtr_context* ctx, const float* host_model_xyz1, int n_model, const float* host_scene_xyz1, int n_scene,
                      const rtr_register_params* p, rtr_pose_result* host_result) {
    if (!host_result) return rtr_fail("register", "bad argument", RTR_ERR_INVALID);
    if (int e = register_host_begin_impl(ctx, host_model_xyz1, n_model, host_scene_xyz1, n_scene, p, true)) return e;
    return rtr_register_end(ctx, host_result);
}

}  // extern "C"
