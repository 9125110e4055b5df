// bvh.cuh — exact 1-NN over a SMALL target cloud (a scan of a few thousand points) for one WARP per query.
//
// Why: ICP's correspondence step against the reference's scans (1.8 k - 3.2 k points, RealTimeRobot.cpp:34-35) ran as a
// warp-cooperative brute-force scan — every query tests every target point, ~0.7 warp instructions per point pair, 36.6 M
// warp instructions per iteration for the bench batch (26 k active queries x 1909 targets; ncu, profiles/r02) — and was
// instruction-bound.  A uniform grid does not help there: most of a model's points have no scan point nearby and the ring
// walks of far queries cost as much as the scan (measured).  A binary BVH with one THREAD per query does not help either:
// far queries visit hundreds of nodes and 32 divergent traversals per warp cost more issue slots than the scan (measured:
// 116 us per iteration against 60).
//
// What does: a TWO-LEVEL, 32-WIDE hierarchy for the warp.  The points are sorted by 30-bit Morton code and cut into leaves
// of 32 consecutive points (a compact patch each) with an axis-aligned box per leaf — at most 128 leaves.  A warp tests all
// leaf boxes of a query at once (one box per lane and round), then scans only the leaves whose box can beat or tie the best
// distance so far, 32 points per step (one per lane), tightening the bound after every leaf.  A converging ICP starts from
// the query's previous neighbour and typically scans 1-3 leaves instead of all 60.
//
// Exactness: the lower bound of a box is built from the same float subtract / multiply / add sequence as dist2f, which is
// monotone under round-to-nearest, so bound <= dist2f(q, p) for every p in the box; a leaf is skipped only when its bound is
// STRICTLY above the best distance, and candidates compare as (d2 bits << 32 | original index) keys like the brute-force
// scan: the result is the scan's, ties -> lowest index.
#pragma once
#include "common.cuh"

#define WBVH_LEAF 32
#define WBVH_MAX_POINTS 4096           // == RTR_BRUTE_NN_MAX: the clouds this path serves (<= 128 leaves)
#define WBVH_ROUNDS (WBVH_MAX_POINTS / WBVH_LEAF / 32)      // leaf boxes per lane: 4
#define WBVH_BUILD_THREADS 1024

struct WbvhView {
    const float4* __restrict__ boxes;  // 2 float4 per leaf: minimum, maximum
    const float4* __restrict__ pts;    // n points in Morton order, .w = original index (int bits)
    int nleaf;
    int n;
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned wbvh_expand10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ float wbvh_box_d2(const float4 lo, const float4 hi, float qx, float qy, float qz) {
    const float ex = qx > hi.x ? __fsub_rn(qx, hi.x) : (qx < lo.x ? __fsub_rn(qx, lo.x) : 0.f);
    const float ey = qy > hi.y ? __fsub_rn(qy, hi.y) : (qy < lo.y ? __fsub_rn(qy, lo.y) : 0.f);
    const float ez = qz > hi.z ? __fsub_rn(qz, hi.z) : (qz < lo.z ? __fsub_rn(qz, lo.z) : 0.f);
    return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}

// One CTA builds the whole structure.  pts: the target's points by local index 0 .. n-1; idx_base is added to the index
// stored in .w (a member of a model set reports set-wide indices).  mn / mx: the cloud's bounding box (Morton quantisation
// only: any box gives a correct structure, a tight one compact leaves).
__global__ void __launch_bounds__(WBVH_BUILD_THREADS) k_wbvh_build(const float4* __restrict__ pts, int n, int idx_base, float mnx, float mny, float mnz,
                                                                   float mxx, float mxy, float mxz, float4* __restrict__ boxes,
                                                                   float4* __restrict__ mpts) {
    __shared__ unsigned long long keys[WBVH_MAX_POINTS];
    const int tid = threadIdx.x;
    int np = 1;
    while (np < n) np <<= 1;
    const float sx = mxx > mnx ? 1023.0f / (mxx - mnx) : 0.f, sy = mxy > mny ? 1023.0f / (mxy - mny) : 0.f, sz = mxz > mnz ? 1023.0f / (mxz - mnz) : 0.f;
    for (int i = tid; i < np; i += WBVH_BUILD_THREADS) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const float4 p = __ldg(pts + i);
            float fx = (p.x - mnx) * sx, fy = (p.y - mny) * sy, fz = (p.z - mnz) * sz;
            fx = fx >= 0.f ? fminf(fx, 1023.f) : 0.f;          // NaN -> 0
            fy = fy >= 0.f ? fminf(fy, 1023.f) : 0.f;
            fz = fz >= 0.f ? fminf(fz, 1023.f) : 0.f;
            const unsigned m = (wbvh_expand10((unsigned)fx) << 2) | (wbvh_expand10((unsigned)fy) << 1) | wbvh_expand10((unsigned)fz);
            k = ((unsigned long long)m << 32) | (unsigned)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np; i += WBVH_BUILD_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    // one warp per leaf: its 32 points (one per lane) and the box over the finite ones
    const int warp = tid >> 5, lane = tid & 31, nleaf = (n + WBVH_LEAF - 1) / WBVH_LEAF;
    const float inf = __int_as_float(0x7f800000);
    for (int l = warp; l < nleaf; l += WBVH_BUILD_THREADS / 32) {
        const int s = l * WBVH_LEAF + lane;
        float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
        if (s < n) {
            const int i = (int)(unsigned)keys[s];
            float4 p = __ldg(pts + i);
            p.w = __int_as_float(idx_base + i);
            mpts[s] = p;
            lx = hx = p.x; ly = hy = p.y; lz = hz = p.z;          // fminf / fmaxf below ignore NaN coordinates
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
            hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
        }
        if (lane == 0) { boxes[2 * l] = make_float4(lx, ly, lz, 0.f); boxes[2 * l + 1] = make_float4(hx, hy, hz, 0.f); }
    }
}

// Exact nearest neighbour of q for a whole warp (every lane passes the same q and receives the same answer), ties -> lowest
// original index.  key: the best candidate so far as (d2 bits << 32 | index), ~0ull for none — a warm start is any target
// point with its dist2f to q.
__device__ __forceinline__ unsigned long long wbvh_nearest_warp(const WbvhView& B, float qx, float qy, float qz, int lane, unsigned long long key) {
    float bnd[WBVH_ROUNDS];
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int r = 0; r < WBVH_ROUNDS; ++r) {
        const int l = lane + 32 * r;
        bnd[r] = inf;
        if (l < B.nleaf) bnd[r] = wbvh_box_d2(__ldg(B.boxes + 2 * l), __ldg(B.boxes + 2 * l + 1), qx, qy, qz);
    }
    // best so far as two words: d2 bits (their unsigned order is the order of the non-negative distances, NaN bits above
    // +inf) and the original index; the warp-wide minimum of a leaf is two redux.sync instructions (min of the distance
    // bits, then min of the indices among the lanes that hold it) instead of a five-step 64-bit shuffle butterfly
    unsigned kd = (unsigned)(key >> 32), ki = (unsigned)key;
    auto scan_leaf = [&](int l) {
        const int s = l * WBVH_LEAF + lane;
        unsigned db = 0xffffffffu, id = 0xffffffffu;
        if (s < B.n) {
            const float4 p = __ldg(B.pts + s);
            db = __float_as_uint(dist2f(qx, qy, qz, p.x, p.y, p.z));
            id = (unsigned)__float_as_int(p.w);
        }
        const unsigned mdb = __reduce_min_sync(0xffffffffu, db);
        if (mdb <= kd) {                       // warp-uniform
            const unsigned mid = __reduce_min_sync(0xffffffffu, db == mdb ? id : 0xffffffffu);
            if (mdb < kd || mid < ki) { kd = mdb; ki = mid; }
        }
        key = ((unsigned long long)kd << 32) | ki;
    };
    if (key == ~0ull) {
        // no candidate yet: the leaf with the nearest box first (its best point is usually the answer or close to it)
        float mb = bnd[0]; int ml = lane;
#pragma unroll
        for (int r = 1; r < WBVH_ROUNDS; ++r) if (bnd[r] < mb) { mb = bnd[r]; ml = lane + 32 * r; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, mb, o);
            const int ol = __shfl_xor_sync(0xffffffffu, ml, o);
            if (ob < mb || (ob == mb && ol < ml)) { mb = ob; ml = ol; }
        }
        if (mb < inf) {
            scan_leaf(ml);
#pragma unroll
            for (int r = 0; r < WBVH_ROUNDS; ++r) if (ml == lane + 32 * r) bnd[r] = inf;       // visited
        }
    }
    // every leaf whose box can beat or tie the best distance, the bound tightening after each
#pragma unroll
    for (int r = 0; r < WBVH_ROUNDS; ++r) {
        if (32 * r >= B.nleaf) break;
        for (;;) {
            const float bd = __uint_as_float((unsigned)(key >> 32));      // NaN bits (no candidate): every comparison below is false ...
            const bool open = bnd[r] < inf && (key == ~0ull || bnd[r] <= bd);        // ... so "no candidate" opens every non-empty, unvisited leaf
            const unsigned m = __ballot_sync(0xffffffffu, open);
            if (!m) break;
            const int src = __ffs(m) - 1;
            scan_leaf(src + 32 * r);
            if (lane == src) bnd[r] = inf;
        }
    }
    return key;
}
#endif  // __CUDACC__

// ----------------------------------------------------------------------------------------------------------------------
// The same idea for a target of ANY size: a 32-ary hierarchy of boxes over the Morton-ordered points.  Level 0 holds the
// leaves (32 consecutive points each), level l + 1 one node per 32 consecutive nodes of level l, up to a top level of at
// most 32 nodes — 3 levels for 100 k points (3125 / 98 / 4 nodes), 3 for 1 M (31 250 / 977 / 31).  A warp tests the 32
// children of a node at once (one box per lane), descends into the child with the smallest lower bound first and comes
// back for every other child whose bound can still beat or tie the best distance, so the first descent already yields a
// near-optimal bound and the rest of the tree is pruned against it.
//
// Why: ICP with PCL's defaults has NO correspondence cap (function.h:111-117), so a source point metres away from the
// target still needs its exact nearest target point.  On a uniform grid that query walks every row of cells inside its
// search sphere — 1 M scan points against a 100 k-point model cost 15 ms per iteration (14.8 ms of it in the far queries'
// ring walks; bench.py icp_1m.scan_to_model_uncapped, round 2) — while here its cost is a handful of 32-wide box tests.
// Exactness as above: bounds from the same monotone float sequence as dist2f, children opened while bound <= best,
// candidates ordered by (d2 bits, original index).
#define WIDE_MAX_LEVELS 6               // 32^6 leaves x 32 points: any int-indexed cloud
struct WideBvh {
    const float4* __restrict__ pts;     // n points in Morton order, .w = original index (int bits)
    const float4* __restrict__ boxes;   // 2 float4 per node (minimum, maximum), level 0 first
    int n;
    int depth;                          // levels in use (>= 1); level depth - 1 has <= 32 nodes
    int count[WIDE_MAX_LEVELS];         // nodes per level
    int offset[WIDE_MAX_LEVELS];        // first node of each level in boxes
};

#ifdef __CUDACC__
__global__ void k_wide_morton(const float4* __restrict__ pts, int n, float mnx, float mny, float mnz, float mxx, float mxy, float mxz,
                              unsigned* __restrict__ keys, int* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sx = mxx > mnx ? 1023.0f / (mxx - mnx) : 0.f, sy = mxy > mny ? 1023.0f / (mxy - mny) : 0.f, sz = mxz > mnz ? 1023.0f / (mxz - mnz) : 0.f;
    const float4 p = __ldg(pts + i);
    float fx = (p.x - mnx) * sx, fy = (p.y - mny) * sy, fz = (p.z - mnz) * sz;
    fx = fx >= 0.f ? fminf(fx, 1023.f) : 0.f;          // NaN -> 0
    fy = fy >= 0.f ? fminf(fy, 1023.f) : 0.f;
    fz = fz >= 0.f ? fminf(fz, 1023.f) : 0.f;
    keys[i] = (wbvh_expand10((unsigned)fx) << 2) | (wbvh_expand10((unsigned)fy) << 1) | wbvh_expand10((unsigned)fz);
    vals[i] = i;
}
// one warp per leaf: its 32 points (one per lane, Morton order) and the box over the finite ones
__global__ void k_wide_leaves(const float4* __restrict__ pts, const int* __restrict__ order, int n, int idx_base, float4* __restrict__ mpts,
                              float4* __restrict__ boxes) {
    const int lane = threadIdx.x & 31;
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (l * WBVH_LEAF >= n) return;
    const int s = l * WBVH_LEAF + lane;
    const float inf = __int_as_float(0x7f800000);
    float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
    if (s < n) {
        const int i = __ldg(order + s);
        float4 p = __ldg(pts + i);
        p.w = __int_as_float(idx_base + i);
        mpts[s] = p;
        lx = hx = p.x; ly = hy = p.y; lz = hz = p.z;              // fminf / fmaxf below ignore NaN coordinates
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
        hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
    }
    if (lane == 0) { boxes[2 * l] = make_float4(lx, ly, lz, 0.f); boxes[2 * l + 1] = make_float4(hx, hy, hz, 0.f); }
}
// one warp per node of the upper level: the union of its (up to 32) children
__global__ void k_wide_level(const float4* __restrict__ lower, int n_lower, float4* __restrict__ upper) {
    const int lane = threadIdx.x & 31;
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u * 32 >= n_lower) return;
    const int c = u * 32 + lane;
    const float inf = __int_as_float(0x7f800000);
    float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
    if (c < n_lower) {
        const float4 lo = __ldg(lower + 2 * c), hi = __ldg(lower + 2 * c + 1);
        lx = lo.x; ly = lo.y; lz = lo.z; hx = hi.x; hy = hi.y; hz = hi.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
        hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
    }
    if (lane == 0) { upper[2 * u] = make_float4(lx, ly, lz, 0.f); upper[2 * u + 1] = make_float4(hx, hy, hz, 0.f); }
}

// the 32 points of one leaf against the best candidate so far (kd: d2 bits, ki: original index)
__device__ __forceinline__ void wide_scan_leaf(const WideBvh& B, int leaf, float qx, float qy, float qz, int lane, unsigned& kd, unsigned& ki) {
    const int s = leaf * WBVH_LEAF + lane;
    unsigned db = 0xffffffffu, id = 0xffffffffu;
    if (s < B.n) {
        const float4 p = __ldg(B.pts + s);
        db = __float_as_uint(dist2f(qx, qy, qz, p.x, p.y, p.z));
        id = (unsigned)__float_as_int(p.w);
    }
    const unsigned mdb = __reduce_min_sync(0xffffffffu, db);
    if (mdb <= kd) {                           // warp-uniform
        const unsigned mid = __reduce_min_sync(0xffffffffu, db == mdb ? id : 0xffffffffu);
        if (mdb < kd || mid < ki) { kd = mdb; ki = mid; }
    }
}
// children [node * 32, node * 32 + 32) of level L, nearest first
template <int L>
__device__ __forceinline__ void wide_visit(const WideBvh& B, int node, float qx, float qy, float qz, int lane, unsigned& kd, unsigned& ki) {
    const int c = node * 32 + lane;
    const float inf = __int_as_float(0x7f800000);
    float b = inf;
    if (c < B.count[L]) {
        const float4* bx = B.boxes + 2 * (size_t)(B.offset[L] + c);
        b = wbvh_box_d2(__ldg(bx), __ldg(bx + 1), qx, qy, qz);
    }
    for (;;) {
        // kd holds the bits of a non-negative float or of +inf ("anything"): unsigned order == float order; an empty or
        // visited child has b == +inf and stays closed even against kd == +inf
        const unsigned ub = (b < inf && __float_as_uint(b) <= kd) ? __float_as_uint(b) : 0xffffffffu;
        const unsigned m = __reduce_min_sync(0xffffffffu, ub);
        if (m == 0xffffffffu) break;
        const int src = __ffs(__ballot_sync(0xffffffffu, ub == m)) - 1;
        if (lane == src) b = inf;
        if constexpr (L == 0) wide_scan_leaf(B, node * 32 + src, qx, qy, qz, lane, kd, ki);
        else wide_visit<L - 1>(B, node * 32 + src, qx, qy, qz, lane, kd, ki);
    }
}
// Exact nearest neighbour of q for a whole warp (every lane passes the same q and key and receives the same answer).
// key: (d2 bits << 32 | index) of the best candidate so far — a warm start, or (bits of the search radius^2 << 32 |
// 0xffffffff) for "anything at or inside the radius", radius +inf for an unbounded search.  An index of 0xffffffff in the
// result means nothing was found.
__device__ __forceinline__ unsigned long long wide_nearest_warp(const WideBvh& B, float qx, float qy, float qz, int lane, unsigned long long key) {
    unsigned kd = (unsigned)(key >> 32), ki = (unsigned)key;
    switch (B.depth) {
        case 1: wide_visit<0>(B, 0, qx, qy, qz, lane, kd, ki); break;
        case 2: wide_visit<1>(B, 0, qx, qy, qz, lane, kd, ki); break;
        case 3: wide_visit<2>(B, 0, qx, qy, qz, lane, kd, ki); break;
        case 4: wide_visit<3>(B, 0, qx, qy, qz, lane, kd, ki); break;
        case 5: wide_visit<4>(B, 0, qx, qy, qz, lane, kd, ki); break;
        case 6: wide_visit<5>(B, 0, qx, qy, qz, lane, kd, ki); break;
        default: break;
    }
    return ((unsigned long long)kd << 32) | ki;
}
#endif  // __CUDACC__
