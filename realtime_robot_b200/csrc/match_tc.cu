// match_tc.cu — descriptor correspondence search on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// The step is a dense contraction: d(i,j) = |a_i|^2 + |b_j|^2 - 2 a_i.b_j over 33-D FPFH signatures (KdTreeFLANN<
// FPFHSignature33>::nearestKSearch inside SampleConsensusPrerejective::findSimilarFeatures).  The
// result must be EXACT (same indices as the fp64 oracle), so the tensor cores are used as a PREFILTER with a certificate:
//
//   1. k_tc_prep   : each fp32 feature value is split into tf32 hi + tf32 lo; operand rows are
//                    A = [hi | hi | lo | 0], B = [hi | lo | hi | 0] (K = 104), so one TF32 GEMM yields
//                    hi.hi + hi.lo + lo.hi = a.b to ~2^-22 relative.  Rows are written in the UMMA canonical K-major
//                    no-swizzle layout (8-row x 16-byte core matrices), 128-row tiles, ready for 1-D bulk copies.
//   2. k_tc_match  : one CTA (18 warps) per (128 source rows, split of the target tiles).  Warp 0 streams 128-row target
//                    tiles into a 2-stage shared-memory ring with cp.async.bulk + mbarrier (optionally as a cluster of
//                    two CTAs that each load half a tile and multicast it to both); warp 1 issues tcgen05.mma (M = N =
//                    128, kind::tf32, 13 k-steps) into a 4-deep ring of TMEM accumulators and commits twice: "accumulator
//                    ready" for the epilogue and "smem stage free" for the loader, so loads never wait for the epilogue;
//                    warps 2-17 are the epilogue: TMEM lane quarter = warp % 4, and each of the four groups scans one
//                    32-column chunk of every accumulator with tcgen05.ld.  The target norm rides in three spare K
//                    columns, so the accumulator is a.b - |b|^2/2 and "closer than the row's threshold" is one compare;
//                    passing elements are appended to a per-thread pending buffer and inserted into the register-resident
//                    16-best list in batches.  Measured per 128x128 tile (clock64 trace, -DRTR_TC_TRACE): MMA ~1180
//                    cycles (74 % of the tf32 peak while active), epilogue ~1210, period ~1775.
//   3. k_tc_rerank : one warp per source row recomputes the kept candidates' distances exactly (the oracle's sequential
//                    fp64 sum), takes the top k, and certifies the row: every rejected candidate had approximate distance
//                    >= T (the 16th kept), hence exact distance >= T - E; if the exact k-th best is < T - E nothing
//                    rejected can enter.  Uncertified rows (rare) are redone by the exact SIMT kernel.
// Tried and dropped: the source tile as a TMEM operand (tcgen05.st, [a_tmem] in the MMA) — correct, but with room for only
// three accumulators it was 5 % slower: the kernel is not bound by shared-memory operand traffic.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#define TC_TILE 128                    // rows per operand tile (UMMA M = N = 128)
#define TC_KF 104                      // floats per operand row: 3 x 33 + 5 pad  (13 k-steps of 8 tf32)
#define TC_NC 26                       // 16-byte chunks per operand row
#define TC_TILE_BYTES (TC_TILE * TC_KF * 4)    // 53248
#define TC_STAGES 2                    // shared-memory ring of target tiles: a stage is free again when its MMAs have completed
#define TC_ACC 4                       // TMEM ring of 128-column accumulators: free again when the epilogue has scanned it
#define TC_PEND 16
#define TC_EPI_GROUPS 4                  // epilogue warp groups: group g handles the g-th 32-column chunk of every accumulator
#define TC_KEEP_MAX 32                 // candidates kept per (source row, split): 16, or 32 when there is a single split
#define TC_THREADS (64 + TC_EPI_GROUPS * 128)      // loader warp + MMA warp + 4 x 4 epilogue warps
#define TC_MAX_SPLITS 8

// defined in features.cu
int rtr_match_exact_launch(rtr_context* ctx, const float* fa, int na, const float* fb, int nb, int k, int* out_idx, float* out_dist,
                           const int* rows, const int* row_count, int max_rows);

// ----------------------------------------------------------------------------- small PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded spin: a protocol bug must end in a trapped kernel (an error status at the C ABI), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (long long spin = 0; spin < 4000000LL; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
#ifdef RTR_TC_TRACE
#define TC_TWAIT(acc, bar, ph) do { long long t_ = clock64(); mbar_wait(bar, ph); acc += clock64() - t_; } while (0)
#else
#define TC_TWAIT(acc, bar, ph) mbar_wait(bar, ph)
#endif
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// half a tile from global into the SAME shared-memory offset of every CTA in ctaMask; each destination CTA's mbarrier (same
// offset) receives the complete_tx for the bytes that landed in it
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE ("interleave"): 8-row x 16-byte core matrices;
// LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups (both in 16-byte units).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(128 >> 4) << 16;                         // LBO: next core matrix along K is 128 B away
    d |= (uint64_t)((TC_NC * 128) >> 4) << 32;               // SBO: next 8-row group is 26 core matrices away
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    return d;                                                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
#define TC_IDESC ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_TILE >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24))

// ----------------------------------------------------------------------------- 1. operand preparation
__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_trunc(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

// role 0: source rows A = [hi | hi | lo | 0];  role 1: target rows B = [hi | lo | hi | 0]
// Features are centred on the targets' mean first (distances are translation invariant): the prefilter's error scales
// with |a - mu||b - mu| instead of |a||b|, which matters when all signatures are alike (dense scenes).
__global__ void k_tc_mean(const float* __restrict__ feat, int n, float* __restrict__ mu) {
    __shared__ double acc[32][33];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s0 = 0, s1 = 0;
    for (int r = warp; r < n; r += 32) {
        float a = __ldg(feat + (size_t)r * 33 + lane), b = (lane == 0) ? __ldg(feat + (size_t)r * 33 + 32) : 0.f;
        if (isfinite(a)) s0 += a;
        if (isfinite(b)) s1 += b;
    }
    acc[warp][lane] = s0;
    if (lane == 0) acc[warp][32] = s1;
    __syncthreads();
    if (threadIdx.x < 33) {
        double t = 0;
        for (int w = 0; w < 32; ++w) t += acc[w][threadIdx.x];
        mu[threadIdx.x] = n > 0 ? (float)(t / n) : 0.f;
    }
}

__global__ void k_tc_prep(const float* __restrict__ feat, int n, int n_pad, int role, const float* __restrict__ mu,
                          float* __restrict__ tiles, float* __restrict__ norms, float* __restrict__ norm_max) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    int tile = r / TC_TILE, rr = r % TC_TILE;
    char* base = (char*)tiles + (size_t)tile * TC_TILE_BYTES + (size_t)(rr / 8) * TC_NC * 128 + (size_t)(rr % 8) * 16;
    bool valid = r < n;
    double nrm = 0;
    float v[33];
    if (valid) {
#pragma unroll
        for (int c = 0; c < 33; ++c) { v[c] = __ldg(feat + (size_t)r * 33 + c) - __ldg(mu + c); nrm += (double)v[c] * (double)v[c]; if (!isfinite(v[c])) valid = false; }
    }
    const float nrm_f = valid ? (float)nrm : FLT_MAX;        // FLT_MAX: a padded / non-finite row can never be selected
    // columns 99..101: the target's -|b|^2 / 2 as three tf32 pieces (11 + 11 + 2 mantissa bits: exact) against ones on the
    // source side, so the accumulator is S' = a.b - |b|^2 / 2 and the epilogue needs no norm lookup: d = -2 S'
    const float hx = -0.5f * nrm_f;
    const float h1 = tf32_trunc(hx), h2 = tf32_trunc(hx - h1), h3 = (hx - h1) - h2;
#pragma unroll
    for (int e = 0; e < TC_KF; ++e) {
        float x = 0.f;
        if (valid && e < 99) {
            int c = e % 33, part = e / 33;
            float hi = tf32_hi(v[c]);
            float lo = tf32_trunc(v[c] - hi);
            bool want_lo = (role == 0) ? (part == 2) : (part == 1);
            x = want_lo ? lo : hi;
        }
        if (e >= 99 && e < 102) x = (role == 0) ? 1.0f : (e == 99 ? h1 : (e == 100 ? h2 : h3));
        *(float*)(base + (size_t)(e / 4) * 128 + (e % 4) * 4) = x;
    }
    norms[r] = nrm_f;
    if (valid && norm_max) atomicMax((int*)norm_max, __float_as_int((float)nrm * 1.000001f));   // non-negative floats order like ints
}

// ----------------------------------------------------------------------------- 2. tcgen05 prefilter
struct TcSmem {
    alignas(128) unsigned char a[TC_TILE_BYTES];
    alignas(128) unsigned char b[TC_STAGES][TC_TILE_BYTES];
    // per epilogue thread: up to TC_PEND candidates (accumulator value, column) that passed the row's threshold and wait for
    // insertion into the register-resident list; [entry][thread] so that a warp's accesses to one entry are conflict free
    alignas(16) float pend_v[TC_PEND][TC_EPI_GROUPS * TC_TILE];
    alignas(16) int pend_i[TC_PEND][TC_EPI_GROUPS * TC_TILE];
    alignas(8) unsigned long long bar_a;
    unsigned long long bar_full[TC_STAGES];      // target tile landed in smem
    unsigned long long bar_sfree[TC_STAGES];     // the MMAs reading this smem stage have completed (in every CTA of the cluster)
    unsigned long long bar_acc[TC_ACC];          // accumulator complete in TMEM
    unsigned long long bar_afree[TC_ACC];        // every epilogue warp is done with this accumulator
    unsigned int tmem_base;
};

// CL = 2: two CTAs (two source tiles) form a cluster and share the stream of target tiles — each loads HALF of every tile
// and multicasts it into both CTAs' shared memory, which halves the L2 -> SM traffic the kernel is bound by.  A stage
// may be overwritten only when the epilogues of BOTH CTAs are done with it, so every epilogue warp arrives on the
// stage's free barrier in both CTAs.
template <int TC_KEEP, int CL, bool MERGE>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_match(const float* __restrict__ a_tiles, const float* __restrict__ b_tiles, const float* __restrict__ b_norms, int ns, int n_src_tiles,
           int n_tgt_tiles, int tiles_per_split, int n_splits, int* __restrict__ cand_idx, float* __restrict__ cand_val,
           float* __restrict__ cand_thr) {
    // (no manual realignment: the pointer must stay visibly derived from the __shared__ array, otherwise every access
    //  becomes a generic load / store on the long scoreboard)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool active = (int)blockIdx.x < n_src_tiles;            // CL = 2 with an odd tile count: the last CTA only helps loading
    const int src_tile = active ? (int)blockIdx.x : n_src_tiles - 1, split = blockIdx.y;
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
    const int t0 = split * tiles_per_split;
    const int nt = max(0, min(tiles_per_split, n_tgt_tiles - t0));

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&sm.bar_a), 1);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&sm.bar_full[s]), 1);
            mbar_init(smem_u32(&sm.bar_sfree[s]), CL);
        }
        for (int a = 0; a < TC_ACC; ++a) {
            mbar_init(smem_u32(&sm.bar_acc[a]), 1);
            mbar_init(smem_u32(&sm.bar_afree[a]), 4 * TC_EPI_GROUPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: TC_ACC accumulators x 128 fp32 columns = all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"((uint32_t)(TC_ACC * TC_TILE)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_barrier();          // the peer's barriers exist before anything is multicast to / arrives on them
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    long long w_free = 0, w_full = 0, w_acc = 0, t_begin = clock64();
    (void)w_free; (void)w_full; (void)w_acc; (void)t_begin;

    if (warp == 0) {
        // ---------------- loader: source tile once, then the ring of target tiles
        if (lane == 0) {
            mbar_expect_tx(smem_u32(&sm.bar_a), TC_TILE_BYTES);
            bulk_g2s(smem_u32(sm.a), (const char*)a_tiles + (size_t)src_tile * TC_TILE_BYTES, TC_TILE_BYTES, smem_u32(&sm.bar_a));
            for (int t = 0; t < nt; ++t) {
                int s = t % TC_STAGES;
                uint32_t ph = (uint32_t)(t / TC_STAGES) & 1u;
                TC_TWAIT(w_free, smem_u32(&sm.bar_sfree[s]), ph ^ 1u);
                uint32_t full = smem_u32(&sm.bar_full[s]);
                mbar_expect_tx(full, TC_TILE_BYTES);
                if (CL > 1) {
                    const uint32_t half = TC_TILE_BYTES / CL, off = crank * half;
                    bulk_g2s_multicast(smem_u32(sm.b[s]) + off, (const char*)b_tiles + (size_t)(t0 + t) * TC_TILE_BYTES + off, half, full,
                                       (uint16_t)((1u << CL) - 1u));
                } else {
                    bulk_g2s(smem_u32(sm.b[s]), (const char*)b_tiles + (size_t)(t0 + t) * TC_TILE_BYTES, TC_TILE_BYTES, full);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: one thread issues, accumulators live in TMEM
        mbar_wait(smem_u32(&sm.bar_a), 0);
        const uint64_t adesc = umma_desc(smem_u32(sm.a));
        for (int t = 0; t < nt; ++t) {
            int s = t % TC_STAGES, a = t % TC_ACC;
            uint32_t ph = (uint32_t)(t / TC_STAGES) & 1u, pa = (uint32_t)(t / TC_ACC) & 1u;
            TC_TWAIT(w_full, smem_u32(&sm.bar_full[s]), ph);
            TC_TWAIT(w_acc, smem_u32(&sm.bar_afree[a]), pa ^ 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t bdesc = umma_desc(smem_u32(sm.b[s]));
#pragma unroll
                for (int k = 0; k < TC_KF / 8; ++k)      // 8 tf32 = 32 B = two 16-byte chunks = 256 B of core matrices per k-step
                    tc_mma_tf32(tmem + (uint32_t)(a * TC_TILE), adesc + (uint64_t)(k * 16), bdesc + (uint64_t)(k * 16), TC_IDESC, k > 0 ? 1u : 0u);
                tc_commit(smem_u32(&sm.bar_acc[a]));      // arrives when the MMAs above have completed: accumulator ready ...
                if (CL > 1) tc_commit_multicast(smem_u32(&sm.bar_sfree[s]), (uint16_t)((1u << CL) - 1u));     // ... and the smem stage
                else tc_commit(smem_u32(&sm.bar_sfree[s]));                                                     // may be refilled
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue: thread = source row.  Sixteen warps: TMEM lane quarter q = warp % 4 (a warp may only
        // touch its own quarter), group g = which 32-column chunk of every 128-column accumulator this warp scans; each
        // (row, group) keeps its own list of the TC_KEEP smallest distances (unsorted, in registers, thr = current maximum).
        // The accumulator already holds S' = a.b - |b|^2 / 2 (the norm rides in three spare K columns), so d = -2 S' and
        // "d < thr" is "S' > -thr / 2": one compare per element, and a passing element is only APPENDED (two predicated
        // stores) to the thread's pending buffer.  Insertions — 4 instructions per list slot, executed by the whole warp
        // whenever any lane inserts — happen in batches when some lane's buffer runs out of room, with the threshold
        // tightening as they go, instead of once per chunk.
        const int q = warp & 3, g = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int et = g * TC_TILE + row;                     // epilogue thread id 0..511
        float val[TC_KEEP];
        int idx[TC_KEEP];
        float thr = FLT_MAX, thr_s = -0.5f * FLT_MAX;
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < TC_KEEP; ++u) { val[u] = FLT_MAX; idx[u] = -1; }
        auto flush = [&]() {
            const int mx = __reduce_max_sync(0xffffffffu, cnt);
            for (int e = 0; e < mx; ++e) {
                if (e < cnt) {
                    float d = -2.0f * sm.pend_v[e][et];
                    if (d < thr) {                 // thr may have tightened since the element was appended
                        const int col = sm.pend_i[e][et];
                        bool done = false;
                        float nt_ = -FLT_MAX;
#pragma unroll
                        for (int u = 0; u < TC_KEEP; ++u) {
                            bool hit = (val[u] == thr) && !done;
                            val[u] = hit ? d : val[u];
                            idx[u] = hit ? col : idx[u];
                            done = done || hit;
                            nt_ = fmaxf(nt_, val[u]);
                        }
                        thr = nt_;
                    }
                }
            }
            thr_s = -0.5f * thr;
            cnt = 0;
        };
        for (int t = 0; t < nt; ++t) {
            int a = t % TC_ACC;
            uint32_t pa = (uint32_t)(t / TC_ACC) & 1u;
            TC_TWAIT(w_acc, smem_u32(&sm.bar_acc[a]), pa);
            tc_fence_after();
            const int col0 = (t0 + t) * TC_TILE;
#pragma unroll 1
            for (int cc = g; cc < TC_TILE / 32; cc += TC_EPI_GROUPS) {
                uint32_t r[32];
                tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TC_TILE + cc * 32), r);
                tc_ld_wait();
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    if (__any_sync(0xffffffffu, cnt > TC_PEND - 8)) flush();      // room for the next 8 columns
                    // most 8-column blocks hold nothing below the row's threshold: one 3-input-max tree and one branch
                    const float* v = reinterpret_cast<const float*>(&r[j8 * 8]);
                    const float m8 = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
                    if (m8 > thr_s) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float sp = v[jj];
                            if (sp > thr_s) {
                                sm.pend_v[cnt][et] = sp;
                                sm.pend_i[cnt][et] = col0 + cc * 32 + j8 * 8 + jj;
                                ++cnt;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.bar_afree[a]));
        }
        flush();
        int grow = src_tile * TC_TILE + row;
        if (MERGE) {
            // several splits: one list per (row, split) is enough, so the four groups' lists are merged through the (now idle)
            // pending buffers.  Every element a group rejected was >= that group's threshold >= the merged 16th smallest.
            static_assert(TC_PEND >= TC_KEEP, "the pending buffer doubles as the merge scratch");
#pragma unroll
            for (int u = 0; u < TC_KEEP; ++u) { sm.pend_v[u][et] = val[u]; sm.pend_i[u][et] = idx[u]; }
            asm volatile("bar.sync 1, %0;" ::"r"(TC_EPI_GROUPS * TC_TILE) : "memory");       // the 16 epilogue warps only
            if (g == 0) {
                for (int og = 1; og < TC_EPI_GROUPS; ++og)
                    for (int e = 0; e < TC_KEEP; ++e) {
                        float d = sm.pend_v[e][og * TC_TILE + row];
                        if (d < thr) {
                            const int col = sm.pend_i[e][og * TC_TILE + row];
                            bool done = false;
                            float nt_ = -FLT_MAX;
#pragma unroll
                            for (int u = 0; u < TC_KEEP; ++u) {
                                bool hit = (val[u] == thr) && !done;
                                val[u] = hit ? d : val[u];
                                idx[u] = hit ? col : idx[u];
                                done = done || hit;
                                nt_ = fmaxf(nt_, val[u]);
                            }
                            thr = nt_;
                        }
                    }
                if (active && grow < ns) {
                    const size_t list = (size_t)grow * n_splits + split;
                    size_t o = list * TC_KEEP;
#pragma unroll
                    for (int u = 0; u < TC_KEEP; ++u) { cand_idx[o + u] = idx[u]; cand_val[o + u] = val[u]; }
                    cand_thr[list] = thr;
                }
            }
        } else if (active && grow < ns) {
            const size_t list = (size_t)grow * (n_splits * TC_EPI_GROUPS) + (size_t)split * TC_EPI_GROUPS + g;
            size_t o = list * TC_KEEP;
#pragma unroll
            for (int u = 0; u < TC_KEEP; ++u) { cand_idx[o + u] = idx[u]; cand_val[o + u] = val[u]; }
            cand_thr[list] = thr;     // FLT_MAX while the list is not full: nothing was rejected
        }
    }
#ifdef RTR_TC_TRACE
    if (blockIdx.x == 3 && blockIdx.y == 0 && lane == 0 && (warp < 3 || warp == 17))
        printf("tc trace warp %d: total %lld cycles, %d tiles; waits: free %lld full %lld acc %lld\n", warp, clock64() - t_begin, nt, w_free, w_full, w_acc);
#endif
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_barrier();          // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(TC_ACC * TC_TILE)) : "memory");
}

// ----------------------------------------------------------------------------- 3. exact re-rank + certificate
#define RR_WARPS 8
#define RR_KMAX 8
__global__ void __launch_bounds__(RR_WARPS * 32)
k_tc_rerank(const float* __restrict__ fa, int ns, const float* __restrict__ fb, int nt, int k, const int* __restrict__ cand_idx,
            const float* __restrict__ cand_val, const float* __restrict__ cand_thr, int n_splits, int TC_KEEP, const float* __restrict__ a_norms,
            const float* __restrict__ b_norms, const float* __restrict__ nb_max_p, int* __restrict__ out_idx, float* __restrict__ out_dist, int* __restrict__ redo_rows, int* __restrict__ redo_count,
            float* __restrict__ err_ratio_max, int force_redo) {
    __shared__ float src[RR_WARPS][36];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int row = blockIdx.x * RR_WARPS + warp;
    if (row >= ns) return;
    src[warp][lane] = fa[(size_t)row * 33 + lane];
    if (lane == 0) src[warp][32] = fa[(size_t)row * 33 + 32];
    __syncwarp();
    // squared norms of the CENTRED vectors (what the prefilter worked with); a non-finite source row has FLT_MAX here
    double na = (double)a_norms[row];
    double nb_max = (double)*nb_max_p;
    // Error model of the prefilter's distance for a pair with centred norms |a|, |b| (DESIGN.md "tensor-core matching"):
    // tf32 hi/lo split residue + dropped lo.lo term + fp32 accumulation over 13 MMA k-steps <= 3e-6 |a||b| on the dot
    // product (x2 in the distance), plus the fp32 roundings of |b|^2 and of the final fma.
#define TC_ERR(na_, nb_) (6e-6 * sqrt((na_) * (nb_)) + 5e-7 * ((na_) + (nb_)) + 1e-9)
    float bd[RR_KMAX]; int bi[RR_KMAX];
#pragma unroll
    for (int t = 0; t < RR_KMAX; ++t) { bd[t] = FLT_MAX; bi[t] = 0x7fffffff; }
    int C = n_splits * TC_KEEP;
    bool sane = true;
    double T = DBL_MAX;                       // smallest approximate distance any REJECTED candidate can have
    for (int s = lane; s < n_splits; s += 32) {
        float last = cand_thr[(size_t)row * n_splits + s];
        if (last < FLT_MAX) T = fmin(T, (double)last + na);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) T = fmin(T, __shfl_xor_sync(0xffffffffu, T, o));
    for (int c = lane; c < C; c += 32) {
        int id = cand_idx[(size_t)row * C + c];
        if (id < 0 || id >= nt) continue;
        const float* b = fb + (size_t)id * 33;
        double sacc = 0;
#pragma unroll
        for (int e = 0; e < 33; ++e) { double d = (double)src[warp][e] - (double)__ldg(b + e); sacc += d * d; }
        float df = (float)sacc;
        if (!(df == df)) continue;
        // the error model must hold on everything we can check, with the candidate's own norm
        double err = fabs(((double)cand_val[(size_t)row * C + c] + na) - sacc);
        if (err > TC_ERR(na, (double)b_norms[id])) sane = false;
        // observed error in units of |a||b| (diagnostic: shows how much head-room the 3e-6 model constant has)
        double scale = sqrt(na * (double)b_norms[id]);
        if (scale > 0) atomicMax((int*)err_ratio_max, __float_as_int((float)(err / scale)));
        if (df < bd[RR_KMAX - 1] || (df == bd[RR_KMAX - 1] && id < bi[RR_KMAX - 1])) {
            bd[RR_KMAX - 1] = df; bi[RR_KMAX - 1] = id;
#pragma unroll
            for (int t = RR_KMAX - 1; t > 0; --t) {
                bool sw = (bd[t] < bd[t - 1]) || (bd[t] == bd[t - 1] && bi[t] < bi[t - 1]);
                if (sw) { float td = bd[t]; bd[t] = bd[t - 1]; bd[t - 1] = td; int ti = bi[t]; bi[t] = bi[t - 1]; bi[t - 1] = ti; }
            }
        }
    }
    sane = __all_sync(0xffffffffu, sane);
    float kth = FLT_MAX;
    for (int t = 0; t < k; ++t) {
        float hd = bd[0]; int hi = bi[0];
        float md = hd; int mi = hi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, md, o);
            int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
        }
        if (mi == hi && md == hd && hi != 0x7fffffff) {
#pragma unroll
            for (int u = 0; u < RR_KMAX - 1; ++u) { bd[u] = bd[u + 1]; bi[u] = bi[u + 1]; }
            bd[RR_KMAX - 1] = FLT_MAX; bi[RR_KMAX - 1] = 0x7fffffff;
        }
        kth = md;
        if (lane == 0) {
            bool none = (mi == 0x7fffffff);
            out_idx[(size_t)row * k + t] = none ? -1 : mi;
            out_dist[(size_t)row * k + t] = none ? __int_as_float(0x7fc00000) : md;
        }
    }
    // Certificate.  A rejected candidate j has approximate distance >= T.  If |b_j| > |a| + sqrt(kth) its exact distance
    // exceeds kth by the triangle inequality; otherwise its error is at most E = TC_ERR(|a|^2, (|a| + sqrt(kth))^2), so
    // its exact distance is >= T - E.  The row is final iff the exact k-th best is strictly below that.
    double reach = sqrt(na) + sqrt(fmax((double)kth, 0.0));
    double E = TC_ERR(na, fmin(nb_max, reach * reach));
    bool certified = !force_redo && sane && (T == DBL_MAX || (kth < FLT_MAX && (double)kth < T - E));
    if (!certified && lane == 0) redo_rows[atomicAdd(redo_count, 1)] = row;
}

// ----------------------------------------------------------------------------- host driver
static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

bool rtr_match_tc_wanted(long long ns, long long nt) {
    const char* e = getenv("RTR_MATCH_TC");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return ns > 0 && nt > 0;
    return ns * nt >= 2000000LL;          // below ~2 M pairs the exact SIMT kernel is already a few tens of microseconds
}

// fa, fb: device feature arrays (ns x 33, nt x 33).  out_idx / out_dist: device, ns x k.  stats (host, optional):
// [0] rows redone by the exact kernel, [1] splits, [2] target tiles.
int rtr_match_tc_dev(rtr_context* ctx, const float* fa, int ns, const float* fb, int nt, int k, int* out_idx, float* out_dist, int* stats) {
    if (k > RR_KMAX) return rtr_fail("match.tc", "k must be <= 8", RTR_ERR_INVALID);
    int src_tiles = nblk(ns, TC_TILE), tgt_tiles = nblk(nt, TC_TILE);
    int splits = std::min(std::min(TC_MAX_SPLITS, tgt_tiles), std::max(1, (2 * ctx->sm_count + src_tiles - 1) / src_tiles));
    int tiles_per_split = nblk(tgt_tiles, splits);
    splits = nblk(tgt_tiles, tiles_per_split);
    float *a_tiles = nullptr, *b_tiles = nullptr, *a_norms = nullptr, *b_norms = nullptr, *cand_val = nullptr, *cand_thr = nullptr, *d_nbmax = nullptr, *mu = nullptr;
    int *cand_idx = nullptr, *redo_rows = nullptr, *redo_count = nullptr;
    if (int e = tmp_alloc(ctx, &a_tiles, (size_t)src_tiles * TC_TILE * TC_KF, "match.tc")) return e;
    if (int e = tmp_alloc(ctx, &b_tiles, (size_t)tgt_tiles * TC_TILE * TC_KF, "match.tc")) return e;
    if (int e = tmp_alloc(ctx, &a_norms, (size_t)src_tiles * TC_TILE, "match.tc")) return e;
    if (int e = tmp_alloc(ctx, &b_norms, (size_t)tgt_tiles * TC_TILE, "match.tc")) return e;
    // more kept candidates make the certificate succeed more often; with one split there is only one list per row
    // one split: keep the four groups' lists (64 candidates per row make the certificate succeed more often); several splits:
    // the groups' lists are merged in the kernel, one list of 16 per (row, split)
    const bool merge = splits > 1;
    // Short target lists (a scan of a few thousand points against a whole model database: 15 target tiles per source row) never
    // tighten the threshold enough for the one-compare fast path: the epilogue is bound by list insertions (~6 instructions per
    // list slot each, warp-wide).  Shorter lists make an insertion cheaper and rarer (6 ln(n/6) against 16 ln(n/16) per 480
    // columns); the certificate decides as before whether a row is final, uncertified rows are redone exactly.  Measured on the
    // bench batch (74 786 model rows x 1909 scan rows, k = 5): lists of 16 / 8 / 6 / 4 -> 659 / 398 / 338 / 819 us with
    // 0 / 0 / 0 / 15 139 rows redone (4 < k cannot hold the answer of one group).
    const char* keep_e = getenv("RTR_MATCH_KEEP");       // read per call: tools/bench_match_scale.py times several settings in one process
    const int keep_env = keep_e ? atoi(keep_e) : 0;
    const int keep = (keep_env == 4 || keep_env == 6 || keep_env == 8 || keep_env == 16) ? keep_env : ((!merge && tgt_tiles <= 64 && k <= 5) ? 6 : 16);
    const int lists = merge ? splits : TC_EPI_GROUPS;
    if (int e = tmp_alloc(ctx, &cand_idx, (size_t)ns * lists * keep, "match.tc")) return e;
    if (int e = tmp_alloc(ctx, &cand_val, (size_t)ns * lists * keep, "match.tc")) return e;
    if (int e = tmp_alloc(ctx, &cand_thr, (size_t)ns * lists, "match.tc")) return e;
    if (int e = tmp_alloc(ctx, &redo_rows, (size_t)ns, "match.tc")) return e;
    if (int e = tmp_alloc(ctx, &redo_count, 1, "match.tc")) return e;
    if (int e = tmp_alloc(ctx, &d_nbmax, 2, "match.tc")) return e;      // [0] max centred |b|^2, [1] observed error ratio
    if (int e = tmp_alloc(ctx, &mu, 33, "match.tc")) return e;
    RTR_CHECK(cudaMemsetAsync(redo_count, 0, sizeof(int), ctx->stream), "match.tc");
    RTR_CHECK(cudaMemsetAsync(d_nbmax, 0, 2 * sizeof(float), ctx->stream), "match.tc");
    k_tc_mean<<<1, 1024, 0, ctx->stream>>>(fb, std::min(nt, 1024), mu);      // any fixed vector works; a sample mean is enough
    RTR_LAUNCH_CHECK(ctx, "match.tc_mean");
    k_tc_prep<<<nblk((long long)src_tiles * TC_TILE, 128), 128, 0, ctx->stream>>>(fa, ns, src_tiles * TC_TILE, 0, mu, a_tiles, a_norms, nullptr);
    RTR_LAUNCH_CHECK(ctx, "match.tc_prep");
    k_tc_prep<<<nblk((long long)tgt_tiles * TC_TILE, 128), 128, 0, ctx->stream>>>(fb, nt, tgt_tiles * TC_TILE, 1, mu, b_tiles, b_norms, d_nbmax);
    RTR_LAUNCH_CHECK(ctx, "match.tc_prep");
    size_t smem = sizeof(TcSmem);
    if (int e = rtr_kernel_smem(k_tc_match<16, 1, false>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<16, 2, false>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<16, 1, true>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<16, 2, true>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<4, 1, false>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<4, 2, false>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<4, 1, true>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<4, 2, true>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<6, 1, false>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<6, 1, true>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<8, 1, false>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<8, 2, false>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<8, 1, true>, ctx, smem)) return e;
    if (int e = rtr_kernel_smem(k_tc_match<8, 2, true>, ctx, smem)) return e;
    // RTR_MATCH_CLUSTER=1: pairs of source tiles share the target-tile stream through cluster multicast (half the L2 -> SM
    // traffic).  Off by default: since loads stopped waiting for the epilogue the kernel is bound by MMA + epilogue, the two
    // variants time the same (7.6 ms at 262144 x 65536), and independent CTAs need no gang scheduling.
    bool use_cluster = false;
    if (const char* e = getenv("RTR_MATCH_CLUSTER")) use_cluster = (e[0] == '1') && src_tiles >= 2;
    {
        const int cl = use_cluster ? 2 : 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((src_tiles + cl - 1) / cl * cl), (unsigned)splits, 1);
        cfg.blockDim = dim3(TC_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const float *ca = a_tiles, *cb = b_tiles, *cn = b_norms;
        cudaError_t le;
#define TC_LAUNCH(KP_, CL_, MG_) cudaLaunchKernelEx(&cfg, k_tc_match<KP_, CL_, MG_>, ca, cb, cn, ns, src_tiles, tgt_tiles, tiles_per_split, splits, cand_idx, cand_val, cand_thr)
        if (keep == 4) le = use_cluster ? (merge ? TC_LAUNCH(4, 2, true) : TC_LAUNCH(4, 2, false)) : (merge ? TC_LAUNCH(4, 1, true) : TC_LAUNCH(4, 1, false));
        else if (keep == 6) le = merge ? TC_LAUNCH(6, 1, true) : TC_LAUNCH(6, 1, false);
        else if (keep == 8) le = use_cluster ? (merge ? TC_LAUNCH(8, 2, true) : TC_LAUNCH(8, 2, false)) : (merge ? TC_LAUNCH(8, 1, true) : TC_LAUNCH(8, 1, false));
        else le = use_cluster ? (merge ? TC_LAUNCH(16, 2, true) : TC_LAUNCH(16, 2, false)) : (merge ? TC_LAUNCH(16, 1, true) : TC_LAUNCH(16, 1, false));
#undef TC_LAUNCH
        RTR_CHECK(le, "match.tc_mma");
    }
    RTR_LAUNCH_CHECK(ctx, "match.tc_mma");
    // RTR_MATCH_FORCE_REDO=1 (tests): treat every row as uncertified, so the exact redo kernels answer all of them
    const char* fr = getenv("RTR_MATCH_FORCE_REDO");
    const int force_redo = (fr && fr[0] == '1') ? 1 : 0;
    k_tc_rerank<<<nblk(ns, RR_WARPS), RR_WARPS * 32, 0, ctx->stream>>>(fa, ns, fb, nt, k, cand_idx, cand_val, cand_thr, lists, keep, a_norms, b_norms, d_nbmax, out_idx, out_dist, redo_rows, redo_count, d_nbmax + 1, force_redo);
    RTR_LAUNCH_CHECK(ctx, "match.tc_rerank");
    if (int e = rtr_match_exact_launch(ctx, fa, ns, fb, nt, k, out_idx, out_dist, redo_rows, redo_count, ns)) return e;
    if (stats) {
        float ratio = 0.f;
        RTR_CHECK(cudaMemcpyAsync(&stats[0], redo_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "match.tc");
        RTR_CHECK(cudaMemcpyAsync(&ratio, d_nbmax + 1, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "match.tc");
        RTR_CHECK(cudaStreamSynchronize(ctx->stream), "match.tc");
        stats[1] = splits; stats[2] = (int)(ratio * 1e9f);      // observed max |approx - exact| / (|a||b|), in 1e-9 units
    }
    dev_free(ctx, a_tiles); dev_free(ctx, b_tiles); dev_free(ctx, a_norms); dev_free(ctx, b_norms);
    dev_free(ctx, cand_idx); dev_free(ctx, cand_val); dev_free(ctx, cand_thr); dev_free(ctx, redo_rows); dev_free(ctx, redo_count); dev_free(ctx, d_nbmax); dev_free(ctx, mu);
    return 0;
}
