// features.cu — per-cloud stages: normals, Harris-3D (response, NMS, refinement), FPFH, feature k-NN.
//
// These replace PCL classes the reference instantiates inside ModelPoint::getKeypoint / ScanPoint::getKeypoint
// (model_point.h:127-136, scan_point.h:85-94: HarrisKeypoint3D with its implicit NormalEstimation) and the FPFH +
// feature-correspondence stages BASELINE.json's north_star adds.  All gather kernels run one
// thread (or one warp) per point IN CELL ORDER, so neighbouring threads scan the same 9 contiguous ranges of the
// cell-sorted array; accumulation is fp64, storage fp32 (common.cuh, arithmetic contract).
#include "common.cuh"
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <algorithm>
#include <cmath>
#include <cstring>

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }
// Size thresholds between the small-cloud and large-cloud kernel variants; the environment overrides exist so that the
// tests can push a repo-sized cloud through the large-cloud kernels and compare them with the oracle.
// persistent-grid width of the warp-per-point kernels in CTAs per SM (RTR_GRID_MULT: scheduling experiment knob)
static inline int wide_grid_mult() {
    static const int m = []() { const char* e = getenv("RTR_GRID_MULT"); int v = e ? atoi(e) : 16; return v > 0 ? v : 16; }();
    return m;
}
static inline int env_threshold(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

// ----------------------------------------------------------------------------- normals (App. A.2)
// covariance of the offsets from the query (fp64) -> smallest eigenvector (Jacobi) -> flip towards the viewpoint (0,0,0)
__device__ __forceinline__ float4 normal_from_sums(float4 q, int cnt, double sx, double sy, double sz, double cxx, double cxy,
                                                   double cxz, double cyy, double cyz, double czz) {
    float4 o;
    if (cnt < 3) {
        o.x = o.y = o.z = o.w = __int_as_float(0x7fc00000);
        return o;
    }
    double k = (double)cnt;
    double mx = sx / k, my = sy / k, mz = sz / k;
    double a[3][3], v[3][3];
    a[0][0] = cxx / k - mx * mx; a[0][1] = cxy / k - mx * my; a[0][2] = cxz / k - mx * mz;
    a[1][1] = cyy / k - my * my; a[1][2] = cyz / k - my * mz; a[2][2] = czz / k - mz * mz;
    a[1][0] = a[0][1]; a[2][0] = a[0][2]; a[2][1] = a[1][2];
    double trace = a[0][0] + a[1][1] + a[2][2];
    jacobi_eig<3>(a, v);
    double ev = a[0][0], nx = v[0][0], ny = v[1][0], nz = v[2][0];
    if (a[1][1] < ev) { ev = a[1][1]; nx = v[0][1]; ny = v[1][1]; nz = v[2][1]; }
    if (a[2][2] < ev) { ev = a[2][2]; nx = v[0][2]; ny = v[1][2]; nz = v[2][2]; }
    double curv = (trace > 0) ? fabs(ev / trace) : 0.0;
    double dot = (nx * -(double)q.x + ny * -(double)q.y) + nz * -(double)q.z;     // flipNormalTowardsViewpoint, vp = 0
    if (dot < 0) { nx = -nx; ny = -ny; nz = -nz; }
    o.x = (float)nx; o.y = (float)ny; o.z = (float)nz; o.w = (float)curv;
    return o;
}

// large clouds: one thread per point (neighbouring threads walk the same cell ranges).  ncu at 4 M points: issue-bound
// (67 % issue active, fp64 pipe 58 %), the in-radius block runs with ~11 of 32 lanes because which candidates are inside
// differs from lane to lane.  Parking accepted candidates in a per-thread shared-memory ring and draining the rings
// warp-wide (dense fp64 blocks) was built and measured: 1.25 -> 1.76 ms — the extra votes, ring traffic and second load
// cost more than the divergence they remove — so the plain loop stays.
__global__ void __launch_bounds__(128) k_normals(GridView g, float r2, float4* __restrict__ normals) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n) return;
    float4 q = __ldg(g.sorted + s);
    double sx = 0, sy = 0, sz = 0, cxx = 0, cxy = 0, cxz = 0, cyy = 0, cyz = 0, czz = 0;
    int cnt = 0;
    for_block27(g, q.x, q.y, q.z, [&](int, float4 p, float d2) {
        if (d2 < r2) {
            double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
            sx += dx; sy += dy; sz += dz;
            cxx += dx * dx; cxy += dx * dy; cxz += dx * dz; cyy += dy * dy; cyz += dy * dz; czz += dz * dz;
            ++cnt;
        }
    });
    normals[__float_as_int(q.w)] = normal_from_sums(q, cnt, sx, sy, sz, cxx, cxy, cxz, cyy, cyz, czz);
}

// Visit the 27-cell block with the lanes of a warp striding over each of the 9 ranges (all lanes share the query).
template <typename F>
__device__ __forceinline__ void for_block27_warp(const GridView& g, float qx, float qy, float qz, int lane, F&& f) {
    int cx = cell_coord(qx, g.mnx, g.inv_h), cy = cell_coord(qy, g.mny, g.inv_h), cz = cell_coord(qz, g.mnz, g.inv_h);
    if (cx < -1 || cy < -1 || cz < -1 || cx > g.dx || cy > g.dy || cz > g.dz) return;
    cx = clampi(cx, 0, g.dx - 1); cy = clampi(cy, 0, g.dy - 1); cz = clampi(cz, 0, g.dz - 1);
    BlockRanges br = warp_block_ranges(g, cx, cy, cz, lane);
    for (int z = br.z0; z <= br.z1; ++z)
        for (int y = br.y0; y <= br.y1; ++y) {
            int r = (z - br.z0) * 3 + (y - br.y0);
            int s0 = __shfl_sync(0xffffffffu, br.bound, r);
            int s1 = __shfl_sync(0xffffffffu, br.bound, 9 + r);
            for (int s = s0 + lane; s < s1; s += 32) {
                float4 p = __ldg(g.sorted + s);
                f(s, p, dist2f(qx, qy, qz, p.x, p.y, p.z));
            }
        }
}

// repo-sized clouds (a few thousand points) cannot fill the GPU with one thread per point: one WARP per point, lanes
// stride over the candidates, fp64 partial sums folded with a warp-shuffle tree.
#define PW_WARPS 8
// The eigen-solve is serial fp64 (cyclic Jacobi: divisions, square roots) — run by lane 0 of a warp it costs the whole warp
// ~1500 issue slots per point, half of this kernel's instructions on the bench batch (ncu: 221 M warp instructions for 75 k
// points).  So the warps only gather and reduce (the ten sums go to global memory, 80 B per point), and k_normals_solve
// runs the solve with one THREAD per point: every lane busy, same sums, same operations, same bits.
template <class GS>
__global__ void __launch_bounds__(PW_WARPS * 32) k_normals_warp(const __grid_constant__ GS gs, float r2, double* __restrict__ sums /* [10][total] */) {
    int lane = threadIdx.x & 31;
    int nwarps = gridDim.x * PW_WARPS;
    const int total = gs.total();
    for (int s = blockIdx.x * PW_WARPS + (threadIdx.x >> 5); s < total; s += nwarps) {
        const GridView g = gs.at(s);
        float4 q = __ldg(g.sorted + s);
        double sx = 0, sy = 0, sz = 0, cxx = 0, cxy = 0, cxz = 0, cyy = 0, cyz = 0, czz = 0;
        int cnt = 0;
        for_block27_warp(g, q.x, q.y, q.z, lane, [&](int, float4 p, float d2) {
            if (d2 < r2) {
                double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
                sx += dx; sy += dy; sz += dz;
                cxx += dx * dx; cxy += dx * dy; cxz += dx * dz; cyy += dy * dy; cyz += dy * dz; czz += dz * dz;
                ++cnt;
            }
        });
        cnt = warp_sum(cnt);
        sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
        cxx = warp_sum(cxx); cxy = warp_sum(cxy); cxz = warp_sum(cxz); cyy = warp_sum(cyy); cyz = warp_sum(cyz); czz = warp_sum(czz);
        // lane k stores sum k (one coalescing-friendly store per lane instead of ten by lane 0)
        double v = lane == 0 ? sx : lane == 1 ? sy : lane == 2 ? sz : lane == 3 ? cxx : lane == 4 ? cxy : lane == 5 ? cxz : lane == 6 ? cyy
                 : lane == 7 ? cyz : lane == 8 ? czz : (double)cnt;
        if (lane < 10) sums[(size_t)lane * total + s] = v;
    }
}
__global__ void __launch_bounds__(128) k_normals_solve(const float4* __restrict__ sorted, int total, const double* __restrict__ sums,
                                                       float4* __restrict__ normals) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const float4 q = __ldg(sorted + s);
    const size_t t = (size_t)total;
    normals[__float_as_int(q.w)] = normal_from_sums(q, (int)sums[9 * t + s], sums[s], sums[t + s], sums[2 * t + s], sums[3 * t + s], sums[4 * t + s],
                                                    sums[5 * t + s], sums[6 * t + s], sums[7 * t + s], sums[8 * t + s]);
}

// ---- PCL-float-faithful normals (mode 1) ------------------------------------------------
// What pcl::NormalEstimation itself does: nine SINGLE-PASS float accumulators over the raw coordinates (sum xx, xy, xz, yy,
// yz, zz, x, y, z — no de-meaning pass, so cancellation grows with the distance from the origin), C = E[pp^T] - mu mu^T,
// smallest eigenpair by pcl::eigen33's closed-form trigonometric roots in float.  Closer to the reference than the exact
// mode (fp64 sums of offsets + Jacobi) and off the fp64 pipe, which bounds k_normals at 58 % (ncu, 4 M points).  One thread
// per point in cell order; each thread adds its neighbours in the grid's candidate order, so the result is run-to-run
// reproducible.  It differs from the oracle's mode 1 by float summation order (the oracle adds in ascending index) and by
// the device's atan2f / sincosf — parity is by tolerance (tests/test_gpu_parity.py::test_normals_pcl_float_mode).
__device__ __forceinline__ void pcl_eigen33_smallest_dev(const float* cov, float& eval, float* evec) {
    float scale = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(cov[i]));
    if (scale <= FLT_MIN) scale = 1.0f;
    float m[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = cov[i] / scale;
    const float c0 = m[0] * m[4] * m[8] + 2.f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
    const float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
    const float c2 = m[0] + m[4] + m[8];
    float r0, r1, r2;
    auto quad = [&](float b, float c) {        // roots of x^2 - b x + c, with root 0 = 0
        r0 = 0.f;
        float d = b * b - 4.f * c; if (d < 0.f) d = 0.f;
        const float sd = sqrtf(d);
        r2 = 0.5f * (b + sd); r1 = 0.5f * (b - sd);
    };
    if (fabsf(c0) < FLT_EPSILON) quad(c2, c1);
    else {
        const float s_inv3 = 1.0f / 3.0f, s_sqrt3 = sqrtf(3.0f);
        const float c2_over_3 = c2 * s_inv3;
        float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3; if (a_over_3 > 0.f) a_over_3 = 0.f;
        const float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
        float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3; if (q > 0.f) q = 0.f;
        const float rho = sqrtf(-a_over_3);
        const float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
        float st, ct;
        sincosf(theta, &st, &ct);
        r0 = c2_over_3 + 2.f * rho * ct;
        r1 = c2_over_3 - rho * (ct + s_sqrt3 * st);
        r2 = c2_over_3 - rho * (ct - s_sqrt3 * st);
        float t;
        if (r0 >= r1) { t = r0; r0 = r1; r1 = t; }
        if (r1 >= r2) { t = r1; r1 = r2; r2 = t; if (r0 >= r1) { t = r0; r0 = r1; r1 = t; } }
        if (r0 <= 0.f) quad(c2, c1);
    }
    eval = r0 * scale;
    const float d0 = m[0] - r0, d1 = m[4] - r0, d2 = m[8] - r0;
    // rows of (C / scale - lambda I): (d0, m1, m2), (m1, d1, m5), (m2, m5, d2); the largest of the three cross products
    const float v0x = m[1] * m[5] - m[2] * d1, v0y = m[2] * m[1] - d0 * m[5], v0z = d0 * d1 - m[1] * m[1];
    const float v1x = m[1] * d2 - m[2] * m[5], v1y = m[2] * m[2] - d0 * d2, v1z = d0 * m[5] - m[1] * m[2];
    const float v2x = d1 * d2 - m[5] * m[5], v2y = m[5] * m[2] - m[1] * d2, v2z = m[1] * m[5] - d1 * m[2];
    const float l0 = v0x * v0x + v0y * v0y + v0z * v0z, l1 = v1x * v1x + v1y * v1y + v1z * v1z, l2 = v2x * v2x + v2y * v2y + v2z * v2z;
    float vx, vy, vz, l;
    if (l0 >= l1 && l0 >= l2) { vx = v0x; vy = v0y; vz = v0z; l = l0; }
    else if (l1 >= l0 && l1 >= l2) { vx = v1x; vy = v1y; vz = v1z; l = l1; }
    else { vx = v2x; vy = v2y; vz = v2z; l = l2; }
    const float inv = 1.0f / sqrtf(l);
    evec[0] = vx * inv; evec[1] = vy * inv; evec[2] = vz * inv;
}

template <class GS>
__global__ void __launch_bounds__(128) k_normals_pcl_float(const __grid_constant__ GS gs, float r2, float4* __restrict__ normals) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= gs.total()) return;
    const GridView g = gs.at(s);
    const float4 q = __ldg(g.sorted + s);
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;
    int cnt = 0;
    for_block27(g, q.x, q.y, q.z, [&](int, float4 p, float d2) {
        if (d2 < r2) {
            a0 += p.x * p.x; a1 += p.x * p.y; a2 += p.x * p.z; a3 += p.y * p.y; a4 += p.y * p.z; a5 += p.z * p.z;
            a6 += p.x; a7 += p.y; a8 += p.z;
            ++cnt;
        }
    });
    float4 o;
    if (cnt < 3) o.x = o.y = o.z = o.w = __int_as_float(0x7fc00000);
    else {
        const float k = (float)cnt;
        a0 /= k; a1 /= k; a2 /= k; a3 /= k; a4 /= k; a5 /= k; a6 /= k; a7 /= k; a8 /= k;
        float cov[9];
        cov[0] = a0 - a6 * a6; cov[1] = a1 - a6 * a7; cov[2] = a2 - a6 * a8;
        cov[4] = a3 - a7 * a7; cov[5] = a4 - a7 * a8; cov[8] = a5 - a8 * a8;
        cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
        float ev, e[3];
        pcl_eigen33_smallest_dev(cov, ev, e);
        const float tr = cov[0] + cov[4] + cov[8];
        const double curv = (tr != 0.f) ? fabs((double)(ev / tr)) : 0.0;
        double nx = e[0], ny = e[1], nz = e[2];
        const double dot = (nx * -(double)q.x + ny * -(double)q.y) + nz * -(double)q.z;     // flipNormalTowardsViewpoint, vp = 0
        if (dot < 0) { nx = -nx; ny = -ny; nz = -nz; }
        o.x = (float)nx; o.y = (float)ny; o.z = (float)nz; o.w = (float)curv;
    }
    normals[__float_as_int(q.w)] = o;
}

// ----------------------------------------------------------------------------- Harris 3D (App. A.3)
__device__ __forceinline__ float harris_from_sums(int cnt, double c0, double c1, double c2, double c3, double c4, double c5);
__global__ void __launch_bounds__(128) k_harris_response(GridView g, const float4* __restrict__ sn, float r2,
                                                         float* __restrict__ resp, float* __restrict__ resp_sorted) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n) return;
    float4 q = __ldg(g.sorted + s);
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
    int cnt = 0;
    for_block27(g, q.x, q.y, q.z, [&](int sp, float4, float d2) {
        if (d2 < r2) {
            float4 nj = __ldg(sn + sp);
            if (finite3(nj)) {
                double x = nj.x, y = nj.y, z = nj.z;
                c0 += x * x; c1 += x * y; c2 += x * z; c3 += y * y; c4 += y * z; c5 += z * z; ++cnt;
            }
        }
    });
    float r = harris_from_sums(cnt, c0, c1, c2, c3, c4, c5);
    resp[__float_as_int(q.w)] = r;
    resp_sorted[s] = r;
}

__global__ void __launch_bounds__(128) k_harris_nms(GridView g, const float* __restrict__ resp_sorted, float r2, float thr,
                                                    int nms, unsigned char* __restrict__ flags) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n) return;
    float4 q = __ldg(g.sorted + s);
    float r = resp_sorted[s];
    bool keep = isfinite(r) && (r >= thr);
    if (keep && nms) {
        bool is_max = true;
        for_block27(g, q.x, q.y, q.z, [&](int sp, float4, float d2) {
            if (d2 < r2 && __ldg(resp_sorted + sp) > r) is_max = false;
        });
        keep = is_max;
    }
    flags[__float_as_int(q.w)] = keep ? 1 : 0;
}

__device__ __forceinline__ float harris_from_sums(int cnt, double c0, double c1, double c2, double c3, double c4, double c5) {
    float r = 0.f;
    if (cnt > 0) {
        double k = (double)cnt;
        double xx = c0 / k, xy = c1 / k, xz = c2 / k, yy = c3 / k, yz = c4 / k, zz = c5 / k;
        double trace = xx + yy + zz;
        if (trace != 0) {
            double det = xx * yy * zz + 2.0 * xy * xz * yz - xz * xz * yy - xy * xy * zz - yz * yz * xx;
            r = (float)(0.04 + det - 0.04 * trace * trace);
        }
    }
    return r;
}

template <class GS>
__global__ void __launch_bounds__(PW_WARPS * 32) k_harris_response_warp(const __grid_constant__ GS gs, const float4* __restrict__ sn, float r2,
                                                                       float* __restrict__ resp, float* __restrict__ resp_sorted) {
    int lane = threadIdx.x & 31;
    int nwarps = gridDim.x * PW_WARPS;
    const int total = gs.total();
    for (int s = blockIdx.x * PW_WARPS + (threadIdx.x >> 5); s < total; s += nwarps) {
        const GridView g = gs.at(s);
        float4 q = __ldg(g.sorted + s);
        double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
        int cnt = 0;
        for_block27_warp(g, q.x, q.y, q.z, lane, [&](int sp, float4, float d2) {
            if (d2 < r2) {
                float4 nj = __ldg(sn + sp);
                if (finite3(nj)) {
                    double x = nj.x, y = nj.y, z = nj.z;
                    c0 += x * x; c1 += x * y; c2 += x * z; c3 += y * y; c4 += y * z; c5 += z * z; ++cnt;
                }
            }
        });
        cnt = warp_sum(cnt);
        c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); c3 = warp_sum(c3); c4 = warp_sum(c4); c5 = warp_sum(c5);
        if (lane == 0) {
            float r = harris_from_sums(cnt, c0, c1, c2, c3, c4, c5);
            resp[__float_as_int(q.w)] = r;
            resp_sorted[s] = r;
        }
    }
}

template <class GS>
__global__ void __launch_bounds__(PW_WARPS * 32) k_harris_nms_warp(const __grid_constant__ GS gs, const float* __restrict__ resp_sorted, float r2, float thr,
                                                                  int nms, unsigned char* __restrict__ flags) {
    int lane = threadIdx.x & 31;
    int nwarps = gridDim.x * PW_WARPS;
    const int total = gs.total();
    for (int s = blockIdx.x * PW_WARPS + (threadIdx.x >> 5); s < total; s += nwarps) {
        const GridView g = gs.at(s);
        float4 q = __ldg(g.sorted + s);
        float r = resp_sorted[s];
        bool keep = isfinite(r) && (r >= thr);
        if (keep && nms) {       // warp-uniform branch: all lanes share r
            bool bigger = false;
            for_block27_warp(g, q.x, q.y, q.z, lane, [&](int sp, float4, float d2) {
                if (d2 < r2 && __ldg(resp_sorted + sp) > r) bigger = true;
            });
            keep = !__any_sync(0xffffffffu, bigger);
        }
        if (lane == 0) flags[__float_as_int(q.w)] = keep ? 1 : 0;
    }
}

// refineCorners: one warp per corner, lanes stride over the 9 candidate ranges.  The sums A = sum n n^T and
// b = sum n n^T p are accumulated in 64-bit FIXED POINT (each fp64 term rounded to 2^-40 resp. 2^-34): integer addition
// is associative, so the warp-shuffle reduction gives the same bits as the oracle's sequential loop in any order.  This
// stage is ill-conditioned (corners can move by tens of centimetres), which is why it is made exactly reproducible.
#define REFINE_SCALE_A 1099511627776.0     /* 2^40 */
#define REFINE_SCALE_B 17179869184.0       /* 2^34 */
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// One CTA (4 warps) per corner: warp 0 fetches the 18 range bounds of the 27-cell block in one round trip, all 128
// threads stride over the 9 ranges, the nine 64-bit integer sums are folded by shuffles and through shared memory
// (any order: integer addition), thread 0 solves the 3x3 system and publishes the moved corner.
#define REFINE_THREADS 128
// Model sets: blockIdx.y = member cloud; its corner list starts at the cloud's first point index (capacity = its size).
template <class GS>
__global__ void __launch_bounds__(REFINE_THREADS) k_harris_refine(const __grid_constant__ GS gs, const float4* __restrict__ sn, const float4* __restrict__ pts, float r2,
                                                                  const int* __restrict__ kp_idx_all, const int* __restrict__ kp_count, int capacity,
                                                                  float4* __restrict__ kp_xyz_all, int refine) {
    __shared__ int bounds[18];
    __shared__ long long wsum[REFINE_THREADS / 32][9];
    __shared__ float4 c_sh;
    __shared__ int again;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = blockIdx.y;
    const GridView g = gs.view(seg);
    const int* kp_idx = kp_idx_all + gs.seg_start(seg);
    float4* kp_xyz = kp_xyz_all + gs.seg_start(seg);
    int m = min(kp_count[seg], min(capacity, gs.seg_size(seg)));
    for (int t = blockIdx.x; t < m; t += gridDim.x) {
        float4 c = __ldg(pts + kp_idx[t]);
        c.w = 1.0f;
        if (refine) {
            int it = 0;
            for (;;) {
                float4 cur = c;
                long long q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                int cx = cell_coord(cur.x, g.mnx, g.inv_h), cy = cell_coord(cur.y, g.mny, g.inv_h), cz = cell_coord(cur.z, g.mnz, g.inv_h);
                bool near_grid = !(cx < -1 || cy < -1 || cz < -1 || cx > g.dx || cy > g.dy || cz > g.dz);
                if (near_grid) {
                    cx = clampi(cx, 0, g.dx - 1); cy = clampi(cy, 0, g.dy - 1); cz = clampi(cz, 0, g.dz - 1);
                    if (warp == 0) {
                        BlockRanges br = warp_block_ranges(g, cx, cy, cz, lane);
                        if (lane < 18) bounds[lane] = br.bound;      // slots of rows outside the grid hold 0 / 0: empty ranges
                    }
                    __syncthreads();
                    // the 9 ranges as ONE candidate list: every thread issues all its loads before the first use
                    int pre[10];
                    pre[0] = 0;
#pragma unroll
                    for (int r = 0; r < 9; ++r) pre[r + 1] = pre[r] + max(bounds[9 + r] - bounds[r], 0);
                    for (int j = tid; j < pre[9]; j += REFINE_THREADS) {
                        int start = 0, first = bounds[0];
#pragma unroll
                        for (int k = 1; k < 9; ++k) if (j >= pre[k]) { start = pre[k]; first = bounds[k]; }
                        int s = first + (j - start);
                        float4 p = __ldg(g.sorted + s);
                        float4 nj = __ldg(sn + s);          // fetched with the point, not after the distance test
                        if (dist2f(cur.x, cur.y, cur.z, p.x, p.y, p.z) < r2 && finite3(nj)) {
                            double x = nj.x, y2 = nj.y, z2 = nj.z;
                            double xx = x * x, xy = x * y2, xz = x * z2, yy = y2 * y2, yz = y2 * z2, zz = z2 * z2;
                            double px = p.x, py = p.y, pz = p.z;
                            q[0] += __double2ll_rn(xx * REFINE_SCALE_A); q[1] += __double2ll_rn(xy * REFINE_SCALE_A);
                            q[2] += __double2ll_rn(xz * REFINE_SCALE_A); q[3] += __double2ll_rn(yy * REFINE_SCALE_A);
                            q[4] += __double2ll_rn(yz * REFINE_SCALE_A); q[5] += __double2ll_rn(zz * REFINE_SCALE_A);
                            q[6] += __double2ll_rn(((xx * px + xy * py) + xz * pz) * REFINE_SCALE_B);
                            q[7] += __double2ll_rn(((xy * px + yy * py) + yz * pz) * REFINE_SCALE_B);
                            q[8] += __double2ll_rn(((xz * px + yz * py) + zz * pz) * REFINE_SCALE_B);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 9; ++k) { long long v = warp_sum_ll(q[k]); if (lane == 0) wsum[warp][k] = v; }
                __syncthreads();
                if (tid == 0) {
                    long long Q[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) Q[k] = ((wsum[0][k] + wsum[1][k]) + wsum[2][k]) + wsum[3][k];
                    double A0 = (double)Q[0] / REFINE_SCALE_A, A1 = (double)Q[1] / REFINE_SCALE_A, A2 = (double)Q[2] / REFINE_SCALE_A;
                    double A3 = (double)Q[3] / REFINE_SCALE_A, A4 = (double)Q[4] / REFINE_SCALE_A, A5 = (double)Q[5] / REFINE_SCALE_A;
                    double b0 = (double)Q[6] / REFINE_SCALE_B, b1 = (double)Q[7] / REFINE_SCALE_B, b2 = (double)Q[8] / REFINE_SCALE_B;
                    double c00 = A3 * A5 - A4 * A4, c01 = A2 * A4 - A1 * A5, c02 = A1 * A4 - A2 * A3;
                    double c11 = A0 * A5 - A2 * A2, c12 = A1 * A2 - A0 * A4, c22 = A0 * A3 - A1 * A1;
                    double det = (A0 * c00 + A1 * c01) + A2 * c02;
                    float4 nc = cur;
                    if (det != 0) {
                        nc.x = (float)(((c00 * b0 + c01 * b1) + c02 * b2) / det);
                        nc.y = (float)(((c01 * b0 + c11 * b1) + c12 * b2) / det);
                        nc.z = (float)(((c02 * b0 + c12 * b1) + c22 * b2) / det);
                    }
                    double ddx = (double)nc.x - (double)cur.x, ddy = (double)nc.y - (double)cur.y, ddz = (double)nc.z - (double)cur.z;
                    double diff = (ddx * ddx + ddy * ddy) + ddz * ddz;
                    c_sh = nc;
                    again = (diff > 1e-6 && it + 1 < 10) ? 1 : 0;
                }
                __syncthreads();
                c = c_sh;
                ++it;
                int go = again;
                __syncthreads();            // c_sh / again / bounds / wsum are rewritten by the next iteration
                if (!go) break;
            }
        }
        if (tid == 0) kp_xyz[t] = c;
    }
}

// ----------------------------------------------------------------------------- FPFH (App. A.4)
// computePairFeatures in fp64 with the |a1| < |a2| role swap (== acos|a1| > acos|a2|); returns the three bin indices.
__device__ __forceinline__ bool pair_bins(float4 p1, float4 n1f, float4 p2, float4 n2f, int& b0, int& b1, int& b2) {
    double d0 = (double)p2.x - (double)p1.x, d1 = (double)p2.y - (double)p1.y, d2 = (double)p2.z - (double)p1.z;
    double f4 = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    if (f4 == 0.0) return false;
    double n10 = n1f.x, n11 = n1f.y, n12 = n1f.z, n20 = n2f.x, n21 = n2f.y, n22 = n2f.z;
    double a1 = ((n10 * d0 + n11 * d1) + n12 * d2) / f4;
    double a2 = ((n20 * d0 + n21 * d1) + n22 * d2) / f4;
    double f3;
    if (fabs(a1) < fabs(a2)) {
        double t;
        t = n10; n10 = n20; n20 = t; t = n11; n11 = n21; n21 = t; t = n12; n12 = n22; n22 = t;
        d0 = -d0; d1 = -d1; d2 = -d2;
        f3 = -a2;
    } else f3 = a1;
    double v0 = d1 * n12 - d2 * n11, v1 = d2 * n10 - d0 * n12, v2 = d0 * n11 - d1 * n10;
    double vn = sqrt((v0 * v0 + v1 * v1) + v2 * v2);
    if (vn == 0.0) return false;
    v0 /= vn; v1 /= vn; v2 /= vn;
    double w0 = n11 * v2 - n12 * v1, w1 = n12 * v0 - n10 * v2, w2 = n10 * v1 - n11 * v0;
    double f2 = (v0 * n20 + v1 * n21) + v2 * n22;
    double f1 = atan2((w0 * n20 + w1 * n21) + w2 * n22, (n10 * n20 + n11 * n21) + n12 * n22);
    const double kPi = 3.14159265358979323846;
    int i0 = (int)floor(11.0 * ((f1 + kPi) * (1.0 / (2.0 * kPi))));
    int i1 = (int)floor(11.0 * ((f2 + 1.0) * 0.5));
    int i2 = (int)floor(11.0 * ((f3 + 1.0) * 0.5));
    b0 = min(max(i0, 0), 10); b1 = min(max(i1, 0), 10); b2 = min(max(i2, 0), 10);
    return true;
}

// fp32 screen for the three bin indices.  The bins are integers, so a cheap single-precision evaluation decides them
// whenever every feature sits safely inside a bin; only pairs within SCREEN_MARGIN of a bin edge, near-degenerate frames
// (d almost parallel to n1, n2 almost parallel to v) or an undecidable role swap go through the fp64 pair_bins above.
// Error budget (t = 11 x normalised feature): fp32 rounding of d, the dot / cross products and rsqrtf / atan2f gives
// |dt| < 1e-4 under the degeneracy guards below (sin^2 > 1e-2); the margin is 20x that.  The result is identical to
// pair_bins by construction — the tests compare whole histograms with the oracle bit for bit.
#define SCREEN_MARGIN 2e-3f
__device__ __forceinline__ float dot3f(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}
__device__ __forceinline__ bool screen_roles(float n10, float n11, float n12, float n20, float n21, float n22, float u0, float u1, float u2,
                                             float f3, int& b0, int& b1, int& b2) {
    float v0 = fmaf(u1, n12, -u2 * n11), v1 = fmaf(u2, n10, -u0 * n12), v2 = fmaf(u0, n11, -u1 * n10);
    float vn2 = dot3f(v0, v1, v2, v0, v1, v2);
    float iv = rsqrtf(vn2);
    v0 *= iv; v1 *= iv; v2 *= iv;
    float w0 = fmaf(n11, v2, -n12 * v1), w1 = fmaf(n12, v0, -n10 * v2), w2 = fmaf(n10, v1, -n11 * v0);
    float f2 = dot3f(v0, v1, v2, n20, n21, n22);
    float y = dot3f(w0, w1, w2, n20, n21, n22), x = dot3f(n10, n11, n12, n20, n21, n22);
    float f1 = atan2f(y, x);
    float t0 = fmaf(f1, 11.0f * 0.159154943f, 5.5f);
    float t1 = fmaf(f2, 5.5f, 5.5f);
    float t2 = fmaf(f3, 5.5f, 5.5f);
    float l0 = floorf(t0), l1 = floorf(t1), l2 = floorf(t2);
    float r0 = t0 - l0, r1 = t1 - l1, r2 = t2 - l2;
    // distance of every t to the nearest bin edge, and t inside (0, 11): one min chain, NaNs fail the comparison
    float m = fminf(fminf(fminf(r0, 1.0f - r0), fminf(r1, 1.0f - r1)), fminf(r2, 1.0f - r2));
    float lo = fminf(fminf(t0, t1), t2), hi = fmaxf(fmaxf(t0, t1), t2);
    bool ok = (m > SCREEN_MARGIN) & (lo > 0.f) & (hi < 11.f) & (vn2 > 1e-2f) & (fmaf(x, x, y * y) > 1e-2f);
    b0 = (int)l0; b1 = (int)l1; b2 = (int)l2;
    return ok;
}
__device__ __forceinline__ bool pair_bins_screen(float4 p1, float4 n1, float4 p2, float4 n2, int& b0, int& b1, int& b2) {
    float d0 = p2.x - p1.x, d1 = p2.y - p1.y, d2 = p2.z - p1.z;
    float ss = dot3f(d0, d1, d2, d0, d1, d2);
    if (!(ss > 1e-20f)) return false;
    float inv = rsqrtf(ss);
    float u0 = d0 * inv, u1 = d1 * inv, u2 = d2 * inv;
    float a1 = dot3f(n1.x, n1.y, n1.z, u0, u1, u2), a2 = dot3f(n2.x, n2.y, n2.z, u0, u1, u2);
    float diff = fabsf(a1) - fabsf(a2);
    if (fabsf(diff) > 1e-5f) {
        // role swap (|a1| < |a2|) decided: select the operands, one evaluation
        bool sw = diff < 0.f;
        float sg = sw ? -1.f : 1.f;
        return screen_roles(sw ? n2.x : n1.x, sw ? n2.y : n1.y, sw ? n2.z : n1.z, sw ? n1.x : n2.x, sw ? n1.y : n2.y, sw ? n1.z : n2.z,
                            sg * u0, sg * u1, sg * u2, sw ? -a2 : a1, b0, b1, b2);
    }
    // the swap cannot be decided in fp32 (flat neighbourhoods: both ~ 0): accept only if both roles give the same bins
    int c0, c1, c2;
    bool ok = screen_roles(n1.x, n1.y, n1.z, n2.x, n2.y, n2.z, u0, u1, u2, a1, b0, b1, b2) &
              screen_roles(n2.x, n2.y, n2.z, n1.x, n1.y, n1.z, -u0, -u1, -u2, -a2, c0, c1, c2);
    return ok && b0 == c0 && b1 == c1 && b2 == c2;
}

// computePointSPFHSignature: one warp per point, lanes stride over the 9 candidate ranges; the 3 x 11 bins are integer
// counters in shared memory (integer atomics: order independent), scaled by 100 / (|N| - 1) at the end.  Pairs the fp32
// screen cannot decide are queued per warp and evaluated in fp64 32 at a time (dense lanes instead of divergence).
#define SPFH_WARPS 8
__device__ __forceinline__ void spfh_exact_one(const GridView& g, const float4* __restrict__ sn, float4 q, float4 nq, int sp, int* cnt) {
    float4 p = __ldg(g.sorted + sp);
    float4 nj = __ldg(sn + sp);
    int b0, b1, b2;
    if (pair_bins(q, nq, p, nj, b0, b1, b2)) {
        atomicAdd(&cnt[b0], 1);
        atomicAdd(&cnt[11 + b1], 1);
        atomicAdd(&cnt[22 + b2], 1);
    }
}
// one in-radius candidate per lane: fp32 screen -> histogram, or flag it for the fp64 queue
__device__ __forceinline__ bool spfh_screen_one(const GridView& g, const float4* __restrict__ sn, float4 q, float4 nq, int sp, int use_screen, int* cnt) {
    float4 p = __ldg(g.sorted + sp);
    float4 nj = __ldg(sn + sp);
    if (!finite3(nj)) return false;
    int b0, b1, b2;
    if (use_screen && pair_bins_screen(q, nq, p, nj, b0, b1, b2)) {
        atomicAdd(&cnt[b0], 1);
        atomicAdd(&cnt[11 + b1], 1);
        atomicAdd(&cnt[22 + b2], 1);
        return false;
    }
    return true;
}
// MIN_CTAS 3: 80 registers, 24 warps / SM (4 M points: 2 -> 14.9 ms, 3 -> 12.5 ms, 4 -> spills, no faster)
template <int MIN_CTAS, class GS>
__global__ void __launch_bounds__(SPFH_WARPS * 32, MIN_CTAS) k_spfh(const __grid_constant__ GS gs, const float4* __restrict__ sn, float r2,
                                                             float* __restrict__ spfh_sorted, int use_screen,
                                                             const int* __restrict__ list, const int* __restrict__ list_count) {
    __shared__ int cnt[SPFH_WARPS][36];
    __shared__ int cq[SPFH_WARPS][64];      // in-radius candidates waiting for a dense batch of 32
    __shared__ int fq[SPFH_WARPS][64];      // pairs the screen could not decide
    __shared__ int wtab[SPFH_WARPS][18];    // flat candidate list of the current query (warp_candidates_smem)
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nwarps = gridDim.x * SPFH_WARPS;
    const unsigned lt = (1u << lane) - 1u;
    const int n_items = list ? *list_count : gs.total();       // list: sorted positions whose SPFH is wanted (rtr_fpfh_at), else all points
    for (int it = blockIdx.x * SPFH_WARPS + warp; it < n_items; it += nwarps) {
        const int s = list ? __ldg(list + it) : it;
        const GridView g = gs.at(s);
        cnt[warp][lane] = 0;
        if (lane < 4) cnt[warp][32 + lane] = 0;
        __syncwarp();
        float4 q = __ldg(g.sorted + s);
        float4 nq = __ldg(sn + s);
        int qi = __float_as_int(q.w);
        bool qfin = finite3(nq);
        int nb = 0, n_cand = 0, n_exact = 0;
        // drains 32 (or the last < 32) queued candidates; uniform control flow across the warp
        auto drain = [&](int take) {
            n_cand -= take;
            bool exact = false;
            int sp = 0;
            if (lane < take) { sp = cq[warp][n_cand + lane]; exact = spfh_screen_one(g, sn, q, nq, sp, use_screen, cnt[warp]); }
            unsigned em = __ballot_sync(0xffffffffu, exact);
            if (em) {
                if (exact) fq[warp][n_exact + __popc(em & lt)] = sp;
                n_exact += __popc(em);
                __syncwarp();
                if (n_exact >= 32) {
                    n_exact -= 32;
                    spfh_exact_one(g, sn, q, nq, fq[warp][n_exact + lane], cnt[warp]);
                    __syncwarp();
                }
            }
        };
        int cx = clampi(cell_coord(q.x, g.mnx, g.inv_h), 0, g.dx - 1);
        int cy = clampi(cell_coord(q.y, g.mny, g.inv_h), 0, g.dy - 1);
        int cz = clampi(cell_coord(q.z, g.mnz, g.inv_h), 0, g.dz - 1);
        const int total = warp_candidates_smem(g, cx, cy, cz, lane, wtab[warp]);   // the 9 ranges as one flat list: dense batches of 32
        // this lane's current range (its j only grows, so the range only advances): begin / end of the range in list positions
        // and its first cell-order position live in registers, so the common step is one compare (re-reading the table for
        // every candidate was 10.7 % of this kernel's instructions on the bench batch)
        int kr = 0, rbeg = 0, rend = wtab[warp][1], rfirst = wtab[warp][9];
        for (int j0 = 0; j0 < total; j0 += 32) {
            const int j = j0 + lane;
            int sp = 0;
            bool in = false, want = false;
            if (j < total) {
                while (j >= rend && kr < 8) { ++kr; rbeg = rend; rend = wtab[warp][kr + 1]; rfirst = wtab[warp][9 + kr]; }
                sp = rfirst + (j - rbeg);
                float4 p = __ldg(g.sorted + sp);
                in = dist2f(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
                want = in && qfin && __float_as_int(p.w) != qi;
            }
            nb += __popc(__ballot_sync(0xffffffffu, in));          // warp-uniform count
            unsigned wm = __ballot_sync(0xffffffffu, want);
            if (wm) {
                if (want) cq[warp][n_cand + __popc(wm & lt)] = sp;
                n_cand += __popc(wm);
                __syncwarp();
                if (n_cand >= 32) { drain(32); __syncwarp(); }
            }
        }
        if (n_cand > 0) { drain(n_cand); __syncwarp(); }
        if (n_exact > 0) {
            if (lane < n_exact) spfh_exact_one(g, sn, q, nq, fq[warp][lane], cnt[warp]);
        }
        __syncwarp();
        float* o = spfh_sorted + (size_t)s * 33;
        if (nb < 2 || !qfin) {
            o[lane] = 0.f;
            if (lane == 0) o[32] = 0.f;
        } else {
            double incr = 100.0 / (double)(nb - 1);
            o[lane] = (float)((double)cnt[warp][lane] * incr);
            if (lane == 0) o[32] = (float)((double)cnt[warp][32] * incr);
        }
        __syncwarp();
    }
}

// weightPointSPFHSignature: one warp per point, two roles.  Lane = candidate: distance test, the fp64 weight 1/d2 and
// bin 32's term, computed once per pair; accepted candidates are appended IN ORDER to a 64-entry ring in shared memory.
// Lane = histogram bin: whenever 32 entries wait, the warp walks them in order (broadcast reads of (w, term32, row),
// four SPFH rows in flight), so the additions happen in ascending candidate position exactly as the oracle's loop.
#define FPFH_WARPS 8
struct __align__(16) FwEntry { double w, t32; };
template <class GS>
__global__ void __launch_bounds__(FPFH_WARPS * 32) k_fpfh_weight(const __grid_constant__ GS gs, const float* __restrict__ spfh_sorted, float r2,
                                                                 float* __restrict__ fpfh, const int* __restrict__ list, int n_list) {
    __shared__ double hs[FPFH_WARPS][36];
    __shared__ FwEntry wq[FPFH_WARPS][64];
    __shared__ int rq[FPFH_WARPS][64];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nwarps = gridDim.x * FPFH_WARPS;
    const unsigned lt = (1u << lane) - 1u;
    const float* rows = spfh_sorted + lane;
    const int n_items = list ? n_list : gs.total();            // list: sorted positions of the query points (rtr_fpfh_at), row it of the output
    for (int it = blockIdx.x * FPFH_WARPS + warp; it < n_items; it += nwarps) {
        const int s = list ? __ldg(list + it) : it;
        const GridView g = gs.at(s);
        float4 q = __ldg(g.sorted + s);
        double acc = 0, acc32 = 0;
        int nb = 0, head = 0, nq = 0;
        auto consume = [&](int cnt) {
            int k = 0;
            for (; k + 4 <= cnt; k += 4) {
                int e = (head + k) & 63;        // head is a multiple of 32 and k of 4: the four entries do not wrap
                int r0 = rq[warp][e], r1 = rq[warp][e + 1], r2i = rq[warp][e + 2], r3 = rq[warp][e + 3];
                float v0 = __ldg(rows + (size_t)r0 * 33), v1 = __ldg(rows + (size_t)r1 * 33);
                float v2 = __ldg(rows + (size_t)r2i * 33), v3 = __ldg(rows + (size_t)r3 * 33);
                FwEntry a0 = wq[warp][e], a1 = wq[warp][e + 1], a2 = wq[warp][e + 2], a3 = wq[warp][e + 3];
                acc += (double)v0 * a0.w; acc += (double)v1 * a1.w; acc += (double)v2 * a2.w; acc += (double)v3 * a3.w;
                acc32 += a0.t32; acc32 += a1.t32; acc32 += a2.t32; acc32 += a3.t32;
            }
            for (; k < cnt; ++k) {
                int e = (head + k) & 63;
                float v = __ldg(rows + (size_t)rq[warp][e] * 33);
                FwEntry a = wq[warp][e];
                acc += (double)v * a.w;
                acc32 += a.t32;
            }
            head = (head + cnt) & 63;
            nq -= cnt;
        };
        int cx = clampi(cell_coord(q.x, g.mnx, g.inv_h), 0, g.dx - 1);
        int cy = clampi(cell_coord(q.y, g.mny, g.inv_h), 0, g.dy - 1);
        int cz = clampi(cell_coord(q.z, g.mnz, g.inv_h), 0, g.dz - 1);
        int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
        for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dz - 1); ++z)
            for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dy - 1); ++y) {
                int s0 = __ldg(g.cell_begin + cell_key(g, x0, y, z));
                int s1 = __ldg(g.cell_begin + cell_key(g, x1, y, z) + 1);
                for (int base = s0; base < s1; base += 32) {
                    int sp = base + lane;
                    bool in = false, use = false;
                    float d2 = 0.f;
                    if (sp < s1) {
                        float4 p = __ldg(g.sorted + sp);
                        d2 = dist2f(q.x, q.y, q.z, p.x, p.y, p.z);
                        in = d2 < r2;
                        use = in && d2 != 0.f;      // "minus the query point itself": dists == 0 skipped
                    }
                    nb += __popc(__ballot_sync(0xffffffffu, in));
                    unsigned mask = __ballot_sync(0xffffffffu, use);
                    if (mask) {
                        if (use) {
                            double w = __drcp_rn((double)d2);        // == 1.0 / d2 (both correctly rounded), without the general division's fix-up path
                            int e = (head + nq + __popc(mask & lt)) & 63;
                            FwEntry en;
                            en.w = w;
                            en.t32 = (double)__ldg(spfh_sorted + (size_t)sp * 33 + 32) * w;      // bin 32's term
                            wq[warp][e] = en;
                            rq[warp][e] = sp;
                        }
                        nq += __popc(mask);
                        __syncwarp();
                        if (nq >= 32) { consume(32); __syncwarp(); }
                    }
                }
            }
        if (nq > 0) { consume(nq); __syncwarp(); }
        hs[warp][lane] = acc;
        if (lane == 0) hs[warp][32] = acc32;
        __syncwarp();
        if (lane < 3) {
            double sum = 0;
            for (int k = 0; k < 11; ++k) sum += hs[warp][lane * 11 + k];
            hs[warp][33 + lane] = (sum != 0) ? 100.0 / sum : 0.0;
        }
        __syncwarp();
        float* o = fpfh + (size_t)(list ? it : __float_as_int(q.w)) * 33;
        if (nb == 0) {
            o[lane] = __int_as_float(0x7fc00000);
            if (lane == 0) o[32] = __int_as_float(0x7fc00000);
        } else {
            o[lane] = (float)(hs[warp][lane] * hs[warp][33 + lane / 11]);
            if (lane == 0) o[32] = (float)(hs[warp][32] * hs[warp][35]);
        }
        __syncwarp();
    }
}

// Tiled form for larger / denser clouds: all points of a cell share the same 27-cell candidate block, so one CTA per
// occupied cell stages the candidates' positions and SPFH rows through shared memory ONCE per tile instead of once per
// query (the untiled kernel moves 132 B x neighbours per query through L2: 2.8 GB for an 18 779-point dense cloud).
// Every query's 33 running sums live in shared memory between tiles, so the additions happen in exactly the same order
// (ranges in (z, y) order, positions ascending) as in k_fpfh_weight: the results are bit-identical.
#define FW_TILE 128
#define FW_QCHUNK 128
#define FW_WARPS 8
struct FwSmem {
    float4 tpos[FW_TILE];
    float tsp[FW_TILE * 33];
    double acc[FW_QCHUNK][33];
    int nbc[FW_QCHUNK];
    FwEntry wq[FW_WARPS][64];       // per warp: accepted candidates of the current (query, tile), in candidate order
    int rq[FW_WARPS][64];
    int tidx[FW_TILE];              // cell-order position of every tile entry
    int rb[9], pre[10];             // the cell's 9 candidate ranges: begin, exclusive prefix of the lengths
};
__global__ void k_occupied_cells(const int* __restrict__ cell_begin, int ncells, int* __restrict__ cells, int* __restrict__ count) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells && cell_begin[c + 1] > cell_begin[c]) cells[atomicAdd(count, 1)] = c;
}
__global__ void __launch_bounds__(FW_WARPS * 32) k_fpfh_weight_tiled(GridView g, const float* __restrict__ spfh_sorted, float r2,
                                                                     float* __restrict__ fpfh, const int* __restrict__ cells,
                                                                     const int* __restrict__ n_cells) {
    extern __shared__ __align__(16) unsigned char fw_raw[];
    FwSmem& sm = *reinterpret_cast<FwSmem*>(fw_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int nc = *n_cells;
    for (int ci = blockIdx.x; ci < nc; ci += gridDim.x) {
        const int c = cells[ci];
        const int cz = c / (g.dx * g.dy), cy = (c - cz * g.dx * g.dy) / g.dx, cx = c - cz * g.dx * g.dy - cy * g.dx;
        const int qb = __ldg(g.cell_begin + c), qe = __ldg(g.cell_begin + c + 1);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
        // the cell's 27-block as ONE flat candidate list (9 ranges in (z, y) order), cut into tiles of FW_TILE that may span
        // ranges: at ~35 candidates per range a tile per range meant 9 small loads and 27 barriers per cell instead of 3 and 9
        __syncthreads();
        if (threadIdx.x < 9) {
            const int zi = threadIdx.x / 3, z = cz - 1 + zi, y = cy - 1 + (threadIdx.x - zi * 3);
            int s0 = 0, s1 = 0;
            if (z >= 0 && z < g.dz && y >= 0 && y < g.dy) {
                s0 = __ldg(g.cell_begin + cell_key(g, x0, y, z));
                s1 = __ldg(g.cell_begin + cell_key(g, x1, y, z) + 1);
            }
            sm.rb[threadIdx.x] = s0;
            sm.pre[threadIdx.x + 1] = s1 - s0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            sm.pre[0] = 0;
            for (int r = 0; r < 9; ++r) sm.pre[r + 1] += sm.pre[r];
        }
        __syncthreads();
        const int total = sm.pre[9];
        for (int q0 = qb; q0 < qe; q0 += FW_QCHUNK) {
            const int nq = min(FW_QCHUNK, qe - q0);
            __syncthreads();
            for (int e = threadIdx.x; e < nq * 33; e += blockDim.x) (&sm.acc[0][0])[e] = 0.0;
            for (int e = threadIdx.x; e < nq; e += blockDim.x) sm.nbc[e] = 0;
                {
                    for (int T0 = 0; T0 < total; T0 += FW_TILE) {
                        const int nt = min(FW_TILE, total - T0);
                        __syncthreads();
                        for (int i = threadIdx.x; i < nt; i += blockDim.x) {
                            const int j = T0 + i;
                            int k = 0;
#pragma unroll
                            for (int r = 1; r < 9; ++r) if (j >= sm.pre[r]) k = r;
                            const int pos = sm.rb[k] + (j - sm.pre[k]);
                            sm.tidx[i] = pos;
                            sm.tpos[i] = __ldg(g.sorted + pos);
                        }
                        __syncthreads();
                        for (int e = threadIdx.x; e < nt * 33; e += blockDim.x) {
                            const int i = e / 33;
                            sm.tsp[e] = __ldg(spfh_sorted + (size_t)sm.tidx[i] * 33 + (e - i * 33));
                        }
                        __syncthreads();
                        for (int qi = warp; qi < nq; qi += FW_WARPS) {
                            const float4 q = __ldg(g.sorted + q0 + qi);
                            double a = sm.acc[qi][lane], a32 = sm.acc[qi][32];
                            int nb = 0, head = 0, nring = 0;
                            // lane = bin: drains cnt ring entries in order (broadcast reads of (w, term32, row), four rows in flight)
                            auto consume = [&](int cnt) {
                                int k = 0;
                                for (; k + 4 <= cnt; k += 4) {
                                    const int e = (head + k) & 63;      // head is 0 or 32 and k a multiple of 4: no wrap inside the four
                                    const int r0 = sm.rq[warp][e], r1 = sm.rq[warp][e + 1], r2i = sm.rq[warp][e + 2], r3 = sm.rq[warp][e + 3];
                                    const float v0 = sm.tsp[r0 * 33 + lane], v1 = sm.tsp[r1 * 33 + lane];
                                    const float v2 = sm.tsp[r2i * 33 + lane], v3 = sm.tsp[r3 * 33 + lane];
                                    const FwEntry e0 = sm.wq[warp][e], e1 = sm.wq[warp][e + 1], e2 = sm.wq[warp][e + 2], e3 = sm.wq[warp][e + 3];
                                    a += (double)v0 * e0.w; a += (double)v1 * e1.w; a += (double)v2 * e2.w; a += (double)v3 * e3.w;
                                    a32 += e0.t32; a32 += e1.t32; a32 += e2.t32; a32 += e3.t32;
                                }
                                for (; k < cnt; ++k) {
                                    const int e = (head + k) & 63;
                                    const FwEntry en = sm.wq[warp][e];
                                    a += (double)sm.tsp[sm.rq[warp][e] * 33 + lane] * en.w;
                                    a32 += en.t32;
                                }
                                head = (head + cnt) & 63;
                                nring -= cnt;
                            };
                            for (int base = 0; base < nt; base += 32) {
                                const int i = base + lane;
                                bool in = false, use = false;
                                float d2 = 0.f;
                                if (i < nt) {
                                    const float4 p = sm.tpos[i];
                                    d2 = dist2f(q.x, q.y, q.z, p.x, p.y, p.z);
                                    in = d2 < r2;
                                    use = in && d2 != 0.f;      // "minus the query point itself": dists == 0 skipped
                                }
                                nb += __popc(__ballot_sync(0xffffffffu, in));
                                const unsigned mask = __ballot_sync(0xffffffffu, use);
                                if (mask) {
                                    if (use) {                  // lane = candidate: weight and bin 32's term, once per pair
                                        const double w = __drcp_rn((double)d2);
                                        const int e = (head + nring + __popc(mask & lt)) & 63;
                                        FwEntry en;
                                        en.w = w;
                                        en.t32 = (double)sm.tsp[i * 33 + 32] * w;
                                        sm.wq[warp][e] = en;
                                        sm.rq[warp][e] = i;
                                    }
                                    nring += __popc(mask);
                                    __syncwarp();
                                    if (nring >= 32) { consume(32); __syncwarp(); }
                                }
                            }
                            if (nring > 0) consume(nring);
                            __syncwarp();               // every lane has read acc[qi][32] before lane 0 rewrites it
                            sm.acc[qi][lane] = a;
                            if (lane == 0) { sm.acc[qi][32] = a32; sm.nbc[qi] += nb; }
                        }
                    }
                }
            __syncthreads();
            // normalise each third to 100 and store (same arithmetic as k_fpfh_weight)
            for (int qi = warp; qi < nq; qi += FW_WARPS) {
                double sc = 0.0;
                if (lane < 3) {
                    double sum = 0;
                    for (int k = 0; k < 11; ++k) sum += sm.acc[qi][lane * 11 + k];
                    sc = (sum != 0) ? 100.0 / sum : 0.0;
                }
                const double s0_ = __shfl_sync(0xffffffffu, sc, 0), s1_ = __shfl_sync(0xffffffffu, sc, 1), s2_ = __shfl_sync(0xffffffffu, sc, 2);
                const double mine = lane < 11 ? s0_ : (lane < 22 ? s1_ : s2_);
                float* o = fpfh + (size_t)__float_as_int(__ldg(g.sorted + q0 + qi).w) * 33;
                if (sm.nbc[qi] == 0) {
                    o[lane] = __int_as_float(0x7fc00000);
                    if (lane == 0) o[32] = __int_as_float(0x7fc00000);
                } else {
                    o[lane] = (float)(sm.acc[qi][lane] * mine);
                    if (lane == 0) o[32] = (float)(sm.acc[qi][32] * s2_);
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------- feature k-NN (App. A.5), exact
// One warp per source feature; targets are staged through shared memory in tiles of 64 rows (row stride 33 floats is
// conflict-free); each lane keeps a sorted top-K of the rows it scanned; the warp then merges the 32 lists.
// Distances are the oracle's: sequential fp64 sum of squared fp64 differences, rounded to fp32; ties -> lowest index.
#define MATCH_WARPS 8
#define MATCH_TILE 64
#define MATCH_KMAX 8
__global__ void __launch_bounds__(MATCH_WARPS * 32) k_match(const float* __restrict__ fa, int na, const float* __restrict__ fb,
                                                            int nb, int k, int* __restrict__ out_idx, float* __restrict__ out_dist,
                                                            const int* __restrict__ rows, const int* __restrict__ row_count, int row_offset) {
    __shared__ float tile[MATCH_TILE * 33];
    __shared__ float src[MATCH_WARPS][33];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // optional row list (rows the tensor-core prefilter could not certify): same kernel, indirect row index
    int nrows = rows ? *row_count : na;
    for (int rbase = row_offset + blockIdx.x * MATCH_WARPS; rbase < nrows; rbase += gridDim.x * MATCH_WARPS) {
        int li = rbase + warp;
        bool active = li < nrows;
        int row = active ? (rows ? rows[li] : li) : 0;
        __syncwarp();
        if (active) {
            src[warp][lane] = fa[(size_t)row * 33 + lane];
            if (lane == 0) src[warp][32] = fa[(size_t)row * 33 + 32];
        }
        float bd[MATCH_KMAX];
        int bi[MATCH_KMAX];
#pragma unroll
        for (int t = 0; t < MATCH_KMAX; ++t) { bd[t] = FLT_MAX; bi[t] = 0x7fffffff; }
        for (int base = 0; base < nb; base += MATCH_TILE) {
            __syncthreads();
            int rows_in_tile = min(MATCH_TILE, nb - base);
            for (int e = threadIdx.x; e < rows_in_tile * 33; e += blockDim.x) tile[e] = __ldg(fb + (size_t)base * 33 + e);
            __syncthreads();
            if (active) {
#pragma unroll
                for (int half = 0; half < MATCH_TILE / 32; ++half) {
                    int r = half * 32 + lane;
                    if (r < rows_in_tile) {
                        const float* b = tile + r * 33;
                        double sacc = 0;
#pragma unroll
                        for (int c = 0; c < 33; ++c) { double d = (double)src[warp][c] - (double)b[c]; sacc += d * d; }
                        float df = (float)sacc;
                        int id = base + r;
                        if (df == df && (df < bd[MATCH_KMAX - 1] || (df == bd[MATCH_KMAX - 1] && id < bi[MATCH_KMAX - 1]))) {
                            bd[MATCH_KMAX - 1] = df; bi[MATCH_KMAX - 1] = id;
#pragma unroll
                            for (int t = MATCH_KMAX - 1; t > 0; --t) {
                                bool sw = (bd[t] < bd[t - 1]) || (bd[t] == bd[t - 1] && bi[t] < bi[t - 1]);
                                if (sw) { float td = bd[t]; bd[t] = bd[t - 1]; bd[t - 1] = td; int ti = bi[t]; bi[t] = bi[t - 1]; bi[t - 1] = ti; }
                            }
                        }
                    }
                }
            }
        }
        if (!active) continue;
        // merge: k rounds of a warp-wide lexicographic (dist, idx) minimum over the list heads
        for (int t = 0; t < k; ++t) {
            float hd = bd[0]; int hi = bi[0];
            float md = hd; int mi = hi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float od = __shfl_xor_sync(0xffffffffu, md, o);
                int oi = __shfl_xor_sync(0xffffffffu, mi, o);
                if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
            }
            if (mi == hi && md == hd && hi != 0x7fffffff) {   // this lane owned the winner: pop it
#pragma unroll
                for (int u = 0; u < MATCH_KMAX - 1; ++u) { bd[u] = bd[u + 1]; bi[u] = bi[u + 1]; }
                bd[MATCH_KMAX - 1] = FLT_MAX; bi[MATCH_KMAX - 1] = 0x7fffffff;
            }
            if (lane == 0) {
                bool none = (mi == 0x7fffffff);
                out_idx[(size_t)row * k + t] = none ? -1 : mi;
                out_dist[(size_t)row * k + t] = none ? __int_as_float(0x7fc00000) : md;
            }
        }
    }
}

// Redo of the few rows the tensor-core prefilter could not certify.  k_match gives a row to ONE warp, fine when every row is
// searched (thousands of warps), but 240 uncertified rows of the 262 144 x 65 536 search were 240 warps scanning 65 536
// targets each: 5.0 ms of the 12.0 ms search with 12 % of the warp slots filled (ncu, profiles/r02).  Here a row goes to
// MS_SPLITS CTAs of 8 warps, every warp scanning its own slice of the targets (same sequential fp64 distance, same
// per-lane top-8), the CTA's 256 lists are merged by block-wide lexicographic (distance, index) minima, and the last CTA of
// a row to finish (ticket) merges the MS_SPLITS partial lists.  Same distances, same tie rule: same result.
#define MS_SPLITS 16
__global__ void __launch_bounds__(MATCH_WARPS * 32) k_match_rows_split(const float* __restrict__ fa, const float* __restrict__ fb, int nb, int k,
                                                                       int* __restrict__ out_idx, float* __restrict__ out_dist,
                                                                       const int* __restrict__ rows, const int* __restrict__ row_count,
                                                                       float* __restrict__ part_d, int* __restrict__ part_i, unsigned* __restrict__ tickets,
                                                                       int max_rows) {
    __shared__ float src[33];
    __shared__ float red_d[MATCH_WARPS];
    __shared__ int red_i[MATCH_WARPS];
    __shared__ int is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nrows = min(*row_count, max_rows);      // rows beyond the partial-list capacity are left to k_match
    for (int li = blockIdx.x; li < nrows; li += gridDim.x) {
        const int row = rows[li], split = blockIdx.y;
        __syncthreads();
        if (threadIdx.x < 33) src[threadIdx.x] = fa[(size_t)row * 33 + threadIdx.x];
        __syncthreads();
        float bd[MATCH_KMAX]; int bi[MATCH_KMAX];
#pragma unroll
        for (int t = 0; t < MATCH_KMAX; ++t) { bd[t] = FLT_MAX; bi[t] = 0x7fffffff; }
        // slice of the targets of this CTA, strided over its 256 threads
        const int per = (nb + MS_SPLITS - 1) / MS_SPLITS, j0 = split * per, j1 = min(nb, j0 + per);
        for (int j = j0 + (int)threadIdx.x; j < j1; j += MATCH_WARPS * 32) {
            const float* b = fb + (size_t)j * 33;
            double sacc = 0;
#pragma unroll
            for (int c = 0; c < 33; ++c) { double d = (double)src[c] - (double)__ldg(b + c); sacc += d * d; }
            const float df = (float)sacc;
            if (df == df && (df < bd[MATCH_KMAX - 1] || (df == bd[MATCH_KMAX - 1] && j < bi[MATCH_KMAX - 1]))) {
                bd[MATCH_KMAX - 1] = df; bi[MATCH_KMAX - 1] = j;
#pragma unroll
                for (int t = MATCH_KMAX - 1; t > 0; --t) {
                    bool sw = (bd[t] < bd[t - 1]) || (bd[t] == bd[t - 1] && bi[t] < bi[t - 1]);
                    if (sw) { float td = bd[t]; bd[t] = bd[t - 1]; bd[t - 1] = td; int ti = bi[t]; bi[t] = bi[t - 1]; bi[t - 1] = ti; }
                }
            }
        }
        // k rounds of a block-wide lexicographic minimum over the 256 list heads; the owner of the winner pops it
        float* pd = part_d + ((size_t)li * MS_SPLITS + split) * MATCH_KMAX;
        int* pi = part_i + ((size_t)li * MS_SPLITS + split) * MATCH_KMAX;
        for (int t = 0; t < k; ++t) {
            float md = bd[0]; int mi = bi[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float od = __shfl_xor_sync(0xffffffffu, md, o);
                int oi = __shfl_xor_sync(0xffffffffu, mi, o);
                if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
            }
            if (lane == 0) { red_d[warp] = md; red_i[warp] = mi; }
            __syncthreads();
            float gd = red_d[0]; int gi = red_i[0];
#pragma unroll
            for (int w = 1; w < MATCH_WARPS; ++w) if (red_d[w] < gd || (red_d[w] == gd && red_i[w] < gi)) { gd = red_d[w]; gi = red_i[w]; }
            if (gi != 0x7fffffff && bi[0] == gi && bd[0] == gd) {      // target indices are unique: exactly one thread owns the winner
#pragma unroll
                for (int u = 0; u < MATCH_KMAX - 1; ++u) { bd[u] = bd[u + 1]; bi[u] = bi[u + 1]; }
                bd[MATCH_KMAX - 1] = FLT_MAX; bi[MATCH_KMAX - 1] = 0x7fffffff;
            }
            if (threadIdx.x == 0) { pd[t] = gd; pi[t] = gi; }
            __syncthreads();
        }
        // the last CTA of this row merges the MS_SPLITS sorted partial lists (k entries each)
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned tk = atomicAdd(tickets + li, 1u);
            is_last = (tk == MS_SPLITS - 1);
        }
        __syncthreads();
        if (is_last && warp == 0) {
            __threadfence();
            const float* rd = part_d + (size_t)li * MS_SPLITS * MATCH_KMAX;
            const int* ri = part_i + (size_t)li * MS_SPLITS * MATCH_KMAX;
            int head = 0;                           // lane s < MS_SPLITS walks the sorted list of split s
            for (int t = 0; t < k; ++t) {
                float md = FLT_MAX; int mi = 0x7fffffff;
                if (lane < MS_SPLITS && head < k) { md = __ldcg(rd + lane * MATCH_KMAX + head); mi = __ldcg(ri + lane * MATCH_KMAX + head); }
                float gd = md; int gi = mi;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    float od = __shfl_xor_sync(0xffffffffu, gd, o);
                    int oi = __shfl_xor_sync(0xffffffffu, gi, o);
                    if (od < gd || (od == gd && oi < gi)) { gd = od; gi = oi; }
                }
                if (gi != 0x7fffffff && mi == gi && md == gd) ++head;
                if (lane == 0) {
                    const bool none = (gi == 0x7fffffff);
                    out_idx[(size_t)row * k + t] = none ? -1 : gi;
                    out_dist[(size_t)row * k + t] = none ? __int_as_float(0x7fc00000) : gd;
                }
            }
            if (lane == 0) tickets[li] = 0u;
        }
    }
}

// exact search for all rows (rows == nullptr) or for a device-side list of rows
int rtr_match_exact_launch(rtr_context* ctx, const float* fa, int na, const float* fb, int nb, int k, int* out_idx, float* out_dist,
                           const int* rows, const int* row_count, int max_rows) {
    if (max_rows <= 0) return 0;
    if (rows && nb >= 4096) {
        // a (short) row list against many targets: MS_SPLITS CTAs per row, for the first `cap` rows of the list (the partial
        // lists are sized for that many; a longer list — the certificate failing wholesale — finishes in k_match below)
        const int cap = std::min(max_rows, 16384);
        float* part_d = nullptr; int* part_i = nullptr; unsigned* tickets = nullptr;
        if (int e = tmp_alloc(ctx, &part_d, (size_t)cap * MS_SPLITS * MATCH_KMAX, "match")) return e;
        if (int e = tmp_alloc(ctx, &part_i, (size_t)cap * MS_SPLITS * MATCH_KMAX, "match")) return e;
        if (int e = tmp_alloc(ctx, &tickets, (size_t)cap, "match")) return e;
        RTR_CHECK(cudaMemsetAsync(tickets, 0, sizeof(unsigned) * (size_t)cap, ctx->stream), "match");
        k_match_rows_split<<<dim3(std::min(cap, ctx->sm_count * 2), MS_SPLITS), MATCH_WARPS * 32, 0, ctx->stream>>>(fa, fb, nb, k, out_idx, out_dist, rows, row_count,
                                                                                                              part_d, part_i, tickets, cap);
        RTR_LAUNCH_CHECK(ctx, "match.redo");
        dev_free(ctx, part_d); dev_free(ctx, part_i); dev_free(ctx, tickets);
        if (max_rows > cap) {
            int grid = std::min(nblk(max_rows - cap, MATCH_WARPS), ctx->sm_count * 8);
            k_match<<<grid, MATCH_WARPS * 32, 0, ctx->stream>>>(fa, na, fb, nb, k, out_idx, out_dist, rows, row_count, cap);
            RTR_LAUNCH_CHECK(ctx, "match.redo");
        }
        return 0;
    }
    int grid = std::min(nblk(max_rows, MATCH_WARPS), ctx->sm_count * 8);
    k_match<<<grid, MATCH_WARPS * 32, 0, ctx->stream>>>(fa, na, fb, nb, k, out_idx, out_dist, rows, row_count, 0);
    RTR_LAUNCH_CHECK(ctx, rows ? "match.redo" : "match");
    return 0;
}

bool rtr_match_tc_wanted(long long ns, long long nt);
int rtr_match_tc_dev(rtr_context* ctx, const float* fa, int ns, const float* fb, int nt, int k, int* out_idx, float* out_dist, int* stats);

// feature k-NN dispatcher: tensor-core prefilter + exact re-rank for large problems, exact SIMT kernel otherwise
static int match_dispatch(rtr_context* ctx, const float* fa, int na, const float* fb, int nb, int k, int* out_idx, float* out_dist, int* stats) {
    RtrRange nvtx_range("rtr.match_features");
    if (stats) { stats[0] = -1; stats[1] = 0; stats[2] = 0; }
    if (rtr_match_tc_wanted(na, nb)) return rtr_match_tc_dev(ctx, fa, na, fb, nb, k, out_idx, out_dist, stats);
    return rtr_match_exact_launch(ctx, fa, na, fb, nb, k, out_idx, out_dist, nullptr, nullptr, na);
}

// Model sets: stable compaction of the corner flags (original order) per member cloud, one CTA per cloud — the ascending
// index lists cub::DeviceSelect::Flagged gives a single cloud, for all clouds in one launch.  The list of cloud k starts at
// out_idx[begin[k]]; counts[k] receives its length.
__global__ void __launch_bounds__(1024) k_select_segments(const __grid_constant__ ManyGrids mg, const unsigned char* __restrict__ flags,
                                                          int* __restrict__ out_idx, int* __restrict__ counts) {
    __shared__ int warp_tot[32];
    const int seg = blockIdx.x, p0 = mg.begin[seg], p1 = mg.begin[seg + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int base = 0;
    for (int c0 = p0; c0 < p1; c0 += 1024) {
        const int i = c0 + tid;
        const bool f = i < p1 && flags[i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
        for (int w = 0; w < 32; ++w) { const int v = warp_tot[w]; all += v; if (w < warp) before += v; }
        if (f) out_idx[p0 + base + before + __popc(m & ((1u << lane) - 1u))] = i;
        base += all;
        __syncthreads();
    }
    if (tid == 0) counts[seg] = base;
}

int rtr_match_features_dev(rtr_context* ctx, const float* fa, int na, const float* fb, int nb, int k, int* out_idx, float* out_dist) {
    if (k < 1 || k > MATCH_KMAX) return rtr_fail("match", "k must be in [1, 8]", RTR_ERR_INVALID);
    if (na <= 0) return 0;
    return match_dispatch(ctx, fa, na, fb, nb, k, out_idx, out_dist, nullptr);
}

// ----------------------------------------------------------------------------- host drivers
int rtr_normals_dev(rtr_cloud* c, float radius) { return rtr_normals_mode_dev(c, radius, 0); }

// mode 0: exact (fp64 sums of offsets, Jacobi) — the parity mode every downstream test is pinned to; mode 1: PCL-float-faithful
int rtr_normals_mode_dev(rtr_cloud* c, float radius, int mode) {
    RtrRange nvtx_range("rtr.normals");
    rtr_context* ctx = c->ctx;
    if (c->normals && c->normals_radius == radius && c->normals_mode == mode) return 0;
    if (mode == 1) {
        DevGrid* g1;
        if (int e = rtr_get_grid(c, radius, &g1)) return e;
        if (!c->normals) if (int e = dev_alloc(ctx, &c->normals, c->n, "normals")) return e;
        if (c->n > 0) {
            if (c->nseg() > 0) k_normals_pcl_float<ManyGrids><<<nblk(c->n, 128), 128, 0, ctx->stream>>>(rtr_many(g1, c), radius * radius, c->normals);
            else k_normals_pcl_float<OneGrid><<<nblk(c->n, 128), 128, 0, ctx->stream>>>(rtr_one(g1), radius * radius, c->normals);
            RTR_LAUNCH_CHECK(ctx, "normals.pcl_float");
        }
        c->normals_radius = radius; c->normals_mode = 1;
        c->normals_version++;
        return 0;
    }
    DevGrid* g;
    if (int e = rtr_get_grid(c, radius, &g)) return e;
    if (!c->normals) if (int e = dev_alloc(ctx, &c->normals, c->n, "normals")) return e;
    if (c->n > 0) {
        const bool many = c->nseg() > 0;
        if (many || c->n <= env_threshold("RTR_WARP_PER_POINT_MAX", RTR_WARP_PER_POINT_MAX)) {
            double* sums = nullptr;
            if (int e = tmp_alloc(ctx, &sums, (size_t)c->n * 10, "normals")) return e;
            const int grid = std::min(nblk(c->n, PW_WARPS), ctx->sm_count * wide_grid_mult());
            if (many) k_normals_warp<ManyGrids><<<grid, PW_WARPS * 32, 0, ctx->stream>>>(rtr_many(g, c), radius * radius, sums);
            else k_normals_warp<OneGrid><<<grid, PW_WARPS * 32, 0, ctx->stream>>>(rtr_one(g), radius * radius, sums);
            RTR_LAUNCH_CHECK(ctx, "normals");
            k_normals_solve<<<nblk(c->n, 128), 128, 0, ctx->stream>>>(g->sorted, c->n, sums, c->normals);
            RTR_LAUNCH_CHECK(ctx, "normals.solve");
            dev_free(ctx, sums);
        } else {
            k_normals<<<nblk(c->n, 128), 128, 0, ctx->stream>>>(rtr_view(g), radius * radius, c->normals);
            RTR_LAUNCH_CHECK(ctx, "normals");
        }
    }
    c->normals_radius = radius; c->normals_mode = 0;
    c->normals_version++;
    return 0;
}

int rtr_harris_dev(rtr_cloud* c, float radius, float threshold, int nms, int refine, int** d_kp_idx, float4** d_kp_xyz,
                   int** d_count) {
    RtrRange nvtx_range("rtr.harris3d");
    rtr_context* ctx = c->ctx;
    if (!c->normals) return rtr_fail("harris", "rtr_normals must run first", RTR_ERR_NOT_READY);
    DevGrid* g;
    if (int e = rtr_get_grid(c, radius, &g)) return e;
    if (int e = rtr_grid_normals(c, g)) return e;
    int n = c->n;
    float r2 = radius * radius;
    if (!c->response) if (int e = dev_alloc(ctx, &c->response, n, "harris")) return e;
    float* resp_sorted = nullptr; unsigned char* flags = nullptr;
    if (int e = tmp_alloc(ctx, &resp_sorted, n, "harris")) return e;
    if (int e = tmp_alloc(ctx, &flags, n, "harris")) return e;
    if (int e = tmp_alloc(ctx, d_kp_idx, n, "harris")) return e;
    if (int e = tmp_alloc(ctx, d_kp_xyz, n, "harris")) return e;
    if (int e = tmp_alloc(ctx, d_count, 1, "harris")) return e;
    const int nseg = c->nseg();
    if (nseg > 0) {
        // model set: per-cloud corner lists (list k starts at the cloud's first point index), counts[nseg]
        dev_free(ctx, *d_count);
        if (int e = tmp_alloc(ctx, d_count, nseg, "harris")) return e;
        if (n > 0) {
            const ManyGrids mg = rtr_many(g, c);
            int grid = std::min(nblk(n, PW_WARPS), ctx->sm_count * wide_grid_mult());
            k_harris_response_warp<ManyGrids><<<grid, PW_WARPS * 32, 0, ctx->stream>>>(mg, g->sorted_normals, r2, c->response, resp_sorted);
            RTR_LAUNCH_CHECK(ctx, "harris.response");
            k_harris_nms_warp<ManyGrids><<<grid, PW_WARPS * 32, 0, ctx->stream>>>(mg, resp_sorted, r2, threshold, nms, flags);
            RTR_LAUNCH_CHECK(ctx, "harris.nms");
            k_select_segments<<<nseg, 1024, 0, ctx->stream>>>(mg, flags, *d_kp_idx, *d_count);
            RTR_LAUNCH_CHECK(ctx, "harris.select");
            k_harris_refine<ManyGrids><<<dim3(32, nseg), REFINE_THREADS, 0, ctx->stream>>>(mg, g->sorted_normals, c->pts, r2, *d_kp_idx, *d_count, n, *d_kp_xyz, refine);
            RTR_LAUNCH_CHECK(ctx, "harris.refine");
        } else {
            RTR_CHECK(cudaMemsetAsync(*d_count, 0, sizeof(int) * nseg, ctx->stream), "harris");
        }
        dev_free(ctx, resp_sorted); dev_free(ctx, flags);
        return 0;
    }
    RTR_CHECK(cudaMemsetAsync(*d_count, 0, sizeof(int), ctx->stream), "harris");
    if (n > 0) {
        GridView v = rtr_view(g);
        const OneGrid ov = rtr_one(g);
        if (n <= env_threshold("RTR_WARP_PER_POINT_MAX", RTR_WARP_PER_POINT_MAX)) {
            int grid = std::min(nblk(n, PW_WARPS), ctx->sm_count * wide_grid_mult());
            k_harris_response_warp<OneGrid><<<grid, PW_WARPS * 32, 0, ctx->stream>>>(ov, g->sorted_normals, r2, c->response, resp_sorted);
            RTR_LAUNCH_CHECK(ctx, "harris.response");
            k_harris_nms_warp<OneGrid><<<grid, PW_WARPS * 32, 0, ctx->stream>>>(ov, resp_sorted, r2, threshold, nms, flags);
            RTR_LAUNCH_CHECK(ctx, "harris.nms");
        } else {
            k_harris_response<<<nblk(n, 128), 128, 0, ctx->stream>>>(v, g->sorted_normals, r2, c->response, resp_sorted);
            RTR_LAUNCH_CHECK(ctx, "harris.response");
            k_harris_nms<<<nblk(n, 128), 128, 0, ctx->stream>>>(v, resp_sorted, r2, threshold, nms, flags);
            RTR_LAUNCH_CHECK(ctx, "harris.nms");
        }
        size_t tb = 0;
        thrust::counting_iterator<int> iota(0);
        cub::DeviceSelect::Flagged(nullptr, tb, iota, flags, *d_kp_idx, *d_count, n, ctx->stream);
        char* temp = nullptr;
        if (int e = tmp_alloc(ctx, &temp, tb, "harris")) return e;
        RTR_CHECK(cub::DeviceSelect::Flagged(temp, tb, iota, flags, *d_kp_idx, *d_count, n, ctx->stream), "harris.select");
        RTR_MARK(ctx, "harris.cub_select");
        dev_free(ctx, temp);
        // the corner count lives on the device: persistent grid of warps striding over the corner list
        k_harris_refine<OneGrid><<<std::min(n, ctx->sm_count * 4), REFINE_THREADS, 0, ctx->stream>>>(ov, g->sorted_normals, c->pts, r2, *d_kp_idx, *d_count, n, *d_kp_xyz, refine);
        RTR_LAUNCH_CHECK(ctx, "harris.refine");
    }
    dev_free(ctx, resp_sorted); dev_free(ctx, flags);
    return 0;
}

int rtr_fpfh_dev(rtr_cloud* c, float radius) {
    RtrRange nvtx_range("rtr.fpfh");
    rtr_context* ctx = c->ctx;
    if (!c->normals) return rtr_fail("fpfh", "rtr_normals must run first", RTR_ERR_NOT_READY);
    if (c->fpfh && c->fpfh_radius == radius) return 0;
    DevGrid* g;
    if (int e = rtr_get_grid(c, radius, &g)) return e;
    if (int e = rtr_grid_normals(c, g)) return e;
    int n = c->n;
    float r2 = radius * radius;
    if (!c->fpfh) if (int e = dev_alloc(ctx, &c->fpfh, (size_t)n * 33, "fpfh")) return e;
    float* spfh = nullptr;
    if (int e = tmp_alloc(ctx, &spfh, (size_t)n * 33, "fpfh")) return e;
    if (n > 0) {
        GridView v = rtr_view(g);
        // RTR_SPFH_EXACT=1 sends every pair through the fp64 evaluation (the tests use it to show the fp32 screen changes nothing)
        const char* ex = getenv("RTR_SPFH_EXACT");
        int use_screen = (ex && ex[0] == '1') ? 0 : 1;
        if (c->nseg() > 0) {
            const ManyGrids mg = rtr_many(g, c);
            k_spfh<3, ManyGrids><<<std::min(nblk(n, SPFH_WARPS), ctx->sm_count * wide_grid_mult()), SPFH_WARPS * 32, 0, ctx->stream>>>(mg, g->sorted_normals, r2, spfh, use_screen, nullptr, nullptr);
            RTR_LAUNCH_CHECK(ctx, "fpfh.spfh");
            k_fpfh_weight<ManyGrids><<<std::min(nblk(n, FPFH_WARPS), ctx->sm_count * wide_grid_mult()), FPFH_WARPS * 32, 0, ctx->stream>>>(mg, spfh, r2, c->fpfh, nullptr, 0);
            RTR_LAUNCH_CHECK(ctx, "fpfh.weight");
            dev_free(ctx, spfh);
            c->fpfh_radius = radius;
            c->feature_gen = rtr_next_generation();
            dev_free(ctx, c->knn); c->knn = nullptr; dev_free(ctx, c->knn_dist); c->knn_dist = nullptr; c->knn_k = 0;
            return 0;
        }
        const OneGrid ov = rtr_one(g);
        k_spfh<3, OneGrid><<<std::min(nblk(n, SPFH_WARPS), ctx->sm_count * wide_grid_mult()), SPFH_WARPS * 32, 0, ctx->stream>>>(ov, g->sorted_normals, r2, spfh, use_screen, nullptr, nullptr);
        RTR_LAUNCH_CHECK(ctx, "fpfh.spfh");
        if (n >= env_threshold("RTR_FPFH_TILED_MIN", 1 << 20)) {
            // one CTA per occupied cell, candidates staged through shared memory (needs many occupied cells to fill the GPU)
            int *cells = nullptr, *n_cells = nullptr;
            if (int e = tmp_alloc(ctx, &cells, (size_t)std::min(g->ncells, n), "fpfh")) return e;
            if (int e = tmp_alloc(ctx, &n_cells, 1, "fpfh")) return e;
            RTR_CHECK(cudaMemsetAsync(n_cells, 0, sizeof(int), ctx->stream), "fpfh");
            k_occupied_cells<<<nblk(g->ncells, 256), 256, 0, ctx->stream>>>(g->cell_begin, g->ncells, cells, n_cells);
            RTR_LAUNCH_CHECK(ctx, "fpfh.cells");
            int per_sm = 1;               // persistent CTAs: exactly as many as are resident at once (a partial second wave would idle most SMs)
            if (int e = rtr_kernel_smem(k_fpfh_weight_tiled, ctx, sizeof(FwSmem))) return e;
            if (int e = rtr_func_occupancy((const void*)k_fpfh_weight_tiled, ctx->device, FW_WARPS * 32, sizeof(FwSmem), &per_sm)) return e;
            int grid = std::min(std::min(g->ncells, n), ctx->sm_count * per_sm);
            k_fpfh_weight_tiled<<<grid, FW_WARPS * 32, sizeof(FwSmem), ctx->stream>>>(v, spfh, r2, c->fpfh, cells, n_cells);
        } else {
            k_fpfh_weight<OneGrid><<<std::min(nblk(n, FPFH_WARPS), ctx->sm_count * wide_grid_mult()), FPFH_WARPS * 32, 0, ctx->stream>>>(ov, spfh, r2, c->fpfh, nullptr, 0);
        }
        RTR_LAUNCH_CHECK(ctx, "fpfh.weight");
    }
    dev_free(ctx, spfh);
    c->fpfh_radius = radius;
    c->feature_gen = rtr_next_generation();
    // features changed: cached correspondences are stale
    dev_free(ctx, c->knn); c->knn = nullptr; dev_free(ctx, c->knn_dist); c->knn_dist = nullptr; c->knn_k = 0;
    return 0;
}

// FPFH at a subset of the points (PCL: setInputCloud(keypoints) + setSearchSurface(cloud)): SPFH only where a query's
// neighbourhood needs it.  inverse permutation -> queries' sorted positions; one warp per query flags its in-radius
// points; the flagged positions are compacted (ascending) into the SPFH work list; the weighting pass runs over the
// queries.  Every row equals the row rtr_fpfh computes for that point.
__global__ void k_inverse_perm(GridView g, int* __restrict__ inv) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < g.n) inv[__float_as_int(__ldg(g.sorted + s).w)] = s;
}
__global__ void k_query_positions(const int* __restrict__ qidx, int nq, const int* __restrict__ inv, int* __restrict__ qpos) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) qpos[i] = inv[qidx[i]];
}
__global__ void __launch_bounds__(PW_WARPS * 32) k_mark_neighbours(GridView g, const int* __restrict__ qpos, int nq, float r2,
                                                                   unsigned char* __restrict__ flags) {
    int lane = threadIdx.x & 31;
    int nwarps = gridDim.x * PW_WARPS;
    for (int i = blockIdx.x * PW_WARPS + (threadIdx.x >> 5); i < nq; i += nwarps) {
        float4 q = __ldg(g.sorted + qpos[i]);
        for_block27_warp(g, q.x, q.y, q.z, lane, [&](int sp, float4, float d2) { if (d2 < r2) flags[sp] = 1; });
    }
}

int rtr_fpfh_at_dev(rtr_cloud* c, float radius, const int* d_query_index, int nq, float* d_out) {
    rtr_context* ctx = c->ctx;
    if (!c->normals) return rtr_fail("fpfh_at", "rtr_normals must run first", RTR_ERR_NOT_READY);
    if (c->nseg() > 0) return rtr_fail("fpfh_at", "not available on a model set", RTR_ERR_INVALID);
    if (nq <= 0 || c->n <= 0) return 0;
    DevGrid* g;
    if (int e = rtr_get_grid(c, radius, &g)) return e;
    if (int e = rtr_grid_normals(c, g)) return e;
    const int n = c->n;
    const float r2 = radius * radius;
    GridView v = rtr_view(g);
    int *inv = nullptr, *qpos = nullptr, *list = nullptr, *count = nullptr; unsigned char* flags = nullptr; float* spfh = nullptr; char* temp = nullptr;
    if (int e = tmp_alloc(ctx, &inv, n, "fpfh_at")) return e;
    if (int e = tmp_alloc(ctx, &qpos, nq, "fpfh_at")) return e;
    if (int e = tmp_alloc(ctx, &list, n, "fpfh_at")) return e;
    if (int e = tmp_alloc(ctx, &count, 1, "fpfh_at")) return e;
    if (int e = tmp_alloc(ctx, &flags, n, "fpfh_at")) return e;
    if (int e = tmp_alloc(ctx, &spfh, (size_t)n * 33, "fpfh_at")) return e;
    k_inverse_perm<<<nblk(n, 256), 256, 0, ctx->stream>>>(v, inv);
    RTR_LAUNCH_CHECK(ctx, "fpfh_at.inverse");
    k_query_positions<<<nblk(nq, 256), 256, 0, ctx->stream>>>(d_query_index, nq, inv, qpos);
    RTR_LAUNCH_CHECK(ctx, "fpfh_at.positions");
    RTR_CHECK(cudaMemsetAsync(flags, 0, (size_t)n, ctx->stream), "fpfh_at");
    k_mark_neighbours<<<std::min(nblk(nq, PW_WARPS), ctx->sm_count * wide_grid_mult()), PW_WARPS * 32, 0, ctx->stream>>>(v, qpos, nq, r2, flags);
    RTR_LAUNCH_CHECK(ctx, "fpfh_at.mark");
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, thrust::counting_iterator<int>(0), flags, list, count, n, ctx->stream);
    if (int e = tmp_alloc(ctx, &temp, tb, "fpfh_at")) return e;
    RTR_CHECK(cub::DeviceSelect::Flagged(temp, tb, thrust::counting_iterator<int>(0), flags, list, count, n, ctx->stream), "fpfh_at.select");
    RTR_MARK(ctx, "fpfh_at.cub_select");
    const char* ex = getenv("RTR_SPFH_EXACT");
    int use_screen = (ex && ex[0] == '1') ? 0 : 1;
    const OneGrid ov = rtr_one(g);
    k_spfh<3, OneGrid><<<std::min(nblk(n, SPFH_WARPS), ctx->sm_count * wide_grid_mult()), SPFH_WARPS * 32, 0, ctx->stream>>>(ov, g->sorted_normals, r2, spfh, use_screen, list, count);
    RTR_LAUNCH_CHECK(ctx, "fpfh.spfh");
    k_fpfh_weight<OneGrid><<<std::min(nblk(nq, FPFH_WARPS), ctx->sm_count * wide_grid_mult()), FPFH_WARPS * 32, 0, ctx->stream>>>(ov, spfh, r2, d_out, qpos, nq);
    RTR_LAUNCH_CHECK(ctx, "fpfh.weight");
    dev_free(ctx, inv); dev_free(ctx, qpos); dev_free(ctx, list); dev_free(ctx, count); dev_free(ctx, flags); dev_free(ctx, spfh); dev_free(ctx, temp);
    return 0;
}

int rtr_match_dev(rtr_cloud* src, rtr_cloud* tgt, int k) {
    rtr_context* ctx = src->ctx;
    if (!src->fpfh || !tgt->fpfh) return rtr_fail("match", "rtr_fpfh must run on both clouds first", RTR_ERR_NOT_READY);
    if (k < 1 || k > MATCH_KMAX) return rtr_fail("match", "k must be in [1, 8]", RTR_ERR_INVALID);
    dev_free(ctx, src->knn); dev_free(ctx, src->knn_dist);
    src->knn = nullptr; src->knn_dist = nullptr;
    if (int e = dev_alloc(ctx, &src->knn, (size_t)src->n * k, "match")) return e;
    if (int e = dev_alloc(ctx, &src->knn_dist, (size_t)src->n * k, "match")) return e;
    if (src->n > 0) if (int e = match_dispatch(ctx, src->fpfh, src->n, tgt->fpfh, tgt->n, k, src->knn, src->knn_dist, nullptr)) return e;
    src->knn_k = k; src->knn_target = tgt; src->knn_target_gen = tgt->feature_gen;
    return 0;
}

extern "C" {

int rtr_normals(rtr_cloud* c, float radius, float* host_normals4) {
    if (!c || !(radius > 0.f)) return rtr_fail("normals", "bad argument", RTR_ERR_INVALID);
    TmpScope tmp_scope(c->ctx);
    RTR_CHECK(cudaSetDevice(c->ctx->device), "normals");
    if (int e = rtr_normals_dev(c, radius)) return e;
    if (host_normals4 && c->n > 0) RTR_CHECK(cudaMemcpyAsync(host_normals4, c->normals, (size_t)c->n * 16, cudaMemcpyDeviceToHost, c->ctx->stream), "normals");
    RTR_CHECK(cudaStreamSynchronize(c->ctx->stream), "normals");
    return 0;
}

int rtr_normals_mode(rtr_cloud* c, float radius, int mode, float* host_normals4) {
    if (!c || !(radius > 0.f) || (mode != 0 && mode != 1)) return rtr_fail("normals", "bad argument (mode 0: exact, 1: PCL-float)", RTR_ERR_INVALID);
    TmpScope tmp_scope(c->ctx);
    RTR_CHECK(cudaSetDevice(c->ctx->device), "normals");
    if (int e = rtr_normals_mode_dev(c, radius, mode)) return e;
    if (host_normals4 && c->n > 0) RTR_CHECK(cudaMemcpyAsync(host_normals4, c->normals, (size_t)c->n * 16, cudaMemcpyDeviceToHost, c->ctx->stream), "normals");
    RTR_CHECK(cudaStreamSynchronize(c->ctx->stream), "normals");
    return 0;
}

int rtr_harris3d(rtr_cloud* c, float radius, float threshold, int nms, int refine, float* host_response, int* host_kp_index,
                 float* host_kp_xyz1, int capacity, int* n_keypoints) {
    if (!c || !(radius > 0.f) || !n_keypoints || capacity < 0) return rtr_fail("harris", "bad argument", RTR_ERR_INVALID);
    rtr_context* ctx = c->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "harris");
    int* d_idx = nullptr; float4* d_xyz = nullptr; int* d_count = nullptr;
    if (int e = rtr_harris_dev(c, radius, threshold, nms, refine, &d_idx, &d_xyz, &d_count)) return e;
    int m = 0;
    RTR_CHECK(cudaMemcpyAsync(&m, d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "harris");
    if (host_response && c->n > 0) RTR_CHECK(cudaMemcpyAsync(host_response, c->response, (size_t)c->n * 4, cudaMemcpyDeviceToHost, ctx->stream), "harris");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "harris");
    *n_keypoints = m;
    c->n_keypoints = m;
    int rc = 0;
    int take = m < capacity ? m : capacity;
    if (take > 0) {
        if (host_kp_index) RTR_CHECK(cudaMemcpyAsync(host_kp_index, d_idx, (size_t)take * 4, cudaMemcpyDeviceToHost, ctx->stream), "harris");
        if (host_kp_xyz1) RTR_CHECK(cudaMemcpyAsync(host_kp_xyz1, d_xyz, (size_t)take * 16, cudaMemcpyDeviceToHost, ctx->stream), "harris");
        RTR_CHECK(cudaStreamSynchronize(ctx->stream), "harris");
    }
    if (m > capacity && (host_kp_index || host_kp_xyz1)) rc = RTR_ERR_CAPACITY;
    dev_free(ctx, d_idx); dev_free(ctx, d_xyz); dev_free(ctx, d_count);
    return rc;
}

int rtr_fpfh(rtr_cloud* c, float radius, float* host_fpfh) {
    if (!c || !(radius > 0.f)) return rtr_fail("fpfh", "bad argument", RTR_ERR_INVALID);
    TmpScope tmp_scope(c->ctx);
    RTR_CHECK(cudaSetDevice(c->ctx->device), "fpfh");
    if (int e = rtr_fpfh_dev(c, radius)) return e;
    if (host_fpfh && c->n > 0) RTR_CHECK(cudaMemcpyAsync(host_fpfh, c->fpfh, (size_t)c->n * 33 * 4, cudaMemcpyDeviceToHost, c->ctx->stream), "fpfh");
    RTR_CHECK(cudaStreamSynchronize(c->ctx->stream), "fpfh");
    return 0;
}

int rtr_fpfh_at(rtr_cloud* c, float radius, const int* host_query_index, int n_query, float* host_fpfh) {
    if (!c || !(radius > 0.f) || n_query < 0 || (n_query > 0 && (!host_query_index || !host_fpfh))) return rtr_fail("fpfh_at", "bad argument", RTR_ERR_INVALID);
    for (int i = 0; i < n_query; ++i)
        if (host_query_index[i] < 0 || host_query_index[i] >= c->n) return rtr_fail("fpfh_at", "query index out of range", RTR_ERR_INVALID);
    if (n_query == 0) return 0;
    rtr_context* ctx = c->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "fpfh_at");
    int* d_q = nullptr; float* d_out = nullptr;
    if (int e = tmp_alloc(ctx, &d_q, n_query, "fpfh_at")) return e;
    if (int e = tmp_alloc(ctx, &d_out, (size_t)n_query * 33, "fpfh_at")) return e;
    RTR_CHECK(cudaMemcpyAsync(d_q, host_query_index, (size_t)n_query * sizeof(int), cudaMemcpyHostToDevice, ctx->stream), "fpfh_at");
    if (int e = rtr_fpfh_at_dev(c, radius, d_q, n_query, d_out)) return e;
    RTR_CHECK(cudaMemcpyAsync(host_fpfh, d_out, (size_t)n_query * 33 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "fpfh_at");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "fpfh_at");
    return 0;
}

int rtr_match_features_raw(rtr_context* ctx, const float* host_source_feat, int ns, const float* host_target_feat, int nt,
                           int k, int* host_idx, float* host_dist, float* kernel_ms) {
    if (!ctx || ns < 0 || nt < 0 || k < 1 || k > MATCH_KMAX || (ns > 0 && (!host_source_feat || !host_idx)) || (nt > 0 && !host_target_feat))
        return rtr_fail("match_raw", "bad argument", RTR_ERR_INVALID);
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "match_raw");
    float *fa = nullptr, *fb = nullptr, *dd = nullptr; int* di = nullptr;
    if (int e = tmp_alloc(ctx, &fa, (size_t)ns * 33, "match_raw")) return e;
    if (int e = tmp_alloc(ctx, &fb, (size_t)nt * 33, "match_raw")) return e;
    if (int e = tmp_alloc(ctx, &di, (size_t)ns * k, "match_raw")) return e;
    if (int e = tmp_alloc(ctx, &dd, (size_t)ns * k, "match_raw")) return e;
    if (ns > 0) RTR_CHECK(cudaMemcpyAsync(fa, host_source_feat, (size_t)ns * 132, cudaMemcpyHostToDevice, ctx->stream), "match_raw");
    if (nt > 0) RTR_CHECK(cudaMemcpyAsync(fb, host_target_feat, (size_t)nt * 132, cudaMemcpyHostToDevice, ctx->stream), "match_raw");
    RTR_CHECK(cudaEventRecord(ctx->events[RTR_NUM_EVENTS - 2], ctx->stream), "match_raw");
    if (ns > 0) if (int e = match_dispatch(ctx, fa, ns, fb, nt, k, di, dd, ctx->match_stats)) return e;
    RTR_CHECK(cudaEventRecord(ctx->events[RTR_NUM_EVENTS - 1], ctx->stream), "match_raw");
    if (ns > 0) {
        RTR_CHECK(cudaMemcpyAsync(host_idx, di, (size_t)ns * k * 4, cudaMemcpyDeviceToHost, ctx->stream), "match_raw");
        if (host_dist) RTR_CHECK(cudaMemcpyAsync(host_dist, dd, (size_t)ns * k * 4, cudaMemcpyDeviceToHost, ctx->stream), "match_raw");
    }
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "match_raw");
    if (kernel_ms) RTR_CHECK(cudaEventElapsedTime(kernel_ms, ctx->events[RTR_NUM_EVENTS - 2], ctx->events[RTR_NUM_EVENTS - 1]), "match_raw");
    dev_free(ctx, fa); dev_free(ctx, fb); dev_free(ctx, di); dev_free(ctx, dd);
    return 0;
}

// [0] rows the tensor-core prefilter could not certify and the exact kernel redid (-1: exact kernel used for everything),
// [1] target splits, [2] observed prefilter error / (|a||b|) in 1e-9 units — of the last rtr_match_features_raw call on this context.
int rtr_match_last_stats(rtr_context* ctx, int* stats3) {
    if (!ctx || !stats3) return rtr_fail("match_stats", "bad argument", RTR_ERR_INVALID);
    for (int i = 0; i < 3; ++i) stats3[i] = ctx->match_stats[i];
    return 0;
}

int rtr_match_features(rtr_cloud* source, rtr_cloud* target, int k, int* host_idx, float* host_dist) {
    if (!source || !target || source->ctx != target->ctx) return rtr_fail("match", "bad argument", RTR_ERR_INVALID);
    rtr_context* ctx = source->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "match");
    if (int e = rtr_match_dev(source, target, k)) return e;
    size_t cnt = (size_t)source->n * k;
    if (host_idx && cnt) RTR_CHECK(cudaMemcpyAsync(host_idx, source->knn, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream), "match");
    if (host_dist && cnt) RTR_CHECK(cudaMemcpyAsync(host_dist, source->knn_dist, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream), "match");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "match");
    return 0;
}

}  // extern "C"
