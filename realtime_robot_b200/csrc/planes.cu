// planes.cu — ModelPoint::getArea / ScanPoint::get_Area (model_point.h:170-245, scan_point.h:117-188) on the device:
// iterative plane peel (pcl::SACSegmentation PLANE / RANSAC, 150 iterations, 5 mm, optimised coefficients) + convex-hull
// area (pcl::ConvexHull::getTotalArea) of every plane, until <= 15 % of the points remain.
//
// Per plane, one host round trip:
//   host   : the 151 index triples RANSAC would draw — boost::mt19937(12345) >> 1 and the partial Fisher-Yates shuffle of
//            SampleConsensusModel::drawIndexSample depend only on the number of remaining points
//   k_plane_hypotheses : one CTA per hypothesis: plane through the three points (fp32, as PCL), inlier count over all points
//   k_plane_pick       : the sequential adaptive-k loop of RandomSampleConsensus::computeModel over those counts
//   k_plane_flags + CUB select : inliers of the winner -> k_plane_moments / k_plane_refit: least-squares plane (fp64 moments,
//            Jacobi) -> k_plane_flags + select again: refined inliers (the plane's cloud) and the rest (next iteration's cloud)
//   k_hull_frame / k_hull_keys / CUB sort / k_hull_chain : hull dimension (fp64 eigenvalue ratio), PCL's choice of projection
//            plane, lexicographic sort of the 2-D points, monotone chain + shoelace in fp64
#include "common.cuh"
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>

#define PL_HYP 160      // 151 iterations at most, plus head-room for degenerate (collinear) samples, which are skipped

__device__ __forceinline__ float plane_dist(const float* c, float4 p) {
    return fabsf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c[0], p.x), __fmul_rn(c[1], p.y)), __fmul_rn(c[2], p.z)), c[3]));
}

// SampleConsensusModelPlane::computeModelCoefficients + countWithinDistance
__global__ void __launch_bounds__(256) k_plane_hypotheses(const float4* __restrict__ pts, int n, const int* __restrict__ samples, double threshold,
                                                          float* __restrict__ coeffs, int* __restrict__ counts, int* __restrict__ bad) {
    __shared__ float c[4];
    __shared__ int ok;
    __shared__ int wcnt[8];
    int h = blockIdx.x;
    if (threadIdx.x == 0) {
        float4 p0 = pts[samples[h * 3 + 0]], p1 = pts[samples[h * 3 + 1]], p2 = pts[samples[h * 3 + 2]];
        float a0 = __fsub_rn(p1.x, p0.x), a1 = __fsub_rn(p1.y, p0.y), a2 = __fsub_rn(p1.z, p0.z);
        float b0 = __fsub_rn(p2.x, p0.x), b1 = __fsub_rn(p2.y, p0.y), b2 = __fsub_rn(p2.z, p0.z);
        float r0 = __fdiv_rn(a0, b0), r1 = __fdiv_rn(a1, b1), r2 = __fdiv_rn(a2, b2);
        ok = !(r0 == r1 && r2 == r1);
        float c0 = __fsub_rn(__fmul_rn(a1, b2), __fmul_rn(a2, b1));
        float c1 = __fsub_rn(__fmul_rn(a2, b0), __fmul_rn(a0, b2));
        float c2 = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
        float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fmul_rn(c2, c2)));
        c[0] = __fdiv_rn(c0, nrm); c[1] = __fdiv_rn(c1, nrm); c[2] = __fdiv_rn(c2, nrm);
        c[3] = __fmul_rn(-1.0f, __fadd_rn(__fadd_rn(__fmul_rn(c[0], p0.x), __fmul_rn(c[1], p0.y)), __fmul_rn(c[2], p0.z)));
        for (int i = 0; i < 4; ++i) coeffs[h * 4 + i] = c[i];
        bad[h] = ok ? 0 : 1;
    }
    __syncthreads();
    int cnt = 0;
    if (ok) for (int i = threadIdx.x; i < n; i += blockDim.x) cnt += ((double)plane_dist(c, __ldg(pts + i)) < threshold) ? 1 : 0;
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += wcnt[w]; counts[h] = t; }
}

struct PlaneState {
    float coeff[4];        // current model (best RANSAC model, then the refined one)
    int best, iterations, error;
    int n_inliers, n_rest;
    double area; int dimension;
    int axis_u, axis_v;    // 2-D hull: coordinate axes of the projection; 3-D: -1
    double mean[3], eu[3], ev[3];
};

// RandomSampleConsensus::computeModel, replayed over the precomputed counts
__device__ void plane_pick(const float* coeffs, const int* counts, const int* bad, int n, int max_iterations, PlaneState* st) {
    int best = -1, best_h = -1, iterations = 0;
    unsigned skipped = 0, max_skip = (unsigned)max_iterations * 10u;
    double k = 1.0;
    const double log_probability = log(1.0 - 0.99), one_over = 1.0 / (double)n;
    int error = 0, h = 0;
    for (; h < PL_HYP && (double)iterations < k && skipped < max_skip; ++h) {
        if (bad[h]) { ++skipped; continue; }                  // collinear sample: drawn again (the draw sequence simply continues)
        int cnt = counts[h];
        if (cnt > best) {
            best = cnt; best_h = h;
            double w = (double)best * one_over;
            double p_no = 1.0 - pow(w, 3.0);
            p_no = fmax(2.220446049250313e-16, p_no);
            p_no = fmin(1.0 - 2.220446049250313e-16, p_no);
            k = log_probability / log(p_no);
        }
        ++iterations;
        if (iterations > max_iterations) break;
    }
    if (h == PL_HYP && (double)iterations < k && iterations <= max_iterations) error = 1;     // ran out of precomputed draws
    st->best = best_h; st->iterations = iterations; st->error = error;
    if (best_h >= 0) for (int i = 0; i < 4; ++i) st->coeff[i] = coeffs[best_h * 4 + i];
}
__global__ void k_plane_pick(const float* __restrict__ coeffs, const int* __restrict__ counts, const int* __restrict__ bad, int n,
                             int max_iterations, PlaneState* st) {
    if (threadIdx.x == 0) plane_pick(coeffs, counts, bad, n, max_iterations, st);
}

__global__ void k_plane_flags(const float4* __restrict__ pts, int n, const PlaneState* __restrict__ st, double threshold,
                              unsigned char* __restrict__ in_flag, unsigned char* __restrict__ out_flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool in = st->best >= 0 && ((double)plane_dist(st->coeff, __ldg(pts + i)) < threshold);
    in_flag[i] = in ? 1 : 0;
    if (out_flag) out_flag[i] = in ? 0 : 1;
}

// fp64 moments (sum x, y, z, xx, xy, xz, yy, yz, zz) of pts[idx[0..m)): per-CTA partials, folded in fixed order by the consumer
__global__ void __launch_bounds__(256) k_plane_moments(const float4* __restrict__ pts, const int* __restrict__ idx, const int* __restrict__ m_ptr,
                                                       double* __restrict__ partials) {
    __shared__ double red[8][9];
    int m = *m_ptr;
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        float4 p = __ldg(pts + (idx ? idx[i] : i));
        double x = p.x, y = p.y, z = p.z;
        a[0] += x; a[1] += y; a[2] += z; a[3] += x * x; a[4] += x * y; a[5] += x * z; a[6] += y * y; a[7] += y * z; a[8] += z * z;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) { double v = warp_sum(a[k]); if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v; }
    __syncthreads();
    if (threadIdx.x < 9) { double v = 0; for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x]; partials[blockIdx.x * 9 + threadIdx.x] = v; }
}

__device__ void sums_eig(const double* s, double m, double mean[3], double evals[3], double evecs[3][3]);
__device__ void moments_eig(const double* partials, int nparts, double m, double mean[3], double evals[3], double evecs[3][3]) {
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nparts; ++b) for (int k = 0; k < 9; ++k) s[k] += partials[b * 9 + k];
    sums_eig(s, m, mean, evals, evecs);
}
__device__ void sums_eig(const double* s, double m, double mean[3], double evals[3], double evecs[3][3]) {
    for (int a = 0; a < 3; ++a) mean[a] = s[a] / m;
    double A[3][3], V[3][3];
    A[0][0] = s[3] / m - mean[0] * mean[0]; A[0][1] = s[4] / m - mean[0] * mean[1]; A[0][2] = s[5] / m - mean[0] * mean[2];
    A[1][1] = s[6] / m - mean[1] * mean[1]; A[1][2] = s[7] / m - mean[1] * mean[2]; A[2][2] = s[8] / m - mean[2] * mean[2];
    A[1][0] = A[0][1]; A[2][0] = A[0][2]; A[2][1] = A[1][2];
    jacobi_eig<3>(A, V);
    int o[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j)
        if (A[o[j + 1]][o[j + 1]] < A[o[j]][o[j]]) { int t = o[j]; o[j] = o[j + 1]; o[j + 1] = t; }     // stable ascending
    for (int e = 0; e < 3; ++e) { evals[e] = A[o[e]][o[e]]; for (int r = 0; r < 3; ++r) evecs[e][r] = V[r][o[e]]; }
}

// SampleConsensusModelPlane::optimizeModelCoefficients
__device__ void plane_refit_sums(const double* sums, int m, PlaneState* st) {
    if (st->best < 0 || m <= 3) return;
    double mean[3], ev[3], evec[3][3];
    sums_eig(sums, (double)m, mean, ev, evec);
    st->coeff[0] = (float)evec[0][0]; st->coeff[1] = (float)evec[0][1]; st->coeff[2] = (float)evec[0][2];
    st->coeff[3] = (float)(-1.0 * ((evec[0][0] * mean[0] + evec[0][1] * mean[1]) + evec[0][2] * mean[2]));
}
__global__ void k_plane_refit(const double* __restrict__ partials, int nparts, const int* __restrict__ m_ptr, PlaneState* st) {
    if (threadIdx.x != 0) return;
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nparts; ++b) for (int k = 0; k < 9; ++k) s[k] += partials[b * 9 + k];
    plane_refit_sums(s, *m_ptr, st);
}

// ConvexHull::calculateInputDimension + the projection choice of performReconstruction2D
__device__ void hull_frame_sums(const float4* plane_pts, const double* sums, int m, PlaneState* st);
__global__ void k_hull_frame(const float4* __restrict__ plane_pts, const double* __restrict__ partials, int nparts, const int* __restrict__ m_ptr,
                             PlaneState* st) {
    if (threadIdx.x != 0) return;
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nparts; ++b) for (int k = 0; k < 9; ++k) s[k] += partials[b * 9 + k];
    hull_frame_sums(plane_pts, s, *m_ptr, st);
}
__device__ void hull_frame_sums(const float4* plane_pts, const double* sums, int m, PlaneState* st) {
    st->n_inliers = m;
    st->area = 0.0; st->dimension = 2; st->axis_u = 0; st->axis_v = 1;
    if (m < 3) return;
    double ev[3], evec[3][3];
    sums_eig(sums, (double)m, st->mean, ev, evec);
    int dim = (fabs(ev[0]) < 2.220446049250313e-16 || fabs(ev[0] / ev[2]) < 1.0e-3) ? 2 : 3;
    st->dimension = dim;
    if (dim == 2) {
        float4 p0 = plane_pts[0], p1 = plane_pts[m - 1], pm = plane_pts[m / 2];
        double a[3] = {(double)p1.x - p0.x, (double)p1.y - p0.y, (double)p1.z - p0.z}, b[3] = {(double)pm.x - p0.x, (double)pm.y - p0.y, (double)pm.z - p0.z};
        double nx = a[1] * b[2] - a[2] * b[1], ny = a[2] * b[0] - a[0] * b[2], nz = a[0] * b[1] - a[1] * b[0];
        double nn = sqrt((nx * nx + ny * ny) + nz * nz);
        if (nn == 0) { nx = evec[0][0]; ny = evec[0][1]; nz = evec[0][2]; nn = 1; }
        double tx = fabs(nx / nn), ty = fabs(ny / nn), tz = fabs(nz / nn);
        const double thresh = 0.984807753012208;          // cos(0.174532925)
        bool xy = true, yz = true, xz = true;
        if (tz > thresh) { xz = false; yz = false; }
        if (tx > thresh) { xz = false; xy = false; }
        if (ty > thresh) { xy = false; yz = false; }
        if (xy) { st->axis_u = 0; st->axis_v = 1; } else if (yz) { st->axis_u = 1; st->axis_v = 2; } else if (xz) { st->axis_u = 0; st->axis_v = 2; }
    } else {
        st->axis_u = -1; st->axis_v = -1;
        for (int r = 0; r < 3; ++r) { st->eu[r] = evec[2][r]; st->ev[r] = evec[1][r]; }
    }
}

__device__ __forceinline__ unsigned f2key(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(unsigned k) { unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; return __uint_as_float(u); }

// 2-D points of the hull problem as sortable 64-bit keys (u major, v minor)
__device__ __forceinline__ void hull_uv(float4 p, const PlaneState* st, float& u, float& v) {
    const float c[3] = {p.x, p.y, p.z};
    if (st->axis_u >= 0) { u = c[st->axis_u]; v = c[st->axis_v]; }
    else {
        double d[3] = {(double)p.x - st->mean[0], (double)p.y - st->mean[1], (double)p.z - st->mean[2]};
        u = (float)((d[0] * st->eu[0] + d[1] * st->eu[1]) + d[2] * st->eu[2]);
        v = (float)((d[0] * st->ev[0] + d[1] * st->ev[1]) + d[2] * st->ev[2]);
    }
    if (u == 0.f) u = 0.f;      // -0 -> +0 so that equal values get equal keys
    if (v == 0.f) v = 0.f;
}
__global__ void k_hull_keys(const float4* __restrict__ plane_pts, const int* __restrict__ m_ptr, const PlaneState* __restrict__ st,
                            unsigned long long* __restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int m = *m_ptr;
    if (i >= m) return;
    float u, v;
    hull_uv(plane_pts[i], st, u, v);
    keys[i] = ((unsigned long long)f2key(u) << 32) | (unsigned long long)f2key(v);
}

// monotone chain over the sorted keys + shoelace (fp64); one thread: the chain is inherently sequential
__device__ void hull_chain(const unsigned long long* keys, int m, double2* stack, PlaneState* st);
__global__ void k_hull_chain(const unsigned long long* __restrict__ keys, const int* __restrict__ m_ptr, double2* __restrict__ stack, PlaneState* st) {
    if (threadIdx.x == 0) hull_chain(keys, *m_ptr, stack, st);
}
__device__ void hull_chain(const unsigned long long* keys, int m, double2* stack, PlaneState* st) {
    if (m < 3) { st->area = 0.0; return; }
    auto pt = [&](int i) { unsigned long long k = keys[i]; return make_double2((double)key2f((unsigned)(k >> 32)), (double)key2f((unsigned)(k & 0xffffffffu))); };
    auto cross = [](double2 o, double2 a, double2 b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); };
    int k = 0;
    unsigned long long prev = 0; bool have = false;
    int uniq = 0;
    for (int i = 0; i < m; ++i) {                                   // lower hull (duplicates skipped)
        if (have && keys[i] == prev) continue;
        prev = keys[i]; have = true; ++uniq;
        double2 p = pt(i);
        while (k >= 2 && cross(stack[k - 2], stack[k - 1], p) <= 0) --k;
        stack[k++] = p;
    }
    if (uniq < 3) { st->area = 0.0; return; }
    int t = k + 1;
    have = false;
    bool first = true;
    for (int i = m - 1; i >= 0; --i) {                              // upper hull
        if (have && keys[i] == prev) continue;
        prev = keys[i]; have = true;
        if (first) { first = false; continue; }                     // the last unique point is already on the stack
        double2 p = pt(i);
        while (k >= t && cross(stack[k - 2], stack[k - 1], p) <= 0) --k;
        stack[k++] = p;
    }
    double a = 0;
    for (int i = 0; i + 1 < k; ++i) a += stack[i].x * stack[i + 1].y - stack[i + 1].x * stack[i].y;
    a = 0.5 * fabs(a);
    st->area = (st->dimension == 2) ? a : 2.0 * a;
}

__global__ void k_gather4(const float4* __restrict__ src, const int* __restrict__ idx, const int* __restrict__ m_ptr, float4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *m_ptr) dst[i] = __ldg(src + idx[i]);
}
__global__ void k_plane_counts(const int* n_in, const int* n_rest, PlaneState* st) { if (threadIdx.x == 0) { st->n_inliers = *n_in; st->n_rest = *n_rest; } }


// ---------------------------------------------------------------------------------------------------------------------
// One plane of the peel in ONE single-CTA launch, for clouds of at most PL_SMALL_MAX points (every cloud the reference
// ships): pick -> inliers of the RANSAC model -> least-squares refit -> refined inliers / rest (stable partition) -> hull
// frame -> hull area.  Replaces ~20 launches (three cub::DeviceSelect, a 64-bit radix sort of every inlier, a one-thread
// chain over global memory) and one of the two host round trips.
//  * the fp64 moment sums reproduce the shape of k_plane_moments (64 virtual CTAs of 256 threads, shuffle tree, warps
//    then CTAs folded in order) so both paths give bit-identical planes;
//  * hull: points strictly inside the octagon of the 8 extreme points (min / max of u, v, u + v, u - v) can never be hull
//    vertices and are dropped (Akl-Toussaint); the survivors (~sqrt(m)) are bitonic-sorted and chained in shared memory.
#define PL_SMALL_MAX 32768
#define PL_HULL_CAP 4096
#define PL_FT 1024
struct PlaneSmall {
    unsigned long long keys[PL_HULL_CAP];
    double2 stack[PL_HULL_CAP + 2];
    double red[32][9];
    double part[64][9];
    double sums[9];
    unsigned long long ext[8];
    double2 oct[8];
    int scan[PL_FT];
    int n_keys, overflow, m1, m2;
    float hyp_coeffs[PL_HYP * 4];
    int hyp_counts[PL_HYP], hyp_bad[PL_HYP];
    PlaneState st;
};

// k_plane_moments' summation shape on one CTA of 1024 threads: 4 virtual CTAs per pass
__device__ void moments_small(const float4* pts, const int* idx, int m, PlaneSmall& S) {
    const int t = threadIdx.x;
    for (int pass = 0; pass < 16; ++pass) {
        if (pass * 1024 >= m) { if (t < 36) S.part[pass * 4 + t / 9][t % 9] = 0.0; continue; }      // no element: the partial is +0
        int g = pass * 1024 + t;
        double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = g; i < m; i += 64 * 256) {
            float4 p = pts[idx ? idx[i] : i];
            double x = p.x, y = p.y, z = p.z;
            a[0] += x; a[1] += y; a[2] += z; a[3] += x * x; a[4] += x * y; a[5] += x * z; a[6] += y * y; a[7] += y * z; a[8] += z * z;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) { double v = warp_sum(a[k]); if ((t & 31) == 0) S.red[t >> 5][k] = v; }
        __syncthreads();
        if (t < 36) { int vc = t / 9, k = t % 9; double v = 0; for (int w = 0; w < 8; ++w) v += S.red[vc * 8 + w][k]; S.part[pass * 4 + vc][k] = v; }
        __syncthreads();
    }
    __syncthreads();
    if (t < 9) { double v = 0; for (int b = 0; b < 64; ++b) v += S.part[b][t]; S.sums[t] = v; }
    __syncthreads();
}

// stable partition of cur[0..n) by |plane . p| < threshold: inlier indices -> idx_in (ascending), optionally the points
// themselves -> in_pts / rest_pts.  Returns the inlier count (uniform).
__device__ int partition_small(const float4* cur, int n, const float* coeff, double threshold, int* idx_in, float4* in_pts, float4* rest_pts,
                               PlaneSmall& S) {
    const int t = threadIdx.x;
    const int per = (n + PL_FT - 1) / PL_FT, b = min(t * per, n), e = min(b + per, n);
    int cnt = 0;
    for (int i = b; i < e; ++i) cnt += ((double)plane_dist(coeff, cur[i]) < threshold) ? 1 : 0;
    S.scan[t] = cnt;
    __syncthreads();
    for (int off = 1; off < PL_FT; off <<= 1) {
        int v = (t >= off) ? S.scan[t - off] : 0;
        __syncthreads();
        S.scan[t] += v;
        __syncthreads();
    }
    int total = S.scan[PL_FT - 1];
    int pos_in = S.scan[t] - cnt, pos_rest = b - pos_in;
    for (int i = b; i < e; ++i) {
        float4 p = cur[i];
        if ((double)plane_dist(coeff, p) < threshold) { if (idx_in) idx_in[pos_in] = i; if (in_pts) in_pts[pos_in] = p; ++pos_in; }
        else { if (rest_pts) rest_pts[pos_rest] = p; ++pos_rest; }
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(PL_FT) k_plane_step_small(const float4* __restrict__ cur, int n, const float* __restrict__ coeffs,
                                                            const int* __restrict__ counts, const int* __restrict__ bad, double threshold,
                                                            int max_iterations, int* idx_in, float4* plane_pts, float4* next,
                                                            int* n_in, int* n_rest, PlaneState* st_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PlaneSmall& S = *reinterpret_cast<PlaneSmall*>(smem_raw);
    const int t = threadIdx.x;
    for (int i = t; i < PL_HYP * 4; i += PL_FT) S.hyp_coeffs[i] = coeffs[i];
    for (int i = t; i < PL_HYP; i += PL_FT) { S.hyp_counts[i] = counts[i]; S.hyp_bad[i] = bad[i]; }
    if (t == 0) { S.n_keys = 0; S.overflow = 0; S.st.n_inliers = 0; S.st.n_rest = n; S.st.area = 0.0; S.st.dimension = 2; S.st.axis_u = 0; S.st.axis_v = 1; }
    __syncthreads();
    if (t == 0) plane_pick(S.hyp_coeffs, S.hyp_counts, S.hyp_bad, n, max_iterations, &S.st);
    __syncthreads();
    if (S.st.best < 0 || S.st.error) {                        // "Could not estimate a planar model" / out of draws: the host stops
        if (t == 0) { *st_out = S.st; *n_in = 0; *n_rest = n; }
        return;
    }
    // inliers of the RANSAC model -> least-squares refit
    int m1 = partition_small(cur, n, S.st.coeff, threshold, idx_in, nullptr, nullptr, S);
    moments_small(cur, idx_in, m1, S);
    if (t == 0) plane_refit_sums(S.sums, m1, &S.st);
    __syncthreads();
    // refined inliers (the plane's cloud) and the rest (the next iteration's cloud)
    int m = partition_small(cur, n, S.st.coeff, threshold, nullptr, plane_pts, next, S);
    if (t == 0) { *n_in = m; *n_rest = n - m; S.st.n_inliers = m; S.st.n_rest = n - m; }
    __syncthreads();
    if (m == 0) { if (t == 0) *st_out = S.st; return; }
    // hull frame
    moments_small(plane_pts, nullptr, m, S);
    if (t == 0) { hull_frame_sums(plane_pts, S.sums, m, &S.st); S.st.n_rest = n - m; }
    __syncthreads();
    if (m >= 3) {
        // 8 extreme points: (ordered value bits << 32 | index), min for W SW S NW-ish directions as laid out below
        unsigned long long e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = (k & 1) ? 0ull : ~0ull;       // even slots: minima, odd slots: maxima
        for (int i = t; i < m; i += PL_FT) {
            float u, v;
            hull_uv(plane_pts[i], &S.st, u, v);
            float d[4] = {u, v, u + v, u - v};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned long long key = ((unsigned long long)f2key(d[k]) << 32) | (unsigned)i;
                e[2 * k] = key < e[2 * k] ? key : e[2 * k];
                e[2 * k + 1] = key > e[2 * k + 1] ? key : e[2 * k + 1];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, e[k], o);
                e[k] = (k & 1) ? (other > e[k] ? other : e[k]) : (other < e[k] ? other : e[k]);
            }
        }
        unsigned long long* wred = reinterpret_cast<unsigned long long*>(&S.red[0][0]);     // 32 warps x 8 keys (2 KB of the 2.3 KB)
        if ((t & 31) == 0) for (int k = 0; k < 8; ++k) wred[(t >> 5) * 8 + k] = e[k];
        __syncthreads();
        if (t < 8) {
            unsigned long long r = wred[t];
            for (int w = 1; w < 32; ++w) { unsigned long long o = wred[w * 8 + t]; r = (t & 1) ? (o > r ? o : r) : (o < r ? o : r); }
            S.ext[t] = r;
        }
        __syncthreads();
        if (t == 0) {
            // counter-clockwise octagon: min u (W), min u+v (SW), min v (S), max u-v (SE), max u (E), max u+v (NE), max v (N), min u-v (NW)
            const int order[8] = {0, 4, 2, 7, 1, 5, 3, 6};
            for (int k = 0; k < 8; ++k) {
                float u, v;
                hull_uv(plane_pts[(unsigned)(S.ext[order[k]] & 0xffffffffull)], &S.st, u, v);
                S.oct[k] = make_double2((double)u, (double)v);
            }
        }
        __syncthreads();
        double ext2 = 0;
        {
            double du = S.oct[4].x - S.oct[0].x, dv = S.oct[6].y - S.oct[2].y;
            ext2 = du * du + dv * dv;
        }
        const double margin = 1e-9 * ext2;           // >> the rounding of a fp64 cross product of float coordinates
        for (int i = t; i < m; i += PL_FT) {
            float u, v;
            hull_uv(plane_pts[i], &S.st, u, v);
            double x = u, y = v;
            bool inside = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                double2 a = S.oct[k], b2 = S.oct[(k + 1) & 7];
                double ex = b2.x - a.x, ey = b2.y - a.y;
                if (ex == 0.0 && ey == 0.0) continue;                                     // coincident extreme points: no edge
                double cr = ex * (y - a.y) - ey * (x - a.x);
                inside = inside && (cr > margin);
            }
            if (!inside) {
                int slot = atomicAdd(&S.n_keys, 1);
                if (slot < PL_HULL_CAP) S.keys[slot] = ((unsigned long long)f2key(u) << 32) | (unsigned long long)f2key(v);
                else S.overflow = 1;
            }
        }
        __syncthreads();
        if (!S.overflow) {
            int nk = S.n_keys, P = 1;
            while (P < nk) P <<= 1;
            for (int i = nk + t; i < P; i += PL_FT) S.keys[i] = ~0ull;
            __syncthreads();
            for (int k = 2; k <= P; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = t; i < P; i += PL_FT) {
                        int l = i ^ j;
                        if (l > i) {
                            unsigned long long a = S.keys[i], b2 = S.keys[l];
                            bool up = (i & k) == 0;
                            if ((a > b2) == up) { S.keys[i] = b2; S.keys[l] = a; }
                        }
                    }
                    __syncthreads();
                }
            if (t == 0) hull_chain(S.keys, nk, S.stack, &S.st);
        } else if (t == 0) S.st.area = -1.0;         // the host finishes this hull on the general path
    }
    __syncthreads();
    if (t == 0) *st_out = S.st;
}

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

// is_h_plane / is_v_plane (model_point.h:48-79; PI is the macro 3.1415926)
static int plane_class(const float* c) {
    double b = ((double)c[0] * c[0]) + ((double)c[1] * c[1]) + ((double)c[2] * c[2]);
    double angle = std::acos((double)c[2] / std::sqrt(b));
    if ((angle > 2.9670597 && angle < 3.1415926) || (angle > 0 && angle < 0.1745329)) return 1;
    if (angle > 1.3962634 && angle < 1.7453292) return 2;
    return 0;
}

extern "C" int rtr_plane_areas(rtr_cloud* c, rtr_surface* host_surfaces, int capacity, int* n_planes) {
    if (!c || !n_planes || capacity < 0 || (capacity > 0 && !host_surfaces)) return rtr_fail("planes", "bad argument", RTR_ERR_INVALID);
    rtr_context* ctx = c->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "planes");
    *n_planes = 0;
    const int n0 = c->n;
    if (n0 < 3) return 0;
    const double threshold = 0.005;        // seg.setDistanceThreshold(0.005), model_point.h:187
    const int max_iterations = 150;        // seg.setMaxIterations(150), model_point.h:186
    float4 *cur = nullptr, *next = nullptr, *plane_pts = nullptr;
    int *d_samples = nullptr, *counts = nullptr, *bad = nullptr, *idx_in = nullptr, *idx_rest = nullptr, *n_in = nullptr, *n_rest = nullptr;
    float* coeffs = nullptr; unsigned char *f_in = nullptr, *f_out = nullptr; double* partials = nullptr; PlaneState* st = nullptr;
    unsigned long long *keys = nullptr, *keys2 = nullptr; double2* stack = nullptr; char* cub_tmp = nullptr;
    const int NP = 64;
    if (int e = tmp_alloc(ctx, &cur, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &next, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &plane_pts, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &d_samples, PL_HYP * 3, "planes")) return e;
    if (int e = tmp_alloc(ctx, &counts, PL_HYP, "planes")) return e;
    if (int e = tmp_alloc(ctx, &bad, PL_HYP, "planes")) return e;
    if (int e = tmp_alloc(ctx, &coeffs, PL_HYP * 4, "planes")) return e;
    if (int e = tmp_alloc(ctx, &idx_in, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &idx_rest, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &n_in, 1, "planes")) return e;
    if (int e = tmp_alloc(ctx, &n_rest, 1, "planes")) return e;
    if (int e = tmp_alloc(ctx, &f_in, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &f_out, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &partials, NP * 9, "planes")) return e;
    if (int e = tmp_alloc(ctx, &st, 1, "planes")) return e;
    if (int e = tmp_alloc(ctx, &keys, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &keys2, n0, "planes")) return e;
    if (int e = tmp_alloc(ctx, &stack, 2 * (size_t)n0 + 2, "planes")) return e;
    size_t tb_sel = 0, tb_sort = 0;
    thrust::counting_iterator<int> iota(0);
    cub::DeviceSelect::Flagged(nullptr, tb_sel, iota, f_in, idx_in, n_in, n0, ctx->stream);
    cub::DeviceRadixSort::SortKeys(nullptr, tb_sort, keys, keys2, n0, 0, 64, ctx->stream);
    size_t tb = std::max(tb_sel, tb_sort);
    if (int e = tmp_alloc(ctx, &cub_tmp, tb, "planes")) return e;
    RTR_CHECK(cudaMemcpyAsync(cur, c->pts, (size_t)n0 * 16, cudaMemcpyDeviceToDevice, ctx->stream), "planes");   // copyPointCloud, model_point.h:173
    int n_cur = n0, planes = 0, rc = 0;
    std::vector<int> shuffled, samples(PL_HYP * 3);
    PlaneState h_st;
    while ((double)n_cur > 0.15 * (double)n0) {                                  // model_point.h:193
        // the draws of one segment() call: new model -> mt19937 reseeded with 12345, identity index array
        std::mt19937 gen(12345u);
        shuffled.resize(n_cur);
        for (int i = 0; i < n_cur; ++i) shuffled[i] = i;
        if (n_cur < 3) break;
        for (int h = 0; h < PL_HYP; ++h) {
            for (int i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + (int)((unsigned)(gen() >> 1) % (unsigned)(n_cur - i))]);
            for (int i = 0; i < 3; ++i) samples[h * 3 + i] = shuffled[i];
        }
        memcpy(ctx->pinned, samples.data(), sizeof(int) * PL_HYP * 3);
        RTR_CHECK(cudaMemcpyAsync(d_samples, ctx->pinned, sizeof(int) * PL_HYP * 3, cudaMemcpyHostToDevice, ctx->stream), "planes");
        k_plane_hypotheses<<<PL_HYP, 256, 0, ctx->stream>>>(cur, n_cur, d_samples, threshold, coeffs, counts, bad);
        RTR_LAUNCH_CHECK(ctx, "planes.hypotheses");
        int m = 0;
        if (n_cur <= PL_SMALL_MAX) {
            // one single-CTA launch for the rest of this plane, one host round trip
            if (int e = rtr_kernel_smem(k_plane_step_small, ctx, sizeof(PlaneSmall))) return e;
            k_plane_step_small<<<1, PL_FT, sizeof(PlaneSmall), ctx->stream>>>(cur, n_cur, coeffs, counts, bad, threshold, max_iterations, idx_in,
                                                                              plane_pts, next, n_in, n_rest, st);
            RTR_LAUNCH_CHECK(ctx, "planes.step_small");
            RTR_CHECK(cudaMemcpyAsync(ctx->pinned, st, sizeof(PlaneState), cudaMemcpyDeviceToHost, ctx->stream), "planes");
            RTR_CHECK(cudaStreamSynchronize(ctx->stream), "planes");
            memcpy(&h_st, ctx->pinned, sizeof(PlaneState));
            if (h_st.error) { rc = rtr_fail("planes", "too many degenerate (collinear) RANSAC samples", RTR_ERR_INVALID); break; }
            if (h_st.best < 0 || h_st.n_inliers == 0) break;                      // "Could not estimate a planar model" (model_point.h:198-202)
            m = h_st.n_inliers;
            if (h_st.area < 0.0) {                                                // more hull candidates than the shared-memory sort holds
                k_hull_keys<<<nblk(m, 256), 256, 0, ctx->stream>>>(plane_pts, n_in, st, keys);
                RTR_LAUNCH_CHECK(ctx, "planes.hull_keys");
                RTR_CHECK(cub::DeviceRadixSort::SortKeys(cub_tmp, tb_sort, keys, keys2, m, 0, 64, ctx->stream), "planes.sort");
                k_hull_chain<<<1, 32, 0, ctx->stream>>>(keys2, n_in, stack, st);
                RTR_LAUNCH_CHECK(ctx, "planes.hull_chain");
                RTR_CHECK(cudaMemcpyAsync(ctx->pinned, st, sizeof(PlaneState), cudaMemcpyDeviceToHost, ctx->stream), "planes");
                RTR_CHECK(cudaStreamSynchronize(ctx->stream), "planes");
                memcpy(&h_st, ctx->pinned, sizeof(PlaneState));
            }
        } else {
        k_plane_pick<<<1, 32, 0, ctx->stream>>>(coeffs, counts, bad, n_cur, max_iterations, st);
        RTR_LAUNCH_CHECK(ctx, "planes.pick");
        // inliers of the RANSAC model -> least-squares refit -> refined inliers + rest
        k_plane_flags<<<nblk(n_cur, 256), 256, 0, ctx->stream>>>(cur, n_cur, st, threshold, f_in, nullptr);
        RTR_LAUNCH_CHECK(ctx, "planes.flags");
        RTR_CHECK(cub::DeviceSelect::Flagged(cub_tmp, tb_sel, iota, f_in, idx_in, n_in, n_cur, ctx->stream), "planes.select");
        k_plane_moments<<<NP, 256, 0, ctx->stream>>>(cur, idx_in, n_in, partials);
        RTR_LAUNCH_CHECK(ctx, "planes.moments");
        k_plane_refit<<<1, 32, 0, ctx->stream>>>(partials, NP, n_in, st);
        RTR_LAUNCH_CHECK(ctx, "planes.refit");
        k_plane_flags<<<nblk(n_cur, 256), 256, 0, ctx->stream>>>(cur, n_cur, st, threshold, f_in, f_out);
        RTR_LAUNCH_CHECK(ctx, "planes.flags");
        RTR_CHECK(cub::DeviceSelect::Flagged(cub_tmp, tb_sel, iota, f_in, idx_in, n_in, n_cur, ctx->stream), "planes.select");
        RTR_CHECK(cub::DeviceSelect::Flagged(cub_tmp, tb_sel, iota, f_out, idx_rest, n_rest, n_cur, ctx->stream), "planes.select");
        k_gather4<<<nblk(n_cur, 256), 256, 0, ctx->stream>>>(cur, idx_in, n_in, plane_pts);      // extract.filter(*cloud_p)
        RTR_LAUNCH_CHECK(ctx, "planes.gather");
        k_gather4<<<nblk(n_cur, 256), 256, 0, ctx->stream>>>(cur, idx_rest, n_rest, next);       // extract.setNegative(true)
        RTR_LAUNCH_CHECK(ctx, "planes.gather");
        // convex hull area of the plane's cloud
        k_plane_moments<<<NP, 256, 0, ctx->stream>>>(plane_pts, nullptr, n_in, partials);
        RTR_LAUNCH_CHECK(ctx, "planes.moments");
        k_hull_frame<<<1, 32, 0, ctx->stream>>>(plane_pts, partials, NP, n_in, st);
        RTR_LAUNCH_CHECK(ctx, "planes.hull_frame");
        k_hull_keys<<<nblk(n_cur, 256), 256, 0, ctx->stream>>>(plane_pts, n_in, st, keys);
        RTR_LAUNCH_CHECK(ctx, "planes.hull_keys");
        // the inlier count sizes the sort: first of the two host round trips of this plane
        k_plane_counts<<<1, 32, 0, ctx->stream>>>(n_in, n_rest, st);
        RTR_LAUNCH_CHECK(ctx, "planes.counts");
        RTR_CHECK(cudaMemcpyAsync(ctx->pinned, st, sizeof(PlaneState), cudaMemcpyDeviceToHost, ctx->stream), "planes");
        RTR_CHECK(cudaStreamSynchronize(ctx->stream), "planes");
        memcpy(&h_st, ctx->pinned, sizeof(PlaneState));
        if (h_st.error) { rc = rtr_fail("planes", "too many degenerate (collinear) RANSAC samples", RTR_ERR_INVALID); break; }
        if (h_st.best < 0 || h_st.n_inliers == 0) break;                          // "Could not estimate a planar model" (model_point.h:198-202)
        m = h_st.n_inliers;
        RTR_CHECK(cub::DeviceRadixSort::SortKeys(cub_tmp, tb_sort, keys, keys2, m, 0, 64, ctx->stream), "planes.sort");
        k_hull_chain<<<1, 32, 0, ctx->stream>>>(keys2, n_in, stack, st);
        RTR_LAUNCH_CHECK(ctx, "planes.hull_chain");
        RTR_CHECK(cudaMemcpyAsync(ctx->pinned, st, sizeof(PlaneState), cudaMemcpyDeviceToHost, ctx->stream), "planes");
        RTR_CHECK(cudaStreamSynchronize(ctx->stream), "planes");
        memcpy(&h_st, ctx->pinned, sizeof(PlaneState));
        }
        int cls = plane_class(h_st.coeff);
        if (planes < capacity) {
            rtr_surface& s = host_surfaces[planes];
            memset(&s, 0, sizeof(s));
            s.area = h_st.area; memcpy(s.coefficients, h_st.coeff, sizeof(s.coefficients));
            s.is_vertical = (cls == 2) ? 1 : 0; s.inliers = m; s.dimension = h_st.dimension; s.iterations = h_st.iterations;
            s.kept = (cls != 0 && h_st.area >= 0.16) ? 1 : 0;                     // model_point.h:223-233
        }
        ++planes;
        std::swap(cur, next);
        n_cur = h_st.n_rest;
    }
    *n_planes = planes;
    if (rc == 0 && planes > capacity) rc = RTR_ERR_CAPACITY;
    return rc;
}
