// oracle.cpp — CPU restatement of the model-to-scene registration path of ICCD/RealTime_Robot.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under realtime_robot_b200/ links, imports or calls this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.
//
// PARITY STATUS
//   * orc_tdf follows the reference's only CUDA kernel (RealTimeRobot/kernel.cu:8-31) and is PINNED:
//     tests/golden/tdf_ref_*.npz hold outputs of that kernel compiled unmodified (oracle/_ref) and run on a B200.
//   * every other stage restates PCL 1.8.0 classes the reference instantiates (HarrisKeypoint3D incl. its implicit
//     NormalEstimation, IterativeClosestPoint) or that BASELINE.json's north_star names (FPFHEstimation,
//     SampleConsensusPrerejective).  PCL 1.8.0 is NOT vendored in /root/reference (named only in README.md:3 and
//     RealTimeRobot.vcxproj:75,116) and is not installed here, and the reference has no tests or golden vectors:
//     for these stages **parity is unpinned**.  The algorithms follow SURVEY.md Appendix A and are cross-checked
//     against scipy / numpy and closed-form shapes in tests/.
//
// ARITHMETIC CONTRACT (what the CUDA path is held to)
//   * neighbour membership: float d2 = (dx*dx + dy*dy) + dz*dz, no FMA; radius set: d2 < r*r (strict, self included);
//     nearest: min over (d2, index) lexicographic (FLANN's tie order is traversal dependent -> documented tie rule).
//   * reductions (covariances, histograms, ICP sums) accumulate in double from float inputs and are rounded to float
//     when stored; small solves (3x3 / 4x4 symmetric eigenproblems, 3x3 inverse) are cyclic Jacobi / cofactors in
//     double using only + - * / sqrt.  This is the "exact" mode.  mode 1 ("pcl_float") restates PCL's single-pass
//     float accumulation + closed-form eigen33 for the normals so the deviation can be measured.
//   * point transforms are float: ((m00*x + m01*y) + m02*z) + m03, no FMA (pcl::transformPointCloud).
//   * RANSAC draws come from a counter hash of (seed, hypothesis, draw) — PCL uses C rand(), which is not portable.
//
// Build: see oracle/Makefile (g++ -O3 -march=native -ffp-contract=off -fno-fast-math -fopenmp; `make oracle_asan` for the
// AddressSanitizer / UBSan build that tests/test_oracle_sanitizers.py runs).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/rtr.h"

namespace {

struct P4 { float x, y, z, w; };

static int g_threads = 1;

inline float dist2f(const P4& a, const P4& b) {
    // FLANN L2_Simple<float>: result += diff*diff, x then y then z (App. A.1)
    // (-ffp-contract=off in the Makefile keeps these as separate IEEE float operations)
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    float s = dx * dx + dy * dy;
    return s + dz * dz;
}

// ---------------------------------------------------------------------------------------------
// uniform grid over a cloud (oracle's own index; double-precision cell assignment)
struct Grid {
    double h = 0, inv = 0;
    double mn[3] = {0, 0, 0};
    int dim[3] = {1, 1, 1};
    std::vector<int> begin;   // ncells + 1
    std::vector<int> order;   // point indices grouped by cell, ascending inside a cell
    const P4* pts = nullptr;
    int n = 0;

    void cell_of(const P4& p, int c[3]) const {
        const double v[3] = {p.x, p.y, p.z};
        for (int a = 0; a < 3; ++a) {
            double f = std::floor((v[a] - mn[a]) * inv);
            if (!(f >= 0)) f = 0;
            if (f > dim[a] - 1) f = dim[a] - 1;
            c[a] = (int)f;
        }
    }
    long key(int x, int y, int z) const { return ((long)z * dim[1] + y) * dim[0] + x; }

    void build(const P4* p, int count, double cell) {
        pts = p; n = count;
        double mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        mn[0] = mn[1] = mn[2] = DBL_MAX;
        for (int i = 0; i < n; ++i) {
            const double v[3] = {p[i].x, p[i].y, p[i].z};
            for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); }
        }
        if (n == 0) { mn[0] = mn[1] = mn[2] = 0; mx[0] = mx[1] = mx[2] = 0; }
        h = cell;
        for (;;) {
            inv = 1.0 / h;
            double cells = 1;
            for (int a = 0; a < 3; ++a) { dim[a] = (int)std::floor((mx[a] - mn[a]) * inv) + 1; cells *= dim[a]; }
            if (cells <= 64e6) break;
            h *= 1.5;
        }
        long nc = (long)dim[0] * dim[1] * dim[2];
        begin.assign(nc + 1, 0);
        std::vector<long> keys(n);
        for (int i = 0; i < n; ++i) { int c[3]; cell_of(p[i], c); keys[i] = key(c[0], c[1], c[2]); begin[keys[i] + 1]++; }
        for (long k = 0; k < nc; ++k) begin[k + 1] += begin[k];
        order.resize(n);
        std::vector<int> cur(begin.begin(), begin.end() - 1);
        for (int i = 0; i < n; ++i) order[cur[keys[i]]++] = i;   // ascending index inside each cell
    }

    // all i with d2(q, p_i) < r2, ascending index.  Requires h > r.
    void radius(const P4& q, float r2, std::vector<int>& out) const {
        out.clear();
        int c[3]; cell_of(q, c);
        for (int z = std::max(c[2] - 1, 0); z <= std::min(c[2] + 1, dim[2] - 1); ++z)
            for (int y = std::max(c[1] - 1, 0); y <= std::min(c[1] + 1, dim[1] - 1); ++y) {
                long k0 = key(std::max(c[0] - 1, 0), y, z), k1 = key(std::min(c[0] + 1, dim[0] - 1), y, z);
                for (int s = begin[k0]; s < begin[k1 + 1]; ++s) {
                    int i = order[s];
                    if (dist2f(q, pts[i]) < r2) out.push_back(i);
                }
            }
        std::sort(out.begin(), out.end());
    }

    // exact nearest neighbour, ties -> lowest index; expanding Chebyshev rings.
    void nearest(const P4& q, int& best, float& best_d2) const {
        best = -1; best_d2 = FLT_MAX;
        if (n == 0) return;
        int c[3]; cell_of(q, c);
        int maxr = std::max(dim[0], std::max(dim[1], dim[2]));
        for (int R = 0; R <= maxr; ++R) {
            for (int z = c[2] - R; z <= c[2] + R; ++z) {
                if (z < 0 || z >= dim[2]) continue;
                for (int y = c[1] - R; y <= c[1] + R; ++y) {
                    if (y < 0 || y >= dim[1]) continue;
                    bool shell_zy = (std::abs(z - c[2]) == R) || (std::abs(y - c[1]) == R);
                    for (int x = c[0] - R; x <= c[0] + R; ++x) {
                        if (x < 0 || x >= dim[0]) continue;
                        if (!shell_zy && std::abs(x - c[0]) != R) continue;
                        long k = key(x, y, z);
                        for (int s = begin[k]; s < begin[k + 1]; ++s) {
                            int i = order[s];
                            float d = dist2f(q, pts[i]);
                            if (d < best_d2 || (d == best_d2 && i < best)) { best_d2 = d; best = i; }
                        }
                    }
                }
            }
            // every unvisited point is farther than R*h along some axis (minus cell-assignment slack)
            if (best >= 0) {
                double lim = R * h * 0.999;
                if ((double)best_d2 <= lim * lim) return;
            }
        }
    }

    // nearest within radius (d2 < r2); -1 if none.  Requires h > r.
    void nearest_within(const P4& q, float r2, int& best, float& best_d2) const {
        best = -1; best_d2 = FLT_MAX;
        int c[3];
        // a query farther than h outside the box cannot have a neighbour within r < h
        const double v[3] = {q.x, q.y, q.z};
        for (int a = 0; a < 3; ++a) {
            double f = std::floor((v[a] - mn[a]) * inv);
            if (f < -1 || f > dim[a]) return;
        }
        cell_of(q, c);
        for (int z = std::max(c[2] - 1, 0); z <= std::min(c[2] + 1, dim[2] - 1); ++z)
            for (int y = std::max(c[1] - 1, 0); y <= std::min(c[1] + 1, dim[1] - 1); ++y) {
                long k0 = key(std::max(c[0] - 1, 0), y, z), k1 = key(std::min(c[0] + 1, dim[0] - 1), y, z);
                for (int s = begin[k0]; s < begin[k1 + 1]; ++s) {
                    int i = order[s];
                    float d = dist2f(q, pts[i]);
                    if (d < r2 && (d < best_d2 || (d == best_d2 && i < best))) { best_d2 = d; best = i; }
                }
            }
    }
};

// ---------------------------------------------------------------------------------------------
// cyclic Jacobi for a symmetric NxN matrix (double, only + - * / sqrt).  a is destroyed: eigenvalues on the
// diagonal, eigenvectors in the columns of v.
template <int N>
void jacobi_eig(double a[N][N], double v[N][N]) {
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0;
        for (int p = 0; p < N; ++p) for (int q = p + 1; q < N; ++q) off += std::fabs(a[p][q]);
        if (off == 0.0) break;
        for (int p = 0; p < N; ++p)
            for (int q = p + 1; q < N; ++q) {
                double apq = a[p][q];
                if (apq == 0.0) continue;
                double g = 100.0 * std::fabs(apq);
                if (sweep > 3 && std::fabs(a[p][p]) + g == std::fabs(a[p][p]) && std::fabs(a[q][q]) + g == std::fabs(a[q][q])) {
                    a[p][q] = 0.0; a[q][p] = 0.0; continue;
                }
                double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                double t = 1.0 / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                if (theta < 0) t = -t;
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                a[p][p] = a[p][p] - t * apq;
                a[q][q] = a[q][q] + t * apq;
                a[p][q] = 0.0; a[q][p] = 0.0;
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

// Largest eigenpair of the symmetric 4x4 Horn matrix without an iterative diagonalisation: the characteristic quartic's
// top root by Newton from an upper bound (monotone for real-rooted polynomials), then one column of adj(N - lambda I).
// + - * / sqrt only, fixed order.  ~40x shorter dependent chain than cyclic Jacobi (the ICP solve is one serial thread).
// Returns false when the top eigenvalue is not well separated (collinear / symmetric inputs): the caller falls back
// to jacobi_eig, which handles repeated eigenvalues gracefully.
void adj4(double m00, double m01, double m02, double m03, double m10, double m11, double m12, double m13,
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
bool horn_top_eigvec(const double N[4][4], double q[4]) {
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
    double lam = std::sqrt(0.75 * F2) * 1.000000001;
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
    double best = std::fabs(A[0]);
    if (std::fabs(A[5]) > best) { best = std::fabs(A[5]); col = 1; }
    if (std::fabs(A[10]) > best) { best = std::fabs(A[10]); col = 2; }
    if (std::fabs(A[15]) > best) { best = std::fabs(A[15]); col = 3; }
    q[0] = A[0 * 4 + col]; q[1] = A[1 * 4 + col]; q[2] = A[2 * 4 + col]; q[3] = A[3 * 4 + col];
    const double nq2 = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
    // well separated top eigenvalue only (product of gaps > ~1e-4 |N|^3); otherwise the caller runs the Jacobi solve
    return nq2 > 1e-9 * ((F2 * F2) * F2);
}

// Rigid transform maximising sum t_i . (R s_i): Horn's unit-quaternion solution.  Inputs are raw double sums
// over n correspondences: ss[3] = sum s, st[3] = sum t, m[a][b] = sum s_a * t_b.  Output: 16 floats, column-major.
// Same optimum as PCL's TransformationEstimationSVD / Eigen::umeyama without scale (App. A.6).
void horn_pose(const double ss[3], const double st[3], const double m[3][3], double n, float pose[16]) {
    double cs[3], ct[3], S[3][3];
    for (int a = 0; a < 3; ++a) { cs[a] = ss[a] / n; ct[a] = st[a] / n; }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] = m[a][b] - n * cs[a] * ct[b];
    double N[4][4], V[4][4];
    N[0][0] = S[0][0] + S[1][1] + S[2][2];
    N[0][1] = S[1][2] - S[2][1]; N[0][2] = S[2][0] - S[0][2]; N[0][3] = S[0][1] - S[1][0];
    N[1][1] = S[0][0] - S[1][1] - S[2][2];
    N[1][2] = S[0][1] + S[1][0]; N[1][3] = S[2][0] + S[0][2];
    N[2][2] = -S[0][0] + S[1][1] - S[2][2];
    N[2][3] = S[1][2] + S[2][1];
    N[3][3] = -S[0][0] - S[1][1] + S[2][2];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < i; ++j) N[i][j] = N[j][i];
    double q0, q1, q2, q3, qv[4];
    if (horn_top_eigvec(N, qv)) { q0 = qv[0]; q1 = qv[1]; q2 = qv[2]; q3 = qv[3]; }
    else {
        jacobi_eig<4>(N, V);
        int best = 0;
        for (int i = 1; i < 4; ++i) if (N[i][i] > N[best][best]) best = i;
        q0 = V[0][best]; q1 = V[1][best]; q2 = V[2][best]; q3 = V[3][best];
    }
    double nq = std::sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 /= nq; q1 /= nq; q2 /= nq; q3 /= nq;
    double R[3][3];
    R[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3; R[0][1] = 2.0 * (q1 * q2 - q0 * q3); R[0][2] = 2.0 * (q1 * q3 + q0 * q2);
    R[1][0] = 2.0 * (q1 * q2 + q0 * q3); R[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3; R[1][2] = 2.0 * (q2 * q3 - q0 * q1);
    R[2][0] = 2.0 * (q1 * q3 - q0 * q2); R[2][1] = 2.0 * (q2 * q3 + q0 * q1); R[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
    for (int r = 0; r < 3; ++r) {
        double t = ct[r] - ((R[r][0] * cs[0] + R[r][1] * cs[1]) + R[r][2] * cs[2]);
        for (int c = 0; c < 3; ++c) pose[c * 4 + r] = (float)R[r][c];
        pose[12 + r] = (float)t;
    }
    pose[3] = pose[7] = pose[11] = 0.f; pose[15] = 1.f;
}

inline P4 xform(const float m[16], const P4& p) {
    // pcl::transformPointCloud: rows of the column-major 4x4, float, left to right, no FMA
    P4 o;
    o.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12];
    o.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
    o.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14];
    o.w = 1.0f;
    return o;
}

// c = a * b, 4x4 column-major float, k = 0..3 left to right, no FMA (final = step * final, App. A.6)
void matmul4(const float a[16], const float b[16], float c[16]) {
    float out[16];
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            float acc = a[0 * 4 + row] * b[col * 4 + 0];
            for (int k = 1; k < 4; ++k) acc = acc + a[k * 4 + row] * b[col * 4 + k];
            out[col * 4 + row] = acc;
        }
    std::memcpy(c, out, sizeof(out));
}

inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline uint32_t rand_below(uint64_t seed, uint64_t h, uint32_t draw, uint32_t n) {
    uint64_t r = mix64(mix64(seed + 0x9E3779B97F4A7C15ULL * (h + 1)) + 0x9E3779B97F4A7C15ULL * (uint64_t)(draw + 1));
    return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32);
}

// PCL eigen33 restated in float (App. A.2) — used only by mode 1 of orc_normals.
void pcl_eigen33_smallest(const float cov[9], float& eval, float evec[3]) {
    float scale = 0;
    for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(cov[i]));
    if (scale <= FLT_MIN) scale = 1.0f;
    float m[9];
    for (int i = 0; i < 9; ++i) m[i] = cov[i] / scale;
    // characteristic polynomial x^3 - c2 x^2 + c1 x - c0
    float c0 = m[0] * m[4] * m[8] + 2.f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
    float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
    float c2 = m[0] + m[4] + m[8];
    float roots[3];
    auto quad = [&](float b, float c) {   // roots of x^2 - b x + c
        roots[0] = 0.f;
        float d = b * b - 4.f * c; if (d < 0.f) d = 0.f;
        float sd = std::sqrt(d);
        roots[2] = 0.5f * (b + sd); roots[1] = 0.5f * (b - sd);
    };
    if (std::fabs(c0) < FLT_EPSILON) quad(c2, c1);
    else {
        const float s_inv3 = 1.0f / 3.0f, s_sqrt3 = std::sqrt(3.0f);
        float c2_over_3 = c2 * s_inv3;
        float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3; if (a_over_3 > 0.f) a_over_3 = 0.f;
        float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
        float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3; if (q > 0.f) q = 0.f;
        float rho = std::sqrt(-a_over_3);
        float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
        float ct = std::cos(theta), st = std::sin(theta);
        roots[0] = c2_over_3 + 2.f * rho * ct;
        roots[1] = c2_over_3 - rho * (ct + s_sqrt3 * st);
        roots[2] = c2_over_3 - rho * (ct - s_sqrt3 * st);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
        if (roots[1] >= roots[2]) { std::swap(roots[1], roots[2]); if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]); }
        if (roots[0] <= 0) quad(c2, c1);
    }
    eval = roots[0] * scale;
    float d0 = m[0] - roots[0], d1 = m[4] - roots[0], d2 = m[8] - roots[0];
    float r0[3] = {d0, m[1], m[2]}, r1[3] = {m[1], d1, m[5]}, r2[3] = {m[2], m[5], d2};
    auto cross = [](const float* a, const float* b, float* o) {
        o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
    };
    float v0[3], v1[3], v2[3];
    cross(r0, r1, v0); cross(r0, r2, v1); cross(r1, r2, v2);
    float l0 = v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2];
    float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    const float* v = v0; float l = l0;
    if (l0 >= l1 && l0 >= l2) { v = v0; l = l0; } else if (l1 >= l0 && l1 >= l2) { v = v1; l = l1; } else { v = v2; l = l2; }
    float inv = 1.0f / std::sqrt(l);
    evec[0] = v[0] * inv; evec[1] = v[1] * inv; evec[2] = v[2] * inv;
}

const float kNaN = std::numeric_limits<float>::quiet_NaN();

struct RadiusIndex {   // grid for radius r, or brute force for tiny clouds / cross-checks
    Grid g; const P4* pts; int n; float r2; bool brute;
    RadiusIndex(const P4* p, int count, float radius, bool force_brute = false) : pts(p), n(count), brute(force_brute) {
        r2 = radius * radius;
        if (!brute) g.build(p, count, (double)radius * 1.001);
    }
    void query(const P4& q, std::vector<int>& out) const {
        if (!brute) { g.radius(q, r2, out); return; }
        out.clear();
        for (int i = 0; i < n; ++i) if (dist2f(q, pts[i]) < r2) out.push_back(i);
    }
};

// ---------------------------------------------------------------------------------------------
// stages

// pcl::NormalEstimation::computeFeature / computePointNormal (App. A.2)
void normals_impl(const P4* pts, int n, float radius, int mode, float* out4) {
    RadiusIndex idx(pts, n, radius);
#pragma omp parallel for schedule(dynamic, 256) num_threads(g_threads)
    for (int i = 0; i < n; ++i) {
        std::vector<int> nb;
        idx.query(pts[i], nb);
        float* o = out4 + 4 * (size_t)i;
        if (nb.size() < 3) { o[0] = o[1] = o[2] = o[3] = kNaN; continue; }
        double nx, ny, nz, curv;
        if (mode == 0) {
            // exact: double sums of offsets from the query point
            double s[3] = {0, 0, 0}, c[6] = {0, 0, 0, 0, 0, 0};
            for (int j : nb) {
                double dx = (double)pts[j].x - (double)pts[i].x, dy = (double)pts[j].y - (double)pts[i].y, dz = (double)pts[j].z - (double)pts[i].z;
                s[0] += dx; s[1] += dy; s[2] += dz;
                c[0] += dx * dx; c[1] += dx * dy; c[2] += dx * dz; c[3] += dy * dy; c[4] += dy * dz; c[5] += dz * dz;
            }
            double k = (double)nb.size();
            double mx = s[0] / k, my = s[1] / k, mz = s[2] / k;
            double a[3][3], v[3][3];
            a[0][0] = c[0] / k - mx * mx; a[0][1] = c[1] / k - mx * my; a[0][2] = c[2] / k - mx * mz;
            a[1][1] = c[3] / k - my * my; a[1][2] = c[4] / k - my * mz; a[2][2] = c[5] / k - mz * mz;
            a[1][0] = a[0][1]; a[2][0] = a[0][2]; a[2][1] = a[1][2];
            double trace = a[0][0] + a[1][1] + a[2][2];
            jacobi_eig<3>(a, v);
            int b = 0;
            for (int e = 1; e < 3; ++e) if (a[e][e] < a[b][b]) b = e;
            nx = v[0][b]; ny = v[1][b]; nz = v[2][b];
            curv = (trace > 0) ? std::fabs(a[b][b] / trace) : 0.0;
        } else {
            // pcl_float: single pass float accumulators, ascending index order, closed-form eigen33
            float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j : nb) {
                const P4& p = pts[j];
                acc[0] += p.x * p.x; acc[1] += p.x * p.y; acc[2] += p.x * p.z; acc[3] += p.y * p.y; acc[4] += p.y * p.z; acc[5] += p.z * p.z;
                acc[6] += p.x; acc[7] += p.y; acc[8] += p.z;
            }
            float k = (float)nb.size();
            for (int e = 0; e < 9; ++e) acc[e] /= k;
            float cov[9];
            cov[0] = acc[0] - acc[6] * acc[6]; cov[1] = acc[1] - acc[6] * acc[7]; cov[2] = acc[2] - acc[6] * acc[8];
            cov[4] = acc[3] - acc[7] * acc[7]; cov[5] = acc[4] - acc[7] * acc[8]; cov[8] = acc[5] - acc[8] * acc[8];
            cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
            float ev, evec[3];
            pcl_eigen33_smallest(cov, ev, evec);
            nx = evec[0]; ny = evec[1]; nz = evec[2];
            float tr = cov[0] + cov[4] + cov[8];
            curv = (tr != 0) ? std::fabs(ev / tr) : 0.0;
        }
        // flipNormalTowardsViewpoint, viewpoint (0,0,0): n . (vp - p) >= 0
        double dot = (nx * -(double)pts[i].x + ny * -(double)pts[i].y) + nz * -(double)pts[i].z;
        if (dot < 0) { nx = -nx; ny = -ny; nz = -nz; }
        o[0] = (float)nx; o[1] = (float)ny; o[2] = (float)nz; o[3] = (float)curv;
    }
}

inline bool finite3(const float* v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

// pcl::HarrisKeypoint3D::responseHarris (App. A.3)
void harris_response_impl(const P4* pts, int n, const float* nrm4, const RadiusIndex& idx, float* resp) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(g_threads)
    for (int i = 0; i < n; ++i) {
        std::vector<int> nb;
        idx.query(pts[i], nb);
        double c[6] = {0, 0, 0, 0, 0, 0}; int cnt = 0;
        for (int j : nb) {
            const float* nj = nrm4 + 4 * (size_t)j;
            if (!finite3(nj)) continue;
            double x = nj[0], y = nj[1], z = nj[2];
            c[0] += x * x; c[1] += x * y; c[2] += x * z; c[3] += y * y; c[4] += y * z; c[5] += z * z; ++cnt;
        }
        if (cnt == 0) { resp[i] = 0.f; continue; }
        double k = cnt;
        double xx = c[0] / k, xy = c[1] / k, xz = c[2] / k, yy = c[3] / k, yz = c[4] / k, zz = c[5] / k;
        double trace = xx + yy + zz;
        if (trace != 0) {
            double det = xx * yy * zz + 2.0 * xy * xz * yz - xz * xz * yy - xy * xy * zz - yz * yz * xx;
            resp[i] = (float)(0.04 + det - 0.04 * trace * trace);
        } else resp[i] = 0.f;
    }
}

// refineCorners (App. A.3): <= 10 Gauss-Newton steps  corner <- (sum n n^T)^-1 sum n n^T p  over the r-neighbours of the
// moving corner.  The two sums are accumulated in 64-bit fixed point (terms rounded to 2^-40 resp. 2^-34) so that the
// result does not depend on the order of the neighbours — this stage is ill-conditioned, and the CUDA path reduces the
// same integers with warp shuffles.
void harris_refine_one(const P4* pts, const float* nrm4, const RadiusIndex& idx, P4& corner) {
    const double SA = 1099511627776.0, SB = 17179869184.0;   // 2^40, 2^34
    std::vector<int> nb;
    int it = 0; double diff;
    do {
        P4 cur = corner;
        idx.query(cur, nb);
        long long q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j : nb) {
            const float* nj = nrm4 + 4 * (size_t)j;
            if (!finite3(nj)) continue;
            double x = nj[0], y = nj[1], z = nj[2];
            double xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
            double px = pts[j].x, py = pts[j].y, pz = pts[j].z;
            q[0] += std::llrint(xx * SA); q[1] += std::llrint(xy * SA); q[2] += std::llrint(xz * SA);
            q[3] += std::llrint(yy * SA); q[4] += std::llrint(yz * SA); q[5] += std::llrint(zz * SA);
            q[6] += std::llrint(((xx * px + xy * py) + xz * pz) * SB);
            q[7] += std::llrint(((xy * px + yy * py) + yz * pz) * SB);
            q[8] += std::llrint(((xz * px + yz * py) + zz * pz) * SB);
        }
        double A[6], b[3];
        for (int k = 0; k < 6; ++k) A[k] = (double)q[k] / SA;
        for (int k = 0; k < 3; ++k) b[k] = (double)q[6 + k] / SB;
        // symmetric 3x3 inverse by cofactors
        double c00 = A[3] * A[5] - A[4] * A[4], c01 = A[2] * A[4] - A[1] * A[5], c02 = A[1] * A[4] - A[2] * A[3];
        double c11 = A[0] * A[5] - A[2] * A[2], c12 = A[1] * A[2] - A[0] * A[4], c22 = A[0] * A[3] - A[1] * A[1];
        double det = (A[0] * c00 + A[1] * c01) + A[2] * c02;
        if (det != 0) {
            corner.x = (float)(((c00 * b[0] + c01 * b[1]) + c02 * b[2]) / det);
            corner.y = (float)(((c01 * b[0] + c11 * b[1]) + c12 * b[2]) / det);
            corner.z = (float)(((c02 * b[0] + c12 * b[1]) + c22 * b[2]) / det);
        }
        double dx = (double)corner.x - (double)cur.x, dy = (double)corner.y - (double)cur.y, dz = (double)corner.z - (double)cur.z;
        diff = (dx * dx + dy * dy) + dz * dz;
    } while (diff > 1e-6 && ++it < 10);
}

int harris_impl(const P4* pts, int n, const float* nrm4, float radius, float thr, int nms, int refine,
                float* resp_out, int* kp_idx, float* kp_xyz1, int cap) {
    RadiusIndex idx(pts, n, radius);
    std::vector<float> resp(n);
    harris_response_impl(pts, n, nrm4, idx, resp.data());
    if (resp_out) std::memcpy(resp_out, resp.data(), sizeof(float) * n);
    std::vector<char> keep(n, 0);
#pragma omp parallel for schedule(dynamic, 256) num_threads(g_threads)
    for (int i = 0; i < n; ++i) {
        float r = resp[i];
        if (!std::isfinite(r) || !(r >= thr)) continue;
        bool is_max = true;
        if (nms) {
            std::vector<int> nb;
            idx.query(pts[i], nb);
            for (int j : nb) if (resp[j] > r) { is_max = false; break; }
        }
        keep[i] = is_max;
    }
    int m = 0;
    for (int i = 0; i < n; ++i) {
        if (!keep[i]) continue;
        if (m < cap) {
            P4 c = pts[i]; c.w = 1.0f;
            if (refine) harris_refine_one(pts, nrm4, idx, c);
            if (kp_idx) kp_idx[m] = i;
            if (kp_xyz1) { kp_xyz1[4 * m + 0] = c.x; kp_xyz1[4 * m + 1] = c.y; kp_xyz1[4 * m + 2] = c.z; kp_xyz1[4 * m + 3] = 1.0f; }
        }
        ++m;
    }
    return m;
}

// pcl::computePairFeatures in double with the |a1| < |a2| swap criterion (== acos|a1| > acos|a2|), App. A.4
inline bool pair_features(const P4& p1, const float* n1f, const P4& p2, const float* n2f, int bins[3]) {
    double d[3] = {(double)p2.x - (double)p1.x, (double)p2.y - (double)p1.y, (double)p2.z - (double)p1.z};
    double f4 = std::sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    if (f4 == 0.0) return false;
    double n1[3] = {n1f[0], n1f[1], n1f[2]}, n2[3] = {n2f[0], n2f[1], n2f[2]};
    double a1 = ((n1[0] * d[0] + n1[1] * d[1]) + n1[2] * d[2]) / f4;
    double a2 = ((n2[0] * d[0] + n2[1] * d[1]) + n2[2] * d[2]) / f4;
    double f3;
    if (std::fabs(a1) < std::fabs(a2)) {
        for (int a = 0; a < 3; ++a) { std::swap(n1[a], n2[a]); d[a] = -d[a]; }
        f3 = -a2;
    } else f3 = a1;
    double v[3] = {d[1] * n1[2] - d[2] * n1[1], d[2] * n1[0] - d[0] * n1[2], d[0] * n1[1] - d[1] * n1[0]};
    double vn = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    if (vn == 0.0) return false;
    v[0] /= vn; v[1] /= vn; v[2] /= vn;
    double w[3] = {n1[1] * v[2] - n1[2] * v[1], n1[2] * v[0] - n1[0] * v[2], n1[0] * v[1] - n1[1] * v[0]};
    double f2 = (v[0] * n2[0] + v[1] * n2[1]) + v[2] * n2[2];
    double f1 = std::atan2((w[0] * n2[0] + w[1] * n2[1]) + w[2] * n2[2], (n1[0] * n2[0] + n1[1] * n2[1]) + n1[2] * n2[2]);
    const double kPi = 3.14159265358979323846;
    int b0 = (int)std::floor(11.0 * ((f1 + kPi) * (1.0 / (2.0 * kPi))));
    int b1 = (int)std::floor(11.0 * ((f2 + 1.0) * 0.5));
    int b2 = (int)std::floor(11.0 * ((f3 + 1.0) * 0.5));
    bins[0] = std::min(std::max(b0, 0), 10); bins[1] = std::min(std::max(b1, 0), 10); bins[2] = std::min(std::max(b2, 0), 10);
    return true;
}

// pcl::FPFHEstimation::computeFeature, input == surface (App. A.4)
void fpfh_impl(const P4* pts, int n, const float* nrm4, float radius, float* out33) {
    RadiusIndex idx(pts, n, radius);
    std::vector<float> spfh((size_t)n * 33, 0.f);
#pragma omp parallel for schedule(dynamic, 128) num_threads(g_threads)
    for (int i = 0; i < n; ++i) {
        std::vector<int> nb;
        idx.query(pts[i], nb);
        const float* ni = nrm4 + 4 * (size_t)i;
        if (nb.size() < 2 || !finite3(ni)) continue;
        int cnt[33]; std::memset(cnt, 0, sizeof(cnt));
        for (int j : nb) {
            if (j == i) continue;
            const float* nj = nrm4 + 4 * (size_t)j;
            if (!finite3(nj)) continue;
            int b[3];
            if (!pair_features(pts[i], ni, pts[j], nj, b)) continue;
            cnt[b[0]]++; cnt[11 + b[1]]++; cnt[22 + b[2]]++;
        }
        double incr = 100.0 / (double)(nb.size() - 1);
        for (int k = 0; k < 33; ++k) spfh[(size_t)i * 33 + k] = (float)(cnt[k] * incr);
    }
#pragma omp parallel for schedule(dynamic, 128) num_threads(g_threads)
    for (int i = 0; i < n; ++i) {
        std::vector<int> nb;
        idx.query(pts[i], nb);
        double h[33]; for (int k = 0; k < 33; ++k) h[k] = 0;
        float* o = out33 + (size_t)i * 33;
        bool any = false;
        for (int j : nb) {
            float d2 = dist2f(pts[i], pts[j]);
            if (d2 == 0.f) continue;                 // "minus the query point itself"
            any = true;
            double w = 1.0 / (double)d2;             // weight = 1 / squared distance
            const float* s = &spfh[(size_t)j * 33];
            for (int k = 0; k < 33; ++k) h[k] += (double)s[k] * w;
        }
        if (nb.empty()) { for (int k = 0; k < 33; ++k) o[k] = kNaN; continue; }
        (void)any;
        for (int t = 0; t < 3; ++t) {
            double sum = 0;
            for (int k = 0; k < 11; ++k) sum += h[t * 11 + k];
            double sc = (sum != 0) ? 100.0 / sum : 0.0;
            for (int k = 0; k < 11; ++k) o[t * 11 + k] = (float)(h[t * 11 + k] * sc);
        }
    }
}

inline double feat_dist(const float* a, const float* b) {
    double s = 0;
    for (int k = 0; k < 33; ++k) { double d = (double)a[k] - (double)b[k]; s += d * d; }
    return s;
}

// k nearest target features per source feature (App. A.5 findSimilarFeatures), brute force
void match_impl(const float* fa, int na, const float* fb, int nb, int k, int* idx, float* dist) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(g_threads)
    for (int i = 0; i < na; ++i) {
        std::vector<std::pair<float, int>> best;   // ascending (dist, idx)
        const float* a = fa + (size_t)i * 33;
        for (int j = 0; j < nb; ++j) {
            float d = (float)feat_dist(a, fb + (size_t)j * 33);
            if (!(d == d)) continue;
            std::pair<float, int> e(d, j);
            if ((int)best.size() < k) { best.push_back(e); std::sort(best.begin(), best.end()); }
            else if (e < best.back()) { best.back() = e; std::sort(best.begin(), best.end()); }
        }
        for (int t = 0; t < k; ++t) {
            idx[(size_t)i * k + t] = t < (int)best.size() ? best[t].second : -1;
            if (dist) dist[(size_t)i * k + t] = t < (int)best.size() ? best[t].first : kNaN;
        }
    }
}

struct Hyp { int s[3], c[3]; };

inline bool draw_hypothesis(uint64_t seed, uint64_t h, int ns, const int* knn, int knn_stride, int k, Hyp& hy) {
    // three distinct source indices (uniform without replacement), then one of the k most similar target features each
    uint32_t a = rand_below(seed, h, 0, ns);
    uint32_t b = rand_below(seed, h, 1, ns - 1); if (b >= a) ++b;
    uint32_t c = rand_below(seed, h, 2, ns - 2);
    uint32_t lo = std::min(a, b), hi = std::max(a, b);
    if (c >= lo) ++c;
    if (c >= hi) ++c;
    hy.s[0] = a; hy.s[1] = b; hy.s[2] = c;
    for (int t = 0; t < 3; ++t) {
        uint32_t pick = (k > 1) ? rand_below(seed, h, 3 + t, k) : 0;
        hy.c[t] = knn[(size_t)hy.s[t] * knn_stride + pick];
        if (hy.c[t] < 0) return false;
    }
    return true;
}

// CorrespondenceRejectorPoly::thresholdPolygon, cardinality 3 (App. A.5)
inline bool polygon_ok(const P4* src, const P4* tgt, const Hyp& hy, float simsq) {
    for (int e = 0; e < 3; ++e) {
        int f = (e + 1) % 3;
        float ds = dist2f(src[hy.s[e]], src[hy.s[f]]);
        float dt = dist2f(tgt[hy.c[e]], tgt[hy.c[f]]);
        float sim = ds < dt ? ds / dt : dt / ds;
        if (!(sim >= simsq)) return false;
    }
    return true;
}

void pose_from_pairs(const P4* src, const P4* tgt, const int* si, const int* ti, int cnt, float pose[16]) {
    double ss[3] = {0, 0, 0}, st[3] = {0, 0, 0}, m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int e = 0; e < cnt; ++e) {
        const double s[3] = {src[si[e]].x, src[si[e]].y, src[si[e]].z}, t[3] = {tgt[ti[e]].x, tgt[ti[e]].y, tgt[ti[e]].z};
        for (int a = 0; a < 3; ++a) { ss[a] += s[a]; st[a] += t[a]; for (int b = 0; b < 3; ++b) m[a][b] += s[a] * t[b]; }
    }
    horn_pose(ss, st, m, (double)cnt, pose);
}

void identity16(float m[16]) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }

// SampleConsensusPrerejective::computeTransformation + getFitness (App. A.5)
void ransac_impl(const P4* src, int ns, const P4* tgt, int nt, const int* knn, int knn_stride,
                 const rtr_ransac_params* p, rtr_pose_result* res) {
    std::memset(res, 0, sizeof(*res));
    identity16(res->pose);
    res->fitness = FLT_MAX; res->hypothesis = -1;
    long long h0 = p->hypothesis_begin, h1 = (p->hypothesis_end > 0) ? p->hypothesis_end : p->max_iterations;
    if (ns < 3 || nt < 1 || h1 <= h0) return;
    const float dmax2 = p->max_correspondence_distance * p->max_correspondence_distance;
    const float simsq = p->similarity_threshold * p->similarity_threshold;
    Grid g; g.build(tgt, nt, (double)p->max_correspondence_distance * 1.001);
    long long evaluated = 0;
    struct Best { float err; long long h; int inl; float pose[16]; };
    Best best; best.err = FLT_MAX; best.h = -1; best.inl = 0; identity16(best.pose);
#pragma omp parallel num_threads(g_threads)
    {
        Best loc; loc.err = FLT_MAX; loc.h = -1; loc.inl = 0; identity16(loc.pose);
        long long loc_eval = 0;
#pragma omp for schedule(dynamic, 64)
        for (long long h = h0; h < h1; ++h) {
            Hyp hy;
            if (!draw_hypothesis(p->seed, (uint64_t)h, ns, knn, knn_stride, p->correspondence_k, hy)) continue;
            if (!polygon_ok(src, tgt, hy, simsq)) continue;
            ++loc_eval;
            float pose[16];
            pose_from_pairs(src, tgt, hy.s, hy.c, 3, pose);
            int inl = 0; double sum = 0;
            for (int i = 0; i < ns; ++i) {
                P4 q = xform(pose, src[i]);
                int b; float d2;
                g.nearest_within(q, dmax2, b, d2);
                if (b >= 0) { ++inl; sum += (double)d2; }
            }
            float err = inl > 0 ? (float)(sum / (double)inl) : FLT_MAX;
            float frac = (float)inl / (float)ns;
            if (frac >= p->inlier_fraction && (err < loc.err || (err == loc.err && h < loc.h))) {
                loc.err = err; loc.h = h; loc.inl = inl; std::memcpy(loc.pose, pose, sizeof(pose));
            }
        }
#pragma omp critical
        {
            evaluated += loc_eval;
            if (loc.h >= 0 && (loc.err < best.err || (loc.err == best.err && (best.h < 0 || loc.h < best.h)))) best = loc;
        }
    }
    res->evaluated = evaluated;
    if (best.h >= 0) {
        std::memcpy(res->pose, best.pose, sizeof(best.pose));
        res->fitness = best.err; res->inliers = best.inl; res->hypothesis = best.h; res->converged = 1;
    }
}

// getFitnessScore(): mean squared NN distance of (pose o source) to target
float fitness_impl(const P4* src, int ns, const Grid& g, const float pose[16]) {
    double sum = 0; long cnt = 0;
#pragma omp parallel for reduction(+ : sum, cnt) schedule(static) num_threads(g_threads)
    for (int i = 0; i < ns; ++i) {
        P4 q = xform(pose, src[i]);
        int b; float d2; g.nearest(q, b, d2);
        if (b >= 0) { sum += (double)d2; ++cnt; }
    }
    return cnt ? (float)(sum / (double)cnt) : FLT_MAX;
}

// x = A^-1 b by Gaussian elimination with partial pivoting (first largest |pivot|), fp64, fixed operation order
// (stands in for Eigen's ATA.inverse() * ATb of TransformationEstimationPointToPlaneLLS); false if singular
bool solve6(double A[6][6], double b[6], double x[6]) {
    for (int c = 0; c < 6; ++c) {
        int piv = c; double best = std::fabs(A[c][c]);
        for (int r = c + 1; r < 6; ++r) { double v = std::fabs(A[r][c]); if (v > best) { best = v; piv = r; } }
        if (!(best > 0.0)) return false;
        if (piv != c) {
            for (int k = 0; k < 6; ++k) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        for (int r = c + 1; r < 6; ++r) {
            double f = A[r][c] / A[c][c];
            for (int k = c; k < 6; ++k) A[r][k] = A[r][k] - f * A[c][k];
            b[r] = b[r] - f * b[c];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double s = b[r];
        for (int k = r + 1; k < 6; ++k) s = s - A[r][k] * x[k];
        x[r] = s / A[r][r];
    }
    return true;
}

// TransformationEstimationPointToPlaneLLS: sums[0..20] = upper triangle of A^T A (row-major), sums[21..26] = A^T b;
// constructTransformationMatrix(alpha, beta, gamma, tx, ty, tz).  false: singular system (no step)
bool lls_pose(const double* sums, float pose[16]) {
    double A[6][6], b[6], x[6];
    int k = 0;
    for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { A[r][c] = sums[k]; A[c][r] = sums[k]; ++k; }
    for (int r = 0; r < 6; ++r) b[r] = sums[21 + r];
    if (!solve6(A, b, x)) return false;
    const double al = x[0], be = x[1], ga = x[2];
    const double sa = std::sin(al), ca = std::cos(al), sb = std::sin(be), cb = std::cos(be), sg = std::sin(ga), cg = std::cos(ga);
    identity16(pose);
    pose[0] = (float)(cg * cb);  pose[4] = (float)(-sg * ca + cg * sb * sa);  pose[8]  = (float)(sg * sa + cg * sb * ca);   pose[12] = (float)x[3];
    pose[1] = (float)(sg * cb);  pose[5] = (float)(cg * ca + sg * sb * sa);   pose[9]  = (float)(-cg * sa + sg * sb * ca);  pose[13] = (float)x[4];
    pose[2] = (float)(-sb);      pose[6] = (float)(cb * sa);                  pose[10] = (float)(cb * ca);                  pose[14] = (float)x[5];
    return true;
}

// IterativeClosestPoint::computeTransformation + DefaultConvergenceCriteria (App. A.6)
void icp_impl(const P4* src, int ns, const P4* tgt, int nt, const rtr_icp_params* p, const float* init, rtr_pose_result* res,
              const P4* tgt_normals = nullptr) {
    std::memset(res, 0, sizeof(*res));
    float final_[16];
    if (init) std::memcpy(final_, init, sizeof(final_)); else identity16(final_);
    std::memcpy(res->pose, final_, sizeof(final_));
    res->fitness = FLT_MAX; res->hypothesis = -1;
    if (ns < 1 || nt < 1) return;
    double cell;
    {   // cell size ~ 2 x mean spacing of the target, from its bounding box surface estimate
        double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int i = 0; i < nt; ++i) { const double v[3] = {tgt[i].x, tgt[i].y, tgt[i].z}; for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); } }
        double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
        double area = 2.0 * (ex * ey + ey * ez + ex * ez);
        cell = std::max(2.0 * std::sqrt(std::max(area, 1e-12) / std::max(nt, 1)), 1e-4);
    }
    Grid g; g.build(tgt, nt, cell);
    std::vector<P4> cur(ns);
    for (int i = 0; i < ns; ++i) cur[i] = xform(final_, src[i]);
    const double dmax2 = p->max_correspondence_distance > 0 ? (double)p->max_correspondence_distance * (double)p->max_correspondence_distance : DBL_MAX;
    double prev_mse = DBL_MAX;
    int it = 0, state = 0;
    for (;;) {
        double ss[3] = {0, 0, 0}, st[3] = {0, 0, 0}, m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, sumd = 0; long cnt = 0;
        double s0 = 0, s1 = 0, s2 = 0, t0 = 0, t1 = 0, t2 = 0, m00 = 0, m01 = 0, m02 = 0, m10 = 0, m11 = 0, m12 = 0, m20 = 0, m21 = 0, m22 = 0;
#pragma omp parallel for reduction(+ : s0, s1, s2, t0, t1, t2, m00, m01, m02, m10, m11, m12, m20, m21, m22, sumd, cnt) schedule(static) num_threads(g_threads)
        for (int i = 0; i < ns; ++i) {
            int b; float d2; g.nearest(cur[i], b, d2);
            if (b < 0 || (double)d2 > dmax2) continue;
            double sx = cur[i].x, sy = cur[i].y, sz = cur[i].z, tx = tgt[b].x, ty = tgt[b].y, tz = tgt[b].z;
            s0 += sx; s1 += sy; s2 += sz; t0 += tx; t1 += ty; t2 += tz;
            m00 += sx * tx; m01 += sx * ty; m02 += sx * tz; m10 += sy * tx; m11 += sy * ty; m12 += sy * tz; m20 += sz * tx; m21 += sz * ty; m22 += sz * tz;
            sumd += (double)d2; ++cnt;
        }
        ss[0] = s0; ss[1] = s1; ss[2] = s2; st[0] = t0; st[1] = t1; st[2] = t2;
        m[0][0] = m00; m[0][1] = m01; m[0][2] = m02; m[1][0] = m10; m[1][1] = m11; m[1][2] = m12; m[2][0] = m20; m[2][1] = m21; m[2][2] = m22;
        float step[16];
        if (p->estimator == 1) {
            // point-to-plane (IterativeClosestPointWithNormals / TransformationEstimationPointToPlaneLLS): 6x6 A^T A, A^T b in
            // fp64 from the fp32 points and target normals; correspondences with a non-finite normal are skipped
            double S[27]; for (int k = 0; k < 27; ++k) S[k] = 0; sumd = 0; cnt = 0;
            for (int i = 0; i < ns; ++i) {
                int b; float d2; g.nearest(cur[i], b, d2);
                if (b < 0 || (double)d2 > dmax2) continue;
                const P4 nn = tgt_normals[b];
                if (!(std::isfinite(nn.x) && std::isfinite(nn.y) && std::isfinite(nn.z))) continue;
                const double sx = cur[i].x, sy = cur[i].y, sz = cur[i].z, dx = tgt[b].x, dy = tgt[b].y, dz = tgt[b].z, nx = nn.x, ny = nn.y, nz = nn.z;
                const double v[6] = {nz * sy - ny * sz, nx * sz - nz * sx, ny * sx - nx * sy, nx, ny, nz};
                const double dd = ((nx * dx + ny * dy) + nz * dz) - ((nx * sx + ny * sy) + nz * sz);
                int k = 0;
                for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) S[k++] += v[r] * v[c];
                for (int r = 0; r < 6; ++r) S[21 + r] += v[r] * dd;
                sumd += (double)d2; ++cnt;
            }
            res->inliers = (int)cnt;
            if (cnt < 3) { state = 0; break; }
            if (!lls_pose(S, step)) { state = 0; break; }
        } else {
        res->inliers = (int)cnt;
        if (cnt < 3) { state = 0; break; }            // CONVERGENCE_CRITERIA_NO_CORRESPONDENCES
        horn_pose(ss, st, m, (double)cnt, step);
        }
        for (int i = 0; i < ns; ++i) cur[i] = xform(step, cur[i]);
        matmul4(step, final_, final_);
        ++it;
        if (it >= p->max_iterations) { state = 1; break; }
        if (!p->force_iterations) {
            double cos_angle = 0.5 * ((double)step[0] + (double)step[5] + (double)step[10] - 1.0);
            double tsq = (double)step[12] * step[12] + (double)step[13] * step[13] + (double)step[14] * step[14];
            if (cos_angle >= 1.0 && tsq <= 0.0) { state = 2; break; }
            double mse = sumd / (double)cnt;
            if (std::fabs(mse - prev_mse) < p->mse_threshold_absolute) { state = 3; break; }
            prev_mse = mse;
        }
    }
    std::memcpy(res->pose, final_, sizeof(final_));
    res->iterations = it; res->converged = state;
    res->fitness = fitness_impl(src, ns, g, final_);
}

}  // namespace

// =============================================================================================
extern "C" {

void orc_set_threads(int t) {
#ifdef _OPENMP
    g_threads = t > 0 ? t : omp_get_max_threads();
#else
    (void)t; g_threads = 1;
#endif
}
int orc_get_threads() { return g_threads; }

// kernel.cu:8-31 — squared integer voxel distance to the nearest occupied voxel, clamped at 900, stored as float.
// Writes voxels [0, dim^3) (the reference's `>` bound check also lets thread dim^3 write one element more when
// dim^3 < 27000; that extra element is reproduced by the wrapper test, not here).
void orc_tdf(const int* occ, int num_occ, int dim, float* out) {
    for (int v = 0; v < dim * dim * dim; ++v) {
        int z = v / (dim * dim), y = (v - z * dim * dim) / dim, x = v - z * dim * dim - y * dim;
        float best = 900;
        for (int i = 0; i < num_occ; ++i) {
            int dx = x - occ[i * 3 + 0], dy = y - occ[i * 3 + 1], dz = z - occ[i * 3 + 2];
            float t = (float)(dx * dx + dy * dy + dz * dz);
            if (t < best) best = t;
        }
        out[v] = best;
    }
}

// radius neighbour sets of the cloud's own points.  method 0 = brute force, 1 = grid.  CSR output, ascending indices.
long long orc_radius_neighbors(const float* xyz1, int n, float radius, int method, int* counts, long long* offsets,
                               int* indices, long long capacity) {
    const P4* pts = (const P4*)xyz1;
    RadiusIndex idx(pts, n, radius, method == 0);
    long long total = 0;
    std::vector<int> nb;
    for (int i = 0; i < n; ++i) {
        idx.query(pts[i], nb);
        counts[i] = (int)nb.size();
        if (offsets) offsets[i] = total;
        if (indices && total + (long long)nb.size() <= capacity) std::memcpy(indices + total, nb.data(), nb.size() * sizeof(int));
        total += (long long)nb.size();
    }
    if (offsets) offsets[n] = total;
    return total;
}

void orc_nearest(const float* tgt_xyz1, int nt, const float* q_xyz1, int nq, int method, int* idx, float* d2) {
    const P4* t = (const P4*)tgt_xyz1; const P4* q = (const P4*)q_xyz1;
    if (method == 0) {
        for (int i = 0; i < nq; ++i) {
            int b = -1; float bd = FLT_MAX;
            for (int j = 0; j < nt; ++j) { float d = dist2f(q[i], t[j]); if (d < bd) { bd = d; b = j; } }
            idx[i] = b; d2[i] = bd;
        }
        return;
    }
    Grid g;
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = 0; i < nt; ++i) { const double v[3] = {t[i].x, t[i].y, t[i].z}; for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); } }
    double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    double area = 2.0 * (ex * ey + ey * ez + ex * ez);
    g.build(t, nt, std::max(2.0 * std::sqrt(std::max(area, 1e-12) / std::max(nt, 1)), 1e-4));
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int i = 0; i < nq; ++i) g.nearest(q[i], idx[i], d2[i]);
}

void orc_normals(const float* xyz1, int n, float radius, int mode, float* normals4) {
    normals_impl((const P4*)xyz1, n, radius, mode, normals4);
}

int orc_harris3d(const float* xyz1, int n, const float* normals4, float radius, float threshold, int nms, int refine,
                 float* response, int* kp_index, float* kp_xyz1, int capacity) {
    return harris_impl((const P4*)xyz1, n, normals4, radius, threshold, nms, refine, response, kp_index, kp_xyz1, capacity);
}

void orc_fpfh(const float* xyz1, int n, const float* normals4, float radius, float* fpfh33) {
    fpfh_impl((const P4*)xyz1, n, normals4, radius, fpfh33);
}

void orc_match_features(const float* fa, int na, const float* fb, int nb, int k, int* idx, float* dist) {
    match_impl(fa, na, fb, nb, k, idx, dist);
}

void orc_ransac_prerejective(const float* src_xyz1, int ns, const float* tgt_xyz1, int nt, const int* knn, int knn_stride,
                             const rtr_ransac_params* p, rtr_pose_result* res) {
    ransac_impl((const P4*)src_xyz1, ns, (const P4*)tgt_xyz1, nt, knn, knn_stride, p, res);
}

// debugging aid for the parity tests: the samples / correspondences / prerejection verdict / pose of ONE hypothesis
int orc_hypothesis(const float* src_xyz1, int ns, const float* tgt_xyz1, int nt, const int* knn, int knn_stride,
                   const rtr_ransac_params* p, long long h, int* samples6, float* pose16) {
    (void)nt;
    Hyp hy;
    if (!draw_hypothesis(p->seed, (uint64_t)h, ns, knn, knn_stride, p->correspondence_k, hy)) return -1;
    for (int t = 0; t < 3; ++t) { samples6[t] = hy.s[t]; samples6[3 + t] = hy.c[t]; }
    bool ok = polygon_ok((const P4*)src_xyz1, (const P4*)tgt_xyz1, hy, p->similarity_threshold * p->similarity_threshold);
    pose_from_pairs((const P4*)src_xyz1, (const P4*)tgt_xyz1, hy.s, hy.c, 3, pose16);
    return ok ? 1 : 0;
}

void orc_icp(const float* src_xyz1, int ns, const float* tgt_xyz1, int nt, const rtr_icp_params* p, const float* init16,
             rtr_pose_result* res) {
    icp_impl((const P4*)src_xyz1, ns, (const P4*)tgt_xyz1, nt, p, init16, res);
}
// estimator 1 (point-to-plane) needs the target's normals (n x 4 floats, as orc_normals writes them)
void orc_icp_normals(const float* src_xyz1, int ns, const float* tgt_xyz1, int nt, const float* tgt_normals4, const rtr_icp_params* p,
                     const float* init16, rtr_pose_result* res) {
    icp_impl((const P4*)src_xyz1, ns, (const P4*)tgt_xyz1, nt, p, init16, res, (const P4*)tgt_normals4);
}

// least-squares rigid pose from explicit pairs (for cross-checking Horn against numpy's Kabsch)
void orc_pose_from_pairs(const float* src_xyz1, const float* tgt_xyz1, int n, float* pose16) {
    std::vector<int> id(n);
    for (int i = 0; i < n; ++i) id[i] = i;
    pose_from_pairs((const P4*)src_xyz1, (const P4*)tgt_xyz1, id.data(), id.data(), n, pose16);
}

void orc_transform(const float* xyz1, int n, const float* pose16, float* out_xyz1) {
    const P4* p = (const P4*)xyz1; P4* o = (P4*)out_xyz1;
    for (int i = 0; i < n; ++i) o[i] = xform(pose16, p[i]);
}

// symmetric eigen-solver exposed for unit tests (n = 3 or 4; row-major in, eigenvalues + row-major eigenvector columns out)
// test hook: the quartic / adjugate route of horn_pose on a symmetric 4x4 (row-major); returns 1 when it applies
int orc_horn_top_eigvec(const double* n16, double* q4) {
    double N[4][4];
    std::memcpy(N, n16, sizeof(N));
    return horn_top_eigvec(N, q4) ? 1 : 0;
}

void orc_jacobi(const double* a_in, int n, double* evals, double* evecs) {
    if (n == 3) { double a[3][3], v[3][3]; std::memcpy(a, a_in, sizeof(a)); jacobi_eig<3>(a, v); for (int i = 0; i < 3; ++i) { evals[i] = a[i][i]; for (int j = 0; j < 3; ++j) evecs[i * 3 + j] = v[i][j]; } }
    else        { double a[4][4], v[4][4]; std::memcpy(a, a_in, sizeof(a)); jacobi_eig<4>(a, v); for (int i = 0; i < 4; ++i) { evals[i] = a[i][i]; for (int j = 0; j < 4; ++j) evecs[i * 4 + j] = v[i][j]; } }
}

// The whole registration (rtr_register's CPU counterpart): normals -> Harris -> FPFH (both clouds) -> k-NN features
// -> prerejective RANSAC -> ICP.  Sequencing of main(), RealTimeRobot.cpp:39-105, with the north-star stages.
// The scan side (ScanPoint: keypoints + descriptors) is built ONCE per scan and reused for every database model, as the
// reference does (RealTimeRobot.cpp:45-60 before the loops at :62-102): orc_register_many is rtr_register_many's
// counterpart, orc_register the one-model case of it.
struct SceneStages { std::vector<float> normals4, fpfh; int n_keypoints = 0; };

static void scene_stages(const float* scene_xyz1, int nsc, const rtr_register_params* p, SceneStages& sc) {
    sc.normals4.assign((size_t)nsc * 4, 0.f); sc.fpfh.assign((size_t)nsc * 33, 0.f);
    orc_normals(scene_xyz1, nsc, p->normal_radius, 0, sc.normals4.data());
    int kps = orc_harris3d(scene_xyz1, nsc, sc.normals4.data(), p->harris_radius, p->harris_threshold, p->harris_nms, p->harris_refine, nullptr, nullptr, nullptr, 0);
    // refinement of the counted corners (capacity 0 above skips it); run it for real so the CPU baseline pays for it
    std::vector<int> ki(kps + 1); std::vector<float> kx((size_t)(kps + 1) * 4);
    orc_harris3d(scene_xyz1, nsc, sc.normals4.data(), p->harris_radius, p->harris_threshold, p->harris_nms, p->harris_refine, nullptr, ki.data(), kx.data(), kps);
    sc.n_keypoints = kps;
    orc_fpfh(scene_xyz1, nsc, sc.normals4.data(), p->fpfh_radius, sc.fpfh.data());
}

static void register_against(const float* model_xyz1, int nm, const float* scene_xyz1, int nsc, const SceneStages& sc,
                             const rtr_register_params* p, rtr_pose_result* res) {
    std::vector<float> nm4((size_t)nm * 4), fm((size_t)nm * 33);
    orc_normals(model_xyz1, nm, p->normal_radius, 0, nm4.data());
    int kpm = orc_harris3d(model_xyz1, nm, nm4.data(), p->harris_radius, p->harris_threshold, p->harris_nms, p->harris_refine, nullptr, nullptr, nullptr, 0);
    {
        std::vector<int> ki(kpm + 1); std::vector<float> kx((size_t)(kpm + 1) * 4);
        orc_harris3d(model_xyz1, nm, nm4.data(), p->harris_radius, p->harris_threshold, p->harris_nms, p->harris_refine, nullptr, ki.data(), kx.data(), kpm);
    }
    orc_fpfh(model_xyz1, nm, nm4.data(), p->fpfh_radius, fm.data());
    int k = p->ransac.correspondence_k;
    std::vector<int> knn((size_t)nm * k);
    orc_match_features(fm.data(), nm, sc.fpfh.data(), nsc, k, knn.data(), nullptr);
    rtr_pose_result r;
    orc_ransac_prerejective(model_xyz1, nm, scene_xyz1, nsc, knn.data(), k, &p->ransac, &r);
    if (p->run_icp && r.converged) {      // nothing accepted -> no pose to refine (result stays identity / FLT_MAX)
        rtr_pose_result ri;
        orc_icp_normals(model_xyz1, nm, scene_xyz1, nsc, sc.normals4.data(), &p->icp, r.pose, &ri);     // the scene's normals serve estimator 1
        std::memcpy(r.pose, ri.pose, sizeof(r.pose));
        r.fitness = ri.fitness; r.iterations = ri.iterations;
        r.converged = ri.converged;
    }
    r.n_keypoints_src = kpm; r.n_keypoints_tgt = sc.n_keypoints;
    *res = r;
}

void orc_register(const float* model_xyz1, int nm, const float* scene_xyz1, int nsc, const rtr_register_params* p,
                  rtr_pose_result* res) {
    SceneStages sc;
    scene_stages(scene_xyz1, nsc, p, sc);
    register_against(model_xyz1, nm, scene_xyz1, nsc, sc, p, res);
}

// n_models database models (model_xyz1[m], model_n[m]) against one scan; res[m].model_id = m
void orc_register_many(const float* const* model_xyz1, const int* model_n, int n_models, const float* scene_xyz1, int nsc,
                       const rtr_register_params* p, rtr_pose_result* res) {
    SceneStages sc;
    scene_stages(scene_xyz1, nsc, p, sc);
    for (int m = 0; m < n_models; ++m) {
        register_against(model_xyz1[m], model_n[m], scene_xyz1, nsc, sc, p, &res[m]);
        res[m].model_id = m;
    }
}

void orc_default_register_params(rtr_register_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->normal_radius = 0.05f; p->harris_radius = 0.05f; p->harris_threshold = 0.01f; p->harris_nms = 1; p->harris_refine = 1;
    p->fpfh_radius = 0.10f; p->run_icp = 1;
    p->ransac.max_iterations = 50000; p->ransac.hypothesis_begin = 0; p->ransac.hypothesis_end = 0; p->ransac.seed = 20170427ULL;
    p->ransac.correspondence_k = 5; p->ransac.similarity_threshold = 0.9f; p->ransac.max_correspondence_distance = 0.0365f;
    p->ransac.inlier_fraction = 0.25f;
    p->icp.max_iterations = 10; p->icp.force_iterations = 0; p->icp.max_correspondence_distance = 0.f; p->icp.mse_threshold_absolute = 1e-12;
}

}  // extern "C"
