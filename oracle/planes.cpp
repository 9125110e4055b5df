// planes.cpp — CPU restatement of ModelPoint::getArea / ScanPoint::get_Area (model_point.h:170-245, scan_point.h:117-188):
// peel planes off the cloud until <= 15 % of the points remain; every plane is a pcl::SACSegmentation fit
// (SACMODEL_PLANE, SAC_RANSAC, 150 iterations, 5 mm, optimised coefficients) whose inliers go through
// pcl::ConvexHull::getTotalArea; planes within 10 degrees of horizontal / vertical with area >= 0.16 are kept.
// TEST INFRASTRUCTURE (see oracle.cpp header).  PCL 1.8.0 internals follow SURVEY.md Appendix A.7 — **parity unpinned**:
//   * SampleConsensusModel(random = false): boost::mt19937 seeded 12345, rnd() = mt() >> 1 (uniform_int<>(0, INT_MAX) over a
//     32-bit engine), partial Fisher-Yates on a persistent shuffled index array; a new model (same seed) per segment() call;
//   * RandomSampleConsensus::computeModel: adaptive k = log(1 - 0.99) / log(1 - w^3), at most 151 iterations;
//   * plane through three points in float, distance |n.p + d| < 0.005;
//   * optimizeModelCoefficients: least-squares plane of the inliers (here: fp64 covariance + Jacobi, the "exact mode" of this
//     oracle, PCL uses single-pass fp32 + eigen33), inliers re-selected with the refined plane;
//   * ConvexHull: dimension 2 iff |l0| < eps or |l0 / l2| < 1e-3 (fp64 covariance eigenvalues); 2-D hulls are taken in the
//     coordinate plane chosen from the normal of (first, last, middle) inlier with the 10-degree rule and return the
//     PROJECTED area (Qhull's 2-D "volume"); 3-D hulls return the total facet area.  The 3-D case is restated as twice the
//     hull area in the best-fit plane (it only occurs for patches far below the 0.16 gate).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include "../include/rtr.h"

namespace {
struct P4 { float x, y, z, w; };

template <int N>
void jacobi(double a[N][N], double v[N][N]) {   // same routine as oracle.cpp (kept private to this file)
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
                if (sweep > 3 && std::fabs(a[p][p]) + g == std::fabs(a[p][p]) && std::fabs(a[q][q]) + g == std::fabs(a[q][q])) { a[p][q] = 0.0; a[q][p] = 0.0; continue; }
                double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                double t = 1.0 / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                if (theta < 0) t = -t;
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                a[p][p] = a[p][p] - t * apq; a[q][q] = a[q][q] + t * apq; a[p][q] = 0.0; a[q][p] = 0.0;
                for (int r = 0; r < N; ++r) {
                    if (r != p && r != q) { double arp = a[r][p], arq = a[r][q]; a[r][p] = c * arp - s * arq; a[p][r] = a[r][p]; a[r][q] = s * arp + c * arq; a[q][r] = a[r][q]; }
                    double vrp = v[r][p], vrq = v[r][q]; v[r][p] = c * vrp - s * vrq; v[r][q] = s * vrp + c * vrq;
                }
            }
    }
}

inline float plane_dist(const float c[4], const P4& p) { return std::fabs(((c[0] * p.x + c[1] * p.y) + c[2] * p.z) + c[3]); }

// SampleConsensusModelPlane::computeModelCoefficients (float)
bool plane_from_3(const P4& p0, const P4& p1, const P4& p2, float c[4]) {
    float a[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, b[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
    float r0 = a[0] / b[0], r1 = a[1] / b[1], r2 = a[2] / b[2];
    if (r0 == r1 && r2 == r1) return false;                               // collinear
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
    float nrm = std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
    c[0] /= nrm; c[1] /= nrm; c[2] /= nrm;
    c[3] = -1.0f * ((c[0] * p0.x + c[1] * p0.y) + c[2] * p0.z);
    return true;
}

// covariance (fp64) of a point subset: mean, 3x3; eigen-decomposition sorted ascending
void cov_eig(const std::vector<P4>& pts, const std::vector<int>& idx, double mean[3], double evals[3], double evecs[3][3]) {
    double s[3] = {0, 0, 0}, m[6] = {0, 0, 0, 0, 0, 0};
    for (int i : idx) {
        double x = pts[i].x, y = pts[i].y, z = pts[i].z;
        s[0] += x; s[1] += y; s[2] += z;
        m[0] += x * x; m[1] += x * y; m[2] += x * z; m[3] += y * y; m[4] += y * z; m[5] += z * z;
    }
    double k = (double)idx.size();
    for (int a = 0; a < 3; ++a) mean[a] = s[a] / k;
    double A[3][3], V[3][3];
    A[0][0] = m[0] / k - mean[0] * mean[0]; A[0][1] = m[1] / k - mean[0] * mean[1]; A[0][2] = m[2] / k - mean[0] * mean[2];
    A[1][1] = m[3] / k - mean[1] * mean[1]; A[1][2] = m[4] / k - mean[1] * mean[2]; A[2][2] = m[5] / k - mean[2] * mean[2];
    A[1][0] = A[0][1]; A[2][0] = A[0][2]; A[2][1] = A[1][2];
    jacobi<3>(A, V);
    int o[3] = {0, 1, 2};
    std::sort(o, o + 3, [&](int x, int y) { return A[x][x] < A[y][y] || (A[x][x] == A[y][y] && x < y); });
    for (int e = 0; e < 3; ++e) { evals[e] = A[o[e]][o[e]]; for (int r = 0; r < 3; ++r) evecs[e][r] = V[r][o[e]]; }
}

// area of the convex hull of 2-D points (monotone chain + shoelace, fp64) == Qhull's 2-D "volume"
double hull_area_2d(std::vector<std::pair<double, double>> p) {
    std::sort(p.begin(), p.end());
    p.erase(std::unique(p.begin(), p.end()), p.end());
    int n = (int)p.size();
    if (n < 3) return 0.0;
    std::vector<std::pair<double, double>> h(2 * n);
    auto cross = [](const std::pair<double, double>& o, const std::pair<double, double>& a, const std::pair<double, double>& b) {
        return (a.first - o.first) * (b.second - o.second) - (a.second - o.second) * (b.first - o.first);
    };
    int k = 0;
    for (int i = 0; i < n; ++i) { while (k >= 2 && cross(h[k - 2], h[k - 1], p[i]) <= 0) --k; h[k++] = p[i]; }
    for (int i = n - 2, t = k + 1; i >= 0; --i) { while (k >= t && cross(h[k - 2], h[k - 1], p[i]) <= 0) --k; h[k++] = p[i]; }
    double a = 0;
    for (int i = 0; i + 1 < k; ++i) a += h[i].first * h[i + 1].second - h[i + 1].first * h[i].second;
    return 0.5 * std::fabs(a);
}
}  // namespace

extern "C" {

// pcl::ConvexHull<PointXYZ>::getTotalArea of a point set (model_point.h:211-219); dimension receives 2 or 3
double orc_hull_area(const float* xyz1, int n, int* dimension) {
    const P4* q = (const P4*)xyz1;
    std::vector<P4> pts(q, q + n);
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    if (n < 3) { if (dimension) *dimension = 2; return 0.0; }
    double mean[3], ev[3], evec[3][3];
    cov_eig(pts, idx, mean, ev, evec);
    int dim = (std::fabs(ev[0]) < DBL_EPSILON || std::fabs(ev[0] / ev[2]) < 1.0e-3) ? 2 : 3;
    if (dimension) *dimension = dim;
    std::vector<std::pair<double, double>> p2(n);
    if (dim == 2) {
        // projection plane from the normal of (first, last, middle) point, 10-degree rule (convex_hull.hpp, performReconstruction2D)
        const P4 &p0 = pts[0], &p1 = pts[n - 1], &pm = pts[n / 2];
        double a[3] = {(double)p1.x - p0.x, (double)p1.y - p0.y, (double)p1.z - p0.z}, b[3] = {(double)pm.x - p0.x, (double)pm.y - p0.y, (double)pm.z - p0.z};
        double nx = a[1] * b[2] - a[2] * b[1], ny = a[2] * b[0] - a[0] * b[2], nz = a[0] * b[1] - a[1] * b[0];
        double nn = std::sqrt((nx * nx + ny * ny) + nz * nz);
        if (nn == 0) { nx = evec[0][0]; ny = evec[0][1]; nz = evec[0][2]; nn = 1; }
        double tx = std::fabs(nx / nn), ty = std::fabs(ny / nn), tz = std::fabs(nz / nn);
        const double thresh = std::cos(0.174532925);
        bool xy = true, yz = true, xz = true;
        if (tz > thresh) { xz = false; yz = false; }
        if (tx > thresh) { xz = false; xy = false; }
        if (ty > thresh) { xy = false; yz = false; }
        int u = 0, v = 1;
        if (xy) { u = 0; v = 1; } else if (yz) { u = 1; v = 2; } else if (xz) { u = 0; v = 2; }
        for (int i = 0; i < n; ++i) { const float c[3] = {pts[i].x, pts[i].y, pts[i].z}; p2[i] = {(double)c[u], (double)c[v]}; }
        return hull_area_2d(p2);
    }
    // 3-D: twice the hull area in the best-fit plane (basis = the two larger eigenvectors)
    for (int i = 0; i < n; ++i) {
        double d[3] = {pts[i].x - mean[0], pts[i].y - mean[1], pts[i].z - mean[2]};
        p2[i] = {(double)(float)((d[0] * evec[2][0] + d[1] * evec[2][1]) + d[2] * evec[2][2]),      // stored as float like the 2-D case
                 (double)(float)((d[0] * evec[1][0] + d[1] * evec[1][1]) + d[2] * evec[1][2])};
    }
    return 2.0 * hull_area_2d(p2);
}

// one SACSegmentation::segment on a cloud: returns the refined inlier count, fills coeff[4] (refined) and the inlier
// indices (ascending); iterations_used receives the number of RANSAC iterations run
int orc_plane_segment(const float* xyz1, int n, float threshold, int max_iterations, float* coeff4, int* inliers_out, int* iterations_used) {
    const P4* pts = (const P4*)xyz1;
    if (iterations_used) *iterations_used = 0;
    if (n < 3) return 0;
    std::mt19937 gen(12345u);
    std::vector<int> shuffled(n);
    for (int i = 0; i < n; ++i) shuffled[i] = i;
    auto rnd = [&]() -> unsigned { return (unsigned)(gen() >> 1); };
    int best = -1; float best_c[4] = {0, 0, 0, 0};
    double k = 1.0;
    const double log_probability = std::log(1.0 - 0.99), one_over = 1.0 / (double)n;
    int iterations = 0; unsigned skipped = 0; const unsigned max_skip = (unsigned)max_iterations * 10;
    while (iterations < k && skipped < max_skip) {
        for (int i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + (rnd() % (unsigned)(n - i))]);      // drawIndexSample
        float c[4];
        if (!plane_from_3(pts[shuffled[0]], pts[shuffled[1]], pts[shuffled[2]], c)) { ++skipped; continue; }
        int cnt = 0;
        for (int i = 0; i < n; ++i) if ((double)plane_dist(c, pts[i]) < (double)threshold) ++cnt;
        if (cnt > best) {
            best = cnt; std::memcpy(best_c, c, sizeof(c));
            double w = (double)best * one_over;
            double p_no = 1.0 - std::pow(w, 3.0);
            p_no = std::max(std::numeric_limits<double>::epsilon(), p_no);
            p_no = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no);
            k = log_probability / std::log(p_no);
        }
        ++iterations;
        if (iterations > max_iterations) break;
    }
    if (iterations_used) *iterations_used = iterations;
    if (best < 0) return 0;
    std::vector<P4> all(pts, pts + n);
    std::vector<int> in;
    for (int i = 0; i < n; ++i) if ((double)plane_dist(best_c, pts[i]) < (double)threshold) in.push_back(i);
    float refined[4];
    std::memcpy(refined, best_c, sizeof(refined));
    if (in.size() > 3) {                                                     // optimizeModelCoefficients
        double mean[3], ev[3], evec[3][3];
        cov_eig(all, in, mean, ev, evec);
        refined[0] = (float)evec[0][0]; refined[1] = (float)evec[0][1]; refined[2] = (float)evec[0][2];
        refined[3] = (float)(-1.0 * ((evec[0][0] * mean[0] + evec[0][1] * mean[1]) + evec[0][2] * mean[2]));
    }
    int m = 0;
    for (int i = 0; i < n; ++i) if ((double)plane_dist(refined, pts[i]) < (double)threshold) { if (inliers_out) inliers_out[m] = i; ++m; }
    std::memcpy(coeff4, refined, sizeof(refined));
    return m;
}

// is_h_plane / is_v_plane (model_point.h:48-79; PI is the macro 3.1415926): 1 horizontal, 2 vertical, 0 neither
int orc_plane_class(const float* c) {
    double b = ((double)c[0] * c[0]) + ((double)c[1] * c[1]) + ((double)c[2] * c[2]);
    double angle = std::acos((double)c[2] / std::sqrt(b));
    if ((angle > 2.9670597 && angle < 3.1415926) || (angle > 0 && angle < 0.1745329)) return 1;
    if (angle > 1.3962634 && angle < 1.7453292) return 2;
    return 0;
}

// ModelPoint::getArea: every peeled plane in order (rtr_surface records); returns the number of planes peeled
int orc_plane_areas(const float* xyz1, int n, rtr_surface* out, int capacity) {
    std::vector<P4> cur((const P4*)xyz1, (const P4*)xyz1 + n);
    int planes = 0;
    while ((double)cur.size() > 0.15 * (double)n) {
        std::vector<int> in(cur.size());
        float c[4]; int its = 0;
        int m = orc_plane_segment((const float*)cur.data(), (int)cur.size(), 0.005f, 150, c, in.data(), &its);
        if (m == 0) break;                                                    // "Could not estimate a planar model" (model_point.h:198-202)
        in.resize(m);
        std::vector<P4> plane(m);
        for (int i = 0; i < m; ++i) plane[i] = cur[in[i]];
        int dim = 2;
        double area = orc_hull_area((const float*)plane.data(), m, &dim);
        int cls = orc_plane_class(c);
        if (planes < capacity) {
            rtr_surface& s = out[planes];
            s.area = area; std::memcpy(s.coefficients, c, sizeof(c));
            s.is_vertical = (cls == 2) ? 1 : 0; s.inliers = m; s.dimension = dim; s.iterations = its;
            s.kept = (cls != 0 && area >= 0.16) ? 1 : 0;                     // model_point.h:223-233
        }
        ++planes;
        std::vector<char> mark(cur.size(), 0);
        for (int i : in) mark[i] = 1;
        std::vector<P4> rest; rest.reserve(cur.size() - m);
        for (size_t i = 0; i < cur.size(); ++i) if (!mark[i]) rest.push_back(cur[i]);     // extract.setNegative(true)
        cur.swap(rest);
    }
    return planes;
}

}  // extern "C"
