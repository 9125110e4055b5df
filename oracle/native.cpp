// native.cpp — CPU restatement of the reference's own descriptor path (TEST INFRASTRUCTURE, see oracle.cpp header):
//   KeyPoint::getOccupiedGrid   key_point.h:112-161      points in the +-0.1 box, count of occupied 1 cm voxels
//   KeyPoint::get_TSDF          key_point.h:251-318      occupied voxels of that cloud in the +-0.15 box -> int triples -> TDF
//   get_Distance                matching.h:122-222       36 x 10 degree yaw sweep scored against the model keypoint's TDF
//   match_by_height/area/occupied  function.h:158-178    screens
//   pair loop of main()         RealTimeRobot.cpp:70-102
//   Ransac                      function.h:35-109        exhaustive consensus over the screened pairs
//
// PARITY STATUS: the TDF itself is pinned (kernel.cu, see oracle.cpp).  Everything else here goes through
// pcl::octree::OctreePointCloud*, whose source is not in /root/reference: the voxel frame (bounding-box enlargement, key
// computation, depth-first voxel order) follows SURVEY.md Appendix A.8 — **parity unpinned**.
//
// MODES.  The as-committed get_Distance has undefined behaviour (Appendix B#3: it voxelises the MODEL keypoint's own
// cloud translated away from the TDF box, so indices fall outside grid_value[]).  What is restated here is the INTENDED
// algorithm — the older signature kept in comments at matching.h:17-120 scores the SCAN keypoint's occupancy cloud — with
// each remaining quirk behind a flag of rtr_native_params (default off):
//   quirk_skip_first_voxel       B#5  first occupied voxel (lowest Morton key) skipped in TDF build and scoring
//   quirk_running_score          B#4  distance_temp not reset between the 36 angles
//   quirk_integer_screens        B#9/B#10  float(2/3) == 0 and integer division in the match_by_* screens
// Voxels whose index falls outside the 30^3 TDF (index 30 is reachable, B#8) are ignored instead of read out of bounds.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../include/rtr.h"

extern "C" void orc_tdf(const int* occ, int num_occ, int dim, float* out);
extern "C" int orc_plane_areas(const float* xyz1, int n, rtr_surface* out, int capacity);

namespace {

struct P4 { float x, y, z, w; };

// OctreePointCloud::defineBoundingBox + getKeyBitSize (App. A.8): the box is enlarged symmetrically to 2^depth voxels
struct OctreeFrame {
    double origin[3]; double res; int depth; unsigned nkeys;
    void define(const double mn[3], const double mx[3], double resolution) {
        res = resolution;
        const double minValue = (double)FLT_EPSILON;
        unsigned max_voxels = 2;
        for (int a = 0; a < 3; ++a) max_voxels = std::max(max_voxels, (unsigned)std::ceil((mx[a] - mn[a] - minValue) / res));
        depth = std::max((int)std::ceil(std::log2((double)max_voxels) - minValue), 0);
        nkeys = 1u << depth;
        double side = (double)nkeys * res;
        for (int a = 0; a < 3; ++a) origin[a] = mn[a] - (side - (mx[a] - mn[a])) / 2.0;
    }
    bool key(const P4& p, unsigned k[3]) const {          // genOctreeKeyforPoint; false if the point is outside the frame
        const double v[3] = {p.x, p.y, p.z};
        for (int a = 0; a < 3; ++a) {
            double f = (v[a] - origin[a]) / res;
            if (!(f >= 0.0) || f >= (double)nkeys) return false;
            k[a] = (unsigned)f;
        }
        return true;
    }
    float center(unsigned k, int a) const { return (float)(((double)k + 0.5) * res + origin[a]); }   // genLeafNodeCenterFromOctreeKey
};

// depth-first order of getOccupiedVoxelCenters: child index 4*xbit + 2*ybit + zbit at every level
inline unsigned morton(unsigned x, unsigned y, unsigned z, int depth) {
    unsigned c = 0;
    for (int b = depth - 1; b >= 0; --b) c = (c << 3) | (((x >> b) & 1u) << 2) | (((y >> b) & 1u) << 1) | ((z >> b) & 1u);
    return c;
}

struct Voxel { unsigned code, k[3]; };

// occupied voxels of a point set in a frame, in depth-first order
std::vector<Voxel> occupied(const OctreeFrame& f, const P4* pts, int n) {
    std::vector<Voxel> v;
    for (int i = 0; i < n; ++i) {
        Voxel e;
        if (!f.key(pts[i], e.k)) continue;
        e.code = morton(e.k[0], e.k[1], e.k[2], f.depth);
        v.push_back(e);
    }
    std::sort(v.begin(), v.end(), [](const Voxel& a, const Voxel& b) { return a.code < b.code; });
    v.erase(std::unique(v.begin(), v.end(), [](const Voxel& a, const Voxel& b) { return a.code == b.code; }), v.end());
    return v;
}

// the float box of a keypoint: Vector3f v_center -/+ f_adjust, widened to double (key_point.h:118-137)
void box_of(const P4& kp, float half, double mn[3], double mx[3]) {
    const float c[3] = {kp.x, kp.y, kp.z};
    for (int a = 0; a < 3; ++a) { mn[a] = (double)(float)(c[a] - half); mx[a] = (double)(float)(c[a] + half); }
}

void identity16(float m[16]) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
void matmul4(const float a[16], const float b[16], float c[16]) {
    float out[16];
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            float acc = a[row] * b[col * 4];
            for (int k = 1; k < 4; ++k) acc = acc + a[k * 4 + row] * b[col * 4 + k];
            out[col * 4 + row] = acc;
        }
    std::memcpy(c, out, sizeof(out));
}
inline P4 xform(const float m[16], const P4& p) {
    P4 o;
    o.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12];
    o.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
    o.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14];
    o.w = 1.0f;
    return o;
}

}  // namespace

extern "C" {

void orc_native_default_params(rtr_native_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->resolution = 0.01f; p->occ_half = 0.1f; p->tdf_half = 0.15f;
    p->pair_gate = 3.0f; p->consensus_distance = 0.15f; p->consensus_score = 100.0f;
    p->use_plane_areas = 1;
}

// KeyPoint::get_Vector3D (key_point.h:87-111) with getDistance (key_point.h:38-43): areas of <= 1 horizontal plane within
// 5 cm and <= 2 vertical planes within 2 cm of the keypoint, verticals in descending order; defaults 0.16 (key_point.h:83-84)
void orc_native_vector3d(const float* kp_xyz1, const rtr_surface* surfaces, int n_surfaces, double* vector3d) {
    vector3d[0] = vector3d[1] = vector3d[2] = 0.16;
    int vertical = 0, horizontal = 0;
    for (int i = 0; i < n_surfaces; ++i) {
        if (!surfaces[i].kept) continue;
        const float* v = surfaces[i].coefficients;
        double d = (double)std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
        double distance = (double)std::fabs(((v[0] * kp_xyz1[0] + v[1] * kp_xyz1[1]) + v[2] * kp_xyz1[2]) + v[3]) / d;
        if (surfaces[i].is_vertical == 0 && horizontal == 0 && distance <= 0.05) { vector3d[0] = surfaces[i].area; horizontal++; }
        else if (surfaces[i].is_vertical == 1 && vertical <= 1 && distance <= 0.02) { vector3d[1 + vertical] = surfaces[i].area; vertical++; }
    }
    if (vector3d[1] < vector3d[2]) std::swap(vector3d[1], vector3d[2]);
}

// KeyPoint::getOccupiedGrid: indices (ascending) of the points inside the inclusive float box, and Occupiedgrid.Number
int orc_native_occupancy(const float* xyz1, int n, const float* kp_xyz1, const rtr_native_params* p, int* idx_out, int cap, int* number) {
    const P4* pts = (const P4*)xyz1; const P4& kp = *(const P4*)kp_xyz1;
    const float lo[3] = {kp.x - p->occ_half, kp.y - p->occ_half, kp.z - p->occ_half};
    const float hi[3] = {kp.x + p->occ_half, kp.y + p->occ_half, kp.z + p->occ_half};
    std::vector<P4> in;
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const P4& q = pts[i];
        if (q.x >= lo[0] && q.x <= hi[0] && q.y >= lo[1] && q.y <= hi[1] && q.z >= lo[2] && q.z <= hi[2]) {   // boxSearch: inclusive
            if (idx_out && m < cap) idx_out[m] = i;
            in.push_back(q); ++m;
        }
    }
    double mn[3], mx[3];
    box_of(kp, p->occ_half, mn, mx);
    OctreeFrame f; f.define(mn, mx, (double)p->resolution);
    if (number) *number = (int)occupied(f, in.data(), (int)in.size()).size();
    return m;
}

// KeyPoint::get_TSDF up to the FFI call: the int triples handed to ComputeTDFWithCuda (key_point.h:287-308)
int orc_native_tdf_voxels(const float* occ_xyz1, int m, const float* kp_xyz1, const rtr_native_params* p, int* triples, int cap) {
    const P4& kp = *(const P4*)kp_xyz1;
    double mn[3], mx[3];
    box_of(kp, p->tdf_half, mn, mx);
    OctreeFrame f; f.define(mn, mx, (double)p->resolution);
    std::vector<Voxel> v = occupied(f, (const P4*)occ_xyz1, m);
    int out = 0;
    for (size_t i = p->quirk_skip_first_voxel ? 1 : 0; i < v.size(); ++i) {
        if (out < cap) for (int a = 0; a < 3; ++a)
            triples[out * 3 + a] = (int)(((double)f.center(v[i].k[a], a) - mn[a]) / (double)p->resolution);    // truncating cast
        ++out;
    }
    return out;
}

// get_Distance, intended form: score of the scan keypoint's occupancy cloud against the model keypoint's TDF over the
// 36-step yaw sweep; returns the minimum score, the step that produced it and the composed transform (matching.h:204-217).
float orc_native_pair_score(const float* model_kp_xyz1, const float* model_tdf, const float* scan_occ_xyz1, int m, const float* scan_kp_xyz1,
                            const rtr_native_params* p, int* best_step, float* transform16) {
    const P4& k1 = *(const P4*)model_kp_xyz1; const P4& k2 = *(const P4*)scan_kp_xyz1;
    float T[16]; identity16(T);
    T[12] = k1.x - k2.x; T[13] = k1.y - k2.y; T[14] = k1.z - k2.z;
    std::vector<P4> cloud(m);
    for (int i = 0; i < m; ++i) cloud[i] = xform(T, ((const P4*)scan_occ_xyz1)[i]);
    float theta = (float)(M_PI / 18);
    float T1[16], R[16], T3[16], step[16];
    identity16(T1); T1[12] = -k1.x; T1[13] = -k1.y;
    identity16(R); R[0] = (float)std::cos(theta); R[1] = (float)std::sin(theta); R[4] = -(float)std::sin(theta); R[5] = (float)std::cos(theta);
    identity16(T3); T3[12] = k1.x; T3[13] = k1.y;
    matmul4(T3, R, step); matmul4(step, T1, step);                       // transform_3 * transform_2 * transform_1
    double mn[3], mx[3];
    box_of(k1, p->tdf_half, mn, mx);                                     // p1_key.Border (float) widened to double
    OctreeFrame f; f.define(mn, mx, (double)p->resolution);
    const int dim = 30;
    float distance_total = 100000000.f, distance_temp = 0.f;
    int best = 0;
    for (int i = 0; i < 36; ++i) {
        if (i != 0) for (auto& q : cloud) q = xform(step, q);            // cumulative, in place, float (Appendix B#13)
        std::vector<Voxel> v = occupied(f, cloud.data(), m);
        if (!p->quirk_running_score) distance_temp = 0.f;
        int used = 0;
        for (size_t s = p->quirk_skip_first_voxel ? 1 : 0; s < v.size(); ++s) {
            int c[3];
            for (int a = 0; a < 3; ++a) c[a] = (int)(((double)f.center(v[s].k[a], a) - mn[a]) / (double)p->resolution);
            if (c[0] < 0 || c[1] < 0 || c[2] < 0 || c[0] >= dim || c[1] >= dim || c[2] >= dim) continue;
            float g = model_tdf[c[1] * dim + c[2] * dim * dim + c[0]];
            distance_temp += g * g;                                        // float accumulation in depth-first voxel order
            ++used;
        }
        distance_temp = used > 0 ? distance_temp / (float)used : 100000000.f;
        if (distance_temp < distance_total) { distance_total = distance_temp; best = i; }
    }
    if (best_step) *best_step = best;
    if (transform16) {
        double bt = (double)best * (double)theta;
        float R2[16], S[16], out[16];
        identity16(R2); R2[0] = (float)std::cos(bt); R2[1] = (float)std::sin(bt); R2[4] = -(float)std::sin(bt); R2[5] = (float)std::cos(bt);
        identity16(S); float cs = k1.z / k2.z; S[0] = S[5] = S[10] = cs;
        matmul4(S, T3, out); matmul4(out, R2, out); matmul4(out, T1, out); matmul4(out, T, out);   // S * T3 * R * T1 * T
        std::memcpy(transform16, out, sizeof(out));
    }
    return distance_total;
}

// function.h:158-178.  vector3D = the three plane areas of a keypoint (default 0.16 each, key_point.h:83-84).
int orc_native_screens(const float* kp1, const float* kp2, const double* areas1, const double* areas2, int number1, int number2,
                       const rtr_native_params* p) {
    bool height, area = true, occ;
    float hr = kp1[2] / kp2[2];
    if (p->quirk_integer_screens) height = (hr >= 0.0f || hr <= 1.5f);                 // float(2/3) == 0 (B#9)
    else height = (hr >= 2.0f / 3.0f && hr <= 1.5f);
    for (int a = 0; a < 3; ++a) {
        double r = areas1[a] / areas2[a];
        double lo = p->quirk_integer_screens ? 0.0 : 1.0 / 3.0;
        if (r > 3 || r < lo) area = false;
    }
    if (number2 == 0) occ = false;
    else if (p->quirk_integer_screens) { float t = (float)(number1 / number2); occ = !(t > 2 || t < 0.5); }   // integer division (B#10)
    else { float t = (float)number1 / (float)number2; occ = !(t > 2 || t < 0.5); }
    return (height && area && occ) ? 1 : 0;
}

// The reference-native registration: keypoints are given (Harris corners of both clouds); returns the winning pair's
// transform (scan -> model frame, as main() applies it to the scan cloud, RealTimeRobot.cpp:104-105).
// result->inliers = consensus size, result->hypothesis = index of the winning pair in screening order (-1: none),
// result->evaluated = number of screened pairs, result->fitness = the winning pair's sweep score.
void orc_native_register(const float* model_xyz1, int nm, const float* model_kp, int km, const float* scan_xyz1, int ns,
                         const float* scan_kp, int ks, const rtr_native_params* p, rtr_pose_result* res) {
    std::memset(res, 0, sizeof(*res));
    identity16(res->pose);
    res->fitness = FLT_MAX; res->hypothesis = -1;
    std::vector<double> marea((size_t)km * 3, 0.16), sarea((size_t)ks * 3, 0.16);
    if (p->use_plane_areas) {                                              // modelpoint.getArea(mcloud), scanpoint.get_Area(cloud)
        std::vector<rtr_surface> ms(256), ss(256);
        int nms = std::min(orc_plane_areas(model_xyz1, nm, ms.data(), 256), 256), nss = std::min(orc_plane_areas(scan_xyz1, ns, ss.data(), 256), 256);
        for (int k = 0; k < km; ++k) orc_native_vector3d(model_kp + 4 * k, ms.data(), nms, &marea[3 * k]);
        for (int s = 0; s < ks; ++s) orc_native_vector3d(scan_kp + 4 * s, ss.data(), nss, &sarea[3 * s]);
    }
    std::vector<std::vector<P4>> mocc(km), socc(ks);
    std::vector<int> mnum(km), snum(ks);
    std::vector<std::vector<float>> tdf(km, std::vector<float>(27000, 0.f));
    auto gather = [&](const float* xyz1, int n, const float* kp, std::vector<P4>& out, int& number) {
        std::vector<int> idx(n);
        int m = orc_native_occupancy(xyz1, n, kp, p, idx.data(), n, &number);
        out.resize(m);
        for (int i = 0; i < m; ++i) out[i] = ((const P4*)xyz1)[idx[i]];
    };
    for (int s = 0; s < ks; ++s) gather(scan_xyz1, ns, scan_kp + 4 * s, socc[s], snum[s]);
    for (int k = 0; k < km; ++k) {
        gather(model_xyz1, nm, model_kp + 4 * k, mocc[k], mnum[k]);
        std::vector<int> tri(3 * 32768);
        int nt = orc_native_tdf_voxels((const float*)mocc[k].data(), (int)mocc[k].size(), model_kp + 4 * k, p, tri.data(), 32768);
        orc_tdf(tri.data(), nt, 30, tdf[k].data());
    }
    struct Pair { int k, s; float score; float T[16]; };
    std::vector<Pair> pairs;
    for (int k = 0; k < km; ++k)
        for (int s = 0; s < ks; ++s) {
            Pair pr; pr.k = k; pr.s = s;
            pr.score = orc_native_pair_score(model_kp + 4 * k, tdf[k].data(), (const float*)socc[s].data(), (int)socc[s].size(), scan_kp + 4 * s, p, nullptr, pr.T);
            bool gate = pr.score < p->pair_gate;
            if (gate && orc_native_screens(model_kp + 4 * k, scan_kp + 4 * s, &marea[3 * k], &sarea[3 * s], mnum[k], snum[s], p)) pairs.push_back(pr);
        }
    res->evaluated = (long long)pairs.size();
    int best_in = 0;
    for (size_t i = 0; i < pairs.size(); ++i) {
        int in = 0;
        for (size_t j = 0; j < pairs.size(); ++j) {
            P4 a = *(const P4*)(model_kp + 4 * pairs[j].k);
            P4 b = xform(pairs[i].T, *(const P4*)(scan_kp + 4 * pairs[j].s));
            // pointdistance (function.h:27-30): float differences, float products and sum, float sqrt, widened to double
            float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
            float sumf = (dx * dx + dy * dy) + dz * dz;
            double dis = (double)std::sqrt(sumf);
            if (dis < (double)p->consensus_distance && pairs[j].score < p->consensus_score) ++in;
        }
        if (in > best_in) {                                               // first arg-max wins (function.h:83-87)
            best_in = in;
            std::memcpy(res->pose, pairs[i].T, sizeof(res->pose));
            res->inliers = in; res->hypothesis = (long long)i; res->fitness = pairs[i].score; res->converged = 1;
        }
    }
    res->n_keypoints_src = km; res->n_keypoints_tgt = ks;
}

}  // extern "C"
