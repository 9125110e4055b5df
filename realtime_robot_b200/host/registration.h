// registration.h — C++ host mirror of the reference's registration surface, over pcl_compat.h and the C ABI (rtr.h).
//
// Same names, argument meaning and error behaviour as the reference headers for the path this library accelerates:
//   ModelPoint / ScanPoint  (model_point.h:81-96, scan_point.h:43-54): Ptr cloud member, key_coordinates by value,
//                           getKeypoint() mutates members and returns void, prints the corner count (model_point.h:154)
//   keyPointICP             (function.h:111-123): pcl::IterativeClosestPoint defaults, source = mcloud, target = cloud
//   ComputeTDFWithCuda      (key_point.h:35-36): declared in rtr.h, exported unchanged by librtr.so
// plus registerModelToScene(), the north-star pipeline (normals, Harris, FPFH, matching, prerejective RANSAC, ICP) that
// main() (RealTimeRobot.cpp:39-105) sequences.  All arithmetic happens in librtr.so; nothing here computes on the CPU.
// Nothing throws; failures print one line to stderr and leave outputs empty, like key_point.h:315-318.
#pragma once
#include <iostream>
#include "../../include/rtr.h"
#include "pcl_compat.h"

namespace rtr_host {

// process-wide context for device 0, created on first use (the reference hard-wires device 0: kernel.cu:43)
inline rtr_context* default_context() {
    static rtr_context* ctx = nullptr;
    if (!ctx && rtr_context_create(0, &ctx) != 0) { fprintf(stderr, "[rtr_host] no CUDA device: this library has no CPU fallback\n"); ctx = nullptr; }
    return ctx;
}

// Process-wide parameters of the reference-native path behind the reference's own signatures (which carry none): the
// literals of key_point.h / matching.h / function.h / RealTimeRobot.cpp:83 by default; a caller may edit them once, e.g.
// the pair gate or the quirk_* flags.
inline rtr_native_params& native_params() {
    static rtr_native_params p = []() { rtr_native_params q; rtr_native_default_params(&q); return q; }();
    return p;
}

struct DeviceCloud {   // RAII handle around rtr_cloud
    rtr_cloud* h = nullptr;
    explicit DeviceCloud(const pcl::PointCloud<pcl::PointXYZ>& c) {
        rtr_context* ctx = default_context();
        if (ctx) rtr_cloud_upload(ctx, c.points.empty() ? nullptr : &c.points[0].x, (int)c.points.size(), &h);
    }
    ~DeviceCloud() { if (h) rtr_cloud_free(h); }
    DeviceCloud(const DeviceCloud&) = delete;
    DeviceCloud& operator=(const DeviceCloud&) = delete;
};

// shared body of ModelPoint::getKeypoint / ScanPoint::getKeypoint (model_point.h:127-153, scan_point.h:85-112)
inline void harris_keypoints(const pcl::PointCloud<pcl::PointXYZ>& cloud, pcl::PointCloud<pcl::PointXYZ>& key_coordinates) {
    key_coordinates.clear();
    DeviceCloud d(cloud);
    if (!d.h) return;
    const float radius = 0.05f, threshold = 0.01f;           // setRadius(0.05f), setThreshold(0.01f)
    if (rtr_normals(d.h, radius, nullptr) != 0) return;
    int n = (int)cloud.size(), m = 0;
    std::vector<int> idx(n > 0 ? n : 1);
    std::vector<pcl::PointXYZ> xyz(n > 0 ? n : 1);
    if (rtr_harris3d(d.h, radius, threshold, /*nms=*/1, /*refine=*/1, nullptr, idx.data(), &xyz[0].x, n, &m) != 0) return;
    for (int i = 0; i < m; ++i) key_coordinates.push_back(pcl::PointXYZ(xyz[i].x, xyz[i].y, xyz[i].z));
}

}  // namespace rtr_host

struct Surface {                         // key_point.h:47-51 (Coefficients: a, b, c, d of the plane)
    double Area = 0;
    float Coefficients[4] = {0, 0, 0, 0};
    bool IsVertical = false;
};

namespace rtr_host {
// shared body of ModelPoint::getArea / ScanPoint::get_Area (model_point.h:170-245, scan_point.h:117-188)
inline void plane_surfaces(const pcl::PointCloud<pcl::PointXYZ>& cloud, std::vector<Surface>& surface) {
    DeviceCloud d(cloud);
    if (!d.h) return;
    std::vector<rtr_surface> all(256);
    int n = 0;
    int rc = rtr_plane_areas(d.h, all.data(), (int)all.size(), &n);
    if (rc != 0 && rc != RTR_ERR_CAPACITY) { std::cerr << "Could not estimate a planar model for the given dataset." << std::endl; return; }
    for (int i = 0; i < n && i < (int)all.size(); ++i)
        if (all[i].kept) {
            Surface s;
            s.Area = all[i].area; memcpy(s.Coefficients, all[i].coefficients, sizeof(s.Coefficients)); s.IsVertical = all[i].is_vertical != 0;
            surface.push_back(s);
        }
}
}  // namespace rtr_host

class ModelPoint {
public:
    pcl::PointCloud<pcl::PointXYZ>::Ptr Mpoint;
    std::vector<Surface> surface;
    pcl::PointCloud<pcl::PointXYZ> key_coordinates;
    void getArea(pcl::PointCloud<pcl::PointXYZ>::Ptr modelPoint) { if (modelPoint) rtr_host::plane_surfaces(*modelPoint, surface); }
    // Appendix B#1: the as-committed getKeypoint() scales the aliased model cloud by 0.01 IN PLACE before Harris runs
    // (model_point.h:105-111), which leaves no corners.  Default off; set to reproduce the shipped behaviour.
    bool quirk_scale_model_in_place = false;
    ModelPoint() {}
    explicit ModelPoint(pcl::PointCloud<pcl::PointXYZ>::Ptr m) : Mpoint(m) {}
    void getKeypoint() {
        if (!Mpoint) return;
        if (quirk_scale_model_in_place) {
            Eigen::Matrix4f t = Eigen::Matrix4f::Identity();
            t(0, 0) = t(1, 1) = t(2, 2) = 0.01f;
            pcl::transformPointCloud(*Mpoint, *Mpoint, t);
        }
        pcl::PointXYZ mn, mx;
        pcl::getMinMax3D(*Mpoint, mn, mx);
        std::cout << "L: " << mx.x - mn.x << " m   W: " << mx.y - mn.y << " m   H: " << mx.z - mn.z << " m" << std::endl;   // model_point.h:122-124
        rtr_host::harris_keypoints(*Mpoint, key_coordinates);
        std::cout << key_coordinates.size() << std::endl;                                                                 // model_point.h:154
    }
};

class ScanPoint {
public:
    pcl::PointCloud<pcl::PointXYZ>::Ptr Spoint;
    std::vector<Surface> surface;
    pcl::PointCloud<pcl::PointXYZ> key_coordinates;
    void get_Area(pcl::PointCloud<pcl::PointXYZ>::Ptr scanPoint) { if (scanPoint) rtr_host::plane_surfaces(*scanPoint, surface); }
    ScanPoint() {}
    explicit ScanPoint(pcl::PointCloud<pcl::PointXYZ>::Ptr s) : Spoint(s) {}
    void getKeypoint() {
        if (!Spoint) return;
        rtr_host::harris_keypoints(*Spoint, key_coordinates);
        std::cout << key_coordinates.size() << std::endl;                                                                 // scan_point.h:112
    }
};

// function.h:111-123.  The first two arguments are unused there too.  Prints the final transformation (function.h:119)
// and returns it together with the transformed model cloud through the optional outputs instead of writing fixed file
// names and spinning forever (function.h:126-147).
inline void keyPointICP(pcl::PointCloud<pcl::PointXYZ>::Ptr /*SpointCloud*/, pcl::PointCloud<pcl::PointXYZ>::Ptr /*mPointCloud*/,
                        pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, pcl::PointCloud<pcl::PointXYZ>::Ptr mcloud,
                        Eigen::Matrix4f* final_transformation = nullptr, pcl::PointCloud<pcl::PointXYZ>* transformed = nullptr) {
    rtr_host::DeviceCloud src(*mcloud), tgt(*cloud);            // setInputCloud(mcloud), setInputTarget(cloud)
    if (!src.h || !tgt.h) return;
    rtr_register_params p;
    rtr_default_register_params(&p);
    rtr_pose_result r;
    if (rtr_icp(src.h, tgt.h, &p.icp, nullptr, &r) != 0) { fprintf(stderr, "keyPointICP failed!"); return; }
    Eigen::Matrix4f t;
    memcpy(t.data(), r.pose, sizeof(r.pose));
    std::cout << t << std::endl;
    if (final_transformation) *final_transformation = t;
    if (transformed) pcl::transformPointCloud(*mcloud, *transformed, t);
}

// ---------------------------------------------------------------------------------------------------------------------
// The reference's own descriptor records and pair logic (key_point.h, matching.h, function.h), same names and members.
// Arithmetic runs in librtr.so (csrc/native.cu); the semantics are the INTENDED ones (matching.h:17-120) with the
// as-committed quirks behind rtr_native_params flags.

struct OccupiedGrid {                    // key_point.h:53-57
    pcl::PointCloud<pcl::PointXYZ>::Ptr cloud;
    float Border[6] = {0, 0, 0, 0, 0, 0};
    int Number = 0;
};

inline double getDistance(float v1, float v2, float v3, float v4, pcl::PointXYZ point) {     // key_point.h:38-43 (double abs: Appendix B#14)
    double d = std::sqrt((double)(v1 * v1 + v2 * v2 + v3 * v3));
    return std::fabs((double)(v1 * point.x + v2 * point.y + v3 * point.z + v4)) / d;
}

class KeyPoint {                         // key_point.h:59-76
public:
    OccupiedGrid Occupiedgrid;
    pcl::PointXYZ Key_coordinate;
    std::vector<double> vector3D;
    std::vector<float> grid_value;       // 27000 floats (a std::vector instead of float[27000]: KeyPoint is copied by value a lot)
    float Border[6] = {0, 0, 0, 0, 0, 0};
    const rtr_cloud* source_ = nullptr;  // device cloud the descriptors were taken from (set by getOccupiedGrid)

    KeyPoint() : vector3D(3, 0.16), grid_value(RTR_TDF_VOXELS, 0.f) {}
    explicit KeyPoint(pcl::PointXYZ point) : Key_coordinate(point), vector3D(3, 0.16), grid_value(RTR_TDF_VOXELS, 0.f) {
        Occupiedgrid.cloud.reset(new pcl::PointCloud<pcl::PointXYZ>);
    }

    // key_point.h:87-111: areas of <= 1 horizontal plane within 5 cm and <= 2 vertical planes within 2 cm, verticals descending
    void get_Vector3D(std::vector<Surface>& surface) {
        int vertical = 0, horizontal = 0;
        for (size_t i = 0; i < surface.size(); i++) {
            double distance = getDistance(surface[i].Coefficients[0], surface[i].Coefficients[1], surface[i].Coefficients[2], surface[i].Coefficients[3], Key_coordinate);
            if (surface[i].IsVertical == 0 && horizontal == 0 && distance <= 0.05) { vector3D[0] = surface[i].Area; horizontal++; }
            else if (surface[i].IsVertical == 1 && vertical <= 1 && distance <= 0.02) { vector3D[1 + vertical] = surface[i].Area; vertical++; }
        }
        if (vector3D[1] < vector3D[2]) std::swap(vector3D[1], vector3D[2]);
    }

    // key_point.h:112-161 and :251-318, one keypoint at a time like the reference (the batched form is
    // rtr_native_keypoint_descriptors over all keypoints of a cloud)
    void getOccupiedGrid(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, float resolution = 0.01f, float f_adjust = 0.1f) {
        rtr_native_params p; rtr_native_default_params(&p);
        p.resolution = resolution; p.occ_half = f_adjust;
        rtr_host::DeviceCloud d(*cloud);
        if (!d.h) return;
        int number = 0, count = 0;
        if (rtr_native_keypoint_descriptors(d.h, &Key_coordinate.x, 1, &p, &number, &count, nullptr, nullptr) != 0) return;
        Occupiedgrid.Number = number;
        if (!Occupiedgrid.cloud) Occupiedgrid.cloud.reset(new pcl::PointCloud<pcl::PointXYZ>);
        Occupiedgrid.cloud->clear();
        for (const auto& q : cloud->points)     // boxSearch: inclusive float box (key_point.h:118-140)
            if (q.x >= Key_coordinate.x - f_adjust && q.x <= Key_coordinate.x + f_adjust && q.y >= Key_coordinate.y - f_adjust &&
                q.y <= Key_coordinate.y + f_adjust && q.z >= Key_coordinate.z - f_adjust && q.z <= Key_coordinate.z + f_adjust)
                Occupiedgrid.cloud->push_back(q);
    }
    void get_TSDF(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, float resolution = 0.01f, float f_adjust = 0.15f) {
        rtr_native_params p; rtr_native_default_params(&p);
        p.resolution = resolution; p.tdf_half = f_adjust;
        for (int a = 0; a < 3; ++a) { Border[a] = (&Key_coordinate.x)[a] - f_adjust; Border[3 + a] = (&Key_coordinate.x)[a] + f_adjust; }
        rtr_host::DeviceCloud d(*cloud);
        if (!d.h) return;
        std::fill(grid_value.begin(), grid_value.end(), 0.f);
        if (rtr_native_keypoint_descriptors(d.h, &Key_coordinate.x, 1, &p, nullptr, nullptr, grid_value.data(), nullptr) != 0)
            fprintf(stderr, "ComputeTDFWithCuda failed!");     // key_point.h:315-318
    }
};

struct PairPoint { KeyPoint point_i; KeyPoint point_j; };      // function.h:23-26

inline double pointdistance(pcl::PointXYZ p1, pcl::PointXYZ p2) {   // function.h:27-30
    return std::sqrt((p1.x - p2.x) * (p1.x - p2.x) + (p1.y - p2.y) * (p1.y - p2.y) + (p1.z - p2.z) * (p1.z - p2.z));
}

// function.h:158-178 as intended; set quirk = true for the as-committed integer arithmetic (Appendix B#9, B#10)
inline bool match_by_height(pcl::PointXYZ& key1, pcl::PointXYZ& key2, bool quirk = false) {
    float temp = key1.z / key2.z;
    return quirk ? (temp >= float(2 / 3) || temp <= 1.5) : (temp >= 2.0f / 3.0f && temp <= 1.5f);
}
inline bool match_by_area(std::vector<double> v1, std::vector<double> v2, bool quirk = false) {
    double lo = quirk ? (double)float(1 / 3) : 1.0 / 3.0;
    for (int a = 0; a < 3; ++a) if ((v1[a] / v2[a]) > 3 || (v1[a] / v2[a]) < lo) return false;
    return true;
}
inline bool match_by_occupied(OccupiedGrid& o1, OccupiedGrid& o2, bool quirk = false) {
    if (o2.Number == 0) return false;
    float temp = quirk ? (float)(o1.Number / o2.Number) : (float)o1.Number / (float)o2.Number;
    return !(temp > 2 || temp < 0.5);
}

// matching.h:122 — the reference's signature, unchanged:
//     float get_Distance(Eigen::Matrix4f& key_transform, KeyPoint& p1_key, KeyPoint& p2_key, const float resolution = 0.01f)
// p1_key: a MODEL keypoint, p2_key: a SCAN keypoint, both after getOccupiedGrid (RealTimeRobot.cpp:52-69).  Everything the
// sweep needs travels inside the two records, as in the reference: the model keypoint's TDF is built from ITS occupancy
// cloud (key_point.h:251-318) and the scan keypoint contributes its occupancy cloud (the intended semantics, matching.h:17-120;
// Appendix B#3).  The two occupancy clouds go to the device, where cropping an occupancy cloud to its own +-0.1 box is the
// identity, and rtr_native_pair_scores evaluates the pair.  (Batched: rtr_native_pair_scores on the whole clouds.)
inline float get_Distance(Eigen::Matrix4f& key_transform, KeyPoint& p1_key, KeyPoint& p2_key, const float resolution = 0.01f) {
    key_transform = Eigen::Matrix4f::Identity();
    if (!p1_key.Occupiedgrid.cloud || !p2_key.Occupiedgrid.cloud) return 100000000.f;
    rtr_native_params p = rtr_host::native_params();
    p.resolution = resolution;
    rtr_host::DeviceCloud dm(*p1_key.Occupiedgrid.cloud), ds(*p2_key.Occupiedgrid.cloud);
    if (!dm.h || !ds.h) return 100000000.f;
    float score = 100000000.f; int step = 0;
    if (rtr_native_pair_scores(dm.h, &p1_key.Key_coordinate.x, 1, ds.h, &p2_key.Key_coordinate.x, 1, &p, &score, &step, key_transform.data()) != 0)
        return 100000000.f;
    return score;
}

// function.h:35-109 + the pair loop of main() (RealTimeRobot.cpp:70-104): the whole reference-native registration on two
// clouds.  Returns the transform main() applies to `cloud` (the scan); identity when nothing is consistent (the reference
// returns an uninitialised matrix there, Appendix B#2).
inline Eigen::Matrix4f Ransac(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, pcl::PointCloud<pcl::PointXYZ>::Ptr mcloud,
                              const rtr_native_params* params, rtr_pose_result* details = nullptr) {
    Eigen::Matrix4f m = Eigen::Matrix4f::Identity();
    rtr_native_params p;
    if (params) p = *params; else rtr_native_default_params(&p);
    rtr_host::DeviceCloud dm(*mcloud), ds(*cloud);
    if (!dm.h || !ds.h) return m;
    rtr_pose_result r;
    if (rtr_native_register(dm.h, ds.h, &p, &r) != 0) return m;
    memcpy(m.data(), r.pose, sizeof(r.pose));
    std::cout << r.evaluated << "aaaa" << std::endl;                     // RealTimeRobot.cpp:103
    if (details) *details = r;
    return m;
}

// function.h:35 — the reference's signature, unchanged:
//     Eigen::Matrix4f Ransac(vector<PairPoint> pairpoint, int ransac_times, Ptr cloud, Ptr mcloud)
// The reference ignores ransac_times and both clouds and enumerates every pair of `pairpoint`, re-running get_Distance for
// each (Appendix B#12).  Here the device redoes main()'s whole chain — corners, descriptors, all-pairs sweep, screens,
// exhaustive consensus — from the two clouds main() passes (RealTimeRobot.cpp:104), which yields the pair list main() built;
// `pairpoint` is accepted for source compatibility.
inline Eigen::Matrix4f Ransac(std::vector<PairPoint> /*pairpoint*/, int /*ransac_times*/, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud,
                              pcl::PointCloud<pcl::PointXYZ>::Ptr mcloud) {
    if (!cloud || !mcloud) return Eigen::Matrix4f::Identity();
    return Ransac(cloud, mcloud, &rtr_host::native_params(), nullptr);
}

// One scan against a database of models (README.md:10; RealTimeRobot.cpp:45-104 once per model): rtr_register_many_host.
// poses[m] / details[m] belong to models[m]; model_keypoints / scan_keypoints (optional) receive the refined Harris corners
// the batch found — ModelPoint::key_coordinates / ScanPoint::key_coordinates.
inline bool registerModelsToScene(const std::vector<pcl::PointCloud<pcl::PointXYZ>::Ptr>& models, const pcl::PointCloud<pcl::PointXYZ>& scene,
                                  const rtr_register_params& params, std::vector<Eigen::Matrix4f>& poses, std::vector<rtr_pose_result>* details = nullptr,
                                  std::vector<pcl::PointCloud<pcl::PointXYZ> >* model_keypoints = nullptr,
                                  pcl::PointCloud<pcl::PointXYZ>* scan_keypoints = nullptr) {
    rtr_context* ctx = rtr_host::default_context();
    poses.assign(models.size(), Eigen::Matrix4f::Identity());
    if (!ctx || models.empty()) return false;
    std::vector<const float*> ptrs(models.size());
    std::vector<int> ns(models.size());
    for (size_t m = 0; m < models.size(); ++m) {
        if (!models[m]) return false;
        ptrs[m] = models[m]->points.empty() ? nullptr : &models[m]->points[0].x;
        ns[m] = (int)models[m]->size();
    }
    std::vector<rtr_pose_result> rs(models.size());
    if (rtr_register_many_host(ctx, ptrs.data(), ns.data(), (int)models.size(), scene.points.empty() ? nullptr : &scene.points[0].x, (int)scene.size(),
                               &params, rs.data()) != 0) return false;
    for (size_t m = 0; m < models.size(); ++m) memcpy(poses[m].data(), rs[m].pose, sizeof(rs[m].pose));
    auto corners = [&](int member, pcl::PointCloud<pcl::PointXYZ>& out) {
        out.clear();
        std::vector<pcl::PointXYZ> kp(64);
        int n = 0;
        int rc = rtr_register_many_keypoints(ctx, member, &kp[0].x, 64, &n);
        if (rc != 0 && rc != RTR_ERR_CAPACITY) return;
        for (int i = 0; i < n && i < 64; ++i) out.push_back(pcl::PointXYZ(kp[i].x, kp[i].y, kp[i].z));
    };
    if (model_keypoints) { model_keypoints->resize(models.size()); for (size_t m = 0; m < models.size(); ++m) corners((int)m, (*model_keypoints)[m]); }
    if (scan_keypoints) corners((int)models.size(), *scan_keypoints);
    if (details) *details = rs;
    return true;
}

// The offline / online split of the reference's two mains: RealTimeRobot.cpp:124-165 preprocesses every database model once
// ("time to preprocess one database model"), :45-104 matches one scan against them.  add() is the offline half of one model
// (rtr_cloud_prepare: normals, Harris corners, FPFH rows stay on the device); match() the online half (rtr_register_prepared).
// poses[m] / details[m] belong to the m-th model added; the records are those of registerModelsToScene, bit for bit.
class ModelDatabase {
public:
    explicit ModelDatabase(const rtr_register_params& params) : params_(params) {}
    ~ModelDatabase() { for (rtr_cloud* c : models_) rtr_cloud_free(c); }
    ModelDatabase(const ModelDatabase&) = delete;
    ModelDatabase& operator=(const ModelDatabase&) = delete;
    size_t size() const { return models_.size(); }
    bool add(const pcl::PointCloud<pcl::PointXYZ>& model) {
        rtr_context* ctx = rtr_host::default_context();
        if (!ctx || model.points.empty()) return false;
        rtr_cloud* c = nullptr;
        if (rtr_cloud_upload(ctx, &model.points[0].x, (int)model.size(), &c) != 0) return false;
        if (rtr_cloud_prepare(c, &params_) != 0) { rtr_cloud_free(c); return false; }
        models_.push_back(c);
        return true;
    }
    bool match(const pcl::PointCloud<pcl::PointXYZ>& scene, std::vector<Eigen::Matrix4f>& poses, std::vector<rtr_pose_result>* details = nullptr) {
        rtr_context* ctx = rtr_host::default_context();
        poses.assign(models_.size(), Eigen::Matrix4f::Identity());
        if (!ctx || models_.empty() || scene.points.empty()) return false;
        rtr_cloud* scan = nullptr;
        if (rtr_cloud_upload(ctx, &scene.points[0].x, (int)scene.size(), &scan) != 0) return false;
        std::vector<rtr_pose_result> rs(models_.size());
        const int rc = rtr_register_prepared(models_.data(), (int)models_.size(), scan, &params_, rs.data());
        rtr_cloud_free(scan);
        if (rc != 0) return false;
        for (size_t m = 0; m < models_.size(); ++m) memcpy(poses[m].data(), rs[m].pose, sizeof(rs[m].pose));
        if (details) *details = rs;
        return true;
    }
private:
    rtr_register_params params_;
    std::vector<rtr_cloud*> models_;
};

// The north-star pipeline on two pcl clouds: model -> scene pose, mean squared fitness, RANSAC bookkeeping.
inline bool registerModelToScene(const pcl::PointCloud<pcl::PointXYZ>& model, const pcl::PointCloud<pcl::PointXYZ>& scene,
                                 const rtr_register_params& params, Eigen::Matrix4f& pose, rtr_pose_result* details = nullptr) {
    rtr_context* ctx = rtr_host::default_context();
    pose = Eigen::Matrix4f::Identity();
    if (!ctx) return false;
    rtr_pose_result r;
    if (rtr_register_host(ctx, model.points.empty() ? nullptr : &model.points[0].x, (int)model.size(),
                          scene.points.empty() ? nullptr : &scene.points[0].x, (int)scene.size(), &params, &r) != 0) return false;
    memcpy(pose.data(), r.pose, sizeof(r.pose));
    if (details) *details = r;
    return r.converged != 0;
}
