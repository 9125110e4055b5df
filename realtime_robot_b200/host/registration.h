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

class ModelPoint {
public:
    pcl::PointCloud<pcl::PointXYZ>::Ptr Mpoint;
    pcl::PointCloud<pcl::PointXYZ> key_coordinates;
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
    pcl::PointCloud<pcl::PointXYZ> key_coordinates;
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
