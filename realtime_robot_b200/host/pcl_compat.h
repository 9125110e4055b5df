// pcl_compat.h — the sliver of PCL / Eigen surface the reference's registration path touches, so that host code
// written against pcl::PointCloud<pcl::PointXYZ> (RealTimeRobot.cpp:32-35, model_point.h:84-89) compiles unchanged on
// a box without PCL, Eigen or Boost.  Layouts match PCL's: PointXYZ is 16 bytes (x, y, z, pad = 1.0f), i.e. one
// float4 — it is handed to librtr.so without conversion.  Matrix4f is column-major like Eigen::Matrix4f.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {
struct Matrix4f {
    float m[16];                                   // column-major: m[c * 4 + r]
    float& operator()(int r, int c) { return m[c * 4 + r]; }
    float operator()(int r, int c) const { return m[c * 4 + r]; }
    float* data() { return m; }
    const float* data() const { return m; }
    static Matrix4f Identity() { Matrix4f a; for (int i = 0; i < 16; ++i) a.m[i] = (i % 5 == 0) ? 1.f : 0.f; return a; }
    Matrix4f operator*(const Matrix4f& b) const {
        Matrix4f o;
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) {
            float acc = m[r] * b.m[c * 4];
            for (int k = 1; k < 4; ++k) acc += m[k * 4 + r] * b.m[c * 4 + k];
            o.m[c * 4 + r] = acc;
        }
        return o;
    }
};
inline std::ostream& operator<<(std::ostream& os, const Matrix4f& a) {
    for (int r = 0; r < 4; ++r) { for (int c = 0; c < 4; ++c) os << a(r, c) << (c < 3 ? " " : ""); os << "\n"; }
    return os;
}
}  // namespace Eigen

namespace pcl {

struct alignas(16) PointXYZ {
    float x = 0.f, y = 0.f, z = 0.f, pad = 1.0f;
    PointXYZ() = default;
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad(1.0f) {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ must be 16 bytes");

struct alignas(16) PointXYZI { float x = 0.f, y = 0.f, z = 0.f, pad = 1.0f, intensity = 0.f, pad2[3] = {0, 0, 0}; };

template <typename PointT>
class PointCloud {
public:
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    std::vector<PointT> points;
    uint32_t width = 0, height = 1;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = 0; height = 1; }
    void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
    void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
    PointT& at(size_t i) { return points.at(i); }
    const PointT& at(size_t i) const { return points.at(i); }
    PointT& operator[](size_t i) { return points[i]; }
    const PointT& operator[](size_t i) const { return points[i]; }
    typename std::vector<PointT>::iterator begin() { return points.begin(); }
    typename std::vector<PointT>::iterator end() { return points.end(); }
    typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
    typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};

// pcl::transformPointCloud(in, out, m): rows of m applied in float, ((m00 x + m01 y) + m02 z) + m03 (in may alias out)
inline void transformPointCloud(const PointCloud<PointXYZ>& in, PointCloud<PointXYZ>& out, const Eigen::Matrix4f& t) {
    std::vector<PointXYZ> res(in.points.size());
    for (size_t i = 0; i < in.points.size(); ++i) {
        const PointXYZ& p = in.points[i];
        res[i].x = ((t(0, 0) * p.x + t(0, 1) * p.y) + t(0, 2) * p.z) + t(0, 3);
        res[i].y = ((t(1, 0) * p.x + t(1, 1) * p.y) + t(1, 2) * p.z) + t(1, 3);
        res[i].z = ((t(2, 0) * p.x + t(2, 1) * p.y) + t(2, 2) * p.z) + t(2, 3);
    }
    out.points.swap(res);
    out.width = (uint32_t)out.points.size(); out.height = 1; out.is_dense = in.is_dense;
}

inline void getMinMax3D(const PointCloud<PointXYZ>& c, PointXYZ& mn, PointXYZ& mx) {
    mn = PointXYZ(INFINITY, INFINITY, INFINITY); mx = PointXYZ(-INFINITY, -INFINITY, -INFINITY);
    for (const auto& p : c.points) {
        if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
        mn.x = std::min(mn.x, p.x); mn.y = std::min(mn.y, p.y); mn.z = std::min(mn.z, p.z);
        mx.x = std::max(mx.x, p.x); mx.y = std::max(mx.y, p.y); mx.z = std::max(mx.z, p.z);
    }
}

namespace io {

// pcl::io::loadPCDFile for PCD v0.7, DATA ascii | binary | binary_compressed, keeping x y z (RealTimeRobot.cpp:34-35).
// The decoding (multi-threaded, LZF) lives in librtr.so (csrc/pcd_io.cu); PointXYZ is the 16-byte record it fills.
// 0 on success, -1 on error, like PCL.
inline int loadPCDFile(const std::string& path, PointCloud<PointXYZ>& cloud) {
    int n = 0, mode = 0;
    if (rtr_pcd_info(path.c_str(), &n, &mode) != 0) return -1;
    cloud.points.assign((size_t)n, PointXYZ());
    static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ is x, y, z + padding");
    if (rtr_pcd_read(path.c_str(), n ? &cloud.points[0].x : nullptr, n, &n) != 0) return -1;
    cloud.width = (uint32_t)n; cloud.height = 1; cloud.is_dense = true;
    return 0;
}

inline int save_pcd(const std::string& path, const PointCloud<PointXYZ>& cloud, int mode) {
    return rtr_pcd_write(path.c_str(), cloud.empty() ? nullptr : &cloud.points[0].x, (int)cloud.size(), mode) == 0 ? 0 : -1;
}
// pcl::io::savePCDFileASCII (RealTimeRobot.cpp:108-109; 8 significant digits like PCL)
inline int savePCDFileASCII(const std::string& path, const PointCloud<PointXYZ>& cloud) { return save_pcd(path, cloud, 0); }
inline int savePCDFileBinary(const std::string& path, const PointCloud<PointXYZ>& cloud) { return save_pcd(path, cloud, 1); }
inline int savePCDFileBinaryCompressed(const std::string& path, const PointCloud<PointXYZ>& cloud) { return save_pcd(path, cloud, 2); }
}  // namespace io
}  // namespace pcl
