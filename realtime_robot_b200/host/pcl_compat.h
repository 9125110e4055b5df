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

// pcl::io::loadPCDFile for PCD v0.7, DATA ascii | binary, keeping x y z (RealTimeRobot.cpp:34-35).  0 on success, -1 on error.
inline int loadPCDFile(const std::string& path, PointCloud<PointXYZ>& cloud) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { fprintf(stderr, "[pcl_compat] cannot open %s\n", path.c_str()); return -1; }
    std::vector<std::string> fields; std::vector<int> sizes, counts; std::vector<char> types;
    size_t npts = 0, w = 0, h = 1; std::string mode, line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line); std::string key; ss >> key;
        if (key == "FIELDS") { std::string v; while (ss >> v) fields.push_back(v); }
        else if (key == "SIZE") { int v; while (ss >> v) sizes.push_back(v); }
        else if (key == "TYPE") { char v; while (ss >> v) types.push_back(v); }
        else if (key == "COUNT") { int v; while (ss >> v) counts.push_back(v); }
        else if (key == "WIDTH") ss >> w;
        else if (key == "HEIGHT") ss >> h;
        else if (key == "POINTS") ss >> npts;
        else if (key == "DATA") { ss >> mode; break; }
    }
    if (npts == 0) npts = w * h;
    if (counts.empty()) counts.assign(fields.size(), 1);
    if (fields.size() != sizes.size() || fields.size() != types.size() || fields.size() != counts.size()) return -1;
    int col[3] = {-1, -1, -1}, off[3] = {-1, -1, -1}, c = 0, o = 0;
    for (size_t i = 0; i < fields.size(); ++i) {
        for (int a = 0; a < 3; ++a) if (fields[i] == std::string(1, "xyz"[a])) { col[a] = c; off[a] = o; if (types[i] != 'F' || sizes[i] != 4) return -1; }
        c += counts[i]; o += sizes[i] * counts[i];
    }
    if (col[0] < 0 || col[1] < 0 || col[2] < 0) return -1;
    cloud.points.assign(npts, PointXYZ());
    if (mode == "ascii") {
        for (size_t i = 0; i < npts; ++i) {
            if (!std::getline(f, line)) return -1;
            std::istringstream ss(line); std::string tok; int k = 0;
            while (ss >> tok) {
                for (int a = 0; a < 3; ++a) if (k == col[a]) (&cloud.points[i].x)[a] = strtof(tok.c_str(), nullptr);
                ++k;
            }
        }
    } else if (mode == "binary") {
        std::vector<char> rec(o);
        for (size_t i = 0; i < npts; ++i) {
            if (!f.read(rec.data(), o)) return -1;
            for (int a = 0; a < 3; ++a) memcpy(&(&cloud.points[i].x)[a], rec.data() + off[a], 4);
        }
    } else { fprintf(stderr, "[pcl_compat] DATA %s not supported\n", mode.c_str()); return -1; }
    cloud.width = (uint32_t)npts; cloud.height = 1; cloud.is_dense = true;
    return 0;
}

inline std::string pcd_header(size_t n, const char* mode) {
    std::ostringstream h;
    h << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH " << n
      << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA " << mode << "\n";
    return h.str();
}

// pcl::io::savePCDFileASCII (RealTimeRobot.cpp:108-109; 8 significant digits like PCL)
inline int savePCDFileASCII(const std::string& path, const PointCloud<PointXYZ>& cloud) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return -1;
    fputs(pcd_header(cloud.size(), "ascii").c_str(), f);
    for (const auto& p : cloud.points) fprintf(f, "%.8g %.8g %.8g\n", p.x, p.y, p.z);
    fclose(f);
    return 0;
}
inline int savePCDFileBinary(const std::string& path, const PointCloud<PointXYZ>& cloud) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return -1;
    fputs(pcd_header(cloud.size(), "binary").c_str(), f);
    for (const auto& p : cloud.points) fwrite(&p.x, 4, 3, f);
    fclose(f);
    return 0;
}
}  // namespace io
}  // namespace pcl
