// asan_driver.cpp — runs the CPU oracle (TEST INFRASTRUCTURE, see oracle.cpp) under AddressSanitizer + UBSan.
// usage: oracle_asan model.f32 scene.f32   (raw float32 x,y,z,1 records, as tests/test_oracle_sanitizers.py writes them)
// Exercises every stage the registration path has (normals, Harris incl. refinement, FPFH, feature k-NN, prerejective
// RANSAC, ICP both estimators), the reference-native path (occupancy, TDF, yaw sweep, consensus) and the plane peel,
// and prints one line of results; any sanitizer report makes the process fail.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/rtr.h"

extern "C" {
void orc_set_threads(int t);
void orc_default_register_params(rtr_register_params* p);
void orc_register(const float* model_xyz1, int nm, const float* scene_xyz1, int nsc, const rtr_register_params* p, rtr_pose_result* res);
void orc_normals(const float* xyz1, int n, float radius, int mode, float* normals4);
int orc_harris3d(const float* xyz1, int n, const float* normals4, float radius, float threshold, int nms, int refine, float* resp,
                 int* kp_idx, float* kp_xyz1, int capacity);
void orc_native_register(const float* model_xyz1, int nm, const float* model_kp, int km, const float* scan_xyz1, int ns,
                         const float* scan_kp, int ks, const rtr_native_params* p, rtr_pose_result* res);
int orc_plane_areas(const float* xyz1, int n, rtr_surface* out, int capacity);
void orc_tdf(const int* occ, int num_occ, int dim, float* out);
}

static std::vector<float> slurp(const char* path) {
    std::vector<float> v;
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END); long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    v.resize((size_t)bytes / 4);
    if (fread(v.data(), 4, v.size(), f) != v.size()) { fprintf(stderr, "short read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s model.f32 scene.f32\n", argv[0]); return 2; }
    std::vector<float> m = slurp(argv[1]), s = slurp(argv[2]);
    int nm = (int)(m.size() / 4), ns = (int)(s.size() / 4);
    orc_set_threads(2);
    rtr_register_params p;
    orc_default_register_params(&p);
    p.ransac.max_iterations = 3000;
    rtr_pose_result r;
    orc_register(m.data(), nm, s.data(), ns, &p, &r);
    p.icp.estimator = 1;
    rtr_pose_result r1;
    orc_register(m.data(), nm, s.data(), ns, &p, &r1);
    // degenerate inputs the tests also feed the device: empty and 2-point clouds
    rtr_pose_result r2;
    orc_register(m.data(), 0, s.data(), ns, &p, &r2);
    orc_register(m.data(), 2, s.data(), 2, &p, &r2);
    // reference-native path on the Harris corners of both clouds
    std::vector<float> nm4((size_t)nm * 4), ns4((size_t)ns * 4), km((size_t)nm * 4), ks((size_t)ns * 4);
    std::vector<int> ki((size_t)std::max(nm, ns));
    orc_normals(m.data(), nm, 0.05f, 0, nm4.data());
    orc_normals(s.data(), ns, 0.05f, 1, ns4.data());          // mode 1: the PCL-float restatement
    orc_normals(s.data(), ns, 0.05f, 0, ns4.data());
    int nkm = orc_harris3d(m.data(), nm, nm4.data(), 0.05f, 0.01f, 1, 1, nullptr, ki.data(), km.data(), nm);
    int nks = orc_harris3d(s.data(), ns, ns4.data(), 0.05f, 0.01f, 1, 1, nullptr, ki.data(), ks.data(), ns);
    rtr_native_params np;
    np.resolution = 0.01f; np.occ_half = 0.1f; np.tdf_half = 0.15f; np.pair_gate = 30.0f; np.consensus_distance = 0.15f; np.consensus_score = 100.0f;
    np.quirk_skip_first_voxel = 0; np.quirk_running_score = 0; np.quirk_integer_screens = 0; np.use_plane_areas = 1;
    rtr_pose_result rn;
    orc_native_register(m.data(), nm, km.data(), nkm, s.data(), ns, ks.data(), nks, &np, &rn);
    np.quirk_skip_first_voxel = 1; np.quirk_running_score = 1; np.quirk_integer_screens = 1;
    orc_native_register(m.data(), nm, km.data(), nkm, s.data(), ns, ks.data(), nks, &np, &rn);
    std::vector<rtr_surface> surf(64);
    int npl = orc_plane_areas(m.data(), nm, surf.data(), 64);
    int occ[9] = {0, 0, 0, 29, 29, 29, 15, 3, 7};
    std::vector<float> tdf(27000);
    orc_tdf(occ, 3, 30, tdf.data());
    orc_tdf(occ, 0, 30, tdf.data());
    printf("asan ok: hypothesis %lld inliers %d fitness %g | lls iterations %d | native pairs %lld consensus %d | planes %d | corners %d %d\n",
           r.hypothesis, r.inliers, (double)r.fitness, r1.iterations, rn.evaluated, rn.inliers, npl, nkm, nks);
    return 0;
}
