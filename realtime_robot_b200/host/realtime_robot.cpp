// realtime_robot.cpp — Linux driver with the shape of the reference's main() (RealTimeRobot.cpp:27-121): load a model
// and a scan, extract keypoints of both, register the model to the scan, transform, save ASCII PCDs, print the running
// time.  MSVC-isms, hard-coded file names, the pcl_viewer hand-off and the trailing infinite loop (RealTimeRobot.cpp:114-120)
// are gone; everything numerical runs in librtr.so on the GPU.
//
//   realtime_robot <model.pcd> <scan.pcd> [--scale-model S] [--out transformed.pcd] [--hypotheses N] [--icp-only] [--native [--gate G]]
//   --native runs the reference's own descriptor path (occupancy / TDF / yaw sweep / exhaustive consensus) and, like
//   main(), transforms the SCAN into the model frame.
//   --reference-main [--gate G] runs main()'s own loops (RealTimeRobot.cpp:49-104) through the reference's SIGNATURES:
//   KeyPoint::getOccupiedGrid / get_TSDF / get_Vector3D per corner, get_Distance(matrix, model_key, scan_key) per pair, the
//   match_by_* screens, then Ransac(pairpoint, 50, cloud, mcloud) — what a maintainer's unchanged call sites execute.
//   realtime_robot --database <scan.pcd> <model1.pcd> [<model2.pcd> ...] [--hypotheses N]
//   registers every database model against the scan in one batch (registerModelsToScene -> rtr_register_many_host).
//   realtime_robot --database-online <model1.pcd> [<model2.pcd> ...] --scans <scan1.pcd> [<scan2.pcd> ...] [--hypotheses N]
//   the reference's offline / online split: every model preprocessed once (RealTimeRobot.cpp:124-165 -> ModelDatabase::add ->
//   rtr_cloud_prepare), then each scan matched against the prepared database (:45-104 -> ModelDatabase::match).
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>
#include "registration.h"

static int run_database(int argc, char** argv) {
    pcl::PointCloud<pcl::PointXYZ> scan;
    std::vector<pcl::PointCloud<pcl::PointXYZ>::Ptr> models;
    rtr_register_params p;
    rtr_default_register_params(&p);
    if (argc < 4 || pcl::io::loadPCDFile(argv[2], scan) != 0) return 2;
    for (int i = 3; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--hypotheses" && i + 1 < argc) { p.ransac.max_iterations = atoll(argv[++i]); continue; }
        pcl::PointCloud<pcl::PointXYZ>::Ptr m(new pcl::PointCloud<pcl::PointXYZ>);
        if (pcl::io::loadPCDFile(a, *m) != 0) return 1;
        models.push_back(m);
    }
    auto start = std::chrono::steady_clock::now();
    std::vector<Eigen::Matrix4f> poses;
    std::vector<rtr_pose_result> rs;
    std::vector<pcl::PointCloud<pcl::PointXYZ> > mk;
    pcl::PointCloud<pcl::PointXYZ> sk;
    if (!registerModelsToScene(models, scan, p, poses, &rs, &mk, &sk)) return 1;
    for (size_t m = 0; m < models.size(); ++m)
        std::cout << "model " << m << " converged " << (rs[m].converged != 0) << " hypothesis " << rs[m].hypothesis << " inliers " << rs[m].inliers
                  << " fitness " << rs[m].fitness << " keypoints " << mk[m].size() << " scan_keypoints " << sk.size() << "\n";
    std::cout << "Running Time : " << std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() << std::endl;
    return 0;
}

static int run_database_online(int argc, char** argv) {
    rtr_register_params p;
    rtr_default_register_params(&p);
    std::vector<std::string> model_files, scan_files;
    bool scans = false;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--hypotheses" && i + 1 < argc) { p.ransac.max_iterations = atoll(argv[++i]); continue; }
        if (a == "--scans") { scans = true; continue; }
        (scans ? scan_files : model_files).push_back(a);
    }
    if (model_files.empty() || scan_files.empty()) return 2;
    ModelDatabase db(p);
    auto start = std::chrono::steady_clock::now();
    for (const std::string& f : model_files) {
        pcl::PointCloud<pcl::PointXYZ> m;
        if (pcl::io::loadPCDFile(f, m) != 0 || !db.add(m)) return 1;
    }
    std::cout << "time to preprocess " << db.size() << " database models : " << std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() << std::endl;    // RealTimeRobot.cpp:164
    for (const std::string& f : scan_files) {
        pcl::PointCloud<pcl::PointXYZ> scan;
        if (pcl::io::loadPCDFile(f, scan) != 0) return 1;
        start = std::chrono::steady_clock::now();
        std::vector<Eigen::Matrix4f> poses;
        std::vector<rtr_pose_result> rs;
        if (!db.match(scan, poses, &rs)) return 1;
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        for (size_t m = 0; m < rs.size(); ++m)
            std::cout << "scan " << f << " model " << m << " converged " << (rs[m].converged != 0) << " hypothesis " << rs[m].hypothesis << " inliers " << rs[m].inliers
                      << " fitness " << rs[m].fitness << "\n";
        std::cout << "Running Time : " << secs << std::endl;
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "--database") return run_database(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "--database-online") return run_database_online(argc, argv);
    if (argc < 3) { fprintf(stderr, "usage: %s <model.pcd> <scan.pcd> [--scale-model S] [--out file.pcd] [--hypotheses N] [--icp-only]\n", argv[0]); return 2; }
    std::string out;
    float scale = 1.0f, gate = 3.0f; long long hyp = 0; bool icp_only = false, native = false, reference_main = false;
    for (int i = 3; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--scale-model" && i + 1 < argc) scale = strtof(argv[++i], nullptr);
        else if (a == "--out" && i + 1 < argc) out = argv[++i];
        else if (a == "--hypotheses" && i + 1 < argc) hyp = atoll(argv[++i]);
        else if (a == "--icp-only") icp_only = true;
        else if (a == "--native") native = true;
        else if (a == "--reference-main") reference_main = true;
        else if (a == "--gate" && i + 1 < argc) gate = strtof(argv[++i], nullptr);
    }
    pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>), mcloud(new pcl::PointCloud<pcl::PointXYZ>);
    if (pcl::io::loadPCDFile(argv[1], *mcloud) != 0 || pcl::io::loadPCDFile(argv[2], *cloud) != 0) return 1;
    if (scale != 1.0f) {                                    // what model_point.h:106-111 intends for Chair_025.pcd (units x100)
        Eigen::Matrix4f t = Eigen::Matrix4f::Identity();
        t(0, 0) = t(1, 1) = t(2, 2) = scale;
        pcl::transformPointCloud(*mcloud, *mcloud, t);
    }
    ModelPoint modelpoint(mcloud);
    modelpoint.getKeypoint();
    if (native || reference_main) { modelpoint.getArea(mcloud); std::cout << "model surfaces kept: " << modelpoint.surface.size() << std::endl; }     // RealTimeRobot.cpp:41
    auto start = std::chrono::steady_clock::now();          // RealTimeRobot.cpp:43
    ScanPoint scanpoint(cloud);
    scanpoint.getKeypoint();
    Eigen::Matrix4f matrix = Eigen::Matrix4f::Identity();
    pcl::PointCloud<pcl::PointXYZ> moved;
    if (reference_main) {
        // RealTimeRobot.cpp:47-104 with the reference's own call sites; `gate` replaces the literal 3 of :83
        rtr_host::native_params().pair_gate = gate;
        scanpoint.get_Area(cloud);
        std::vector<KeyPoint> model_keys, scan_keys;
        for (const auto& pt : modelpoint.key_coordinates.points) {          // :62-69
            KeyPoint k(pt);
            k.getOccupiedGrid(mcloud); k.get_TSDF(mcloud); k.get_Vector3D(modelpoint.surface);
            model_keys.push_back(k);
        }
        for (const auto& pt : scanpoint.key_coordinates.points) {           // :52-60
            KeyPoint k(pt);
            k.getOccupiedGrid(cloud); k.get_Vector3D(scanpoint.surface);
            scan_keys.push_back(k);
        }
        std::vector<PairPoint> pairpoint;
        for (auto& mkp : model_keys)                                         // :73-102
            for (auto& skp : scan_keys) {
                Eigen::Matrix4f m;
                if (get_Distance(m, mkp, skp) < gate && match_by_height(mkp.Key_coordinate, skp.Key_coordinate) &&
                    match_by_area(mkp.vector3D, skp.vector3D) && match_by_occupied(mkp.Occupiedgrid, skp.Occupiedgrid)) {
                    PairPoint pp; pp.point_i = mkp; pp.point_j = skp;
                    pairpoint.push_back(pp);
                }
            }
        std::cout << "pairs " << pairpoint.size() << std::endl;
        matrix = Ransac(pairpoint, 50, cloud, mcloud);                       // :104
        std::cout << "matrix\n" << matrix;
        pcl::transformPointCloud(*cloud, moved, matrix);                     // :105
    } else if (native) {
        rtr_native_params np; rtr_native_default_params(&np);
        np.pair_gate = gate;
        rtr_pose_result r = rtr_pose_result();
        matrix = Ransac(cloud, mcloud, &np, &r);                 // RealTimeRobot.cpp:104
        std::cout << "consensus " << r.inliers << " of " << r.evaluated << " screened pairs, winner " << r.hypothesis << "\n" << matrix;
        pcl::transformPointCloud(*cloud, moved, matrix);         // RealTimeRobot.cpp:105
    } else if (icp_only) {
        keyPointICP(nullptr, nullptr, cloud, mcloud, &matrix, &moved);
    } else {
        rtr_register_params p;
        rtr_default_register_params(&p);
        if (hyp > 0) p.ransac.max_iterations = hyp;
        rtr_pose_result r = rtr_pose_result();
        bool ok = registerModelToScene(*mcloud, *cloud, p, matrix, &r);
        std::cout << "converged " << ok << "  hypothesis " << r.hypothesis << "  inliers " << r.inliers << " / " << mcloud->size()
                  << "  prerejection survivors " << r.evaluated << "  fitness (mean squared NN distance) " << r.fitness << "\n" << matrix;
        pcl::transformPointCloud(*mcloud, moved, matrix);
    }
    if (!out.empty()) pcl::io::savePCDFileASCII(out, moved);   // RealTimeRobot.cpp:108-109
    auto ends = std::chrono::steady_clock::now();
    std::cout << "Running Time : " << std::chrono::duration<double>(ends - start).count() << std::endl;   // RealTimeRobot.cpp:111
    return 0;
}
