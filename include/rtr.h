/* rtr.h — C ABI of librtr.so: the B200 (sm_100a) drop-in for the model-to-scene registration path
 * of ICCD/RealTime_Robot.
 *
 * Conventions (taken from the only FFI the reference has, `ComputeTDFWithCuda`,
 * RealTimeRobot/key_point.h:35-36 + kernel.cu:34-106):
 *   - plain pointers and sizes, no C++ / torch types;
 *   - every function returns an int status, 0 == success (cudaSuccess); on failure one tagged line
 *     goes to stderr and nothing is thrown across the boundary;
 *   - "host" pointers are caller-owned host memory; "dev" pointers are device memory on the
 *     context's GPU; handles (rtr_context, rtr_cloud) are created / destroyed explicitly.
 *   - points are the 16-byte pcl::PointXYZ layout: float x, y, z, pad (= 1.0f) — i.e. one float4.
 *   - 4x4 poses are 16 floats, COLUMN-major (Eigen::Matrix4f::data() order), mapping source -> target.
 *
 * Each entry point cites the reference interface it replaces (file:line under
 * /root/reference/RealTimeRobot).  Stages that exist only inside PCL 1.8.0 in the reference
 * cite the PCL class the reference instantiates and the reference call site.
 */
#ifndef RTR_H_
#define RTR_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RTR_OK                0
#define RTR_ERR_INVALID       1   /* bad argument (mirrors cudaErrorInvalidValue == 1) */
#define RTR_ERR_CAPACITY      3   /* caller buffer too small; required size reported through the count argument */
#define RTR_ERR_NOT_READY     4   /* a prerequisite stage has not been run on this cloud */
/* any other non-zero value is the cudaError_t of the failing CUDA call */

#define RTR_FPFH_DIM   33
#define RTR_TDF_DIM    30
#define RTR_TDF_VOXELS 27000   /* KeyPoint::grid_value[27000], key_point.h:65 */

typedef struct rtr_context rtr_context;
typedef struct rtr_cloud   rtr_cloud;

/* ------------------------------------------------------------------ parameters */

/* pcl::SampleConsensusPrerejective<PointXYZ,PointXYZ,FPFHSignature33> knobs.
 * The reference has no such stage (its consensus is function.h:35-109); the defaults are the PCL
 * tutorial's and are NOT reference parameters. */
typedef struct rtr_ransac_params {
    long long          max_iterations;               /* 50000 */
    long long          hypothesis_begin;             /* shard [begin, end) of 0..max_iterations; */
    long long          hypothesis_end;               /*   end <= 0 means max_iterations          */
    unsigned long long seed;                         /* counter-RNG key: hypothesis h depends on (seed, h) only */
    int                correspondence_k;             /* setCorrespondenceRandomness, 5 */
    float              similarity_threshold;         /* 0.9  (edge-length ratio, compared squared) */
    float              max_correspondence_distance;  /* 0.0365 (inlier iff nn d^2 < this^2) */
    float              inlier_fraction;              /* 0.25 */
} rtr_ransac_params;

/* pcl::IterativeClosestPoint<PointXYZ,PointXYZ> knobs; defaults = PCL's, which is what
 * keyPointICP() runs (function.h:111-117). */
typedef struct rtr_icp_params {
    int    max_iterations;                /* 10 */
    int    force_iterations;              /* 1: ignore convergence tests, always run max_iterations (bench cfg 3) */
    float  max_correspondence_distance;   /* <= 0: unlimited (PCL default sqrt(DBL_MAX)) */
    int    estimator;                     /* 0: TransformationEstimationSVD, point-to-point (pcl::IterativeClosestPoint, the reference's
                                           *    keyPointICP, function.h:112); 1: TransformationEstimationPointToPlaneLLS — the 6x6
                                           *    A^T A / A^T b system of pcl::IterativeClosestPointWithNormals; needs rtr_normals on the TARGET */
    double mse_threshold_absolute;        /* 1e-12 (DefaultConvergenceCriteria) */
} rtr_icp_params;

typedef struct rtr_register_params {
    float normal_radius;        /* 0.05  == Harris radius, model_point.h:130 */
    float harris_radius;        /* 0.05  model_point.h:130, scan_point.h:88 */
    float harris_threshold;     /* 0.01  model_point.h:131, scan_point.h:89 */
    int   harris_nms;           /* 1     model_point.h:129 */
    int   harris_refine;        /* 1     PCL default */
    float fpfh_radius;          /* 0.10  builder-chosen (no FPFH in the reference) */
    int   run_icp;              /* 1 */
    int   pad_;
    rtr_ransac_params ransac;
    rtr_icp_params    icp;
} rtr_register_params;

/* The reference's own descriptor path (key_point.h, matching.h, function.h); literals are the reference's.
 * The quirk_* flags reproduce as-committed behaviour and default to 0. */
typedef struct rtr_native_params {
    float resolution;              /* 0.01  key_point.h:112,251; matching.h:122 */
    float occ_half;                /* 0.1   KeyPoint::getOccupiedGrid f_adjust, key_point.h:112 */
    float tdf_half;                /* 0.15  KeyPoint::get_TSDF f_adjust, key_point.h:251 (dim = 30) */
    float pair_gate;               /* 3     get_Distance(...) < 3, RealTimeRobot.cpp:83 */
    float consensus_distance;      /* 0.15  function.h:78 */
    float consensus_score;         /* 100   function.h:78 */
    int   quirk_skip_first_voxel;  /* B#5:  key_point.h:298 (i = 1), matching.h:179 (p = 1) */
    int   quirk_running_score;     /* B#4:  distance_temp not reset between angles, matching.h:141,187,189 */
    int   quirk_integer_screens;   /* B#9, B#10: float(2/3) == 0 and integer division, function.h:161,167,174 */
    int   use_plane_areas;         /* 1: getArea + get_Vector3D feed match_by_area (RealTimeRobot.cpp:41,47,57,67); 0: KeyPoint defaults 0.16 */
} rtr_native_params;

/* One peeled plane of ModelPoint::getArea / ScanPoint::get_Area (model_point.h:170-245): the reference's Surface
 * (key_point.h:47-51: Area, Coefficients, IsVertical) plus bookkeeping. */
typedef struct rtr_surface {
    double area;              /* pcl::ConvexHull::getTotalArea of the plane's inliers */
    float  coefficients[4];   /* refined plane a x + b y + c z + d = 0 */
    int    is_vertical;       /* Surface::IsVertical (is_v_plane, model_point.h:65-79) */
    int    inliers;           /* points removed with this plane */
    int    dimension;         /* 2 or 3: the hull dimension ConvexHull detects */
    int    iterations;        /* RANSAC iterations run (adaptive, <= 151) */
    int    kept;              /* 1 if pushed to ModelPoint::surface: horizontal / vertical within 10 degrees and area >= 0.16 */
    int    pad_;
} rtr_surface;

/* fixed 128-byte record: the unit of the multi-GPU all-gather. */
typedef struct rtr_pose_result {
    float     pose[16];        /* column-major source->target */
    float     fitness;         /* mean squared NN distance (RANSAC: over inliers; ICP: all, getFitnessScore) */
    int       inliers;         /* RANSAC inlier count of the winner (ICP: correspondences of last iteration) */
    long long hypothesis;      /* winning hypothesis id, -1 if none accepted */
    long long evaluated;       /* hypotheses that passed polygon prerejection in this shard */
    int       converged;       /* RANSAC: any accepted; ICP: convergence state (1 iterations, 2 transform, 3 abs mse, 0 no) */
    int       iterations;      /* ICP iterations performed */
    int       model_id;        /* caller tag, copied through */
    int       n_keypoints_src; /* Harris corners found (register only) */
    int       n_keypoints_tgt;
    int       pad_[5];
} rtr_pose_result;

void rtr_default_register_params(rtr_register_params* p);

/* ------------------------------------------------------------------ context / clouds */

/* One context per GPU (device id, one stream, a grow-only workspace arena).  Replaces the
 * cudaSetDevice(0) + cudaMalloc/cudaFree done on EVERY call at kernel.cu:43-56,101-102. */
int rtr_context_create(int device, rtr_context** out);
/* Same, with a scheduling hint: urgency 0 = default, larger = served first by the device when several contexts compete
 * (mapped onto the CUDA stream priority range; longest-job-first for a batch of registrations of unequal size). */
int rtr_context_create_prio(int device, int urgency, rtr_context** out);
int rtr_context_destroy(rtr_context* ctx);
int rtr_context_sync(rtr_context* ctx);
/* cudaStream_t of the context, as void* (so callers can record CUDA events on it). */
void* rtr_context_stream(rtr_context* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches). */
long long rtr_context_launches(rtr_context* ctx);
/* CUDA-event timing on the context's own stream: record into slot 0..15, elapsed between two slots (syncs on b). */
int rtr_event_record(rtr_context* ctx, int slot);
int rtr_event_elapsed_ms(rtr_context* ctx, int slot_a, int slot_b, float* ms);

/* Per-operation device timing (CUDA events after every kernel / CUB pass / memset on the context's stream).
 * rtr_profile_end writes "tag count total_ms" lines into buf. */
int rtr_profile_begin(rtr_context* ctx);
int rtr_profile_end(rtr_context* ctx, char* buf, int capacity);

/* Upload a pcl::PointCloud<pcl::PointXYZ>::points array (n x 16 B, host).  Replaces the cloud hand-off
 * ModelPoint(Ptr) / ScanPoint(Ptr), model_point.h:165-168, scan_point.h:43-54. */
int rtr_cloud_upload(rtr_context* ctx, const float* host_xyz1, int n, rtr_cloud** out);
/* Same, from a device buffer (copied). */
int rtr_cloud_from_device(rtr_context* ctx, const float* dev_xyz1, int n, rtr_cloud** out);
int rtr_cloud_free(rtr_cloud* c);
int rtr_cloud_size(const rtr_cloud* c);
/* Drop all cached stages (grids, normals, features, correspondences); keep the points. */
int rtr_cloud_reset(rtr_cloud* c);
/* Apply a 4x4 (column-major) to the cloud in place: pcl::transformPointCloud(*c, *c, m)
 * (model_point.h:111, RealTimeRobot.cpp:105).  Invalidates cached stages. */
int rtr_cloud_transform(rtr_cloud* c, const float* pose16);
/* Copy points back to host (n x 4 floats). */
int rtr_cloud_download(rtr_cloud* c, float* host_xyz1);

/* ------------------------------------------------------------------ PCD v0.7 I/O
 * pcl::io::loadPCDFile (RealTimeRobot.cpp:34-35, scan_point.h:62) and pcl::io::savePCDFileASCII (RealTimeRobot.cpp:108-109,
 * function.h:126-127), for DATA ascii | binary | binary_compressed; only x y z (float32) are kept, as the reference loads
 * every file into pcl::PointCloud<pcl::PointXYZ>.  Host-side, multi-threaded; no GPU needed except for rtr_pcd_load /
 * rtr_cloud_save.  data_mode / mode: 0 ascii (8 significant digits, what savePCDFileASCII prints), 1 binary,
 * 2 binary_compressed (LZF, field-major). */
int rtr_pcd_info(const char* path, int* n_points, int* data_mode);
/* Decode into n x 4 floats (x, y, z, 1).  RTR_ERR_CAPACITY (with *n_points set) when capacity is too small. */
int rtr_pcd_read(const char* path, float* host_xyz1, int capacity, int* n_points);
/* File -> pinned staging owned by the context -> device cloud (one asynchronous copy). */
int rtr_pcd_load(rtr_context* ctx, const char* path, rtr_cloud** out);
int rtr_pcd_write(const char* path, const float* host_xyz1, int n, int mode);
int rtr_cloud_save(rtr_cloud* c, const char* path, int mode);

/* ------------------------------------------------------------------ neighbour index */

/* Exact radius neighbour sets over the uniform grid, for queries == the cloud's own points:
 * pcl::search::KdTree::radiusSearch as used inside HarrisKeypoint3D (model_point.h:127-136):
 * d^2 < r^2 strict, self included.  counts[n]; if indices != NULL it receives, per query i at
 * offsets[i] (exclusive prefix of counts, also returned), the neighbour indices in ascending order.
 * *total receives sum(counts); RTR_ERR_CAPACITY if capacity < *total. */
int rtr_radius_neighbors(rtr_cloud* c, float radius, int* host_counts, long long* host_offsets,
                         int* host_indices, long long capacity, long long* total);

/* Exact 1-NN of each query (host, nq x 16 B) in the cloud: KdTreeFLANN::nearestKSearch(q,1) as used by
 * IterativeClosestPoint (function.h:112-117).  Ties -> lowest index. */
int rtr_nearest(rtr_cloud* target, const float* host_queries_xyz1, int nq, int* host_idx, float* host_d2);

/* ------------------------------------------------------------------ per-cloud stages (results cached on device) */

/* pcl::NormalEstimation (run implicitly by HarrisKeypoint3D, model_point.h:127-136; App. A.2).
 * host_normals4 (optional): n x {nx, ny, nz, curvature}. */
int rtr_normals(rtr_cloud* c, float radius, float* host_normals4);
/* The same stage with the arithmetic selectable:
 *   mode 0  exact — fp64 sums of the offsets from the query point, cyclic Jacobi; what rtr_normals and rtr_register run and
 *           what every parity test is pinned to (GPU == oracle mode 0 bit for bit);
 *   mode 1  PCL-float-faithful — what pcl::NormalEstimation::computePointNormal itself does (App. A.2): nine single-pass FLOAT
 *           accumulators over the raw coordinates (no de-meaning), pcl::eigen33's closed-form roots in float.  Agrees with the
 *           oracle's mode 1 to float summation order (tolerance parity), runs off the fp64 pipe.
 * The normals cached on the cloud are those of the last call; later stages (Harris, FPFH) use whatever is cached. */
int rtr_normals_mode(rtr_cloud* c, float radius, int mode, float* host_normals4);

/* pcl::HarrisKeypoint3D<PointXYZ,PointXYZI,Normal>::compute — ModelPoint::getKeypoint (model_point.h:99-156),
 * ScanPoint::getKeypoint (scan_point.h:57-113).  Needs rtr_normals first (same radius in the reference).
 * host_response (optional) n floats; keypoints: original indices (ascending) + refined xyz1. */
int rtr_harris3d(rtr_cloud* c, float radius, float threshold, int nms, int refine,
                 float* host_response, int* host_kp_index, float* host_kp_xyz1, int capacity, int* n_keypoints);

/* pcl::FPFHEstimation<PointXYZ,Normal,FPFHSignature33> over all points (input == surface; App. A.4).
 * Needs rtr_normals.  host_fpfh (optional): n x 33 floats. */
int rtr_fpfh(rtr_cloud* c, float radius, float* host_fpfh);

/* The same estimator with input != surface: pcl::FPFHEstimation::setInputCloud(keypoints) + setSearchSurface(cloud)
 * (App. A.4; BASELINE.json configs[3]: 100 000 query points on a 4 M-point surface).  host_query_index: n_query ORIGINAL
 * indices into the cloud; host_fpfh: n_query x 33 floats, row i = the row rtr_fpfh gives point host_query_index[i].
 * SPFH signatures are computed only for the points some query's neighbourhood needs.  Needs rtr_normals. */
int rtr_fpfh_at(rtr_cloud* c, float radius, const int* host_query_index, int n_query, float* host_fpfh);

/* k nearest target features of every source feature, 33-D squared L2, ascending, ties -> lowest index:
 * KdTreeFLANN<FPFHSignature33>::nearestKSearch inside SampleConsensusPrerejective::findSimilarFeatures
 * (App. A.5).  Both clouds need rtr_fpfh.  Result cached on `source`; host outputs optional (ns x k). */
int rtr_match_features(rtr_cloud* source, rtr_cloud* target, int k, int* host_idx, float* host_dist);

/* Same search on two bare feature arrays (host, ns x 33 and nt x 33 floats): the descriptor-correspondence stage on its
 * own, for descriptors computed elsewhere.  *kernel_ms (optional) receives the device time of the search itself. */
int rtr_match_features_raw(rtr_context* ctx, const float* host_source_feat, int ns, const float* host_target_feat, int nt,
                           int k, int* host_idx, float* host_dist, float* kernel_ms);

/* stats3[0] = rows the tensor-core prefilter could not certify and the exact kernel redid (-1: the exact kernel did
 * everything), [1] = target splits, [2] = largest observed prefilter error / (|a||b|) in units of 1e-9, for the last rtr_match_features_raw call on this context. */
int rtr_match_last_stats(rtr_context* ctx, int* stats3);

/* ------------------------------------------------------------------ pose stages */

/* SampleConsensusPrerejective::computeTransformation over hypotheses [begin,end) (App. A.5).
 * Needs rtr_match_features(source, target, k >= correspondence_k). */
int rtr_ransac_prerejective(rtr_cloud* source, rtr_cloud* target, const rtr_ransac_params* p,
                            rtr_pose_result* host_result);

/* pcl::IterativeClosestPoint::align with an initial guess — keyPointICP (function.h:111-123).
 * init_pose16 may be NULL (identity). */
int rtr_icp(rtr_cloud* source, rtr_cloud* target, const rtr_icp_params* p, const float* init_pose16,
            rtr_pose_result* host_result);

/* The whole model-to-scene registration the north star names, device-resident end to end:
 * grid -> normals -> Harris -> FPFH (both clouds) -> feature k-NN -> prerejective RANSAC -> ICP (skipped when RANSAC
 * accepted no hypothesis: the result then keeps the identity pose, fitness FLT_MAX, converged 0).
 * Sequencing counterpart of main(), RealTimeRobot.cpp:39-105. */
int rtr_register(rtr_cloud* model, rtr_cloud* scene, const rtr_register_params* p, rtr_pose_result* host_result);

/* One-call form with HOST clouds in and a HOST result out (upload + register + free): the e2e
 * entry point a pcl::PointCloud<PointXYZ> caller uses. */
int rtr_register_host(rtr_context* ctx, const float* host_model_xyz1, int n_model,
                      const float* host_scene_xyz1, int n_scene, const rtr_register_params* p,
                      rtr_pose_result* host_result);

/* The same two calls split in an enqueue half and a wait half: _begin returns as soon as the whole registration is
 * queued on the context's stream (no host synchronisation); _end waits and returns the record.  One registration may be in
 * flight per context, so ONE host thread can keep many contexts (streams) busy: begin on each, then end on each.
 * The synchronous forms above (one registration, latency) queue the scene's stages on the context's second stream so they
 * overlap the model's; the split forms (many in flight, throughput) keep one stream per registration. */
int rtr_register_begin(rtr_cloud* model, rtr_cloud* scene, const rtr_register_params* p);
int rtr_register_host_begin(rtr_context* ctx, const float* host_model_xyz1, int n_model, const float* host_scene_xyz1, int n_scene,
                            const rtr_register_params* p);
int rtr_register_end(rtr_context* ctx, rtr_pose_result* host_result);

/* One scan against MANY database models (README.md:10: "match the segmented objects against the model database").  The reference
 * builds the scan side once — ScanPoint keypoints and descriptors, RealTimeRobot.cpp:45-60 — and then loops over the model's
 * keypoints (:62-102); the per-model work is the offline side (:124-165).  These entry points do the same for a batch: the
 * scan's grid / normals / Harris / FPFH run once, every per-cloud stage of the models and the scan is ONE launch over all of
 * them (a "model set": clouds concatenated, one uniform grid per cloud), descriptor matching is one search of all model
 * features against the scan's, and the models' RANSAC hypotheses and ICP iterations share their launches.
 * host_results[m] is what rtr_register(models[m], scene) returns (bit for bit), with model_id = m.
 * Batches of up to 31 models go through the model-set path; clouds of 65536 points or more, sweeps above 2^20 hypotheses
 * and degenerate clouds (< 3 points) fall back to one rtr_register per model inside the same call. */
int rtr_register_many(rtr_cloud* const* models, int n_models, rtr_cloud* scene, const rtr_register_params* p,
                      rtr_pose_result* host_results);
/* HOST clouds in (n_models pointers + sizes, and the scan), HOST records out: uploads + batch + free in one call. */
int rtr_register_many_host(rtr_context* ctx, const float* const* host_models_xyz1, const int* n_points, int n_models,
                           const float* host_scene_xyz1, int n_scene, const rtr_register_params* p, rtr_pose_result* host_results);
/* Enqueue / wait halves (at most 31 models, model-set shapes only; RTR_ERR_INVALID otherwise).  Host buffers must stay
 * valid until _end. */
int rtr_register_many_begin(rtr_cloud* const* models, int n_models, rtr_cloud* scene, const rtr_register_params* p);
int rtr_register_many_host_begin(rtr_context* ctx, const float* const* host_models_xyz1, const int* n_points, int n_models,
                                 const float* host_scene_xyz1, int n_scene, const rtr_register_params* p);
int rtr_register_many_end(rtr_context* ctx, rtr_pose_result* host_results, int capacity);
/* The Harris corners the last batch found — ModelPoint::key_coordinates / ScanPoint::key_coordinates (model_point.h:146-152,
 * scan_point.h:104-110), refined xyz1 — of member `member` (0 .. n_models-1: the models, n_models: the scan).  They travel
 * back with the records (up to 64 per cloud; *n_keypoints is the full count, RTR_ERR_CAPACITY if not all fit). */
int rtr_register_many_keypoints(rtr_context* ctx, int member, float* host_kp_xyz1, int capacity, int* n_keypoints);

/* The reference's OFFLINE / ONLINE split.  Its commented-out second main (RealTimeRobot.cpp:124-165, "time to preprocess one
 * database model") builds every database model's keypoints and descriptors once; the live main (:45-104) does the per-scan work
 * against them.  rtr_cloud_prepare is the offline half: normals, Harris corners and FPFH rows of one cloud, computed with the
 * stage parameters of *p (normal_radius, harris_*, fpfh_radius) and kept on the cloud until rtr_cloud_reset / rtr_cloud_free.
 * rtr_register_prepared is the online half of rtr_register_many: every model must have been prepared with the same stage
 * parameters (RTR_ERR_NOT_READY otherwise); the scan's stages run inside the call, once (or not at all if the scan was prepared
 * too); then descriptor matching, RANSAC and ICP for all models in shared launches.  host_results[m] is what
 * rtr_register_many / rtr_register return for that model, bit for bit.  At most 31 models per _begin; rtr_register_prepared
 * takes any number (batches of 31).  _begin is completed by rtr_register_many_end; rtr_register_many_keypoints works as above. */
int rtr_cloud_prepare(rtr_cloud* cloud, const rtr_register_params* p);
int rtr_register_prepared(rtr_cloud* const* models, int n_models, rtr_cloud* scene, const rtr_register_params* p,
                          rtr_pose_result* host_results);
int rtr_register_prepared_begin(rtr_cloud* const* models, int n_models, rtr_cloud* scene, const rtr_register_params* p);

/* ------------------------------------------------------------------ multi-GPU: one small all-gather
 * The path shards by candidate model cloud (rank r registers its models against the replicated scan) and by RANSAC hypothesis
 * range [hypothesis_begin, hypothesis_end); the only exchange is ONE ncclAllGather of the 128-byte records.  One process per
 * GPU; rank 0 calls rtr_comm_unique_id and ships the 128 bytes to the others by any means (MPI, a file, torch.distributed),
 * then every rank calls rtr_comm_init on its context.  NCCL is bound at run time (the copy already in the process, else
 * libnccl.so.2); a host without NCCL gets RTR_ERR_NOT_READY here and nothing else changes. */
int rtr_comm_unique_id(char* id128);
int rtr_comm_init(rtr_context* ctx, int world, int rank, const char* id128);
int rtr_comm_destroy(rtr_context* ctx);
int rtr_comm_world(rtr_context* ctx, int* world, int* rank);
/* Every rank contributes n_local (<= 64, the same everywhere) records; host_all receives world x n_local in rank order.
 * Without a communicator it is a copy.  Cost: one H2D, one ncclAllGather on the context's stream, one D2H, one sync. */
int rtr_allgather_results(rtr_context* ctx, const rtr_pose_result* host_local, int n_local, rtr_pose_result* host_all);
/* In-stream form for batches: after rtr_comm_gather_batches(ctx, 1, base) every rtr_register_many* on this context ends with the
 * all-gather of its records queued on the stream right behind the batch (no extra host round trip or synchronisation);
 * model_id of local record k becomes base + k; every rank must run batches of the same size (<= 64 models).
 * rtr_gathered_results returns the world x n_models records (rank order) after rtr_register_many_end. */
int rtr_comm_gather_batches(rtr_context* ctx, int on, int model_id_base);
int rtr_gathered_results(rtr_context* ctx, rtr_pose_result* host_all, int capacity, int* n_records);
/* Winner of a hypothesis-sharded registration: arg-min over (fitness, hypothesis id) among the accepted shard records —
 * identical on every rank and for every world size; `evaluated` is the sum over shards. */
int rtr_select_best_hypothesis(const rtr_pose_result* records, int n, rtr_pose_result* best);

/* ------------------------------------------------------------------ reference-native descriptor path */

/* THE reference FFI, exported unchanged (key_point.h:35-36, kernel.cu:34-35).  Host pointers;
 * voxel_grid_occ = num_occ packed (x,y,z) int32; voxel_grid_TDF = 27000 floats in/out
 * (only the first dim^3 are written); synchronous; returns cudaSuccess (0) or the failing code. */
int ComputeTDFWithCuda(const int* voxel_grid_occ, float* voxel_grid_TDF, int voxel_grid_dim, int num_occ);

/* Batched form: n_grids keypoints in one launch.  occ = concatenated triples, occ_offsets[n_grids+1]
 * in triples; tdf_out = n_grids x dim^3 floats.  All host pointers.  Replaces the per-keypoint loop
 * RealTimeRobot.cpp:62-69 -> key_point.h:313. */
int rtr_tdf_batch(rtr_context* ctx, const int* host_occ, const int* host_occ_offsets, int n_grids,
                  int dim, float* host_tdf_out);
/* Same with device pointers, asynchronous on the context stream. */
int rtr_tdf_batch_dev(rtr_context* ctx, const int* dev_occ, const int* dev_occ_offsets, int n_grids,
                      int dim, float* dev_tdf_out);

/* ModelPoint::getArea / ScanPoint::get_Area (model_point.h:170-245, scan_point.h:117-188): peel planes with
 * SACSegmentation(PLANE, RANSAC, 150 it, 5 mm, optimised) until <= 15 % of the points remain, convex-hull area of each.
 * host_surfaces receives EVERY peeled plane in order (kept = 1 marks those the reference pushes to `surface`). */
int rtr_plane_areas(rtr_cloud* c, rtr_surface* host_surfaces, int capacity, int* n_planes);

/* ---- the reference's own pipeline, batched on the device ---- */
void rtr_native_default_params(rtr_native_params* p);

/* KeyPoint::getOccupiedGrid + KeyPoint::get_TSDF for every keypoint of a cloud in one go (key_point.h:112-161,251-318;
 * loops RealTimeRobot.cpp:52-69).  host_kp_xyz1: n_kp x 16 B keypoint coordinates (ModelPoint::key_coordinates).
 * Outputs (host, each optional): number[n_kp] = Occupiedgrid.Number; occ_count[n_kp] = points in the +-occ_half box;
 * tdf[n_kp x 27000] = KeyPoint::grid_value; voxel_count[n_kp] = triples handed to the TDF kernel. */
int rtr_native_keypoint_descriptors(rtr_cloud* c, const float* host_kp_xyz1, int n_kp, const rtr_native_params* p,
                                    int* host_number, int* host_occ_count, float* host_tdf, int* host_voxel_count);

/* get_Distance (matching.h:122-222) for all Km x Ks pairs: the 36-step yaw sweep of every scan keypoint's occupancy
 * cloud against every model keypoint's TDF.  Outputs (host, Km x Ks, model-major): score, best step, 4x4 transform. */
int rtr_native_pair_scores(rtr_cloud* model, const float* host_model_kp_xyz1, int km, rtr_cloud* scan,
                           const float* host_scan_kp_xyz1, int ks, const rtr_native_params* p, float* host_score,
                           int* host_best_step, float* host_transform16);

/* main() as the reference intends it (RealTimeRobot.cpp:39-105): Harris corners of both clouds -> occupancy / TDF
 * descriptors -> all-pairs sweep -> match_by_* screens -> exhaustive Ransac (function.h:35-109).  The pose maps the SCAN
 * into the model frame (main applies it to `cloud`).  result: inliers = consensus size, hypothesis = winning pair in
 * screening order (-1 if none: pose = identity, where the reference returns an uninitialised matrix, Appendix B#2),
 * evaluated = screened pairs, fitness = the winner's sweep score. */
int rtr_native_register(rtr_cloud* model, rtr_cloud* scan, const rtr_native_params* p, rtr_pose_result* host_result);

#ifdef __cplusplus
}
#endif
#endif /* RTR_H_ */
