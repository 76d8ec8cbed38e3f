/*
 * nrslam_b200.h — C ABI of the B200-native deformable-SLAM optimisation core.
 *
 * This is the drop-in boundary for the ONE hot path of endomapper/NR-SLAM:
 *   - CameraPoseOptimization                 modules/optimization/g2o_optimization.h:27   (.cc:50-146)
 *   - CameraPoseAndDeformationOptimization   modules/optimization/g2o_optimization.h:29-32 (.cc:148-557)
 *   - LocalDeformableBundleAdjustment        modules/optimization/g2o_optimization.h:39-40 (.cc:880-1161)
 *   - LucasKanadeTracker::{SetReferenceImage,Track,Get/InsertPhotometricInformation,clear}
 *                                            modules/matching/lucas_kanade_tracker.h:55-70 (.cc:47-631)
 *   - RegularizationGraph::{GetEdges,UpdateVertex}   modules/map/regularization_graph.h:73,78 (.cc:71-146)
 * The reference has no FFI of its own; INTEGRATION.md shows the C++ shim that forwards the four
 * reference signatures to these entry points; the shim sources are shim/g2o_optimization_b200.cc and
 * shim/lucas_kanade_tracker_b200.cc (compiled in the CPU test suite against stand-in headers, tests/test_shim.py).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, all pointers are HOST pointers unless a name ends in _dev.
 *   - geometry is fp32 at the boundary (the reference stores Eigen::Vector3f / Sophus::SE3f), indices int32.
 *   - a pose is 7 floats [qx qy qz qw tx ty tz] = Sophus::SE3f camera_transform_world (unit quaternion + t).
 *   - statuses use the reference enum values (utilities/landmark_status.h:23-30).
 *   - object lifetime: trackers / extractors / triangulators / pre-processors created on a context hold a pointer
 *     to it and must be destroyed BEFORE nrslam_b200_destroy(ctx) (destroying them afterwards is a use after free).
 *   - return value: 0 ok; < 0 CUDA / allocation / argument error; > 0 numerical condition
 *     (1 = fewer points than the reference needs, nothing done). Never throws, never aborts.
 *     nrslam_b200_last_error(ctx) returns a description of the last non-zero return.
 *   - a ctx is single-caller (not re-entrant); calls block until results are in the host buffers.
 *   - host side effects: nrslam_b200_create() raises glibc's mmap / trim thresholds (mallopt) so that the ~60 MB of
 *     staging vectors of a BA call are served from the heap instead of fresh mappings (process-wide; set
 *     NRSLAM_B200_MALLOC_TUNE=0 to leave the allocator alone), and the tracking calls keep up to 8 staging threads
 *     (NRSLAM_B200_HOST_THREADS caps them) that spin only while a call is staging.
 *   - there is NO CPU fallback: every entry point that computes fails with NRSLAM_B200_ERR_NO_DEVICE
 *     when no sm_100 device is present.
 */
#ifndef NRSLAM_B200_H
#define NRSLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRSLAM_B200_ABI_VERSION 2

#define NRSLAM_B200_OK 0
#define NRSLAM_B200_ERR_NO_DEVICE (-1)
#define NRSLAM_B200_ERR_CUDA (-2)
#define NRSLAM_B200_ERR_ARG (-3)
#define NRSLAM_B200_ERR_ALLOC (-4)
#define NRSLAM_B200_ERR_NCCL (-5)
#define NRSLAM_B200_NUM_TOO_FEW (1)
#define NRSLAM_B200_NUM_NONFINITE (2)

/* LandmarkStatus — utilities/landmark_status.h:23-30 */
enum {
  NRSLAM_TRACKED_WITH_3D = 0,
  NRSLAM_TRACKED = 1,
  NRSLAM_JUST_TRIANGULATED = 2,
  NRSLAM_BAD = 3,
  NRSLAM_OUT_IMAGE_BOUNDARIES = 4,
  NRSLAM_BAD_FEATURE = 5
};
/* RegularizationGraph::Status — map/regularization_graph.h:41-46 */
enum { NRSLAM_EDGE_VERIFIED = 0, NRSLAM_EDGE_NEIGHBOR = 1, NRSLAM_EDGE_NEUTRAL = 2, NRSLAM_EDGE_BAD = 3 };

/* CameraModel — calibration/pin_hole.cc (model 0: fx fy cx cy), calibration/kannala_brandt_8.cc
 * (model 1: fx fy cx cy k0 k1 k2 k3). Evaluated in fp32 like the reference (camera_model.h:89-95). */
typedef struct nrslam_b200_camera {
  int32_t model;
  float params[8];
} nrslam_b200_camera;

/* Every literal of g2o_optimization.cc with the reference value as default. */
typedef struct nrslam_b200_options {
  float th_huber_2dof_sq;        /* 5.99   g2o_optimization.cc:63,197,960 */
  float th_huber_3dof_sq;        /* 0.584  :200,963 */
  float sigma_reprojection;      /* 0.5 px :203,966 */
  float sigma_position;          /* 0.1    :206,969 */
  float sigma_spatial_factor;    /* 0.1 (x scale) :209,972 */
  float spring_k;                /* 1.1    :328,1069 */
  int32_t regularizers_per_point;/* 10 (loop admits 11) :195,958 */
  int32_t pose_only_iterations[3];   /* {10,10,10} :103 */
  int32_t pose_deform_iterations[2]; /* {10,10}    :338 */
  int32_t lost_iterations;       /* 10 :538 */
  int32_t ba_iterations;         /* 5  :1143 */
  int32_t lm_max_trials;         /* 10  optimization_algorithm_levenberg.cpp:52 */
  double lm_tau;                 /* 1e-5 :43 */
  /* Linear solver. The reference factorises exactly (sparse LL^T); the GPU core runs a matrix-free
   * block-preconditioned CG to this relative residual ||r||_M^-1 / ||b||_M^-1. */
  double pcg_rel_tol;            /* default 1e-8 */
  int32_t pcg_max_iterations;    /* default 2000 */
  int32_t device;                /* CUDA device ordinal, default: LOCAL_RANK env or 0 */
  int32_t grid_ctas;             /* persistent-kernel CTAs, 0 = auto from problem size */
} nrslam_b200_options;

/* Regularisation graph (map/regularization_graph.h:48-58,89) as CSR over map points.
 * Vertices are map points in ASCENDING MapPoint id (the btree_map order); row entries ascending too.
 * One attribute record per undirected edge (the reference shares one Edge object between both
 * endpoints, regularization_graph.cc:53-54); eid[] maps each CSR entry to its record. */
typedef struct nrslam_b200_graph {
  int32_t n_vertices;
  int32_t n_edges;            /* undirected */
  const int32_t* rowptr;      /* [n_vertices + 1] */
  const int32_t* col;         /* [2 * n_edges] neighbour vertex */
  const int32_t* eid;         /* [2 * n_edges] undirected edge id */
  float* weight;              /* [n_edges] in/out  Edge::weight */
  const float* first_distance;/* [n_edges]         Edge::first_distance */
  float* min_distance;        /* [n_edges] in/out */
  float* max_distance;        /* [n_edges] in/out */
  uint8_t* status;            /* [n_edges] in/out  NRSLAM_EDGE_* */
  float weight_sigma;         /* Options::weight_sigma */
  float stretching_th;        /* Options::streching_th (1.1, map.cc:28-29) */
} nrslam_b200_graph;

#define NRSLAM_B200_TRACE 64
typedef struct nrslam_b200_stats {
  int32_t lm_iterations;      /* LM iterations (g2o solve() calls) */
  int32_t lm_trials;          /* damped linear solves */
  int32_t pcg_iterations;     /* total CG iterations */
  int32_t n_sweeps;           /* normal-equation sweeps (linearise + assemble) */
  int32_t n_chi2_passes;      /* residual-only passes */
  int32_t n_reproj_edges, n_pair_edges, n_spring_edges, n_damper_edges, n_fixed_edges;
  int32_t n_points, n_poses;
  int32_t kernel_launches;    /* kernels launched by this call */
  int32_t n_trace;
  double chi2_trace[NRSLAM_B200_TRACE]; /* accepted robust chi2 after each LM iteration */
  double lambda_final;
  float gpu_ms;               /* device time of the solve stage (CUDA events on the ctx stream) */
  float host_ms;              /* wall time of the whole call */
  float stage_ms;             /* host edge selection + H2D */
  int64_t h2d_bytes;          /* bytes copied host -> device by this call */
  int64_t d2h_bytes;          /* bytes copied device -> host by this call */
  int32_t grid_ctas, block_threads; /* launch geometry of the LM kernel */
  /* ABI 2: exact-solve engine (block L D L^T, nrs_direct.cu) */
  int32_t direct_solves;      /* damped systems factorised exactly (0: the CG engine ran) */
  int32_t solve_failures;     /* factorisations that met a non-positive pivot / CG break-downs */
  int64_t factor_doubles;     /* doubles of the stored factor L' (nnz incl. the dense-front padding), last launch */
  int64_t update_doubles;     /* doubles of the update (Schur complement) matrices written per factorisation */
  int32_t plan_reused;        /* 1: the symbolic analysis of the previous frame was re-used (same points, pairs inside its adjacency) */
  int32_t reserved0;
} nrslam_b200_stats;

typedef struct nrslam_b200_ctx nrslam_b200_ctx;

void nrslam_b200_default_options(nrslam_b200_options* opt);
int nrslam_b200_abi_version(void);
/* opt == NULL -> defaults */
int nrslam_b200_create(const nrslam_b200_options* opt, nrslam_b200_ctx** out);
void nrslam_b200_destroy(nrslam_b200_ctx* ctx);
const char* nrslam_b200_last_error(const nrslam_b200_ctx* ctx);
/* Device the ctx is bound to and its SM count (diagnostics). */
int nrslam_b200_device_info(const nrslam_b200_ctx* ctx, int32_t* device, int32_t* sm_count);

/* ---- CameraPoseOptimization (g2o_optimization.cc:50-146) -------------------------------------
 * n TRACKED_WITH_3D observations: uv[2n] keypoints, X[3n] world positions. pose_io: seed in, result out.
 * inlier_out[n] (optional): the reference's `inliers` vector after the third round. */
int nrslam_b200_pose_only(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                          const float* X, float* pose_io, uint8_t* inlier_out, nrslam_b200_stats* stats);

/* ---- CameraPoseAndDeformationOptimization (g2o_optimization.cc:148-557) -----------------------
 * n optimised points in frame order (Frame::Get*WithStatus({TRACKED_WITH_3D}), frame.cc:83-118):
 *   uv[2n], X_rest[3n] (landmark positions), point_vertex[n] = graph vertex of each point.
 * vertex_frame_status[g->n_vertices]: -1 if the map point is not in Frame::MapPointIdToIndex(), else its
 *   LandmarkStatus in the frame (drives the "lost" classification, :264-273).
 * last_world_position_io[3 * n_vertices]: MapPoint::GetLastWorldPosition of every graph vertex; updated
 *   for inliers (:446) and for lost points (:550).
 * Outputs (all optional except pose_io):
 *   deformation_out[3n]      optimised deformation per point
 *   X_out[3n]                Frame::LandmarkPositions after the call (rest + d where accepted, :444)
 *   chi2_out[n]              final reprojection chi2 (:424)
 *   status_out[n]            LandmarkStatus after the call (TRACKED_WITH_3D / TRACKED / BAD, :428,435,472)
 *   median_deformation_out   Frame::SetDeformationMaginitud (:455)
 *   lost_vertex_out[n_vertices], n_lost_out : returned set of lost map points (graph vertices, ascending)
 * The graph attribute arrays are updated in place (RegularizationGraph::UpdateVertex, :458-474). */
int nrslam_b200_pose_deform(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                            const float* X_rest, const int32_t* point_vertex,
                            const int8_t* vertex_frame_status, nrslam_b200_graph* g, float scale,
                            float* pose_io, float* last_world_position_io, float* deformation_out,
                            float* X_out, float* chi2_out, uint8_t* status_out,
                            float* median_deformation_out, int32_t* lost_vertex_out, int32_t* n_lost_out,
                            nrslam_b200_stats* stats);

/* ---- Tracking::TrackCameraAndDeformation after the data association (tracking/tracking.cc:291-330) -------------
 * CameraPoseOptimization followed by CameraPoseAndDeformationOptimization on the same TRACKED_WITH_3D points, the
 * second seeded with the first one's pose. Arguments as nrslam_b200_pose_deform (pose_io: the motion-model seed in,
 * the final pose out); pose_only_out[7] / pose_only_inlier_out[n] (optional) return what CameraPoseOptimization alone
 * would have returned. Results equal the two calls in sequence bit for bit; the host staging of the second problem
 * overlaps the first kernel and the seed pose never leaves HBM. */
int nrslam_b200_track_pose_and_deform(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                                      const float* X_rest, const int32_t* point_vertex,
                                      const int8_t* vertex_frame_status, nrslam_b200_graph* g, float scale,
                                      float* pose_io, float* last_world_position_io, float* deformation_out,
                                      float* X_out, float* chi2_out, uint8_t* status_out,
                                      float* median_deformation_out, int32_t* lost_vertex_out, int32_t* n_lost_out,
                                      float* pose_only_out, uint8_t* pose_only_inlier_out,
                                      nrslam_b200_stats* stats_pose_only, nrslam_b200_stats* stats);

/* ---- LocalDeformableBundleAdjustment (g2o_optimization.cc:880-1161) ---------------------------
 * n_kf keyframes of the window, OLDEST FIRST (the reference walks keyframes_in_optimization.rbegin(),
 * :930,982). n_obs observations grouped by keyframe in that order, inside a keyframe in
 * KeyFrame::Get*WithStatus({TRACKED_WITH_3D}) order: obs_kf[n_obs] (non-decreasing slot in
 * [0,n_kf)), obs_vertex[n_obs] graph vertex (map point), uv[2 n_obs], X_io[3 n_obs] per-keyframe
 * landmark positions (in: KeyFrame::LandmarkPositions, out: optimised). kf_pose_io[7 n_kf].
 * The reference hard-codes a 5-keyframe window (:894) and returns when < 3 (:922): window selection is
 * the caller's; n_kf < 3 returns NRSLAM_B200_NUM_TOO_FEW untouched. iterations <= 0 -> options. */
int nrslam_b200_local_ba(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n_kf,
                         float* kf_pose_io, int32_t n_obs, const int32_t* obs_kf,
                         const int32_t* obs_vertex, const float* uv, float* X_io,
                         const nrslam_b200_graph* g, float scale, int32_t iterations,
                         nrslam_b200_stats* stats);

/* ---- Landmark-sharded LocalDeformableBundleAdjustment over up to 8 GPUs of one NVSwitch domain ----
 * Same problem and same results (within the stated FP tolerance) as nrslam_b200_local_ba
 * (g2o_optimization.cc:880-1161); one process and one ctx per GPU. Every rank passes the FULL window and derives the
 * same partition: landmarks are cut along a Morton curve into `world` ranges of equal observation count, a rank
 * optimises every per-keyframe copy of its landmarks, keeps read-only halo copies of the neighbours its springs /
 * dampers reach on other ranks, and the ranks exchange halo rows and partial sums (pose blocks of H and b, chi2,
 * CG scalars) through peer-mapped buffers inside the persistent kernel — no host round trip per iteration.
 *   shard_init   allocates this rank's exchange buffer (capacity: max_rows own + halo rows, max_poses keyframes,
 *                identical on every rank) and returns its CUDA IPC handle;
 *   shard_attach maps the other ranks' buffers; handles = world x NRSLAM_B200_IPC_HANDLE_BYTES in rank order
 *                (exchanged by the caller, e.g. torch.distributed.all_gather);
 *   local_ba_sharded  is collective: every rank must call it with identical arguments. kf_pose_io is updated
 *                on every rank (bit-identical), X_io only for the observations this rank owns; owner_out[n_obs]
 *                (optional) names the owner of every observation so the caller can gather the rest.
 * A rank that fails to arrive makes the others time out (NRSLAM_B200_XTIMEOUT_MS, default 20 s) with
 * NRSLAM_B200_ERR_CUDA instead of hanging; the sharded state is then unusable until the ctx is re-created. */
#define NRSLAM_B200_IPC_HANDLE_BYTES 64
int nrslam_b200_shard_init(nrslam_b200_ctx* ctx, int32_t rank, int32_t world, int32_t max_rows,
                           int32_t max_poses, unsigned char* handle_out);
int nrslam_b200_shard_attach(nrslam_b200_ctx* ctx, const unsigned char* handles);
int nrslam_b200_local_ba_sharded(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n_kf,
                                 float* kf_pose_io, int32_t n_obs, const int32_t* obs_kf,
                                 const int32_t* obs_vertex, const float* uv, float* X_io,
                                 const nrslam_b200_graph* g, float scale, int32_t iterations,
                                 int32_t* owner_out, nrslam_b200_stats* stats);
/* Host-only view of that partition (no GPU needed): owner_out[n_obs]; per rank the rows it owns, its halo rows, the
 * halo copies it refreshes, and n_edges_out[3 world] = (springs, dampers, edges whose chi2 it counts). */
int nrslam_b200_shard_partition(const nrslam_b200_options* opt, int32_t world, int32_t n_kf,
                                const float* kf_pose, int32_t n_obs, const int32_t* obs_kf,
                                const int32_t* obs_vertex, const float* uv, const float* X,
                                const nrslam_b200_graph* g, float scale, int32_t* owner_out,
                                int32_t* n_own_out, int32_t* n_halo_out, int32_t* n_push_out,
                                int32_t* n_edges_out);

/* Re-run the device solve of the most recently staged problem on its HBM-resident inputs (no host<->device
 * copies, no host bookkeeping). which: 0 pose_only, 1 pose_deform (both robust rounds), 2 local_ba,
 * 3 the lost-point stage of the last pose_deform call.
 * Benchmark / profiler hook: results are identical to the staged call's. */
int nrslam_b200_resolve(nrslam_b200_ctx* ctx, int32_t which, nrslam_b200_stats* stats);

/* ---- RegularizationGraph (map/regularization_graph.cc:71-146) ---------------------------------
 * get_edges: neighbours of `vertex` sorted by (status asc, weight desc, neighbour asc — E13 tie-break)
 * truncated at the first weight < min_weight; writes CSR entry indices (into col/eid) to out, returns count.
 * update_vertex: UpdateVertex; returns the number of good connections. positions = last world positions. */
int32_t nrslam_b200_graph_get_edges(const nrslam_b200_graph* g, int32_t vertex, int32_t* out_entries,
                                    int32_t capacity);
int32_t nrslam_b200_graph_update_vertex(nrslam_b200_graph* g, int32_t vertex, const float* positions);

/* ---- RegularizationGraph construction: AddEdge / SetSigma (map/regularization_graph.cc:33-55) ---------------
 * A graph STORE owns the edge records and their adjacency across frames, so a host keeps ONE resident graph instead of
 * rebuilding a CSR per call: Map::InitializeRegularizationGraph (map.cc:139-167) and Mapping's insertion of newly
 * triangulated landmarks (mapping.cc:238-256) become add_edges batches; the optimisation entry points take the store's
 * CSR view and refresh its attribute arrays in place (pose_deform's UpdateVertex loop writes straight into the store).
 *   add_edges: for k < n: AddEdge(v1[k], v2[k], relative_position[3k..]): distance = |relative_position|,
 *     first / min / max distance = distance, weight = exp(-d^2 / (2 sigma^2)), status NEUTRAL. Vertices are dense
 *     indices >= 0 in ascending MapPoint id; the store grows to the largest index seen. An EXISTING pair is replaced
 *     (attributes reset), exactly like graph_[a][b] = edge does; v1 == v2 or a negative index is an argument error.
 *   set_sigma: SetSigma (:33-36): later edges and min_weight use the new sigma, existing weights stay.
 *   view: fills *out with pointers into the store (rows and row entries ascending); valid until the next add_edges /
 *     destroy. Returns 0, < 0 on an argument error. */
typedef struct nrslam_b200_graph_store nrslam_b200_graph_store;
int nrslam_b200_graph_store_create(float weight_sigma, float stretching_th, nrslam_b200_graph_store** out);
void nrslam_b200_graph_store_destroy(nrslam_b200_graph_store* store);
int nrslam_b200_graph_store_add_edges(nrslam_b200_graph_store* store, int32_t n, const int32_t* v1, const int32_t* v2,
                                      const float* relative_position);
int nrslam_b200_graph_store_set_sigma(nrslam_b200_graph_store* store, float weight_sigma);
int nrslam_b200_graph_store_view(nrslam_b200_graph_store* store, nrslam_b200_graph* out);

/* ---- LucasKanadeTracker (matching/lucas_kanade_tracker.h:55-92) -------------------------------
 * One tracker object per reference image. Images are 8-bit single channel, `pitch` bytes per row.
 * win_size must be 21 (the only window the reference uses, modules/SLAM/system.cc:78-83).
 * Point i of every call is point i of SetReferenceImage (+ inserted ones): KLT index == frame index. */
typedef struct nrslam_b200_klt nrslam_b200_klt;
int nrslam_b200_klt_create(nrslam_b200_ctx* ctx, int32_t win_size, int32_t max_level, int32_t max_iters,
                           float epsilon, float min_eig_threshold, nrslam_b200_klt** out);
void nrslam_b200_klt_destroy(nrslam_b200_klt* klt);
/* SetReferenceImage (.cc:47-168). mask may be NULL (cv::Mat()). */
int nrslam_b200_klt_set_reference(nrslam_b200_klt* klt, const uint8_t* image, int32_t width, int32_t height,
                                  int32_t pitch, int32_t n_points, const float* pts_xy, const uint8_t* mask,
                                  int32_t mask_pitch);
/* Track (.cc:170-596). pts_io[2n]: initial flow in (if use_initial_flow) / tracked positions out.
 * status_io[n]: LandmarkStatus in/out. n_tracked_out: the reference's return value. */
int nrslam_b200_klt_track(nrslam_b200_klt* klt, const uint8_t* image, int32_t width, int32_t height,
                          int32_t pitch, int32_t n_points, float* pts_io, uint8_t* status_io,
                          int32_t use_initial_flow, float min_ssim, const uint8_t* mask, int32_t mask_pitch,
                          int32_t* n_tracked_out);
/* Get/InsertPhotometricInformation (.cc:598-620): per level (max_level+1) a win*win int16 patch, a
 * win*win*2 int16 gradient patch, mean and mean^2. valid_out[level] = 0 when the reference holds an empty Mat. */
int nrslam_b200_klt_get_patch(nrslam_b200_klt* klt, int32_t idx, int16_t* gray_out, int16_t* grad_out,
                              float* mean_out, float* mean2_out, uint8_t* valid_out);
int nrslam_b200_klt_insert_patch(nrslam_b200_klt* klt, float x, float y, const int16_t* gray,
                                 const int16_t* grad, const float* mean, const float* mean2,
                                 const uint8_t* valid);
/* Batch form of InsertPhotometricInformation: n points (xy[2n], per point the arrays of insert_patch back to back). */
int nrslam_b200_klt_insert_patches(nrslam_b200_klt* klt, int32_t n, const float* xy, const int16_t* gray,
                                   const int16_t* grad, const float* mean, const float* mean2, const uint8_t* valid);
/* ---- Tracking::PointReuse (modules/tracking/tracking.cc:394-506) ------------------------------------------------
 * n map points: X_world[3n] = MapPoint::GetLastWorldPosition, in_frame[n] = 1 when Frame::LandmarkPosition(id) is ok
 * (:398), forced[n] (optional) = 1 for the lost ids CameraPoseAndDeformationOptimization returned. Per point the
 * PhotometricInformation of its two finest pyramid levels: gray[n][2][21*21] int16, grad[n][2][21*21][2] int16,
 * mean[n][2], mean2[n][2], valid[n][2] (0 = empty Mat). pose = the frame's camera_transform_world. The function
 * projects (fp32), keeps the points with depth >= 0 inside the image, tracks them with a fresh 2-level tracker
 * (window 21, max_iters / epsilon / min_eig_threshold = Tracking::Options, initial flow = projection, SSIM 0.75)
 * and gates by SquaredReprojectionError > 5.99.
 * Outputs, one entry per candidate j < *n_cand_out in ascending point index: cand_out[j] (point index), seed_out
 * (optional, the projection), uv_out[2j..] tracked keypoint, status_out (optional, LandmarkStatus after Track),
 * accepted_out[j] = 1 when the reference inserts / refreshes the observation (:481-498). */
int nrslam_b200_point_reuse(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, const float* pose, const uint8_t* image,
                            int32_t width, int32_t height, int32_t pitch, const uint8_t* mask, int32_t mask_pitch,
                            int32_t n, const float* X_world, const uint8_t* in_frame, const uint8_t* forced,
                            int32_t max_iters, float epsilon, float min_eig_threshold, const int16_t* gray,
                            const int16_t* grad, const float* mean, const float* mean2, const uint8_t* valid,
                            int32_t* cand_out, float* seed_out, float* uv_out, uint8_t* status_out,
                            uint8_t* accepted_out, int32_t* n_cand_out, int32_t* n_reused_out);
int nrslam_b200_klt_clear(nrslam_b200_klt* klt);
int32_t nrslam_b200_klt_num_points(const nrslam_b200_klt* klt);
/* Re-run the last Track (pyramid of the device-resident current image + tracking kernel) on the device-resident
 * inputs (benchmark hook); device ms out. */
int nrslam_b200_klt_retrack(nrslam_b200_klt* klt, float* gpu_ms_out);
/* Diagnostics: read back one level of the reference (which = 0) or current (1) pyramid WITH its winSize border,
 * as cv::buildOpticalFlowPyramid lays it out (lucas_kanade_tracker.cc:50,184): img_out (h+2*win) x (w+2*win) u8,
 * deriv_out same size x 2 int16. Used by the bit-exact pyramid parity test. */
int nrslam_b200_klt_debug_level(nrslam_b200_klt* klt, int32_t which, int32_t level, uint8_t* img_out,
                                int16_t* deriv_out, int32_t* w_out, int32_t* h_out);

/* ---- ShiTomasi (features/shi_tomasi.h:30-60, features/feature.h:34 Feature::Extract) ------------------------
 * One extractor object carries the running class-id counter (shi_tomasi.cc:81). extract(): 8-bit image; existing_xy
 * = keypoints already in the frame (their rounded pixel is excluded with a 15-px window, :89-99); out: the NEW
 * keypoints in raster order (x, y as floats, octave size 1) with consecutive class ids. n_out = number found (may
 * exceed capacity: only `capacity` are written, the counter advances by n_out like the reference's). Scores within
 * 4 rows / 1 column of the image border are defined as 0 (the reference leaves row-rotation artefacts there,
 * DESIGN.md 7b). */
typedef struct nrslam_b200_shi nrslam_b200_shi;
int nrslam_b200_shi_create(nrslam_b200_ctx* ctx, int32_t nms_window, nrslam_b200_shi** out);
void nrslam_b200_shi_destroy(nrslam_b200_shi* shi);
int nrslam_b200_shi_extract(nrslam_b200_shi* shi, const uint8_t* image, int32_t width, int32_t height, int32_t pitch,
                            const float* existing_xy, int32_t n_existing, float* out_xy, int32_t* out_class_id,
                            int32_t capacity, int32_t* n_out);
/* Diagnostics: the score map of the last extract (height x width floats, -1 marks included). */
int nrslam_b200_shi_debug_scores(nrslam_b200_shi* shi, float* scores_out);

/* ---- Image pre-processing ---------------------------------------------------
 * pre_image = System::ImageProcessing (SLAM/system.cc:189-201): cv::cvtColor(RGB2GRAY) then cv::CLAHE(clip_limit,
 * tiles) — the reference uses createCLAHE(3.0, Size(8, 8)) (system.cc:37). rgb: 8-bit interleaved R,G,B rows of
 * `pitch` bytes. gray_out / clahe_out [width * height] (either may be NULL); both stay resident on the device.
 * pre_mask = Masker::mask / GetAllMasks()["Global"] (masking/masker.cc:80-92,94-115): AND of the filters' masks,
 * then erode 10x10. gray == NULL re-uses the gray image of the last pre_image call without a host round trip.
 *   BRIGHT      threshold(gray, th, 255, THRESH_BINARY_INV), erode(ellipse 11x11), GaussianBlur(11x11, sigma 5,
 *               REFLECT_101)                                                (masking/bright_filter.cc:24-39)
 *   BORDER      255 inside [rb, rows - re) x [cb, cols - ce) where gray != 0, erode(rect 21x21)
 *                                                                           (masking/border_filter.cc:24-40)
 *   PREDEFINED  a caller-prepared mask (PredefinedFilter loads and erodes it once at start-up,
 *               masking/predefined_filter.cc:27-41)
 * Results are bit-exact with OpenCV 4 (tests/golden/preproc.npz). */
typedef struct nrslam_b200_pre nrslam_b200_pre;
#define NRSLAM_B200_FILTER_BRIGHT 0
#define NRSLAM_B200_FILTER_BORDER 1
#define NRSLAM_B200_FILTER_PREDEFINED 2
typedef struct nrslam_b200_mask_filter {
  int32_t kind;
  int32_t th;               /* BRIGHT */
  int32_t rb, re, cb, ce;   /* BORDER: rows / columns cut at the beginning / end */
  const uint8_t* mask;      /* PREDEFINED: [width * height] */
} nrslam_b200_mask_filter;
int nrslam_b200_pre_create(nrslam_b200_ctx* ctx, int32_t max_width, int32_t max_height, nrslam_b200_pre** out);
void nrslam_b200_pre_destroy(nrslam_b200_pre* pre);
int nrslam_b200_pre_image(nrslam_b200_pre* pre, const uint8_t* rgb, int32_t width, int32_t height, int32_t pitch,
                          float clip_limit, int32_t tiles_x, int32_t tiles_y, uint8_t* gray_out,
                          uint8_t* clahe_out);
int nrslam_b200_pre_mask(nrslam_b200_pre* pre, const uint8_t* gray, int32_t width, int32_t height,
                         const nrslam_b200_mask_filter* filters, int32_t n_filters, uint8_t* mask_out);
/* Device time (CUDA events) and kernel count of the last pre_image / pre_mask call. */
float nrslam_b200_pre_last_ms(const nrslam_b200_pre* pre);
int32_t nrslam_b200_pre_last_launches(const nrslam_b200_pre* pre);

/* ---- DeformableTriangulation, batched over all candidates of a frame ------------------
 * Replaces the per-candidate calls of
 *   absl::StatusOr<Eigen::Vector3f> DeformableTriangulation(TemporalBuffer&, int candidate_id,
 *                                                           std::shared_ptr<CameraModel>, const float scale)
 * (modules/optimization/g2o_optimization.h:34-37, .cc:559-814) that Mapping::LandmarkTriangulation issues in a loop
 * (modules/mapping/mapping.cc:88-113): hundreds of independent small LM problems per non-keyframe frame, one CTA each.
 * The caller (shim) flattens what the function reads from the TemporalBuffer:
 *   track_ptr[n_cand + 1]  CSR over track entries; candidate c owns entries track_ptr[c] .. track_ptr[c+1]-1 =
 *                          TemporalBuffer::GetFeatureTrack(candidate) (ascending frame id: OLDEST FIRST,
 *                          temporal_buffer.cc:173-183); 1 <= length <= NRSLAM_B200_TRI_MAX_TRACK
 *   track_uv[2 E]          keypoint.pt of the candidate in that frame
 *   track_pose[7 E]        TemporalBuffer::GetCameraTransformWorld(frame) (camera_transform_world)
 *   n_neighbours[n_cand]   size of GetClosestMapPointsToFeature(candidate, 10, 20, 500) (0 .. 11: the loop admits
 *                          num_neighbors + 1, temporal_buffer.cc:134-140); 0 -> "Feature too close to other ones"
 *   nb_pos[3 NB E]         GetLandmarkPosition(frame, neighbour k) (world), NB = NRSLAM_B200_TRI_MAX_NB slots per entry
 *   nb_valid[NB E]         its .ok()
 * (E = track_ptr[n_cand].) `scale` is accepted and unused, like in the reference body.
 * Outputs: position_out[3 n_cand] the triangulated world position (valid when status_out == NRSLAM_B200_TRI_OK),
 * status_out[n_cand] one code per absl::InternalError the reference returns (the StatusOr), lm_iterations_out
 * (optional) the LM iterations g2o ran. The LM solve is exact (dense LL^T of the (3 T)^2 system in shared memory),
 * the reprojection edge is differentiated numerically with delta = 1e-9 through the fp32 camera model exactly like
 * g2o does for ReprojectionErrorOnlyDeformation, which declares no analytic Jacobian
 * (reprojection_error_only_deformation.h:40, base_fixed_sized_edge.hpp:160-199). */
#define NRSLAM_B200_TRI_MAX_TRACK 48
#define NRSLAM_B200_TRI_MAX_NB 12
enum {
  NRSLAM_B200_TRI_OK = 0,
  NRSLAM_B200_TRI_TOO_CLOSE = 1,        /* "Feature too close to other ones."           .cc:569-571 */
  NRSLAM_B200_TRI_HIGH_REPROJ_FIRST = 2,/* "High reprojection error at first camera."   .cc:618-620 */
  NRSLAM_B200_TRI_HIGH_REPROJ_SECOND = 3,/* "High reprojection error at second camera." .cc:625-627 */
  NRSLAM_B200_TRI_LOW_PARALLAX = 4,     /* "Low parallax."                              .cc:633-635 */
  NRSLAM_B200_TRI_NO_NEIGHBOURS = 5,    /* "Found no neighbours in a temporal point."   .cc:653-655 */
  NRSLAM_B200_TRI_NEGATIVE_DEPTH = 6,   /* "Negative initial depth."                    .cc:659-661 */
  NRSLAM_B200_TRI_BAD_NEIGHBOURS = 7,   /* "Triangulation has to many bad neighbors."   .cc:781-783 */
  NRSLAM_B200_TRI_HIGH_ERROR = 8,       /* "Triangulation has to much error."           .cc:794-796 */
  NRSLAM_B200_TRI_NAN = 9,              /* result.hasNaN() (the caller's check, mapping.cc:98-99) */
  NRSLAM_B200_TRI_SHORT_TRACK = 10,     /* "Short track" (TrackLenght < 5, mapping.cc:94,111-113)        tri_run_frame only */
  NRSLAM_B200_TRI_NOT_RIGID = 11,       /* "Rigidity not detected" (mapping.cc:122-125)                  rigid branch */
  NRSLAM_B200_TRI_RIGID_PARALLAX = 12   /* "Parallax error." (parallax window, depth, reprojection; mapping.cc:152-181) */
};
typedef struct nrslam_b200_tri nrslam_b200_tri;
int nrslam_b200_tri_create(nrslam_b200_ctx* ctx, nrslam_b200_tri** out);
void nrslam_b200_tri_destroy(nrslam_b200_tri* tri);
int nrslam_b200_tri_run(nrslam_b200_tri* tri, const nrslam_b200_camera* cam, int32_t n_cand,
                        const int32_t* track_ptr, const float* track_uv, const float* track_pose,
                        const int32_t* n_neighbours, const float* nb_pos, const uint8_t* nb_valid, float scale,
                        float* position_out, int32_t* status_out, int32_t* lm_iterations_out);
/* The whole per-candidate compute of Mapping::LandmarkTriangulation (mapping/mapping.cc:65-212) in one launch: the
 * deformable triangulation above for candidates with track length >= min_track (5 in the reference, :94; shorter ones
 * get NRSLAM_B200_TRI_SHORT_TRACK), the RIGID mid-point triangulation of every candidate (:115-185; rigid_ok[c] =
 * TemporalBuffer::CheckRigidity(first frame, last frame, 0.004) of its track, rad_per_pixel = Mapping::Options) and the
 * reference's vote (:188-212): rigid results are used when n_rigid > 1.5 n_deformable, deformable ones when
 * n_deformable >= 1.5 n_rigid, none otherwise; NaN positions are dropped. selected_out[c] = 1 when candidate c gets a
 * map point at selected_pos_out[3c..]. All output arrays are required. */
int nrslam_b200_tri_run_frame(nrslam_b200_tri* tri, const nrslam_b200_camera* cam, int32_t n_cand,
                              const int32_t* track_ptr, const float* track_uv, const float* track_pose,
                              const int32_t* n_neighbours, const float* nb_pos, const uint8_t* nb_valid,
                              const uint8_t* rigid_ok, float rad_per_pixel, int32_t min_track, float scale,
                              float* deform_pos_out, int32_t* deform_status_out, float* rigid_pos_out,
                              int32_t* rigid_status_out, float* selected_pos_out, uint8_t* selected_out);
/* Device time (CUDA events on the ctx stream) of the last run's kernel, and a re-run of it on the HBM-resident
 * staged batch (benchmark hook; results identical). */
float nrslam_b200_tri_last_ms(const nrslam_b200_tri* tri);
int nrslam_b200_tri_rerun(nrslam_b200_tri* tri, float* gpu_ms_out);

/* ---- RegularizationGraph::UpdateVertex for all updated vertices of a frame, on the device ----
 * The loop CameraPoseAndDeformationOptimization runs after the solve (g2o_optimization.cc:458-474):
 *   for every accepted point i (in frame order): good = graph->UpdateVertex(id_i, positions); if (good < 5) BAD
 * UpdateVertex (map/regularization_graph.cc:130-146) walks the vertex's edges, and UpdateConnection (:107-128)
 * updates min/max distance, weight = exp(-d_max^2 / 2 sigma^2) and the BAD status of one edge from the CURRENT
 * positions of its endpoints. Positions do not change during the loop, so the per-edge update is idempotent and the
 * loop is order-independent: one thread per undirected edge incident to an updated vertex, then one thread per
 * updated vertex counting its non-BAD edges. Statuses, counts and min / max distances are bit-exact with the
 * sequential host version (nrslam_b200_graph_update_vertex); the weight is exp() evaluated in fp64 and rounded to
 * fp32, i.e. the correctly rounded expf, which glibc's faithfully rounded expf misses by 1 ulp for 0.07 % of the
 * arguments. vertices[n] = graph vertices to update (each at most once);
 * positions[3 n_vertices] = MapPoint::GetLastWorldPosition of all graph vertices; good_out[n] = UpdateVertex's
 * return value. The graph attribute arrays (weight, min/max distance, status) are updated in place. */
int nrslam_b200_graph_update_vertices(nrslam_b200_ctx* ctx, nrslam_b200_graph* g, int32_t n,
                                      const int32_t* vertices, const float* positions, int32_t* good_out);

/* ---- RegularizationGraph::GetEdges for many vertices at once, on the device -------------------
 * GetEdges (map/regularization_graph.cc:71-87) copies ALL connections of a vertex, sorts them with EdgeComparator
 * (:61-69: status ascending, weight descending; ties broken by ascending neighbour like nrslam_b200_graph_get_edges)
 * and keeps the prefix before the first weight < min_weight — the reference's O(N^2 log N) per frame,
 * because CameraPoseAndDeformationOptimization calls it for every point (:251-336) but reads only the first
 * ~regularizers_per_point entries that pass its filters. Here: one CTA per listed vertex ranks the row's entries by a
 * 64-bit key (2 bits status | 32 bits inverted weight | 30 bits neighbour) — a segmented top-k without a full sort.
 * out_entries[n * top_k]: the first min(count, top_k) CSR entry indices of the sorted list per vertex (-1 padded);
 * out_count[n]: GetEdges(v).size(). Bit-exact with nrslam_b200_graph_get_edges. */
int nrslam_b200_graph_get_edges_batch(nrslam_b200_ctx* ctx, const nrslam_b200_graph* g, int32_t n,
                                      const int32_t* vertices, int32_t top_k, int32_t* out_entries,
                                      int32_t* out_count);

#ifdef __cplusplus
}
#endif
#endif /* NRSLAM_B200_H */
