// shim/g2o_optimization_b200.cc — drop-in replacement for modules/optimization/g2o_optimization.cc of NR-SLAM.
//
// Keeps the four signatures of modules/optimization/g2o_optimization.h:27-40, marshals Frame / Map / KeyFrame /
// TemporalBuffer state into the flat host buffers of include/nrslam_b200.h, calls the C ABI of libnrslam_b200.so and
// writes the results back exactly where the reference writes them, so modules/tracking and modules/mapping compile
// unchanged (callers: tracking.cc:316,321; mapping.cc:57,260). g2o is no longer needed by this module.
//
// Build (modules/CMakeLists.txt): replace optimization/g2o_optimization.cc by this file, add include/ to the include
// path and link libnrslam_b200.so. The only header change in the reference tree is ONE accessor on RegularizationGraph
// (`const absl::btree_map<ID, VertexConnections>& Connections() const { return graph_; }`): GetEdges filters by
// weight (regularization_graph.cc:71-87) and would hide the far neighbours UpdateVertex counts (:130-146).
//
// In this repository the file is compiled by the CPU test suite against the stand-in headers of shim/standin/
// (tests/test_shim.py): the image has no Eigen / Sophus / OpenCV / abseil headers.
#include "optimization/g2o_optimization.h"

#include <algorithm>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "nrslam_b200.h"

namespace {

// One context per process: the reference calls the optimisation from one thread, strictly sequentially
// (SLAM/system.cc:125-128).
nrslam_b200_ctx* Ctx() {
  static nrslam_b200_ctx* ctx = [] {
    nrslam_b200_ctx* c = nullptr;
    nrslam_b200_create(nullptr, &c);
    return c;
  }();
  return ctx;
}

nrslam_b200_tri* Tri() {
  static nrslam_b200_tri* tri = [] {
    nrslam_b200_tri* t = nullptr;
    if (Ctx()) nrslam_b200_tri_create(Ctx(), &t);
    return t;
  }();
  return tri;
}

// PinHole has 4 parameters, KannalaBrandt8 has 8 (calibration/pin_hole.cc, kannala_brandt_8.cc)
nrslam_b200_camera ToCamera(const std::shared_ptr<CameraModel>& calibration) {
  nrslam_b200_camera cam{};
  const int n = calibration->getNumberOfParameters();
  cam.model = n == 4 ? 0 : 1;
  for (int i = 0; i < n && i < 8; i++) cam.params[i] = calibration->GetParameter(i);
  return cam;
}

void ToPose7(const Sophus::SE3f& T, float* p) {  // [qx qy qz qw tx ty tz]
  const auto q = T.unit_quaternion();
  const auto t = T.translation();
  p[0] = q.x(); p[1] = q.y(); p[2] = q.z(); p[3] = q.w();
  p[4] = t.x(); p[5] = t.y(); p[6] = t.z();
}

Sophus::SE3f FromPose7(const float* p) {
  return Sophus::SE3f(Eigen::Quaternionf(p[3], p[0], p[1], p[2]), Eigen::Vector3f(p[4], p[5], p[6]));
}

// The regularisation graph as the CSR of nrslam_b200_graph: vertices = map points in ascending id (the btree order
// of regularization_graph.h:89), one attribute record per shared Edge object (regularization_graph.cc:53-54).
struct GraphCsr {
  std::vector<ID> ids;                          // vertex -> map point id
  std::unordered_map<ID, int32_t> vertex;       // map point id -> vertex
  std::vector<int32_t> rowptr, col, eid;
  std::vector<float> weight, first_distance, min_distance, max_distance;
  std::vector<uint8_t> status;
  std::vector<std::shared_ptr<RegularizationGraph::Edge>> edges;  // record -> the reference's Edge object
  nrslam_b200_graph g{};

  int32_t VertexOf(ID id) const {
    auto it = vertex.find(id);
    return it == vertex.end() ? -1 : it->second;
  }

  void Build(RegularizationGraph& graph, float weight_sigma, float stretching_th) {
    const auto& conn = graph.Connections();
    ids.clear();
    vertex.clear();
    for (const auto& row : conn) {  // ascending id
      vertex[row.first] = (int32_t)ids.size();
      ids.push_back(row.first);
    }
    rowptr.assign(ids.size() + 1, 0);
    col.clear();
    eid.clear();
    edges.clear();
    std::unordered_map<const RegularizationGraph::Edge*, int32_t> record;
    int32_t v = 0;
    for (const auto& row : conn) {
      for (const auto& nb : row.second) {  // ascending neighbour id
        auto jt = vertex.find(nb.first);
        if (jt == vertex.end()) continue;  // neighbour without a row of its own: not addressable
        auto rt = record.find(nb.second.get());
        int32_t e;
        if (rt == record.end()) {
          e = (int32_t)edges.size();
          record[nb.second.get()] = e;
          edges.push_back(nb.second);
        } else {
          e = rt->second;
        }
        col.push_back(jt->second);
        eid.push_back(e);
      }
      rowptr[++v] = (int32_t)col.size();
    }
    const size_t E = edges.size();
    weight.resize(E); first_distance.resize(E); min_distance.resize(E); max_distance.resize(E); status.resize(E);
    for (size_t e = 0; e < E; e++) {
      weight[e] = edges[e]->weight;
      first_distance[e] = edges[e]->first_distance;
      min_distance[e] = edges[e]->min_distance;
      max_distance[e] = edges[e]->max_distance;
      status[e] = (uint8_t)edges[e]->status;  // VERIFIED, NEIGHBOR, NEUTRAL, BAD = NRSLAM_EDGE_* 0..3
    }
    g.n_vertices = (int32_t)ids.size();
    g.n_edges = (int32_t)E;
    g.rowptr = rowptr.data();
    g.col = col.data();
    g.eid = eid.data();
    g.weight = weight.data();
    g.first_distance = first_distance.data();
    g.min_distance = min_distance.data();
    g.max_distance = max_distance.data();
    g.status = status.data();
    g.weight_sigma = weight_sigma;
    g.stretching_th = stretching_th;
  }

  // What RegularizationGraph::UpdateVertex changed (regularization_graph.cc:107-123), back into the Edge objects
  void WriteBack() {
    for (size_t e = 0; e < edges.size(); e++) {
      edges[e]->weight = weight[e];
      edges[e]->min_distance = min_distance[e];
      edges[e]->max_distance = max_distance[e];
      edges[e]->status = (RegularizationGraph::Status)status[e];
    }
  }
};

GraphCsr BuildCsr(Map& map) {
  GraphCsr csr;
  auto graph = map.GetRegularizationGraph();
  const auto opt = graph->GetOptions();
  csr.Build(*graph, opt.weight_sigma, opt.streching_th);
  return csr;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// g2o_optimization.cc:50-146
// ---------------------------------------------------------------------------------------------------------------
void CameraPoseOptimization(Frame& current_frame, const Sophus::SE3f& /*previous_camera_transform_world: unused*/) {
  const auto keypoints = current_frame.GetKeypointsWithStatus({TRACKED_WITH_3D});  // :75-77
  const auto positions = current_frame.GetLandmarkPositionsWithStatus({TRACKED_WITH_3D});
  const int n = (int)keypoints.size();
  std::vector<float> uv(2 * (size_t)n), X(3 * (size_t)n);
  for (int i = 0; i < n; i++) {
    uv[2 * i] = keypoints[i].pt.x;
    uv[2 * i + 1] = keypoints[i].pt.y;
    for (int k = 0; k < 3; k++) X[3 * i + k] = positions[i][k];
  }
  float pose[7];
  ToPose7(current_frame.CameraTransformationWorld(), pose);
  const nrslam_b200_camera cam = ToCamera(current_frame.GetCalibration());
  if (nrslam_b200_pose_only(Ctx(), &cam, n, uv.data(), X.data(), pose, nullptr, nullptr) == NRSLAM_B200_OK)
    current_frame.MutableCameraTransformationWorld() = FromPose7(pose);  // :144-145
}

// ---------------------------------------------------------------------------------------------------------------
// g2o_optimization.cc:148-557
// ---------------------------------------------------------------------------------------------------------------
absl::flat_hash_set<ID> CameraPoseAndDeformationOptimization(Frame& current_frame, std::shared_ptr<Map> map,
                                                             const Sophus::SE3f& /*unused*/, const float scale) {
  absl::flat_hash_set<ID> lost_ids;
  // 1. graph -> CSR. A production build keeps the CSR alive across frames and only appends (AddEdge) / refreshes
  //    attributes; it is rebuilt per call here (O(edges)).
  GraphCsr csr = BuildCsr(*map);
  // 2. frame -> SoA in Frame::Get*WithStatus({TRACKED_WITH_3D}) order (frame.cc:83-118,182-193)
  const auto keypoints = current_frame.GetKeypointsWithStatus({TRACKED_WITH_3D});  // :174-176
  const auto positions = current_frame.GetLandmarkPositionsWithStatus({TRACKED_WITH_3D});
  const auto ids = current_frame.GetMapPointsIdsWithStatus({TRACKED_WITH_3D});
  const int n = (int)ids.size(), M = csr.g.n_vertices;
  std::vector<float> uv(2 * (size_t)n), X(3 * (size_t)n), X_out(3 * (size_t)n), last(3 * (size_t)M), last_in;
  std::vector<int32_t> point_vertex(n), lost(M > 0 ? M : 1);
  std::vector<int8_t> vertex_frame_status(M, -1);  // drives the "lost" classification (:264-273)
  std::vector<uint8_t> status(n);
  for (int i = 0; i < n; i++) {
    uv[2 * i] = keypoints[i].pt.x;
    uv[2 * i + 1] = keypoints[i].pt.y;
    for (int k = 0; k < 3; k++) X[3 * i + k] = positions[i][k];
    point_vertex[i] = csr.VertexOf(ids[i]);
    if (point_vertex[i] < 0) return lost_ids;  // a tracked map point without a graph vertex: the reference aborts
  }
  for (const auto& entry : current_frame.MapPointIdToIndex()) {
    const int32_t v = csr.VertexOf(entry.first);
    if (v >= 0) vertex_frame_status[v] = (int8_t)current_frame.LandmarkStatuses()[entry.second];
  }
  for (int v = 0; v < M; v++) {
    const Eigen::Vector3f p = map->GetMapPoint(csr.ids[v])->GetLastWorldPosition();
    for (int k = 0; k < 3; k++) last[3 * (size_t)v + k] = p[k];
  }
  last_in = last;
  float pose[7], median = 0;
  int32_t n_lost = 0;
  ToPose7(current_frame.CameraTransformationWorld(), pose);
  const nrslam_b200_camera cam = ToCamera(current_frame.GetCalibration());
  if (nrslam_b200_pose_deform(Ctx(), &cam, n, uv.data(), X.data(), point_vertex.data(), vertex_frame_status.data(),
                              &csr.g, scale, pose, last.data(), nullptr, X_out.data(), nullptr, status.data(), &median,
                              lost.data(), &n_lost, nullptr) != NRSLAM_B200_OK)
    return lost_ids;  // like a run whose every trial was rejected: nothing is written
  // 3. write back what the reference writes (:398-399, 428-455, 470-473, 544-552)
  current_frame.MutableCameraTransformationWorld() = FromPose7(pose);
  for (int i = 0; i < n; i++) {
    const int index = current_frame.MapPointIdToIndex().at(ids[i]);
    current_frame.LandmarkStatuses()[index] = (LandmarkStatus)status[i];
    current_frame.LandmarkPositions()[index] = Eigen::Vector3f(X_out[3 * i], X_out[3 * i + 1], X_out[3 * i + 2]);
  }
  current_frame.SetDeformationMaginitud(median);
  csr.WriteBack();
  for (int v = 0; v < M; v++) {  // MapPoint::SetLastWorldPosition for accepted and lost points (:446,550)
    if (last[3 * (size_t)v] == last_in[3 * (size_t)v] && last[3 * (size_t)v + 1] == last_in[3 * (size_t)v + 1] &&
        last[3 * (size_t)v + 2] == last_in[3 * (size_t)v + 2])
      continue;
    Eigen::Vector3f p(last[3 * (size_t)v], last[3 * (size_t)v + 1], last[3 * (size_t)v + 2]);
    map->GetMapPoint(csr.ids[v])->SetLastWorldPosition(p);
  }
  for (int k = 0; k < n_lost; k++) lost_ids.insert(csr.ids[lost[k]]);
  return lost_ids;
}

// ---------------------------------------------------------------------------------------------------------------
// g2o_optimization.cc:559-814 — one candidate; Mapping::LandmarkTriangulation should batch its candidates through
// nrslam_b200_tri_run_frame instead, this signature stays for other callers.
// ---------------------------------------------------------------------------------------------------------------
absl::StatusOr<Eigen::Vector3f> DeformableTriangulation(TemporalBuffer& temporal_buffer, int candidate_id,
                                                        std::shared_ptr<CameraModel> calibration, const float scale) {
  static const char* const kMessage[] = {"", "Feature too close to other ones.", "High reprojection error at first camera.",
                                         "High reprojection error at second camera.", "Low parallax.",
                                         "Found no neighbours in a temporal point.", "Negative initial depth.",
                                         "Triangulation has to many bad neighbors.", "Triangulation has to much error.",
                                         "NaN", "Short track", "Rigidity not detected", "Parallax error."};
  constexpr int NB = NRSLAM_B200_TRI_MAX_NB;
  const auto track = temporal_buffer.GetFeatureTrack(candidate_id);                               // :564
  const auto nbrs = temporal_buffer.GetClosestMapPointsToFeature(candidate_id, 10, 20, 500);     // :567
  const int T = (int)track.size();
  if (T < 1 || T > NRSLAM_B200_TRI_MAX_TRACK || !Tri()) return absl::InternalError("Track length out of range.");
  std::vector<int32_t> track_ptr{0, T};
  std::vector<float> uv(2 * (size_t)T), pose(7 * (size_t)T), nb_pos((size_t)T * NB * 3, 0.f);
  std::vector<uint8_t> nb_valid((size_t)T * NB, 0);
  int32_t n_nb = (int32_t)std::min<size_t>(nbrs.size(), NB);
  for (int f = 0; f < T; f++) {
    uv[2 * f] = track[f].second.pt.x;
    uv[2 * f + 1] = track[f].second.pt.y;
    const auto Tcw = temporal_buffer.GetCameraTransformWorld((int)track[f].first);
    if (!Tcw.ok()) return absl::InternalError("Missing camera pose.");
    ToPose7(*Tcw, &pose[7 * (size_t)f]);
    for (int k = 0; k < n_nb; k++) {
      const auto p = temporal_buffer.GetLandmarkPosition((int)track[f].first, nbrs[k]);
      nb_valid[(size_t)f * NB + k] = p.ok() ? 1 : 0;
      if (p.ok())
        for (int a = 0; a < 3; a++) nb_pos[((size_t)f * NB + k) * 3 + a] = (*p)[a];
    }
  }
  float X[3] = {0, 0, 0};
  int32_t status = 0;
  const nrslam_b200_camera cam = ToCamera(calibration);
  if (nrslam_b200_tri_run(Tri(), &cam, 1, track_ptr.data(), uv.data(), pose.data(), &n_nb, nb_pos.data(),
                          nb_valid.data(), scale, X, &status, nullptr) != NRSLAM_B200_OK)
    return absl::InternalError("nrslam_b200_tri_run failed.");
  if (status != NRSLAM_B200_TRI_OK) return absl::InternalError(kMessage[status >= 0 && status <= 12 ? status : 0]);
  return Eigen::Vector3f(X[0], X[1], X[2]);
}

// ---------------------------------------------------------------------------------------------------------------
// g2o_optimization.cc:880-1161
// ---------------------------------------------------------------------------------------------------------------
void LocalDeformableBundleAdjustment(std::shared_ptr<Map> map, const float scale) {
  auto keyframes = map->GetKeyFrames();
  const size_t max_keyframes_in_optimization = 5;  // :894
  std::vector<std::shared_ptr<KeyFrame>> window;   // newest first, like keyframes_in_optimization (:899-915)
  for (auto it = keyframes.rbegin(); it != keyframes.rend() && window.size() < max_keyframes_in_optimization; ++it)
    window.push_back(it->second);
  if (window.size() < 3) return;  // :922-924
  std::reverse(window.begin(), window.end());  // the C ABI takes the window OLDEST FIRST (:930,982 walk rbegin())
  GraphCsr csr = BuildCsr(*map);
  const int F = (int)window.size();
  std::vector<float> kf_pose(7 * (size_t)F), uv, X;
  std::vector<int32_t> obs_kf, obs_vertex;
  std::vector<ID> obs_id;
  for (int k = 0; k < F; k++) {
    ToPose7(window[k]->CameraTransformationWorld(), &kf_pose[7 * (size_t)k]);
    const auto keypoints = window[k]->GetKeypointsWithStatus({TRACKED_WITH_3D});  // :1000-1004
    const auto positions = window[k]->GetLandmarkPositionsWithStatus({TRACKED_WITH_3D});
    const auto ids = window[k]->GetMapPointsIdsWithStatus({TRACKED_WITH_3D});
    for (size_t i = 0; i < ids.size(); i++) {
      const int32_t v = csr.VertexOf(ids[i]);
      if (v < 0) continue;  // no graph vertex: the point has no regulariser and the reference's GetEdges would abort
      obs_kf.push_back(k);
      obs_vertex.push_back(v);
      obs_id.push_back(ids[i]);
      uv.push_back(keypoints[i].pt.x);
      uv.push_back(keypoints[i].pt.y);
      for (int a = 0; a < 3; a++) X.push_back(positions[i][a]);
    }
  }
  const nrslam_b200_camera cam = ToCamera(window[0]->GetCalibration());
  if (nrslam_b200_local_ba(Ctx(), &cam, F, kf_pose.data(), (int32_t)obs_kf.size(), obs_kf.data(), obs_vertex.data(),
                           uv.data(), X.data(), &csr.g, scale, /*iterations: options.ba_iterations*/ 0,
                           nullptr) != NRSLAM_B200_OK)
    return;
  for (int k = 0; k < F; k++) window[k]->CameraTransformationWorld() = FromPose7(&kf_pose[7 * (size_t)k]);  // :1148-1151
  for (size_t o = 0; o < obs_kf.size(); o++) {                                                                 // :1153-1159
    auto& kf = window[obs_kf[o]];
    const int idx_in_keyframe = kf->MapPointIdToIndex().at(obs_id[o]);
    kf->LandmarkPositions()[idx_in_keyframe] = Eigen::Vector3f(X[3 * o], X[3 * o + 1], X[3 * o + 2]);
  }
}
