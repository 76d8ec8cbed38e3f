// TEST INFRASTRUCTURE: bodies of the stand-in NR-SLAM containers (shim/standin/nrslam_standin.h) used by the shim test.
#include "nrslam_standin.h"

template <class C>
static std::vector<int> Filter(const std::vector<LandmarkStatus>& st, const C& wanted) {
  std::vector<int> idx;
  for (size_t i = 0; i < st.size(); i++)
    if (wanted.count(st[i])) idx.push_back((int)i);
  return idx;
}
#define FILTERED(T, member)                                                \
  std::vector<T> out;                                                      \
  for (int i : Filter(landmark_status_, st)) out.push_back(member[i]);    \
  return out;
std::vector<cv::KeyPoint> Frame::GetKeypointsWithStatus(const absl::flat_hash_set<LandmarkStatus> st) const { FILTERED(cv::KeyPoint, keypoints_) }
std::vector<Eigen::Vector3f> Frame::GetLandmarkPositionsWithStatus(const absl::flat_hash_set<LandmarkStatus> st) const { FILTERED(Eigen::Vector3f, landmark_positions_) }
std::vector<ID> Frame::GetMapPointsIdsWithStatus(const absl::flat_hash_set<LandmarkStatus> st) { FILTERED(ID, ids_) }
void Frame::InsertObservation(const cv::KeyPoint& kp, const Eigen::Vector3f& X, const ID mp, const LandmarkStatus s) {
  mappoint_id_to_index_[mp] = (int)keypoints_.size();
  keypoints_.push_back(kp); landmark_positions_.push_back(X); landmark_status_.push_back(s); ids_.push_back(mp);
}
std::vector<cv::KeyPoint> KeyFrame::GetKeypointsWithStatus(const absl::flat_hash_set<LandmarkStatus> st) { FILTERED(cv::KeyPoint, keypoints_) }
std::vector<Eigen::Vector3f> KeyFrame::GetLandmarkPositionsWithStatus(const absl::flat_hash_set<LandmarkStatus> st) { FILTERED(Eigen::Vector3f, landmark_positions_) }
std::vector<ID> KeyFrame::GetMapPointsIdsWithStatus(const absl::flat_hash_set<LandmarkStatus> st) { FILTERED(ID, ids_) }
void KeyFrame::Insert(const cv::KeyPoint& kp, const Eigen::Vector3f& X, ID mp, LandmarkStatus s) {
  mappoint_id_to_index_[mp] = (int)keypoints_.size();
  keypoints_.push_back(kp); landmark_positions_.push_back(X); landmark_status_.push_back(s); ids_.push_back(mp);
}
void RegularizationGraph::AddEdge(ID a, ID b, Eigen::Vector3f& rel) {
  const float d = std::sqrt(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]);
  auto e = std::make_shared<Edge>();
  e->vertex_id_1 = a; e->vertex_id_2 = b; e->distance = e->first_distance = e->max_distance = e->min_distance = d;
  e->weight = std::exp(-(d * d) / (2 * options_.weight_sigma * options_.weight_sigma));
  e->status = NEUTRAL; e->last_relative_position = rel;
  graph_[a][b] = e; graph_[b][a] = e;
}
std::vector<std::pair<ID, cv::KeyPoint>> TemporalBuffer::GetFeatureTrack(const int id) { return tracks[id]; }
std::vector<int> TemporalBuffer::GetClosestMapPointsToFeature(const int id, const int, const int, const int) { return neighbours[id]; }
absl::StatusOr<Sophus::SE3f> TemporalBuffer::GetCameraTransformWorld(const int f) {
  auto it = poses.find(f);
  if (it == poses.end()) return absl::InternalError("no pose");
  return it->second;
}
absl::StatusOr<Eigen::Vector3f> TemporalBuffer::GetLandmarkPosition(const int f, const int k) {
  auto it = positions.find({f, k});
  if (it == positions.end()) return absl::InternalError("no position");
  return it->second;
}
