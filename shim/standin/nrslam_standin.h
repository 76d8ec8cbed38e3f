// STAND-IN declarations of the NR-SLAM types the shim touches — TEST INFRASTRUCTURE ONLY.
// The image has no Eigen / Sophus / OpenCV / abseil headers, so the shim sources (shim/*.cc) are compiled in the CPU
// test suite against this minimal re-declaration of the public interface they use (same class, method and member
// names as the reference headers: modules/map/frame.h, keyframe.h, map.h, mappoint.h, regularization_graph.h,
// temporal_buffer.h, calibration/camera_model.h, matching/lucas_kanade_tracker.h, utilities/landmark_status.h).
// The bodies are simple containers written for the test driver; nothing here ships in libnrslam_b200.so.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace Eigen {
struct Vector3f {
  float v[3] = {0, 0, 0};
  Vector3f() {}
  Vector3f(float a, float b, float c) { v[0] = a; v[1] = b; v[2] = c; }
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float z() const { return v[2]; }
  bool hasNaN() const { return std::isnan(v[0]) || std::isnan(v[1]) || std::isnan(v[2]); }
};
struct Quaternionf {
  float q[4] = {0, 0, 0, 1};  // x y z w
  Quaternionf() {}
  Quaternionf(float w, float x, float y, float z) { q[0] = x; q[1] = y; q[2] = z; q[3] = w; }
  float x() const { return q[0]; }
  float y() const { return q[1]; }
  float z() const { return q[2]; }
  float w() const { return q[3]; }
};
}  // namespace Eigen

namespace Sophus {
class SE3f {
 public:
  SE3f() {}
  SE3f(const Eigen::Quaternionf& q, const Eigen::Vector3f& t) : q_(q), t_(t) {}
  Eigen::Quaternionf unit_quaternion() const { return q_; }
  Eigen::Vector3f translation() const { return t_; }
 private:
  Eigen::Quaternionf q_;
  Eigen::Vector3f t_;
};
}  // namespace Sophus

namespace cv {
struct Point2f {
  float x = 0, y = 0;
};
struct KeyPoint {
  Point2f pt;
  int class_id = -1;
};
struct Size {
  int width = 0, height = 0;
  Size() {}
  Size(int w, int h) : width(w), height(h) {}
};
// just enough of cv::Mat for 8-bit single-channel images and 16-bit patches
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;  // bytes per row
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int elem_bytes) : rows(r), cols(c), step((size_t)c * elem_bytes), store_((size_t)r * c * elem_bytes) {
    data = store_.data();
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  template <typename T>
  T* ptr(int r = 0) { return reinterpret_cast<T*>(data + step * r); }
  template <typename T>
  const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + step * r); }
  Mat(const Mat& o) : rows(o.rows), cols(o.cols), step(o.step), store_(o.store_) { data = store_.empty() ? nullptr : store_.data(); }
  Mat& operator=(const Mat& o) {
    rows = o.rows; cols = o.cols; step = o.step; store_ = o.store_;
    data = store_.empty() ? nullptr : store_.data();
    return *this;
  }
 private:
  std::vector<unsigned char> store_;
};
}  // namespace cv

namespace absl {
template <class K, class V> using flat_hash_map = std::unordered_map<K, V>;
template <class K> using flat_hash_set = std::unordered_set<K>;
template <class K, class V> using btree_map = std::map<K, V>;
struct Status {
  bool ok_ = true;
  std::string msg;
  bool ok() const { return ok_; }
  std::string message() const { return msg; }
};
inline Status InternalError(const std::string& m) { Status s; s.ok_ = false; s.msg = m; return s; }
template <class T>
class StatusOr {
 public:
  StatusOr(const T& v) : v_(v) {}
  StatusOr(const Status& s) : s_(s) {}
  bool ok() const { return s_.ok(); }
  const T& operator*() const { return v_; }
  const T& value() const { return v_; }
  Status status() const { return s_; }
 private:
  T v_{};
  Status s_;
};
}  // namespace absl

typedef long unsigned int ID;

enum LandmarkStatus { TRACKED_WITH_3D, TRACKED, JUST_TRIANGULATED, BAD, OUT_IMAGE_BOUNDARIES, BAD_FEATURE };

class CameraModel {
 public:
  explicit CameraModel(const std::vector<float>& p) : calibration_parameters_(p) {}
  float GetParameter(const int i) { return calibration_parameters_[i]; }
  int getNumberOfParameters() { return (int)calibration_parameters_.size(); }
 private:
  std::vector<float> calibration_parameters_;
};

class LucasKanadeTracker {
 public:
  struct PhotometricInformation {
    std::vector<float> mean_gray_per_level;
    std::vector<float> squared_mean_gray_per_level;
    std::vector<cv::Mat> gray_reference;
    std::vector<cv::Mat> gradient_reference;
  };
  LucasKanadeTracker();
  LucasKanadeTracker(const cv::Size _winSize, const int _maxLevel, const int _maxIters, const float _epsilon,
                     const float _minEigThreshold);
  ~LucasKanadeTracker();
  void SetReferenceImage(const cv::Mat& refIm, const std::vector<cv::KeyPoint>& refPts, const cv::Mat& mask = cv::Mat());
  int Track(const cv::Mat& newIm, std::vector<cv::KeyPoint>& nextPts, std::vector<LandmarkStatus>& vMatched,
            const bool bInitialFlow, const float minSSIM, const cv::Mat& mask);
  PhotometricInformation GetPhotometricInformationOfPoint(const int idx);
  void InsertPhotometricInformation(cv::KeyPoint& keypoint, PhotometricInformation& photometric_information);
  void clear();
  cv::Size winSize_;
  int maxLevel_ = 4, maxIters_ = 10;
  float epsilon_ = 1e-4f, minEigThreshold_ = 1e-4f;
  std::vector<cv::KeyPoint> prevPts_;
  void* b200_ = nullptr;  // the one member the B200 build adds: the device-side tracker (nrslam_b200_klt*)
};

class Frame;

class KeyFrame {
 public:
  explicit KeyFrame(ID id) : id_(id) {}
  std::vector<cv::KeyPoint> GetKeypointsWithStatus(const absl::flat_hash_set<LandmarkStatus> st);
  std::vector<Eigen::Vector3f>& LandmarkPositions() { return landmark_positions_; }
  std::vector<Eigen::Vector3f> GetLandmarkPositionsWithStatus(const absl::flat_hash_set<LandmarkStatus> st);
  std::vector<LandmarkStatus>& LandmarkStatuses() { return landmark_status_; }
  Sophus::SE3f& CameraTransformationWorld() { return pose_; }
  const absl::flat_hash_map<ID, int>& MapPointIdToIndex() const { return mappoint_id_to_index_; }
  std::shared_ptr<CameraModel> GetCalibration() { return calibration_; }
  std::vector<ID> GetMapPointsIdsWithStatus(const absl::flat_hash_set<LandmarkStatus> st);
  long unsigned int GetId() { return id_; }
  // test-driver helpers
  void Insert(const cv::KeyPoint& kp, const Eigen::Vector3f& X, ID mp, LandmarkStatus s);
  std::shared_ptr<CameraModel> calibration_;
 private:
  std::vector<cv::KeyPoint> keypoints_;
  std::vector<Eigen::Vector3f> landmark_positions_;
  std::vector<LandmarkStatus> landmark_status_;
  std::vector<ID> ids_;
  absl::flat_hash_map<ID, int> mappoint_id_to_index_;
  Sophus::SE3f pose_;
  ID id_;
};

class Frame {
 public:
  std::vector<cv::KeyPoint> GetKeypointsWithStatus(const absl::flat_hash_set<LandmarkStatus> st) const;
  std::vector<Eigen::Vector3f>& LandmarkPositions() { return landmark_positions_; }
  std::vector<Eigen::Vector3f> GetLandmarkPositionsWithStatus(const absl::flat_hash_set<LandmarkStatus> st) const;
  std::vector<LandmarkStatus>& LandmarkStatuses() { return landmark_status_; }
  void InsertObservation(const cv::KeyPoint& kp, const Eigen::Vector3f& X, const ID mp, const LandmarkStatus s);
  Sophus::SE3f& MutableCameraTransformationWorld() { return pose_; }
  Sophus::SE3f CameraTransformationWorld() const { return pose_; }
  const absl::flat_hash_map<ID, int>& MapPointIdToIndex() const { return mappoint_id_to_index_; }
  void SetCalibration(std::shared_ptr<CameraModel> c) { calibration_ = c; }
  std::shared_ptr<CameraModel> GetCalibration() { return calibration_; }
  std::vector<ID> GetMapPointsIdsWithStatus(const absl::flat_hash_set<LandmarkStatus> st);
  void SetDeformationMaginitud(const float m) { median_deformation_magnitud_ = m; }
  float GetDeformationMagnitud() { return median_deformation_magnitud_; }
 private:
  std::vector<cv::KeyPoint> keypoints_;
  std::vector<Eigen::Vector3f> landmark_positions_;
  std::vector<LandmarkStatus> landmark_status_;
  std::vector<ID> ids_;
  absl::flat_hash_map<ID, int> mappoint_id_to_index_;
  Sophus::SE3f pose_;
  std::shared_ptr<CameraModel> calibration_;
  float median_deformation_magnitud_ = 0;
};

class MapPoint {
 public:
  MapPoint(const Eigen::Vector3f& p, ID id) : last_(p), id_(id) {}
  Eigen::Vector3f GetLastWorldPosition() { return last_; }
  void SetLastWorldPosition(Eigen::Vector3f& p) { last_ = p; n_set_++; }
  long unsigned int GetId() { return id_; }
  int n_set_ = 0;  // test-driver helper: how often the position was written
 private:
  Eigen::Vector3f last_;
  ID id_;
};

class Map;

class RegularizationGraph {
 public:
  struct Options {
    float weight_sigma;
    float streching_th;
  };
  enum Status { VERIFIED, NEIGHBOR, NEUTRAL, BAD };
  struct Edge {
    ID vertex_id_1, vertex_id_2;
    float distance, first_distance, weight;
    Status status;
    float max_distance, min_distance;
    Eigen::Vector3f last_relative_position;
  };
  typedef absl::btree_map<ID, std::shared_ptr<Edge>> VertexConnections;
  RegularizationGraph(Options& o, Map* m) : options_(o), map_(m) {}
  void AddEdge(ID a, ID b, Eigen::Vector3f& rel);
  Options GetOptions() const { return options_; }
  // THE ONE ACCESSOR THE INTEGRATION ADDS to modules/map/regularization_graph.h (GetEdges filters by weight and
  // hides the far neighbours UpdateVertex counts, regularization_graph.cc:71-87,130-146):
  const absl::btree_map<ID, VertexConnections>& Connections() const { return graph_; }
 private:
  absl::btree_map<ID, VertexConnections> graph_;
  Options options_;
  Map* map_;
};

class TemporalBuffer {
 public:
  std::vector<std::pair<ID, cv::KeyPoint>> GetFeatureTrack(const int keypoint_id);
  std::vector<int> GetClosestMapPointsToFeature(const int keypoint_id, const int num_neighbors,
                                                const int min_image_distance, const int max_image_distance);
  absl::StatusOr<Sophus::SE3f> GetCameraTransformWorld(const int frame_id);
  absl::StatusOr<Eigen::Vector3f> GetLandmarkPosition(const int frame_id, const int keypoint_id);
  // test-driver storage
  std::map<int, std::vector<std::pair<ID, cv::KeyPoint>>> tracks;
  std::map<int, std::vector<int>> neighbours;
  std::map<int, Sophus::SE3f> poses;
  std::map<std::pair<int, int>, Eigen::Vector3f> positions;
};

class Map {
 public:
  Map() { RegularizationGraph::Options o{1.0f, 1.1f}; graph_ = std::make_shared<RegularizationGraph>(o, this); }
  void InsertKeyFrame(std::shared_ptr<KeyFrame> kf) { keyframes_[kf->GetId()] = kf; }
  void InsertMapPoint(std::shared_ptr<MapPoint> mp) { mappoints_[mp->GetId()] = mp; }
  absl::btree_map<ID, std::shared_ptr<KeyFrame>> GetKeyFrames() { return keyframes_; }
  absl::flat_hash_map<ID, std::shared_ptr<MapPoint>>& GetMapPoints() { return mappoints_; }
  std::shared_ptr<MapPoint> GetMapPoint(ID id) { return mappoints_.at(id); }
  std::shared_ptr<RegularizationGraph> GetRegularizationGraph() { return graph_; }
  void SetSigma(float s) { RegularizationGraph::Options o{s, 1.1f}; graph_ = std::make_shared<RegularizationGraph>(o, this); }
 private:
  absl::flat_hash_map<ID, std::shared_ptr<MapPoint>> mappoints_;
  absl::btree_map<ID, std::shared_ptr<KeyFrame>> keyframes_;
  std::shared_ptr<RegularizationGraph> graph_;
};
