// TEST INFRASTRUCTURE: drives the shim (shim/*.cc) with fake Frame / Map / KeyFrame / TemporalBuffer objects built from
// the stand-in headers. `driver mock` links against tests/shim/mock_abi.cc and checks that every recognisable result
// lands where the reference writes it; `driver real` links against libnrslam_b200.so and runs a consistent synthetic
// scene through the real CUDA path (GPU test).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "matching/lucas_kanade_tracker.h"
#include "optimization/g2o_optimization.h"

extern "C" int mock_errors(void) __attribute__((weak));

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "driver: check failed: %s (line %d)\n", #c, __LINE__); g_fail++; } } while (0)

int main(int argc, char** argv) {
  const bool real = argc > 1 && !strcmp(argv[1], "real");
  const float fx = 500.f, cx = 320.f, cy = 240.f;
  auto calib = std::make_shared<CameraModel>(std::vector<float>{fx, fx, cx, cy});
  auto map = std::make_shared<Map>();
  map->SetSigma(1.0f);
  // map points on a 8 x 6 grid of the plane z = 3, ids 10, 13, 16, ... (not contiguous: exercises the id <-> vertex maps)
  const int W = 8, H = 6, N = W * H;
  std::vector<ID> ids(N);
  std::vector<Eigen::Vector3f> P(N);
  for (int r = 0; r < H; r++)
    for (int c = 0; c < W; c++) {
      const int i = r * W + c;
      ids[i] = 10 + 3 * i;
      P[i] = Eigen::Vector3f(0.3f * (c - 3.5f), 0.3f * (r - 2.5f), 3.0f + 0.02f * c);
      map->InsertMapPoint(std::make_shared<MapPoint>(P[i], ids[i]));
    }
  auto graph = map->GetRegularizationGraph();
  int n_edges = 0;
  for (int r = 0; r < H; r++)
    for (int c = 0; c < W; c++)
      for (int dr = 0; dr <= 2; dr++)
        for (int dc = -2; dc <= 2; dc++) {
          if (dr == 0 && dc <= 0) continue;
          const int r2 = r + dr, c2 = c + dc;
          if (r2 >= H || c2 < 0 || c2 >= W) continue;
          const int a = r * W + c, b = r2 * W + c2;
          Eigen::Vector3f rel(P[b][0] - P[a][0], P[b][1] - P[a][1], P[b][2] - P[a][2]);
          graph->AddEdge(ids[a], ids[b], rel);
          n_edges++;
        }
  auto project = [&](const Eigen::Vector3f& X, float du) {
    cv::KeyPoint kp;
    kp.pt.x = fx * X[0] / X[2] + cx + du;
    kp.pt.y = fx * X[1] / X[2] + cy;
    return kp;
  };
  // ---- frame: every 5th point only TRACKED (no 3-D this frame -> "lost" candidates), the rest TRACKED_WITH_3D
  Frame frame;
  frame.SetCalibration(calib);
  int n3d = 0;
  for (int i = 0; i < N; i++) {
    const LandmarkStatus st = (i % 5 == 4) ? TRACKED : TRACKED_WITH_3D;
    n3d += st == TRACKED_WITH_3D;
    frame.InsertObservation(project(P[i], 0.4f), P[i], ids[i], st);
  }
  frame.MutableCameraTransformationWorld() = Sophus::SE3f(Eigen::Quaternionf(1, 0, 0, 0), Eigen::Vector3f(0, 0, 0));

  // 1. CameraPoseOptimization
  CameraPoseOptimization(frame, Sophus::SE3f());
  const auto t1 = frame.CameraTransformationWorld().translation();
  if (!real) CHECK(std::fabs(t1.x() - 0.25f) < 1e-6f);
  CHECK(std::isfinite(t1.x()) && std::isfinite(t1.y()) && std::isfinite(t1.z()));

  // 2. CameraPoseAndDeformationOptimization
  const float w0 = graph->Connections().begin()->second.begin()->second->weight;
  auto lost = CameraPoseAndDeformationOptimization(frame, map, Sophus::SE3f(), 1.0f);
  const auto t2 = frame.CameraTransformationWorld().translation();
  CHECK(std::isfinite(t2.x()) && std::isfinite(t2.y()) && std::isfinite(t2.z()));
  if (!real) {
    CHECK(std::fabs(t2.y() - 0.75f) < 1e-6f);
    CHECK((int)lost.size() == N - n3d);
    for (int i = 0; i < N; i++)
      if (i % 5 == 4) CHECK(lost.count(ids[i]) == 1);
    CHECK(std::fabs(frame.GetDeformationMagnitud() - 0.125f) < 1e-6f);
    int k3d = 0;
    for (int i = 0; i < N; i++) {
      if (i % 5 == 4) {  // untouched observation; its map point was moved by the lost-point stage
        CHECK(frame.LandmarkStatuses()[i] == TRACKED);
        CHECK(std::fabs(map->GetMapPoint(ids[i])->GetLastWorldPosition()[1] - (P[i][1] + 2.0f)) < 1e-5f);
        continue;
      }
      CHECK(std::fabs(frame.LandmarkPositions()[i][0] - (P[i][0] + 0.5f)) < 1e-6f);
      CHECK(frame.LandmarkStatuses()[i] == ((k3d % 3 == 2) ? TRACKED : TRACKED_WITH_3D));
      CHECK(std::fabs(map->GetMapPoint(ids[i])->GetLastWorldPosition()[0] - (P[i][0] + 1.0f)) < 1e-5f);
      CHECK(map->GetMapPoint(ids[i])->n_set_ == 1);
      k3d++;
    }
    const auto first = graph->Connections().begin()->second.begin()->second;
    CHECK(std::fabs(first->weight - 0.5f * w0) < 1e-6f);
    CHECK(first->status == RegularizationGraph::BAD);
    CHECK(first->max_distance > first->first_distance + 0.9f);
  } else {
    for (int i = 0; i < N; i++) {
      const auto s = frame.LandmarkStatuses()[i];
      CHECK(s == TRACKED_WITH_3D || s == TRACKED || s == BAD);
      CHECK(!frame.LandmarkPositions()[i].hasNaN());
    }
  }

  // 3. LocalDeformableBundleAdjustment: 6 keyframes in the map, the newest 5 form the window
  for (int k = 0; k < 6; k++) {
    auto kf = std::make_shared<KeyFrame>((ID)(100 + k));
    kf->calibration_ = calib;
    // tx encodes the age inside the window (keyframe 101 is the oldest one used -> 0)
    kf->CameraTransformationWorld() = Sophus::SE3f(Eigen::Quaternionf(1, 0, 0, 0), Eigen::Vector3f(real ? 0.01f * k : (float)(k - 1), 0, 0));
    for (int i = 0; i < N; i++) {
      Eigen::Vector3f X = P[i];
      if (real) X[0] += 0.01f * k;  // keeps the projections consistent with the pose
      Eigen::Vector3f Xc(X[0] + (real ? 0.01f * k : 0.f), X[1], X[2]);
      kf->Insert(project(Xc, 0.0f), X, ids[i], (i % 7 == 3) ? TRACKED : TRACKED_WITH_3D);
    }
    map->InsertKeyFrame(kf);
  }
  LocalDeformableBundleAdjustment(map, 1.0f);
  {
    auto kfs = map->GetKeyFrames();
    if (!real) {
      CHECK(std::fabs(kfs.at(100)->CameraTransformationWorld().translation().z()) < 1e-6f);  // outside the window
      for (int k = 1; k < 6; k++) {
        auto kf = kfs.at(100 + k);
        CHECK(std::fabs(kf->CameraTransformationWorld().translation().z() - 1.0f) < 1e-6f);
        for (int i = 0; i < N; i++) {
          const float want = (i % 7 == 3) ? P[i][2] : P[i][2] + 0.01f * k;  // window slot k-1 -> +0.01 (slot + 1)
          CHECK(std::fabs(kf->LandmarkPositions()[i][2] - want) < 1e-5f);
        }
      }
    } else {
      for (int k = 1; k < 6; k++)
        for (int i = 0; i < N; i++) CHECK(!kfs.at(100 + k)->LandmarkPositions()[i].hasNaN());
    }
  }

  // 4. DeformableTriangulation of one candidate
  if (!real) {
    TemporalBuffer tb;
    for (int f = 0; f < 6; f++) {
      cv::KeyPoint kp;
      kp.pt.x = 100.f + f;
      kp.pt.y = 50.f;
      tb.tracks[7].push_back({(ID)(20 + f), kp});
      tb.poses[20 + f] = Sophus::SE3f(Eigen::Quaternionf(1, 0, 0, 0), Eigen::Vector3f(0.01f * f, 0, 0));
      tb.positions[{20 + f, 3}] = Eigen::Vector3f(0, 0, 3);
      if (f != 2) tb.positions[{20 + f, 4}] = Eigen::Vector3f(0.1f, 0, 3);
    }
    tb.neighbours[7] = {3, 4};
    auto r = DeformableTriangulation(tb, 7, calib, 1.0f);
    CHECK(r.ok() && (*r)[0] == 1.f && (*r)[1] == 2.f && (*r)[2] == 3.f);
    for (auto& t : tb.tracks[7]) t.second.pt.x = -t.second.pt.x;
    auto r2 = DeformableTriangulation(tb, 7, calib, 1.0f);
    CHECK(!r2.ok() && r2.status().message() == "Low parallax.");
  }

  // 5. LucasKanadeTracker
  if (!real) {
    LucasKanadeTracker klt(cv::Size(21, 21), 4, 10, 1e-4f, 1e-4f);
    cv::Mat im(48, 64, 1);
    memset(im.data, 128, 48 * 64);
    std::vector<cv::KeyPoint> pts(3);
    for (int i = 0; i < 3; i++) { pts[i].pt.x = 20.f + 5 * i; pts[i].pt.y = 24.f; }
    klt.SetReferenceImage(im, pts);
    CHECK(klt.prevPts_.size() == 3);
    std::vector<cv::KeyPoint> next(3);
    std::vector<LandmarkStatus> st(3, TRACKED_WITH_3D);
    const int nt = klt.Track(im, next, st, false, 0.7f, cv::Mat());
    CHECK(nt == 2 && st[1] == BAD && std::fabs(next[2].pt.x - 31.5f) < 1e-6f && std::fabs(next[2].pt.y - 23.5f) < 1e-6f);
    auto info = klt.GetPhotometricInformationOfPoint(1);
    CHECK(info.gray_reference.size() == 5 && info.gray_reference[4].empty() && !info.gray_reference[0].empty());
    CHECK(info.gray_reference[1].ptr<int16_t>(0)[5] == 1005 && info.mean_gray_per_level[2] == 21.f);
    info.mean_gray_per_level[1] = 10.f;
    cv::KeyPoint extra;
    extra.pt.x = 40.f;
    extra.pt.y = 30.f;
    klt.InsertPhotometricInformation(extra, info);
    CHECK(klt.prevPts_.size() == 4);
    klt.clear();
    CHECK(klt.prevPts_.empty());
  }
  const int merr = (!real && mock_errors) ? mock_errors() : 0;
  printf("shim driver (%s): %d driver failures, %d ABI-contract failures, %d graph edges\n", real ? "real" : "mock", g_fail,
         merr, n_edges);
  return (g_fail || merr) ? 1 : 0;
}
