// shim/lucas_kanade_tracker_b200.cc — drop-in replacement for modules/matching/lucas_kanade_tracker.cc of NR-SLAM.
//
// Keeps the class and its public interface (modules/matching/lucas_kanade_tracker.h:55-92): the constructor, 
// SetReferenceImage, Track, Get/InsertPhotometricInformation, clear and the public member prevPts_ that tracking.cc
// reads. The per-level patches live on the device; they are copied to the host only when a MapPoint stores them
// (GetPhotometricInformationOfPoint, mappoint.h:69) and back when PointReuse re-inserts them
// (InsertPhotometricInformation, tracking.cc:447). The one member the B200 build adds to the header is `void* b200_`.
//
// Compiled by the CPU test suite against shim/standin/ (tests/test_shim.py).
#include "matching/lucas_kanade_tracker.h"

#include <cstdint>
#include <vector>

#include "nrslam_b200.h"

namespace {
nrslam_b200_ctx* KltCtx() {  // shares nothing with the optimisation context: trackers may outlive a Frame
  static nrslam_b200_ctx* ctx = [] {
    nrslam_b200_ctx* c = nullptr;
    nrslam_b200_create(nullptr, &c);
    return c;
  }();
  return ctx;
}
nrslam_b200_klt* Dev(void* p) { return static_cast<nrslam_b200_klt*>(p); }
}  // namespace

LucasKanadeTracker::LucasKanadeTracker() : winSize_(cv::Size(21, 21)), maxLevel_(3), maxIters_(30), epsilon_(0.01f),
                                           minEigThreshold_(1e-4f) {}  // lucas_kanade_tracker.cc:28-30

LucasKanadeTracker::LucasKanadeTracker(const cv::Size _winSize, const int _maxLevel, const int _maxIters,
                                       const float _epsilon, const float _minEigThreshold)
    : winSize_(_winSize), maxLevel_(_maxLevel), maxIters_(_maxIters), epsilon_(_epsilon),
      minEigThreshold_(_minEigThreshold) {}

LucasKanadeTracker::~LucasKanadeTracker() {
  if (b200_) nrslam_b200_klt_destroy(Dev(b200_));
}

// lucas_kanade_tracker.cc:47-168
void LucasKanadeTracker::SetReferenceImage(const cv::Mat& refIm, const std::vector<cv::KeyPoint>& refPts,
                                           const cv::Mat& mask) {
  if (!b200_) {
    nrslam_b200_klt* k = nullptr;
    if (!KltCtx() || nrslam_b200_klt_create(KltCtx(), winSize_.width, maxLevel_, maxIters_, epsilon_, minEigThreshold_,
                                            &k) != NRSLAM_B200_OK)
      return;
    b200_ = k;
  }
  prevPts_ = refPts;  // :52 — read by tracking.cc
  std::vector<float> xy(2 * refPts.size());
  for (size_t i = 0; i < refPts.size(); i++) {
    xy[2 * i] = refPts[i].pt.x;
    xy[2 * i + 1] = refPts[i].pt.y;
  }
  nrslam_b200_klt_set_reference(Dev(b200_), refIm.ptr<uint8_t>(), refIm.cols, refIm.rows, (int32_t)refIm.step,
                                (int32_t)refPts.size(), xy.data(), mask.empty() ? nullptr : mask.ptr<uint8_t>(),
                                mask.empty() ? 0 : (int32_t)mask.step);
}

// lucas_kanade_tracker.cc:170-596
int LucasKanadeTracker::Track(const cv::Mat& newIm, std::vector<cv::KeyPoint>& nextPts,
                              std::vector<LandmarkStatus>& vMatched, const bool bInitialFlow, const float minSSIM,
                              const cv::Mat& mask) {
  if (!b200_) return 0;
  const size_t n = prevPts_.size();
  if (nextPts.size() != n) nextPts.resize(n);  // :176-178
  std::vector<float> xy(2 * n);
  std::vector<uint8_t> st(n);
  for (size_t i = 0; i < n; i++) {
    xy[2 * i] = nextPts[i].pt.x;
    xy[2 * i + 1] = nextPts[i].pt.y;
    st[i] = (uint8_t)vMatched[i];
  }
  int32_t n_tracked = 0;
  if (nrslam_b200_klt_track(Dev(b200_), newIm.ptr<uint8_t>(), newIm.cols, newIm.rows, (int32_t)newIm.step, (int32_t)n,
                            xy.data(), st.data(), bInitialFlow ? 1 : 0, minSSIM,
                            mask.empty() ? nullptr : mask.ptr<uint8_t>(), mask.empty() ? 0 : (int32_t)mask.step,
                            &n_tracked) != NRSLAM_B200_OK)
    return 0;
  for (size_t i = 0; i < n; i++) {
    nextPts[i].pt.x = xy[2 * i];
    nextPts[i].pt.y = xy[2 * i + 1];
    vMatched[i] = (LandmarkStatus)st[i];
  }
  return n_tracked;
}

// lucas_kanade_tracker.cc:598-608
LucasKanadeTracker::PhotometricInformation LucasKanadeTracker::GetPhotometricInformationOfPoint(const int idx) {
  PhotometricInformation info;
  if (!b200_) return info;
  const int L = maxLevel_ + 1, w = winSize_.width, h = winSize_.height;
  std::vector<int16_t> gray((size_t)L * w * h), grad((size_t)L * w * h * 2);
  std::vector<float> mean(L), mean2(L);
  std::vector<uint8_t> valid(L);
  if (nrslam_b200_klt_get_patch(Dev(b200_), idx, gray.data(), grad.data(), mean.data(), mean2.data(), valid.data()) !=
      NRSLAM_B200_OK)
    return info;
  for (int l = 0; l < L; l++) {
    info.mean_gray_per_level.push_back(mean[l]);
    info.squared_mean_gray_per_level.push_back(mean2[l]);
    if (!valid[l]) {  // the reference keeps an empty Mat for a level it could not sample (:160-164)
      info.gray_reference.push_back(cv::Mat());
      info.gradient_reference.push_back(cv::Mat());
      continue;
    }
    cv::Mat g(h, w, 2), d(h, w, 4);  // CV_16S, CV_16SC2
    for (int r = 0; r < h; r++) {
      memcpy(g.ptr<int16_t>(r), &gray[((size_t)l * h + r) * w], sizeof(int16_t) * w);
      memcpy(d.ptr<int16_t>(r), &grad[(((size_t)l * h + r) * w) * 2], sizeof(int16_t) * 2 * w);
    }
    info.gray_reference.push_back(g);
    info.gradient_reference.push_back(d);
  }
  return info;
}

// lucas_kanade_tracker.cc:610-620
void LucasKanadeTracker::InsertPhotometricInformation(cv::KeyPoint& keypoint,
                                                      PhotometricInformation& photometric_information) {
  if (!b200_) return;
  const int L = maxLevel_ + 1, w = winSize_.width, h = winSize_.height;
  std::vector<int16_t> gray((size_t)L * w * h, 0), grad((size_t)L * w * h * 2, 0);
  std::vector<float> mean(L, 0.f), mean2(L, 0.f);
  std::vector<uint8_t> valid(L, 0);
  for (int l = 0; l < L && l < (int)photometric_information.gray_reference.size(); l++) {
    mean[l] = photometric_information.mean_gray_per_level[l];
    mean2[l] = photometric_information.squared_mean_gray_per_level[l];
    const cv::Mat& g = photometric_information.gray_reference[l];
    const cv::Mat& d = photometric_information.gradient_reference[l];
    if (g.empty() || d.empty()) continue;
    valid[l] = 1;
    for (int r = 0; r < h; r++) {
      memcpy(&gray[((size_t)l * h + r) * w], g.ptr<int16_t>(r), sizeof(int16_t) * w);
      memcpy(&grad[(((size_t)l * h + r) * w) * 2], d.ptr<int16_t>(r), sizeof(int16_t) * 2 * w);
    }
  }
  if (nrslam_b200_klt_insert_patch(Dev(b200_), keypoint.pt.x, keypoint.pt.y, gray.data(), grad.data(), mean.data(),
                                   mean2.data(), valid.data()) == NRSLAM_B200_OK)
    prevPts_.push_back(keypoint);  // :612
}

// lucas_kanade_tracker.cc:622-631
void LucasKanadeTracker::clear() {
  prevPts_.clear();
  if (b200_) nrslam_b200_klt_clear(Dev(b200_));
}
