// stand-in (test infrastructure): forwards to the consolidated declarations
#include "nrslam_standin.h"
// the four signatures of modules/optimization/g2o_optimization.h:27-40
void CameraPoseOptimization(Frame& frame, const Sophus::SE3f& previous_camera_transform_world);
absl::flat_hash_set<ID> CameraPoseAndDeformationOptimization(Frame& current_frame, std::shared_ptr<Map> map,
                                                             const Sophus::SE3f& previous_camera_transform_world,
                                                             const float scale);
absl::StatusOr<Eigen::Vector3f> DeformableTriangulation(TemporalBuffer& temporal_buffer, int candidate_id,
                                                        std::shared_ptr<CameraModel> calibration, const float scale);
void LocalDeformableBundleAdjustment(std::shared_ptr<Map> map, const float scale);
