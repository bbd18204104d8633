// oracle/ref_system.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// The reference's own `System` (src/slam/src/system.{hpp,cpp} and everything behind it, compiled unmodified by
// oracle/build_ref.sh) behind a C ABI, with the determinism pins SURVEY.md section 8c / Appendix B prescribe:
//   * state_->multiViewRandomEnabled_ = false  -> OpenGV samplers seeded 12345 instead of the clock (state.hpp:67);
//   * time stamps injected through processCameraPose(image, t) instead of system_clock (system.cpp:114) -- two frames
//     inside one millisecond would otherwise give dt = 0 and a NaN motion model (visual_frontend.hpp:17-56);
//   * cv::setNumThreads(1) (ref_config) for the racy parallel_for_ in feature_extractor.cpp:45.
// The private members are reached without touching the sources: every std / third-party header is included first, then
// `private` is redefined for the reference's own headers only.
#include <opencv2/core.hpp>
#include <opencv2/core/utility.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/highgui.hpp>
#include <opencv2/features2d.hpp>
#include <opencv2/video/tracking.hpp>
#include <opencv2/calib3d.hpp>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/LU>
#include <opencv2/core/eigen.hpp>
#include <sophus/se3.hpp>
#include <ceres/ceres.h>
#include <chrono>
#include <iostream>
#include <memory>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <string>
#include <cstring>
#define private public
#define protected public
#include "system.hpp"
#undef private
#undef protected

extern "C" {

void* ref_system_create(int w, int h, double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2) {
    std::cout.setstate(std::ios_base::failbit);   // System::configure prints its settings
    System* s = new System();
    s->configure(w, h, fx, fy, cx, cy, k1, k2, p1, p2);
    std::cout.clear();
    s->state_->multiViewRandomEnabled_ = false;
    return s;
}
void ref_system_destroy(void* h) { delete (System*)h; }
void ref_system_reset(void* h) { ((System*)h)->reset(); }

// System::findCameraPose (system.cpp:106-121) with the time stamp injected: RGBA -> gray, processCameraPose, pose export.
int ref_system_find_camera_pose(void* h, const uint8_t* rgba, double t_ms, float* pose16) {
    System* s = (System*)h;
    cv::Mat image((int)s->state_->imgHeight_, (int)s->state_->imgWidth_, CV_8UC4, (void*)rgba);
    cv::cvtColor(image, image, cv::COLOR_RGBA2GRAY);
    int status = s->processCameraPose(image, t_ms);
    Utils::toPoseArray(s->currFrame_->getTwc(), pose16);
    return status;
}

// System::getFramePoints (system.cpp:139-154) without its 4096-int overrun: (x, y) = truncated unpx_ of the 2-D keypoints;
// ids (optional) = their keypoint ids (== map point ids == the track ids).  Returns the true count.
int ref_system_get_frame_points(void* h, int32_t* xy, int32_t* ids, float* px, int cap) {
    System* s = (System*)h;
    auto kps = s->currFrame_->getKeypoints2d();
    int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        xy[2 * i] = (int)kps[i].unpx_.x; xy[2 * i + 1] = (int)kps[i].unpx_.y;
        if (ids) ids[i] = kps[i].keypointId_;
        if (px) { px[2 * i] = kps[i].px_.x; px[2 * i + 1] = kps[i].px_.y; }
    }
    return n;
}

int ref_system_info(void* h, int32_t* out6) {
    System* s = (System*)h;
    out6[0] = s->currFrame_->id_; out6[1] = s->currFrame_->keyframeId_; out6[2] = (int)s->currFrame_->numKeypoints_;
    out6[3] = (int)s->currFrame_->numKeypoints3d_; out6[4] = s->state_->slamReadyForInit_ ? 1 : 0; out6[5] = (int)s->mapManager_->numKeyframes_;
    return 0;
}

}  // extern "C"
