// oracle/ref_system.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// The reference's own `System` (src/slam/src/system.{hpp,cpp} and everything behind it, compiled unmodified by
// oracle/build_ref.sh) behind a C ABI, with the determinism pins SURVEY.md section 8c / Appendix B prescribe:
//   * state_->multiViewRandomEnabled_ = false  -> OpenGV samplers seeded 12345 instead of the clock (state.hpp:67);
//   * time stamps injected through processCameraPose(image, t) instead of system_clock (system.cpp:114) -- two frames
//     inside one millisecond would otherwise give dt = 0 and a NaN motion model (visual_frontend.hpp:17-56);
//   * cv::setNumThreads(1) (ref_config) for the racy parallel_for_ in feature_extractor.cpp:45;
//   * the three wall-clock caps of the Ceres solves (optimizer.cpp:258, :322, multi_view_geometry.cpp:185) lifted on request:
//     alva_ref_time_cap() below is what oracle/build_ref.sh's build-time copies of those two files call in place of the
//     literals -- identity by default, 1e9 s after ref_config_time_caps(1) (golden generation only).
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

static bool g_lift_time_caps = false;
double alva_ref_time_cap(double seconds) { return g_lift_time_caps ? 1e9 : seconds; }

extern "C" {

// 1: the Ceres solves run to their iteration limit / convergence whatever the host load (golden generation); 0: the reference's own caps
void ref_config_time_caps(int lift) { g_lift_time_caps = lift != 0; }

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

// Keypoint::desc_ of the frame's 2-D keypoints, same order as ref_system_get_frame_points: desc [cap][32], has [cap]
int ref_system_get_descriptors(void* h, uint8_t* desc, uint8_t* has, int cap) {
    System* s = (System*)h;
    auto kps = s->currFrame_->getKeypoints2d();
    int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        has[i] = kps[i].desc_.empty() ? 0 : 1;
        if (has[i]) memcpy(desc + 32 * (size_t)i, kps[i].desc_.ptr(), 32);
    }
    return n;
}

int ref_system_info(void* h, int32_t* out6) {
    System* s = (System*)h;
    out6[0] = s->currFrame_->id_; out6[1] = s->currFrame_->keyframeId_; out6[2] = (int)s->currFrame_->numKeypoints_;
    out6[3] = (int)s->currFrame_->numKeypoints3d_; out6[4] = s->state_->slamReadyForInit_ ? 1 : 0; out6[5] = (int)s->mapManager_->numKeyframes_;
    return 0;
}

// Mapper::matchToMap (src/slam/src/mapper.cpp:354-587) -- private; reached through the same `private public` switch -- on a map
// rebuilt from flat arrays with the reference's own classes (Frame, MapPoint, MapManager, Mapper), nothing re-implemented:
//   current frame: pose Twc [t, q(x,y,z,w)], keypoints (id, px) added in the given order (that order is the grid-cell order
//                  getSurroundingKeypoints walks), nkp3d = Frame::numKeypoints3d_ (doubles the pixel gate below 30)
//   keyframes    : id + pose; their keypoints come from the observations below
//   map points   : id, world point, is3d; observations (keyframe index, pixel in that keyframe) = observedKeyframeIds_ and
//                  the keyframe's keypoint; descriptors per keyframe = mapKeyframeDescriptors_ (desc_ = the last one)
//   local map    : the ids matchToMap iterates -- an std::unordered_set<int>; its iteration order decides ties, so the
//                  order the reference used is returned in order_out (the oracle and the GPU take it as input)
// Returns the number of (keypoint id -> map point id) pairs written to match_kp / match_mp (ascending keypoint id).
int ref_match_to_map(int w, int h, double fx, double fy, double cx, double cy, const double* Twc_cur, int n_kp, const int32_t* kp_id,
                     const float* kp_px, int nkp3d, int n_kf, const int32_t* kf_id, const double* kf_Twc, int n_mp,
                     const int32_t* mp_id, const double* mp_wpt, const uint8_t* mp_is3d, const int32_t* obs_start,
                     const int32_t* obs_kf, const float* obs_px, const int32_t* desc_start, const int32_t* desc_kf,
                     const uint8_t* desc, int n_local, const int32_t* local_ids, float max_proj_err, float dist_ratio,
                     int32_t* order_out, int32_t* match_kp, int32_t* match_mp) {
    std::cout.setstate(std::ios_base::failbit);
    auto state = std::make_shared<State>(w, h, 40);
    auto calib = std::make_shared<CameraCalibration>(fx, fy, cx, cy, 0., 0., 0., 0., w, h, 20);
    auto frame = std::make_shared<Frame>(calib, state->frameMaxCellSize_);
    auto extractor = std::make_shared<FeatureExtractor>(state->extractorMaxQuality_);
    auto manager = std::make_shared<MapManager>(state, frame, extractor);
    Mapper mapper(state, manager, frame);
    std::cout.clear();
    auto pose = [](const double* p) { return Sophus::SE3d(Eigen::Quaterniond(p[6], p[3], p[4], p[5]), Eigen::Vector3d(p[0], p[1], p[2])); };
    std::vector<std::shared_ptr<Frame>> kfs(n_kf);
    for (int k = 0; k < n_kf; k++) {
        kfs[k] = std::allocate_shared<Frame>(Eigen::aligned_allocator<Frame>(), calib, state->frameMaxCellSize_);
        kfs[k]->keyframeId_ = kf_id[k];
        kfs[k]->setTwc(pose(kf_Twc + 7 * k));
        manager->mapKeyframes_.emplace(kf_id[k], kfs[k]);
    }
    for (int m = 0; m < n_mp; m++) {
        auto mp = std::allocate_shared<MapPoint>(Eigen::aligned_allocator<MapPoint>());
        mp->mapPointId_ = mp_id[m];
        mp->isObserved_ = false;
        mp->is3d_ = mp_is3d[m] != 0;
        mp->point3d_ = Eigen::Vector3d(mp_wpt[3 * m], mp_wpt[3 * m + 1], mp_wpt[3 * m + 2]);
        mp->keyframeId_ = obs_start[m + 1] > obs_start[m] ? kf_id[obs_kf[obs_start[m]]] : 0;
        mp->invDepth_ = -1.;
        for (int o = obs_start[m]; o < obs_start[m + 1]; o++) {
            mp->observedKeyframeIds_.insert(kf_id[obs_kf[o]]);
            kfs[obs_kf[o]]->addKeypoint(cv::Point2f(obs_px[2 * o], obs_px[2 * o + 1]), mp_id[m]);
        }
        for (int d = desc_start[m]; d < desc_start[m + 1]; d++) {
            cv::Mat dm(1, 32, CV_8U);
            memcpy(dm.ptr(), desc + (size_t)32 * d, 32);
            mp->mapKeyframeDescriptors_.emplace(kf_id[desc_kf[d]], dm);
            mp->desc_ = dm;
        }
        manager->mapMapPoints_.emplace(mp_id[m], mp);
    }
    frame->keyframeId_ = n_kf ? kf_id[n_kf - 1] + 1 : 0;
    frame->setTwc(pose(Twc_cur));
    for (int i = 0; i < n_kp; i++) frame->addKeypoint(cv::Point2f(kp_px[2 * i], kp_px[2 * i + 1]), kp_id[i]);
    frame->numKeypoints3d_ = nkp3d;
    std::unordered_set<int> local;
    for (int i = 0; i < n_local; i++) local.insert(local_ids[i]);
    int k = 0;
    for (int id : local) order_out[k++] = id;
    std::map<int, int> res = mapper.matchToMap(*frame, max_proj_err, dist_ratio, local);
    int n = 0;
    for (auto& kv : res) { match_kp[n] = kv.first; match_mp[n] = kv.second; n++; }
    return n;
}

// every keypoint of the current frame in the iteration order of Frame::mapKeypoints_ (the order the reference feeds its
// solvers in): ids, px [n][2], is3d, world point of the 3-D ones; Twc7 = [t, q(x,y,z,w)] in double
int ref_system_keypoints(void* h, int32_t* ids, float* px, uint8_t* is3d, double* wpt, int cap, double* Twc7) {
    System* s = (System*)h;
    int n = 0;
    for (const auto& kv : s->currFrame_->mapKeypoints_) {
        if (n < cap) {
            ids[n] = kv.second.keypointId_; px[2 * n] = kv.second.px_.x; px[2 * n + 1] = kv.second.px_.y; is3d[n] = kv.second.is3d_;
            auto mp = s->mapManager_->getMapPoint(kv.second.keypointId_);
            for (int k = 0; k < 3; k++) wpt[3 * n + k] = (mp && mp->is3d_) ? mp->point3d_[k] : 0.0;
        }
        n++;
    }
    Sophus::SE3d T = s->currFrame_->getTwc();
    Eigen::Quaterniond q = T.unit_quaternion();
    Twc7[0] = T.translation()[0]; Twc7[1] = T.translation()[1]; Twc7[2] = T.translation()[2];
    Twc7[3] = q.x(); Twc7[4] = q.y(); Twc7[5] = q.z(); Twc7[6] = q.w();
    return n;
}
int ref_system_info8(void* h, int32_t* out8) {
    System* s = (System*)h;
    out8[0] = s->currFrame_->id_; out8[1] = s->currFrame_->keyframeId_; out8[2] = (int)s->currFrame_->numKeypoints_;
    out8[3] = (int)s->currFrame_->numKeypoints3d_; out8[4] = s->state_->slamReadyForInit_ ? 1 : 0; out8[5] = (int)s->mapManager_->numKeyframes_;
    out8[6] = (int)s->currFrame_->numOccupiedCells_; out8[7] = s->mapManager_->numMapPointIds_;
    return 0;
}

// System::findPlane (system.cpp:123-137) -> processPlane (:177-342), unmodified: RANSAC over the current frame's observed 3-D
// map points (seeded from std::random_device: not repeatable).  Returns the reference's 0 / 1; out16 as Utils::toPoseArray(Mat).
int ref_system_find_plane(void* h, float* out16, int iterations) {
    System* s = (System*)h;
    cv::Mat mat = s->processPlane(s->mapManager_->getCurrentFrameMapPoints(), s->currFrame_->getTwc(), iterations);
    if (mat.empty()) return 0;
    Utils::toPoseArray(mat, out16);
    return 1;
}

}  // extern "C"
