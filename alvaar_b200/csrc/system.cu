// system.cu -- host-side `System`: the reference's public class (src/slam/src/system.hpp:19-56) re-hosted on the
// B200 hot path, plus its C ABI (alva_system_*).  Same method names, argument meaning and return conventions as the
// reference so that embind.cpp / system.js stay source-compatible:
//
//   configure(w, h, fx, fy, cx, cy, k1, k2, p1, p2)      system.cpp:13-40
//   reset()                                               system.cpp:42-55
//   findCameraPose(rgba, pose16) -> 1 tracking / 2 reset / 3 not initialised     system.cpp:106-121, 156-175
//   findCameraPoseWithIMU(rgba, imu, pose16) -> 1        system.cpp:57-104
//   findPlane(out16, iterations) -> 0 / 1                 system.cpp:123-137
//   getFramePoints(xy) -> count                           system.cpp:139-154
//
// What runs per call: the reference's own per-frame sequence up to map initialisation, every pixel stage on the GPU --
//   System::findCameraPose        RGBA -> gray (system.cpp:112) + VisualFrontend::preprocessImage: pyramid + Scharr levels
//                                 (visual_frontend.cpp:672-698)                         -> alva_k_frontend, Scharr levels
//   first frame = keyframe        MapManager::createKeyframe -> extractKeypoints (map_manager.cpp:193-222):
//                                 FeatureExtractor::detectFeaturePoints + ids           -> alva_k_detect_grid
//   every other frame             VisualFrontend::kltTrackingFromMotionPrior (visual_frontend.cpp:103-243): all keypoints are
//                                 2-D before initialisation, so one forward-backward KLT on 3 levels from the previous
//                                 positions; failed tracks are dropped                   -> alva_k_klt_fb
//                                 < 50 keypoints -> reset, status 2 (visual_frontend.cpp:54-58, system.cpp:163-167)
//                                 median parallax to the keyframe (computeParallax, :596-670) decides when the reference
//                                 attempts its 5-point initialisation (checkReadyForInit, :419-552)
// Status, track ids and keypoint positions of this phase are bit-identical to the reference System (tests/test_gpu_system.py
// against tests/golden/system.npz, dumped from the reference's own System).  Map initialisation (5-point essential matrix,
// triangulation) and the mapper's bookkeeping are NOT built yet: once the parallax test fires the
// reference initialises and starts reporting poses, while this class keeps tracking and keeps reporting status 3 ("not
// initialised", pose = identity) -- it never fabricates a pose.  The per-frame pose kernels it will call then exist and are
// parity-tested on their own (alva_k_p3p_lmeds, alva_k_pnp, alva_k_ba_local).
#include "alva_common.cuh"
#include "../../include/alva_b200.h"
#include <math.h>
#include <string.h>
#include <algorithm>
#include <set>
#include <vector>

int alva_scharr_levels_launch(alva_ctx* ctx, int nlev, const uint8_t* const* src, int16_t* const* dst, const int* w, const int* h,
                              int nframes);

class System {
public:
    System() {}
    ~System() { release(); }

    int configure(int imageWidth, int imageHeight, double fx, double fy, double cx, double cy, double k1, double k2, double p1,
                  double p2) {
        release();
        w_ = imageWidth; h_ = imageHeight;
        K_[0] = fx; K_[1] = fy; K_[2] = cx; K_[3] = cy; dist_[0] = k1; dist_[1] = k2; dist_[2] = p1; dist_[3] = p2;
        ctx_ = alva_ctx_create(device_, nullptr);
        if (!ctx_) return ALVA_E_CUDA;
        // State(w, h, 40): frameMaxNumKeypoints = ceil(w/40) * ceil(h/40) (src/slam/src/state.cpp:3-12)
        cell_ = 40;
        max_kps_ = ((w_ + cell_ - 1) / cell_) * ((h_ + cell_ - 1) / cell_);
        cap_ = std::max(64, 2 * (w_ / cell_) * (h_ / cell_));   // a cell yields at most a primary and a secondary corner
        int ww = w_, hh = h_;
        nlev_ = 0;
        for (int k = 0; k < 4; k++) {   // buildOpticalFlowPyramid(win 9, maxLevel 3) stops when a level is not larger than the window
            lw_[k] = ww; lh_[k] = hh; nlev_ = k + 1;
            ww = (ww + 1) / 2; hh = (hh + 1) / 2;
            if (ww <= 9 || hh <= 9) break;
        }
        bool ok = cudaMalloc(&rgba_dev_, (size_t)w_ * h_ * 4) == cudaSuccess;
        for (int s = 0; s < 2 && ok; s++)
            for (int k = 0; k < 4 && ok; k++) {
                const size_t px = (size_t)lw_[k < nlev_ ? k : nlev_ - 1] * lh_[k < nlev_ ? k : nlev_ - 1];
                ok = cudaMalloc(&img_[s][k], px) == cudaSuccess && cudaMalloc(&der_[s][k], px * 4) == cudaSuccess;
            }
        ok = ok && cudaMalloc(&pts_dev_, (size_t)cap_ * 8) == cudaSuccess && cudaMalloc(&pri_dev_, (size_t)cap_ * 8) == cudaSuccess &&
             cudaMalloc(&good_dev_, cap_) == cudaSuccess && cudaMalloc(&cnt_dev_, 16) == cudaSuccess &&
             cudaMalloc(&quality_dev_, 8) == cudaSuccess && cudaMalloc(&blur_dev_, (size_t)w_ * h_) == cudaSuccess &&
             cudaMalloc(&desc_dev_, (size_t)cap_ * 32) == cudaSuccess && cudaMalloc(&kept_dev_, cap_) == cudaSuccess;
        if (!ok) { alva_set_error("System::configure: cudaMalloc failed"); return ALVA_E_CUDA; }
        const double q0 = 0.001;   // State::extractorMaxQuality_ (state.hpp:59); FeatureExtractor keeps adapting it across resets
        ALVA_CUDA(cudaMemcpy(quality_dev_, &q0, 8, cudaMemcpyHostToDevice));
        host_pts_.assign((size_t)cap_ * 2, 0.f);
        host_good_.assign(cap_, 0);
        host_desc_.assign((size_t)cap_ * 32, 0);
        configured_ = true;
        reset();
        return 0;
    }

    void reset() {   // system.cpp:42-55: frame, front end, map and state flags
        frame_id_ = -1;
        kps_.clear();
        next_id_ = 0;
        cur_ = 0;
        ready_for_init_ = false;
        init_due_ = false;
    }

    // returns the reference's status codes; pose16 layout as Utils::toPoseArray (src/slam/src/utils.cpp:3-27)
    int findCameraPose(const uint8_t* rgba, float* pose16) {
        if (!configured_) { alva_set_error("System: not configured"); return ALVA_E_STATE; }
        const int st = processCameraPose(rgba);
        writeIdentity(pose16);   // Twc of a frame that is not initialised (or was just reset) is the identity
        return st;
    }

    int findCameraPoseWithIMU(const uint8_t* rgba, const double* imu, float* pose16) {
        if (!configured_) { alva_set_error("System: not configured"); return ALVA_E_STATE; }
        const int st = processCameraPose(rgba);
        if (st < 0) return st;
        // system.cpp:66-69: quaternion (w, -x, y, z) -> R, inverted; translation only follows SLAM when status == 1
        const double qw = imu[0], qx = -imu[1], qy = imu[2], qz = imu[3];
        const double n = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
        const double w = qw / n, x = qx / n, y = qy / n, z = qz / n;
        const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                             2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                             2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        for (int i = 0; i < 16; i++) pose16[i] = 0.f;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) pose16[4 * r + c] = (float)R[3 * c + r];
        pose16[15] = 1.f;
        return 1;
    }

    int findPlane(float* /*out16*/, int /*numIterations*/) { return 0; }   // needs map points: none before initialisation

    // (x, y) = truncated undistorted position of the frame's 2-D keypoints; writes min(n, cap) pairs, returns the true count
    // (the reference overruns its 4096-int buffer here)
    int getFramePoints(int32_t* xy, int cap_pairs) {
        const int n = (int)kps_.size();
        for (int i = 0; i < n && i < cap_pairs; i++) { xy[2 * i] = (int)undist_x(kps_[i]); xy[2 * i + 1] = (int)undist_y(kps_[i]); }
        return n;
    }
    // the same keypoints with their track ids (== keypoint ids == map point ids, map_manager.cpp:166-191) and pixel positions
    int getTracks(int32_t* ids, float* px, int cap) {
        const int n = (int)kps_.size();
        for (int i = 0; i < n && i < cap; i++) { ids[i] = kps_[i].id; px[2 * i] = kps_[i].x; px[2 * i + 1] = kps_[i].y; }
        return n;
    }

    // the keypoints' ORB descriptors (same order as getTracks): FeatureExtractor::describeFeaturePoints at keyframe creation
    // (feature_extractor.cpp:160-214; empty for points within 31 px of the border, orb.cpp:1130)
    int getDescriptors(uint8_t* desc, uint8_t* has, int cap) {
        const int n = (int)kps_.size();
        for (int i = 0; i < n && i < cap; i++) { memcpy(desc + 32 * (size_t)i, kps_[i].desc, 32); has[i] = kps_[i].has_desc ? 1 : 0; }
        return n;
    }

    int numMatched() const { return (int)kps_.size(); }
    int initDue() const { return init_due_ ? 1 : 0; }
    int device_ = 0;

private:
    struct Kp { int id; float x, y, kfx, kfy; bool has_desc; uint8_t desc[32]; };

    // CameraCalibration::undistortImagePoint (camera_calibration.cpp:57-72): with the zero distortion the JS shim always passes
    // (system.js:84-141) cv::undistortPoints returns the input to float precision; non-zero coefficients are not supported yet
    float undist_x(const Kp& k) const { return k.x; }
    float undist_y(const Kp& k) const { return k.y; }

    int buildPyramid(const uint8_t* rgba) {
        cudaStream_t st = ctx_->stream;
        cur_ ^= 1;   // VisualFrontend::preprocessImage swaps prev / cur pyramids (visual_frontend.cpp:672-698)
        ALVA_CUDA(cudaMemcpyAsync(rgba_dev_, rgba, (size_t)w_ * h_ * 4, cudaMemcpyHostToDevice, st));
        uint8_t** L = img_[cur_];
        if (int e = alva_k_frontend(ctx_, rgba_dev_, w_, h_, 1, L[0], nlev_ > 1 ? L[1] : nullptr, nlev_ > 2 ? L[2] : nullptr,
                                    nlev_ > 3 ? L[3] : nullptr, 20, nullptr, nullptr, 0, 0))
            return e;
        const uint8_t* srcs[4] = {L[0], L[1], L[2], L[3]};
        return alva_scharr_levels_launch(ctx_, nlev_, srcs, der_[cur_], lw_, lh_, 1);
    }

    // MapManager::createKeyframe on the current frame (map_manager.cpp:12-22): extractKeypoints -> new keypoints with fresh ids
    int createKeyframe() {
        cudaStream_t st = ctx_->stream;
        const int n0 = (int)kps_.size();
        for (int i = 0; i < n0; i++) { host_pts_[2 * i] = kps_[i].x; host_pts_[2 * i + 1] = kps_[i].y; }
        int32_t ncur = n0;
        if (n0) ALVA_CUDA(cudaMemcpyAsync(pts_dev_, host_pts_.data(), (size_t)n0 * 8, cudaMemcpyHostToDevice, st));
        ALVA_CUDA(cudaMemcpyAsync(cnt_dev_ + 1, &ncur, 4, cudaMemcpyHostToDevice, st));
        const int32_t roi[4] = {20, 20, w_ - 40, h_ - 40};   // CameraCalibration(..., imgBorder 20): roi_rect_ (camera_calibration.cpp:21)
        // numToDetect = frameMaxNumKeypoints - occupied cells > 0 always holds before initialisation (map_manager.cpp:206-208)
        if (int e = alva_k_detect_grid(ctx_, img_[cur_][0], w_, h_, 1, cell_, pts_dev_, cnt_dev_ + 1, cap_, roi, quality_dev_, pri_dev_,
                                       nullptr, cnt_dev_, cap_))
            return e;
        int32_t n = 0;
        ALVA_CUDA(cudaMemcpyAsync(&n, cnt_dev_, 4, cudaMemcpyDeviceToHost, st));
        ALVA_CUDA(cudaMemcpyAsync(host_pts_.data(), pri_dev_, (size_t)cap_ * 8, cudaMemcpyDeviceToHost, st));
        ALVA_CUDA(cudaStreamSynchronize(st));
        if (n > cap_) n = cap_;
        // describeFeaturePoints(imageRaw, newPoints) (map_manager.cpp:215-219): ORB::create(500, 1, 0)->compute at the new
        // points = 7x7 blur (the shipped build's unfused arithmetic) + rBRIEF-256 steered by KeyPoint::convert's -1 degree
        if (n > 0) {
            if (int e = alva_k_orb_blur(ctx_, img_[cur_][0], blur_dev_, w_, h_, 1, 0)) return e;
            if (int e = alva_k_orb_describe(ctx_, img_[cur_][0], blur_dev_, w_, h_, 1, pri_dev_, cnt_dev_, cap_, 0, desc_dev_, kept_dev_, nullptr)) return e;
            ALVA_CUDA(cudaMemcpyAsync(host_desc_.data(), desc_dev_, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
            ALVA_CUDA(cudaMemcpyAsync(host_good_.data(), kept_dev_, n, cudaMemcpyDeviceToHost, st));
            ALVA_CUDA(cudaStreamSynchronize(st));
        }
        for (int i = 0; i < n; i++) {   // addKeypointsToFrame: id = running map point counter (map_manager.cpp:166-191)
            Kp k{next_id_++, host_pts_[2 * i], host_pts_[2 * i + 1], 0.f, 0.f, host_good_[i] != 0, {0}};
            if (k.has_desc) memcpy(k.desc, host_desc_.data() + 32 * (size_t)i, 32);
            kps_.push_back(k);
        }
        for (auto& k : kps_) { k.kfx = k.x; k.kfy = k.y; }   // the keyframe is a copy of the frame (addKeyframe, :243-252)
        return 0;
    }

    // VisualFrontend::kltTrackingFromMotionPrior with 2-D keypoints only: fbKltTracking(levels 3, error 30, fb distance 0.5)
    int kltTrack() {
        cudaStream_t st = ctx_->stream;
        const int n = (int)kps_.size();
        if (!n) return 0;
        for (int i = 0; i < n; i++) { host_pts_[2 * i] = kps_[i].x; host_pts_[2 * i + 1] = kps_[i].y; }
        ALVA_CUDA(cudaMemcpyAsync(pts_dev_, host_pts_.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        ALVA_CUDA(cudaMemcpyAsync(pri_dev_, pts_dev_, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
        const int prev = cur_ ^ 1;
        if (int e = alva_k_klt_fb(ctx_, img_[prev], der_[prev], img_[cur_], der_[cur_], w_, h_, 1, nlev_ - 1, 3, 9, 30.0f, 0.5f, pts_dev_,
                                  pri_dev_, nullptr, n, good_dev_))
            return e;
        ALVA_CUDA(cudaMemcpyAsync(host_pts_.data(), pri_dev_, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        ALVA_CUDA(cudaMemcpyAsync(host_good_.data(), good_dev_, n, cudaMemcpyDeviceToHost, st));
        ALVA_CUDA(cudaStreamSynchronize(st));
        std::vector<Kp> kept;
        kept.reserve(n);
        for (int i = 0; i < n; i++)
            if (host_good_[i]) { Kp k = kps_[i]; k.x = host_pts_[2 * i]; k.y = host_pts_[2 * i + 1]; kept.push_back(k); }   // updateKeypoint
        kps_.swap(kept);                                                                          // removeObsFromCurrFrameById
        return 0;
    }

    // VisualFrontend::computeParallax(keyframe, doUnRotate = false, doMedian = true) (visual_frontend.cpp:596-670): the "median"
    // is the element size/2 of a std::set<float>, i.e. of the DISTINCT parallax values
    float medianParallax() const {
        std::set<float> s;
        for (const auto& k : kps_) {
            const float dx = undist_x(k) - k.kfx, dy = undist_y(k) - k.kfy;
            s.insert((float)sqrt((double)dx * dx + (double)dy * dy));   // cv::norm(Point2f) is computed in double
        }
        if (s.empty()) return 0.f;
        auto it = s.begin();
        std::advance(it, s.size() / 2);
        return *it;
    }

    int processCameraPose(const uint8_t* rgba) {   // system.cpp:156-175 + VisualFrontend::track / process
        frame_id_++;
        if (int e = buildPyramid(rgba)) return e;
        if (frame_id_ == 0) {   // first frame -> keyframe (visual_frontend.cpp:42-45, 27)
            if (int e = createKeyframe()) return e;
            return 3;
        }
        if (int e = kltTrack()) return e;
        if (!ready_for_init_) {
            if ((int)kps_.size() < 50) { reset(); return 2; }        // visual_frontend.cpp:54-58 -> system.cpp:163-167
            // checkReadyForInit (visual_frontend.cpp:419-552): the 5-point initialisation is due once the median parallax
            // exceeds State::minAvgRotationParallax_ = 40 px; it is not built yet, so the frame stays "not initialised"
            init_due_ = medianParallax() > 40.0f;
        }
        return 3;
    }

    static void writeIdentity(float* p) {
        for (int i = 0; i < 16; i++) p[i] = (i % 5 == 0) ? 1.f : 0.f;
    }

    void release() {
        if (rgba_dev_) { cudaFree(rgba_dev_); rgba_dev_ = nullptr; }
        for (int s = 0; s < 2; s++)
            for (int k = 0; k < 4; k++) {
                if (img_[s][k]) { cudaFree(img_[s][k]); img_[s][k] = nullptr; }
                if (der_[s][k]) { cudaFree(der_[s][k]); der_[s][k] = nullptr; }
            }
        if (pts_dev_) { cudaFree(pts_dev_); pts_dev_ = nullptr; }
        if (pri_dev_) { cudaFree(pri_dev_); pri_dev_ = nullptr; }
        if (good_dev_) { cudaFree(good_dev_); good_dev_ = nullptr; }
        if (cnt_dev_) { cudaFree(cnt_dev_); cnt_dev_ = nullptr; }
        if (quality_dev_) { cudaFree(quality_dev_); quality_dev_ = nullptr; }
        if (blur_dev_) { cudaFree(blur_dev_); blur_dev_ = nullptr; }
        if (desc_dev_) { cudaFree(desc_dev_); desc_dev_ = nullptr; }
        if (kept_dev_) { cudaFree(kept_dev_); kept_dev_ = nullptr; }
        if (ctx_) { alva_ctx_destroy(ctx_); ctx_ = nullptr; }
        configured_ = false;
    }

    alva_ctx* ctx_ = nullptr;
    uint8_t* rgba_dev_ = nullptr;
    uint8_t* img_[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    int16_t* der_[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    float *pts_dev_ = nullptr, *pri_dev_ = nullptr;
    uint8_t* good_dev_ = nullptr;
    int32_t* cnt_dev_ = nullptr;
    double* quality_dev_ = nullptr;
    uint8_t *blur_dev_ = nullptr, *desc_dev_ = nullptr, *kept_dev_ = nullptr;
    std::vector<uint8_t> host_desc_;
    std::vector<float> host_pts_;
    std::vector<uint8_t> host_good_;
    std::vector<Kp> kps_;
    int w_ = 0, h_ = 0, cell_ = 40, max_kps_ = 0, cap_ = 0, nlev_ = 0, cur_ = 0, next_id_ = 0;
    int lw_[4] = {0, 0, 0, 0}, lh_[4] = {0, 0, 0, 0};
    long long frame_id_ = -1;
    double K_[4] = {0, 0, 0, 0}, dist_[4] = {0, 0, 0, 0};
    bool configured_ = false, ready_for_init_ = false, init_due_ = false;
};

struct alva_system { System sys; };

extern "C" alva_system* alva_system_create(int device) {
    alva_system* s = new alva_system();
    s->sys.device_ = device;
    return s;
}
extern "C" void alva_system_destroy(alva_system* s) { delete s; }
extern "C" int alva_system_configure(alva_system* s, int w, int h, double fx, double fy, double cx, double cy, double k1,
                                     double k2, double p1, double p2) {
    if (!s || w < 64 || h < 64) { alva_set_error("alva_system_configure: bad argument"); return ALVA_E_INVALID; }
    return s->sys.configure(w, h, fx, fy, cx, cy, k1, k2, p1, p2);
}
extern "C" int alva_system_reset(alva_system* s) { if (!s) return ALVA_E_INVALID; s->sys.reset(); return 0; }
extern "C" int alva_system_find_camera_pose(alva_system* s, const uint8_t* rgba, float* pose16) {
    if (!s || !rgba || !pose16) { alva_set_error("alva_system_find_camera_pose: bad argument"); return ALVA_E_INVALID; }
    return s->sys.findCameraPose(rgba, pose16);
}
extern "C" int alva_system_find_camera_pose_imu(alva_system* s, const uint8_t* rgba, const double* imu, float* pose16) {
    if (!s || !rgba || !imu || !pose16) { alva_set_error("alva_system_find_camera_pose_imu: bad argument"); return ALVA_E_INVALID; }
    return s->sys.findCameraPoseWithIMU(rgba, imu, pose16);
}
extern "C" int alva_system_find_plane(alva_system* s, float* out16, int iterations) {
    if (!s || !out16) return ALVA_E_INVALID;
    return s->sys.findPlane(out16, iterations);
}
extern "C" int alva_system_get_frame_points(alva_system* s, int32_t* xy, int cap_pairs) {
    if (!s || !xy || cap_pairs < 0) return ALVA_E_INVALID;
    return s->sys.getFramePoints(xy, cap_pairs);
}
extern "C" int alva_system_num_matched(alva_system* s) { return s ? s->sys.numMatched() : ALVA_E_INVALID; }
extern "C" int alva_system_get_tracks(alva_system* s, int32_t* ids, float* px, int cap) {
    if (!s || !ids || !px || cap < 0) return ALVA_E_INVALID;
    return s->sys.getTracks(ids, px, cap);
}
extern "C" int alva_system_get_descriptors(alva_system* s, uint8_t* desc, uint8_t* has, int cap) {
    if (!s || !desc || !has || cap < 0) return ALVA_E_INVALID;
    return s->sys.getDescriptors(desc, has, cap);
}
extern "C" int alva_system_init_due(alva_system* s) { return s ? s->sys.initDue() : ALVA_E_INVALID; }
