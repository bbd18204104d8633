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
// What runs per call today: the whole GPU front end of the north-star path (gray + pyramid + FAST + retainBest + ORB +
// Hamming 2-NN against the current local map) through alva_pipeline with batch = 1, the first frame seeding the local
// map with its descriptors (the reference's "first frame is a keyframe", visual_frontend.cpp:42).  Pose estimation
// (P3P-LMedS + PnP) is NOT built yet, so findCameraPose keeps reporting status 3
// ("not initialised", pose = identity) exactly as the reference does before its map is initialised -- it never
// fabricates a pose.  getFramePoints returns the features of the current frame.
#include "alva_common.cuh"
#include "../../include/alva_b200.h"
#include <math.h>
#include <string.h>
#include <vector>

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
        max_kps_ = ((w_ + 39) / 40) * ((h_ + 39) / 40);
        alva_pipeline_config c{};
        c.w = w_; c.h = h_; c.batch = 1; c.fast_thr = 20; c.nfeatures = max_kps_; c.orb_flags = 0;   // AlvaAR: -1 degree
        c.map_size = 0; c.kf_interval = 0;
        pipe_ = alva_pipeline_create(ctx_, &c);
        if (!pipe_) return ALVA_E_CUDA;
        int32_t info[4];
        alva_pipeline_info(pipe_, info);
        fcap_ = info[0];
        if (cudaMalloc(&rgba_dev_, (size_t)w_ * h_ * 4) != cudaSuccess || cudaMalloc(&map_dev_, (size_t)fcap_ * 32) != cudaSuccess ||
            cudaMalloc(&match_dev_, (size_t)fcap_ * 16) != cudaSuccess) {
            alva_set_error("System::configure: cudaMalloc failed");
            return ALVA_E_CUDA;
        }
        sel_host_.assign(fcap_, 0);
        match_host_.assign((size_t)fcap_ * 4, -1);
        configured_ = true;
        reset();
        return 0;
    }

    void reset() {
        frame_id_ = -1;
        map_n_ = 0;
        nkp_ = 0;
        n_matched_ = 0;
    }

    // returns the reference's status codes; pose16 layout as Utils::toPoseArray (src/slam/src/utils.cpp:3-27)
    int findCameraPose(const uint8_t* rgba, float* pose16) {
        if (!configured_) { alva_set_error("System: not configured"); return ALVA_E_STATE; }
        if (int e = processFrame(rgba)) return e;
        writeIdentity(pose16);
        return 3;   // not initialised yet: pose estimation is a "next" row; never fabricate a pose
    }

    int findCameraPoseWithIMU(const uint8_t* rgba, const double* imu, float* pose16) {
        if (!configured_) { alva_set_error("System: not configured"); return ALVA_E_STATE; }
        if (int e = processFrame(rgba)) return e;
        // system.cpp:66-69: quaternion (w, -x, y, z) -> R, inverted; translation only follows SLAM when status == 1
        const double qw = imu[0], qx = -imu[1], qy = imu[2], qz = imu[3];
        const double n = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
        const double w = qw / n, x = qx / n, y = qy / n, z = qz / n;
        const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                             2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                             2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        // inverse rotation = transpose; pose array stores R row-major in [0..2],[4..6],[8..10], t in [12..14]
        for (int i = 0; i < 16; i++) pose16[i] = 0.f;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) pose16[4 * r + c] = (float)R[3 * c + r];
        pose16[15] = 1.f;
        return 1;
    }

    int findPlane(float* /*out16*/, int /*numIterations*/) { return 0; }   // needs map points: none before initialisation

    // writes min(n, cap) (x, y) pairs, returns the true count (the reference overruns its buffer here)
    int getFramePoints(int32_t* xy, int cap_pairs) {
        const int n = nkp_;
        for (int i = 0; i < n && i < cap_pairs; i++) {
            xy[2 * i] = (int)ALVA_KEY_X(sel_host_[i]);
            xy[2 * i + 1] = (int)ALVA_KEY_Y(sel_host_[i]);
        }
        return n;
    }

    int numMatched() const { return n_matched_; }
    int device_ = 0;

private:
    int processFrame(const uint8_t* rgba) {
        cudaStream_t st = ctx_->stream;
        frame_id_++;
        ALVA_CUDA(cudaMemcpyAsync(rgba_dev_, rgba, (size_t)w_ * h_ * 4, cudaMemcpyHostToDevice, st));
        if (int e = alva_pipeline_step_dev(pipe_, rgba_dev_)) return e;
        int32_t* selcounts = (int32_t*)alva_pipeline_buffer(pipe_, 8);
        uint32_t* sel = (uint32_t*)alva_pipeline_buffer(pipe_, 7);
        uint8_t* desc = (uint8_t*)alva_pipeline_buffer(pipe_, 11);
        int32_t n = 0;
        ALVA_CUDA(cudaMemcpyAsync(&n, selcounts, 4, cudaMemcpyDeviceToHost, st));
        ALVA_CUDA(cudaMemcpyAsync(sel_host_.data(), sel, (size_t)fcap_ * 4, cudaMemcpyDeviceToHost, st));
        if (map_n_ > 0) {
            if (int e = alva_k_hamming_knn2_batch(ctx_, desc, selcounts, 1, fcap_, map_dev_, map_n_, match_dev_)) return e;
            ALVA_CUDA(cudaMemcpyAsync(match_host_.data(), match_dev_, (size_t)fcap_ * 16, cudaMemcpyDeviceToHost, st));
        }
        ALVA_CUDA(cudaStreamSynchronize(st));
        nkp_ = n < fcap_ ? n : fcap_;
        n_matched_ = 0;
        if (map_n_ > 0) {
            // Mapper::matchToMap thresholds (mapper.cpp:436-441, 540-546): dist <= 0.2*256, ratio 0.9
            for (int i = 0; i < nkp_; i++) {
                const int d0 = match_host_[4 * i + 1], d1 = match_host_[4 * i + 3];
                if (d0 >= 0 && d0 <= 51 && (d1 < 0 || d0 <= 0.9 * d1)) n_matched_++;
            }
        } else {
            // first frame: it becomes the keyframe whose descriptors seed the local map (visual_frontend.cpp:42)
            ALVA_CUDA(cudaMemcpyAsync(map_dev_, desc, (size_t)nkp_ * 32, cudaMemcpyDeviceToDevice, st));
            ALVA_CUDA(cudaStreamSynchronize(st));
            map_n_ = nkp_;
        }
        return 0;
    }

    static void writeIdentity(float* p) {
        for (int i = 0; i < 16; i++) p[i] = (i % 5 == 0) ? 1.f : 0.f;
    }

    void release() {
        if (pipe_) { alva_pipeline_destroy(pipe_); pipe_ = nullptr; }
        if (rgba_dev_) { cudaFree(rgba_dev_); rgba_dev_ = nullptr; }
        if (map_dev_) { cudaFree(map_dev_); map_dev_ = nullptr; }
        if (match_dev_) { cudaFree(match_dev_); match_dev_ = nullptr; }
        if (ctx_) { alva_ctx_destroy(ctx_); ctx_ = nullptr; }
        configured_ = false;
    }

    alva_ctx* ctx_ = nullptr;
    alva_pipeline* pipe_ = nullptr;
    uint8_t *rgba_dev_ = nullptr, *map_dev_ = nullptr;
    int32_t* match_dev_ = nullptr;
    std::vector<uint32_t> sel_host_;
    std::vector<int32_t> match_host_;
    int w_ = 0, h_ = 0, max_kps_ = 0, fcap_ = 0, map_n_ = 0, nkp_ = 0, n_matched_ = 0;
    long long frame_id_ = -1;
    double K_[4] = {0, 0, 0, 0}, dist_[4] = {0, 0, 0, 0};
    bool configured_ = false;
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
