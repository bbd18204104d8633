// tests/host/system_cpu_backend.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libalva_b200.so).
// alvaar_b200/csrc/system_core.h (the host-side state machine of System) instantiated over the CPU ORACLE instead of the CUDA
// kernels, so that the bookkeeping -- keypoint / map point / keyframe state, iteration orders, gates, status codes -- can be
// checked against the reference's own System in the GPU-less CPU suite (tests/test_system_core_cpu.py).  Every oracle function
// used here is bit-identical (integers, KLT, detector) or 1e-9-close (pose solvers) to the kernel the product calls in the same
// place, which the -m gpu tests establish stage by stage.
#include "../../alvaar_b200/csrc/system_core.h"
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {
void orc_gray(const uint8_t* rgba, int w, int h, uint8_t* gray);
void orc_pyrdown(const uint8_t* src, int w, int h, uint8_t* dst);
void orc_scharr(const uint8_t* img, int w, int h, int16_t* out);
void orc_fb_klt(const uint8_t* const* prev_img, const int16_t* const* prev_deriv, const uint8_t* const* cur_img,
                const int16_t* const* cur_deriv, int w0, int h0, int npyr_levels, int levels, int win, float error_value,
                float max_fb_dist, const float* pts, float* priors, uint8_t* good, int n, int max_count, double epsilon);
int orc_detect_points(const uint8_t* img, int w, int h, int cs, const float* cur, int ncur, const int* roi, double* quality,
                      float* out, int32_t* out_int, int cap);
void orc_orb_blur(const uint8_t* img, int w, int h, int fused, uint8_t* out);
int orc_orb_describe(const uint8_t* blurred, int w, int h, const float* pts, const float* angles, int n, uint8_t* desc, uint8_t* kept);
int orc_essential_5pt(const double* bv1, const double* bv2, int n, int max_iter, float err_px, int optimize, float fx, float fy,
                      uint32_t seed, double* Rt_out, uint8_t* outlier, double* info);
int orc_p3p_lmeds(const double* bvs, const double* wpts, int n, int max_iter, float err_px, float fx, float fy, uint32_t seed,
                  double* Twc_out, uint8_t* outlier, double* info);
int orc_pnp(const double* K, const double* uv, const double* X, int n, double* pose, double huber_delta, double chi2_thr,
            int max_iter, int use_robust, int apply_l2, uint8_t* outlier, double* summary);
void orc_triangulate(const double* Tlr, const double* bvl, const double* bvr, int n, double* out);
}

struct CpuBackend {
    int w = 0, h = 0, nlev = 0, cur = 0;
    int lw[4], lh[4];
    std::vector<uint8_t> img[2][4];
    std::vector<int16_t> der[2][4];
    std::vector<uint8_t> blur;
    bool blur_valid = false;
    double quality = 0.001;   // State::extractorMaxQuality_; FeatureExtractor keeps adapting it across resets

    void init(int W, int H) {
        w = W; h = H;
        int ww = W, hh = H;
        nlev = 0;
        for (int k = 0; k < 4; k++) { lw[k] = ww; lh[k] = hh; nlev = k + 1; ww = (ww + 1) / 2; hh = (hh + 1) / 2; if (ww <= 9 || hh <= 9) break; }
        for (int s = 0; s < 2; s++)
            for (int k = 0; k < nlev; k++) { img[s][k].assign((size_t)lw[k] * lh[k], 0); der[s][k].assign((size_t)lw[k] * lh[k] * 2, 0); }
        blur.assign((size_t)W * H, 0);
    }
    int pyramid(const uint8_t* rgba) {
        cur ^= 1;
        orc_gray(rgba, w, h, img[cur][0].data());
        for (int k = 1; k < nlev; k++) orc_pyrdown(img[cur][k - 1].data(), lw[k - 1], lh[k - 1], img[cur][k].data());
        for (int k = 0; k < nlev; k++) orc_scharr(img[cur][k].data(), lw[k], lh[k], der[cur][k].data());
        blur_valid = false;
        return 0;
    }
    int detect(const float* cpts, int ncur, std::vector<float>& fresh) {
        const int roi[4] = {20, 20, w - 40, h - 40};
        const int cap = 2 * (w / 40) * (h / 40) + 64;
        fresh.assign((size_t)cap * 2, 0.f);
        const int n = orc_detect_points(img[cur][0].data(), w, h, 40, cpts, ncur, roi, &quality, fresh.data(), nullptr, cap);
        fresh.resize((size_t)2 * (n < cap ? n : cap));
        return 0;
    }
    int describe(const float* pts, int n, uint8_t* desc, uint8_t* kept) {
        if (!blur_valid) { orc_orb_blur(img[cur][0].data(), w, h, 0, blur.data()); blur_valid = true; }
        orc_orb_describe(blur.data(), w, h, pts, nullptr, n, desc, kept);
        return 0;
    }
    int klt(const float* pts, float* priors, int n, int levels, uint8_t* good) {
        const uint8_t* pi[4]; const uint8_t* ci[4]; const int16_t* pd[4]; const int16_t* cd[4];
        const int prev = cur ^ 1;
        for (int k = 0; k < nlev; k++) { pi[k] = img[prev][k].data(); ci[k] = img[cur][k].data(); pd[k] = der[prev][k].data(); cd[k] = der[cur][k].data(); }
        orc_fb_klt(pi, pd, ci, cd, w, h, nlev - 1, levels, 9, 30.0f, 0.5f, pts, priors, good, n, 30, 0.01);
        return 0;
    }
    // test hook: the live reference's compute5ptEssentialMatrix (oracle/_ref: ref_essential_5pt) in place of the oracle's, to
    // separate the noise-limited refinement of the initialisation from everything downstream
    int (*essential_hook)(const double*, const double*, int, int, float, int, float, float, double*, uint8_t*) = nullptr;
    int essential(const double* b1, const double* b2, int n, float fx, float fy, double* Rt, uint8_t* outl) {
        if (essential_hook) return essential_hook(b1, b2, n, 100, 3.0f, 1, fx, fy, Rt, outl);
        double info[4];
        return orc_essential_5pt(b1, b2, n, 100, 3.0f, 1, fx, fy, 12345u, Rt, outl, info);
    }
    int p3p(const double* bv, const double* X, int n, float fx, float fy, double* T12, uint8_t* outl) {
        double info[3];
        return orc_p3p_lmeds(bv, X, n, 100, 3.0f, fx, fy, 12345u, T12, outl, info);
    }
    int pnp(const double* uv, const double* X, int n, const double* K4, double* pose7, uint8_t* outl) {
        double summary[10];
        const float chi2 = 5.9915f;
        return orc_pnp(K4, uv, X, n, pose7, std::sqrt((double)chi2), (double)chi2, 5, 1, 1, outl, summary);
    }
    int triangulate(const double* T7, const double* bl, const double* br, int n, double* out) { orc_triangulate(T7, bl, br, n, out); return 0; }
};

struct CpuSystem {
    CpuBackend be;
    alva_sys::SystemCore<CpuBackend> core;
    CpuSystem() : core(be) {}
};

extern "C" {
void* cpu_system_create(int w, int h, double fx, double fy, double cx, double cy) {
    CpuSystem* s = new CpuSystem();
    s->be.init(w, h);
    s->core.configure(w, h, fx, fy, cx, cy);
    return s;
}
void cpu_system_set_essential_hook(void* p, void* fn) {
    ((CpuSystem*)p)->be.essential_hook = (int (*)(const double*, const double*, int, int, float, int, float, float, double*, uint8_t*))fn;
}
void cpu_system_destroy(void* p) { delete (CpuSystem*)p; }
int cpu_system_process(void* p, const uint8_t* rgba, double t_ms, double* Twc7) {
    CpuSystem* s = (CpuSystem*)p;
    const int st = s->core.process(rgba, t_ms);
    s->core.cur.Twc.to7(Twc7);
    return st;
}
// keypoints of the current frame in the container's iteration order: ids, px [n][2], is3d, world point of the 3-D ones
int cpu_system_keypoints(void* p, int32_t* ids, float* px, uint8_t* is3d, double* wpt, int cap) {
    CpuSystem* s = (CpuSystem*)p;
    int n = 0;
    for (auto& kv : s->core.cur.kps) {
        if (n < cap) {
            ids[n] = kv.second.id; px[2 * n] = kv.second.px; px[2 * n + 1] = kv.second.py; is3d[n] = kv.second.is3d;
            auto mp = s->core.mappoints.find(kv.second.id);
            for (int k = 0; k < 3; k++) wpt[3 * n + k] = (mp != s->core.mappoints.end() && mp->second.is3d) ? mp->second.p[k] : 0.0;
        }
        n++;
    }
    return n;
}
int cpu_system_info(void* p, int32_t* out8) {
    CpuSystem* s = (CpuSystem*)p;
    out8[0] = s->core.cur.id; out8[1] = s->core.cur.kfid; out8[2] = s->core.cur.n; out8[3] = s->core.cur.n3d;
    out8[4] = s->core.ready_for_init; out8[5] = s->core.n_kf; out8[6] = s->core.cur.nocc; out8[7] = s->core.n_mp_ids;
    return 0;
}
}
