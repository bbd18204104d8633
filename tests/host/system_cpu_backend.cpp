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
int orc_ba_local(const double* calib, double* poses, const uint8_t* pose_const, int nkf, double* invd, const int32_t* anch_kf,
                 const double* anch_uv, int nlm, const int32_t* obs_kf, const int32_t* obs_lm, const double* obs_uv, int nobs,
                 double huber_delta, double chi2_thr, int max_iter, int32_t* flags, double* summary);
int orc_match_to_map(int w, int h, double fx, double fy, double cx, double cy, const double* Twc_cur, int n_kp, const int32_t* kp_id,
                     const float* kp_px, int nkp3d, int n_kf, const int32_t* kf_id, const double* kf_Twc, int n_mp, const int32_t* mp_id,
                     const double* mp_wpt, const uint8_t* mp_is3d, const int32_t* obs_start, const int32_t* obs_kf, const float* obs_px,
                     const int32_t* desc_start, const int32_t* desc_kf, const uint8_t* desc, int n_local, const int32_t* local_ids,
                     float max_proj_err, float dist_ratio, int32_t* match_kp, int32_t* match_mp);
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
        return orc_pnp(K4, uv, X, n, pose7, (double)std::sqrt(chi2), (double)chi2, 5, 1, 1, outl, summary);
    }
    int triangulate(const double* T7, const double* bl, const double* br, int n, double* out) { orc_triangulate(T7, bl, br, n, out); return 0; }
    bool enable_ba = true, enable_match = true;
    double fx = 0, fy = 0, cx = 0, cy = 0;
    bool has_ba_local() const { return enable_ba; }
    bool has_match_to_map() const { return enable_match; }
    int ba_local(alva_sys::BaProblem& bp, int32_t* flags) {
        double summary[10];
        const float chi2 = 5.9915f;   // State::robustCostThreshold_ (float); Huber width std::sqrt(float)
        orc_ba_local(bp.calib, bp.poses.data(), bp.pose_const.data(), bp.nkf, bp.invd.data(), bp.anch_kf.data(), bp.anch_uv.data(), bp.nlm,
                     bp.obs_kf.data(), bp.obs_lm.data(), bp.obs_uv.data(), bp.nobs, (double)std::sqrt(chi2), (double)chi2, 5, flags, summary);
        return 0;
    }
    int match_to_map(const alva_sys::MatchProblem& m, std::vector<int>& kp_match) {
        const int n_kp = (int)m.kp_id.size(), n_mp = (int)m.mp_id.size();
        std::vector<int32_t> obs_kf(m.obs_kfid.size()), local_ids(m.local_mp.size()), mk(n_kp + 1), mm(n_kp + 1);
        for (size_t o = 0; o < obs_kf.size(); o++) { int ki = 0; while (m.kf_id[ki] != m.obs_kfid[o]) ki++; obs_kf[o] = ki; }
        for (size_t i = 0; i < local_ids.size(); i++) local_ids[i] = m.mp_id[m.local_mp[i]];
        const int n = orc_match_to_map(w, h, fx, fy, cx, cy, m.Twc_cur, n_kp, m.kp_id.data(), m.kp_px.data(), m.nkp3d, (int)m.kf_id.size(),
                                       m.kf_id.data(), m.kf_Twc.data(), n_mp, m.mp_id.data(), m.mp_wpt.data(), m.mp_is3d.data(), m.obs_start.data(),
                                       obs_kf.data(), m.obs_px.data(), m.desc_start.data(), m.desc_kfid.data(), m.desc.data(), (int)local_ids.size(),
                                       local_ids.data(), 2.0f, 0.2f, mk.data(), mm.data());
        for (int i = 0; i < n; i++) {
            int ki = 0, mi = 0;
            while (m.kp_id[ki] != mk[i]) ki++;
            while (m.mp_id[mi] != mm[i]) mi++;
            kp_match[ki] = mi;
        }
        return 0;
    }
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
    s->be.fx = fx; s->be.fy = fy; s->be.cx = cx; s->be.cy = cy;
    s->core.configure(w, h, fx, fy, cx, cy);
    return s;
}
void cpu_system_set_essential_hook(void* p, void* fn) {
    ((CpuSystem*)p)->be.essential_hook = (int (*)(const double*, const double*, int, int, float, int, float, float, double*, uint8_t*))fn;
}
void cpu_system_enable(void* p, int ba, int match) { ((CpuSystem*)p)->be.enable_ba = ba != 0; ((CpuSystem*)p)->be.enable_match = match != 0; }
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
int cpu_system_find_plane(void* p, float* out16, int iterations) { return ((CpuSystem*)p)->core.findPlane(out16, iterations); }
int cpu_system_info(void* p, int32_t* out8) {
    CpuSystem* s = (CpuSystem*)p;
    out8[0] = s->core.cur.id; out8[1] = s->core.cur.kfid; out8[2] = s->core.cur.n; out8[3] = s->core.cur.n3d;
    out8[4] = s->core.ready_for_init; out8[5] = s->core.n_kf; out8[6] = s->core.cur.nocc; out8[7] = s->core.n_mp_ids;
    return 0;
}
}
