// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// Thin extern "C" wrapper around the UNMODIFIED reference libraries (vendored OpenCV 4.5.5,
// Ceres 2.0 and AlvaAR's own ceres_parametrization.cpp), compiled where they lie under
// /root/reference by oracle/build_ref.sh into oracle/_ref/libalva_ref.so.
//
// It is used (a) to pin the plain-C restatement in oracle/alva_oracle.c (golden vectors in
// tests/golden are dumped through these entry points by tools/make_golden.py) and (b) as the
// "reference" CPU baseline arm of bench.py.  Each entry point names the reference call it makes.
#include <opencv2/core.hpp>
#include <opencv2/core/utility.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/features2d.hpp>
#include <opencv2/video/tracking.hpp>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <sophus/se3.hpp>
#include <ceres/ceres.h>
#include "ceres_parametrization.hpp"   // /root/reference/src/slam/src (AlvaAR's own cost functors)
#include "feature_tracker.hpp"         // /root/reference/src/slam/src (AlvaAR's own forward-backward KLT wrapper)
#include "feature_extractor.hpp"       // /root/reference/src/slam/src (grid Shi-Tomasi detector)
#include "multi_view_geometry.hpp"     // /root/reference/src/slam/src (P3P-LMedS via OpenGV, Ceres PnP)
#include <cstdint>
#include <cstring>
#include <vector>
#include <memory>
#include <cmath>

extern "C" {

// cv::setUseOptimized / cv::setNumThreads: pin (or release) OpenCV's runtime CPU dispatch.
// optimized=0 forces the SSE3 baseline code paths (no FMA) -- the arithmetic of the shipped WASM
// build, whose simd128 has no fused multiply-add either.
void ref_config(int optimized, int threads) {
    cv::setUseOptimized(optimized != 0);
    cv::setNumThreads(threads);
}

int ref_info(char* buf, int cap) {
    std::string s = std::string("opencv ") + CV_VERSION + " ceres " + CERES_VERSION_STRING +
                    " optimized=" + (cv::useOptimized() ? "1" : "0") +
                    " threads=" + std::to_string(cv::getNumThreads()) +
                    " cpu=" + cv::getCPUFeaturesLine();
    int n = (int)std::min<size_t>(s.size(), cap > 0 ? cap - 1 : 0);
    if (cap > 0) { memcpy(buf, s.data(), n); buf[n] = 0; }
    return (int)s.size();
}

// System::findCameraPose, src/slam/src/system.cpp:111-112 : cv::cvtColor(RGBA2GRAY)
void ref_gray(const uint8_t* rgba, int w, int h, uint8_t* gray) {
    cv::Mat src(h, w, CV_8UC4, (void*)rgba), dst(h, w, CV_8UC1, gray);
    cv::cvtColor(src, dst, cv::COLOR_RGBA2GRAY);
}

// cv::pyrDown (opencv/modules/imgproc/src/pyramids.cpp:1260) -- one level, default border.
void ref_pyrdown(const uint8_t* src, int w, int h, uint8_t* dst) {
    cv::Mat s(h, w, CV_8UC1, (void*)src), d((h + 1) / 2, (w + 1) / 2, CV_8UC1, dst);
    cv::pyrDown(s, d);
}

// VisualFrontend::preprocessImage, src/slam/src/visual_frontend.cpp:696 :
// cv::buildOpticalFlowPyramid(gray, pyr, Size(win,win), maxLevel) -- returns the number of levels
// actually built; level k (u8, tightly packed) is copied to levels[k]; deriv[k] (int16 x2) optional.
int ref_build_pyramid(const uint8_t* gray, int w, int h, int win, int max_level,
                      uint8_t** levels, int16_t** derivs) {
    cv::Mat g(h, w, CV_8UC1, (void*)gray);
    std::vector<cv::Mat> pyr;
    int got = cv::buildOpticalFlowPyramid(g, pyr, cv::Size(win, win), max_level, true);
    for (int k = 0; k <= got; k++) {
        const cv::Mat& L = pyr[2 * k];
        if (levels && levels[k])
            for (int y = 0; y < L.rows; y++) memcpy(levels[k] + (size_t)y * L.cols, L.ptr(y), L.cols);
        const cv::Mat& D = pyr[2 * k + 1];
        if (derivs && derivs[k])
            for (int y = 0; y < D.rows; y++)
                memcpy(derivs[k] + (size_t)y * D.cols * 2, D.ptr(y), (size_t)D.cols * 4);
    }
    return got;
}

// cv::FAST(img, kps, thr, nms, TYPE_9_16)  (opencv/modules/features2d/src/fast.cpp:496)
// out: xys[3*i] = {x, y, response}; returns the true count (writes at most cap).
int ref_fast(const uint8_t* gray, int w, int h, int thr, int nms, int32_t* xys, int cap) {
    cv::Mat g(h, w, CV_8UC1, (void*)gray);
    std::vector<cv::KeyPoint> kps;
    cv::FAST(g, kps, thr, nms != 0, cv::FastFeatureDetector::TYPE_9_16);
    int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        xys[3 * i] = (int)kps[i].pt.x; xys[3 * i + 1] = (int)kps[i].pt.y; xys[3 * i + 2] = (int)kps[i].response;
    }
    return n;
}

// FeatureExtractor::describeFeaturePoints, src/slam/src/feature_extractor.cpp:160-214:
// KeyPoint::convert(points) (angle=-1, octave 0) ; ORB::create(500, 1., 0)->compute(image, kps, desc)
// kept[i] = 1 if point i survived ORB's border filter; desc row i is valid only then.
// angles != NULL: set KeyPoint::angle explicitly (degrees) -- steered-BRIEF with a given orientation.
int ref_orb_compute(const uint8_t* gray, int w, int h, const float* pts, const float* angles, int n,
                    uint8_t* desc, uint8_t* kept) {
    cv::Mat g(h, w, CV_8UC1, (void*)gray);
    std::vector<cv::Point2f> p(n);
    for (int i = 0; i < n; i++) p[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    std::vector<cv::KeyPoint> kps;
    cv::KeyPoint::convert(p, kps);
    if (angles) for (int i = 0; i < n; i++) kps[i].angle = angles[i];
    for (int i = 0; i < n; i++) kps[i].class_id = i;
    cv::Ptr<cv::DescriptorExtractor> orb = cv::ORB::create(500, 1.f, 0);
    cv::Mat d;
    orb->compute(g, kps, d);
    memset(kept, 0, n);
    for (size_t j = 0; j < kps.size(); j++) {
        int i = kps[j].class_id;
        kept[i] = 1;
        memcpy(desc + (size_t)i * 32, d.ptr((int)j), 32);
    }
    return (int)kps.size();
}

// ORB::detectAndCompute (orb.cpp:970) with nlevels=1: FAST -> border -> Harris -> retainBest -> IC angle
// -> blur -> descriptors.  out kp[5*i] = {x, y, response, angle, size} (floats); returns count.
int ref_orb_detect(const uint8_t* gray, int w, int h, int nfeatures, int fast_thr,
                   float* kp, uint8_t* desc, int cap) {
    cv::Mat g(h, w, CV_8UC1, (void*)gray);
    cv::Ptr<cv::ORB> orb = cv::ORB::create(nfeatures, 1.2f, 1, 31, 0, 2, cv::ORB::HARRIS_SCORE, 31, fast_thr);
    std::vector<cv::KeyPoint> kps; cv::Mat d;
    orb->detectAndCompute(g, cv::noArray(), kps, d);
    int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        kp[5 * i] = kps[i].pt.x; kp[5 * i + 1] = kps[i].pt.y; kp[5 * i + 2] = kps[i].response;
        kp[5 * i + 3] = kps[i].angle; kp[5 * i + 4] = kps[i].size;
        memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    }
    return n;
}

// cv::GaussianBlur(7x7, sigma 2) exactly as ORB applies it: in place over a ROI of a REFLECT_101
// bordered copy (orb.cpp:1102,1188).  Output = the blurred ROI (w x h).
void ref_orb_blur(const uint8_t* gray, int w, int h, uint8_t* out) {
    cv::Mat g(h, w, CV_8UC1, (void*)gray), buf;
    const int border = 32;
    cv::copyMakeBorder(g, buf, border, border, border, border, cv::BORDER_REFLECT_101);
    cv::Mat roi = buf(cv::Rect(border, border, w, h));
    cv::GaussianBlur(roi, roi, cv::Size(7, 7), 2, 2, cv::BORDER_REFLECT_101);
    for (int y = 0; y < h; y++) memcpy(out + (size_t)y * w, roi.ptr(y), w);
}

void ref_gaussian_kernel(int n, double sigma, float* out) {
    cv::Mat k = cv::getGaussianKernel(n, sigma, CV_32F);
    for (int i = 0; i < n; i++) out[i] = k.at<float>(i);
}

// cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2)  (features2d/src/matchers.cpp:757)
// out[4*i] = {idx0, dist0, idx1, dist1}; -1 when fewer than k train rows.
void ref_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* out) {
    cv::Mat Q(nq, 32, CV_8UC1, (void*)q), T(nt, 32, CV_8UC1, (void*)t);
    cv::BFMatcher bf(cv::NORM_HAMMING);
    std::vector<std::vector<cv::DMatch>> m;
    bf.knnMatch(Q, T, m, 2);
    for (int i = 0; i < nq; i++) {
        for (int k = 0; k < 2; k++) {
            if (k < (int)m[i].size()) { out[4 * i + 2 * k] = m[i][k].trainIdx; out[4 * i + 2 * k + 1] = (int)m[i][k].distance; }
            else { out[4 * i + 2 * k] = -1; out[4 * i + 2 * k + 1] = -1; }
        }
    }
}

// cv::norm(a, b, NORM_HAMMING) as MapPoint::computeMinDescDist uses it (src/slam/src/map_point.cpp:212)
int ref_hamming(const uint8_t* a, const uint8_t* b, int nbytes) {
    cv::Mat A(1, nbytes, CV_8UC1, (void*)a), B(1, nbytes, CV_8UC1, (void*)b);
    return (int)cv::norm(A, B, cv::NORM_HAMMING);
}

// DirectSE3::ReprojectionErrorKSE3AnchInvDepth::Evaluate (src/slam/src/ceres_parametrization.cpp:157-269)
// calib[4], anch[7], pose[7] ([t, qx qy qz qw]), invd; obs = {u, v, ua, va}.
// res[2]; J_anch[2x7], J_pose[2x7] row-major, J_invd[2]; returns depth-positive flag; chi2 out.
int ref_ba_evaluate(const double* calib, const double* anch, const double* pose, double invd,
                    const double* obs, double* res, double* Ja, double* Jp, double* Jd, double* chi2) {
    DirectSE3::ReprojectionErrorKSE3AnchInvDepth f(obs[0], obs[1], obs[2], obs[3], 1.);
    const double* params[4] = {calib, anch, pose, &invd};
    double Jc[8];
    double* jac[4] = {Jc, Ja, Jp, Jd};
    f.Evaluate(params, res, jac);
    *chi2 = f.chi2err_;
    return f.isDepthPositive_ ? 1 : 0;
}

// Optimizer::localBA's Ceres problem (src/slam/src/optimizer.cpp:20-262) built from flat arrays:
//   poses[nkf*7] ([t, q xyzw]), pose_const[nkf]; invd[nlm]; anchor kf per landmark anch_kf[nlm],
//   anchor pixel anch_uv[nlm*2]; observations obs_kf[nobs], obs_lm[nobs], obs_uv[nobs*2] (non-anchor).
// Options as optimizer.cpp:251-259 but with the wall-clock cap lifted: SPARSE_SCHUR, LM,
// 1 thread, max_iter, function_tolerance 1e-3.  huber_delta<=0 -> plain L2.
// Updates poses/invd in place.  summary[0]=initial cost, [1]=final cost, [2]=#successful steps,
// [3]=#iterations (incl. 0), [4]=termination type.  iter_costs: cost per iteration (cap 64).
int ref_ba_solve(const double* calib, double* poses, const uint8_t* pose_const, int nkf,
                 double* invd, const int32_t* anch_kf, const double* anch_uv, int nlm,
                 const int32_t* obs_kf, const int32_t* obs_lm, const double* obs_uv, int nobs,
                 double huber_delta, int max_iter, double* summary, double* iter_costs) {
    ceres::Problem problem;
    ceres::LossFunction* loss = huber_delta > 0 ? new ceres::HuberLoss(huber_delta) : nullptr;
    auto* ordering = new ceres::ParameterBlockOrdering;
    double K[4] = {calib[0], calib[1], calib[2], calib[3]};
    problem.AddParameterBlock(K, 4);
    ordering->AddElementToGroup(K, 1);
    problem.SetParameterBlockConstant(K);
    for (int i = 0; i < nkf; i++) {
        ceres::LocalParameterization* lp = new SE3Parameterization();
        problem.AddParameterBlock(poses + 7 * i, 7, lp);
        ordering->AddElementToGroup(poses + 7 * i, 1);
        if (pose_const[i]) problem.SetParameterBlockConstant(poses + 7 * i);
    }
    for (int l = 0; l < nlm; l++) {
        problem.AddParameterBlock(invd + l, 1);
        ordering->AddElementToGroup(invd + l, 0);
    }
    for (int o = 0; o < nobs; o++) {
        int l = obs_lm[o];
        auto* f = new DirectSE3::ReprojectionErrorKSE3AnchInvDepth(obs_uv[2 * o], obs_uv[2 * o + 1],
                                                                   anch_uv[2 * l], anch_uv[2 * l + 1], 1.);
        problem.AddResidualBlock(f, loss, K, poses + 7 * anch_kf[l], poses + 7 * obs_kf[o], invd + l);
    }
    ceres::Solver::Options options;
    options.linear_solver_ordering.reset(ordering);
    options.linear_solver_type = ceres::SPARSE_SCHUR;
    options.trust_region_strategy_type = ceres::LEVENBERG_MARQUARDT;
    options.num_threads = 1;
    options.max_num_iterations = max_iter;
    options.function_tolerance = 0.001;
    options.max_solver_time_in_seconds = 1e9;
    options.minimizer_progress_to_stdout = false;
    ceres::Solver::Summary s;
    ceres::Solve(options, &problem, &s);
    summary[0] = s.initial_cost; summary[1] = s.final_cost; summary[2] = s.num_successful_steps;
    summary[3] = (double)s.iterations.size(); summary[4] = (double)s.termination_type;
    if (iter_costs)
        for (size_t i = 0; i < s.iterations.size() && i < 64; i++) iter_costs[i] = s.iterations[i].cost;
    return s.IsSolutionUsable() ? 1 : 0;
}

// Optimizer::localBA numerics, steps 2-4 (src/slam/src/optimizer.cpp:251-359): solve (Huber), drop the residuals whose
// functor reports chi2err_ > chi2_thr or a non-positive depth AT ITS LAST EVALUATION, and -- only if something was
// dropped -- solve again (<= 5 it) and flag once more.  flags[o]: 0 kept, 1 removed after the first solve, 2 flagged
// after the second.  Wall-clock caps lifted.  summary[0..4] first solve, summary[5..9] second solve (zeros if skipped).
int ref_ba_local(const double* calib, double* poses, const uint8_t* pose_const, int nkf,
                 double* invd, const int32_t* anch_kf, const double* anch_uv, int nlm,
                 const int32_t* obs_kf, const int32_t* obs_lm, const double* obs_uv, int nobs,
                 double huber_delta, double chi2_thr, int max_iter, int32_t* flags, double* summary) {
    ceres::Problem problem;
    ceres::LossFunction* loss = huber_delta > 0 ? new ceres::HuberLoss(huber_delta) : nullptr;
    auto* ordering = new ceres::ParameterBlockOrdering;
    double K[4] = {calib[0], calib[1], calib[2], calib[3]};
    problem.AddParameterBlock(K, 4);
    ordering->AddElementToGroup(K, 1);
    problem.SetParameterBlockConstant(K);
    for (int i = 0; i < nkf; i++) {
        problem.AddParameterBlock(poses + 7 * i, 7, new SE3Parameterization());
        ordering->AddElementToGroup(poses + 7 * i, 1);
        if (pose_const[i]) problem.SetParameterBlockConstant(poses + 7 * i);
    }
    for (int l = 0; l < nlm; l++) { problem.AddParameterBlock(invd + l, 1); ordering->AddElementToGroup(invd + l, 0); }
    std::vector<DirectSE3::ReprojectionErrorKSE3AnchInvDepth*> fs(nobs, nullptr);
    std::vector<ceres::ResidualBlockId> rids(nobs);
    for (int o = 0; o < nobs; o++) {
        flags[o] = 0;
        int l = obs_lm[o];
        if (l < 0) continue;
        fs[o] = new DirectSE3::ReprojectionErrorKSE3AnchInvDepth(obs_uv[2 * o], obs_uv[2 * o + 1], anch_uv[2 * l], anch_uv[2 * l + 1], 1.);
        rids[o] = problem.AddResidualBlock(fs[o], loss, K, poses + 7 * anch_kf[l], poses + 7 * obs_kf[o], invd + l);
    }
    ceres::Solver::Options options;
    options.linear_solver_ordering.reset(ordering);
    options.linear_solver_type = ceres::SPARSE_SCHUR;
    options.trust_region_strategy_type = ceres::LEVENBERG_MARQUARDT;
    options.num_threads = 1;
    options.max_num_iterations = max_iter;
    options.function_tolerance = 0.001;
    options.max_solver_time_in_seconds = 1e9;
    ceres::Solver::Summary s;
    ceres::Solve(options, &problem, &s);
    for (int i = 0; i < 10; i++) summary[i] = 0;
    summary[0] = s.initial_cost; summary[1] = s.final_cost; summary[2] = s.num_successful_steps;
    summary[3] = (double)s.iterations.size(); summary[4] = (double)s.termination_type;
    int nbad = 0;
    for (int o = 0; o < nobs; o++) {
        if (!fs[o]) continue;
        if (fs[o]->chi2err_ > chi2_thr || !fs[o]->isDepthPositive_) { problem.RemoveResidualBlock(rids[o]); fs[o] = nullptr; flags[o] = 1; nbad++; }
    }
    if (huber_delta > 0 && nbad > 0) {
        options.max_num_iterations = 5;
        ceres::Solver::Summary s2;
        ceres::Solve(options, &problem, &s2);
        summary[5] = s2.initial_cost; summary[6] = s2.final_cost; summary[7] = s2.num_successful_steps;
        summary[8] = (double)s2.iterations.size(); summary[9] = (double)s2.termination_type;
        for (int o = 0; o < nobs; o++)
            if (fs[o] && (fs[o]->chi2err_ > chi2_thr || !fs[o]->isDepthPositive_)) flags[o] = 2;
    }
    return nbad;
}

// SE3Parameterization::Plus (src/slam/src/ceres_parametrization.hpp:224-240)
void ref_se3_plus(const double* x, const double* delta, double* out) {
    SE3Parameterization p; p.Plus(x, delta, out);
}

// VisualFrontend::preprocessImage + kltTrackingFromMotionPrior's tracker call, unmodified reference code:
//   cv::buildOpticalFlowPyramid(prev / cur, win, pyr_levels)                 src/slam/src/visual_frontend.cpp:696
//   FeatureTracker(30, 0.01).fbKltTracking(prevPyr, curPyr, win, levels, error_value, max_fb_dist, pts, priors, status)
//                                                                            src/slam/src/feature_tracker.cpp:5-111
// pts [n][2] previous positions; priors [n][2] in/out; good [n] out.  Returns the pyramid's max level.
int ref_fb_klt(const uint8_t* prev, const uint8_t* cur, int w, int h, int win, int pyr_levels, int levels,
               float error_value, float max_fb_dist, const float* pts, float* priors, uint8_t* good, int n) {
    cv::Mat gp(h, w, CV_8UC1, (void*)prev), gc(h, w, CV_8UC1, (void*)cur);
    std::vector<cv::Mat> pp, pc;
    int got = cv::buildOpticalFlowPyramid(gp, pp, cv::Size(win, win), pyr_levels);
    cv::buildOpticalFlowPyramid(gc, pc, cv::Size(win, win), pyr_levels);
    std::vector<cv::Point2f> p(n), q(n);
    for (int i = 0; i < n; i++) { p[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]); q[i] = cv::Point2f(priors[2 * i], priors[2 * i + 1]); }
    std::vector<bool> st;
    FeatureTracker tracker(30, 0.01f);
    tracker.fbKltTracking(pp, pc, win, levels, error_value, max_fb_dist, p, q, st);
    for (int i = 0; i < n; i++) {
        priors[2 * i] = q[i].x; priors[2 * i + 1] = q[i].y;
        good[i] = (i < (int)st.size() && st[i]) ? 1 : 0;
    }
    return got;
}

// cv::calcOpticalFlowPyrLK on prebuilt pyramids with the flags AlvaAR uses (feature_tracker.cpp:35-38):
// USE_INITIAL_FLOW (optional) + LK_GET_MIN_EIGENVALS, TermCriteria(COUNT+EPS, max_count, epsilon).
int ref_klt_lk(const uint8_t* prev, const uint8_t* cur, int w, int h, int win, int pyr_levels, int levels, int max_count,
               double epsilon, int use_initial, const float* pts, float* next, uint8_t* status, float* err, int n) {
    cv::Mat gp(h, w, CV_8UC1, (void*)prev), gc(h, w, CV_8UC1, (void*)cur);
    std::vector<cv::Mat> pp, pc;
    int got = cv::buildOpticalFlowPyramid(gp, pp, cv::Size(win, win), pyr_levels);
    cv::buildOpticalFlowPyramid(gc, pc, cv::Size(win, win), pyr_levels);
    std::vector<cv::Point2f> p(n), q(n);
    for (int i = 0; i < n; i++) { p[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]); q[i] = cv::Point2f(next[2 * i], next[2 * i + 1]); }
    std::vector<uchar> st; std::vector<float> er;
    cv::calcOpticalFlowPyrLK(pp, pc, p, q, st, er, cv::Size(win, win), levels,
                             cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, max_count, epsilon),
                             (use_initial ? cv::OPTFLOW_USE_INITIAL_FLOW : 0) + cv::OPTFLOW_LK_GET_MIN_EIGENVALS);
    for (int i = 0; i < n; i++) { next[2 * i] = q[i].x; next[2 * i + 1] = q[i].y; status[i] = st[i]; err[i] = er[i]; }
    return got;
}

// MultiViewGeometry::p3pRansac (src/slam/src/multi_view_geometry.cpp:24-127), unmodified, called as
// VisualFrontend::computePose does (visual_frontend.cpp:299-312): optimize = false; doRandom = false pins the sampler's
// seed to 12345 (SampleConsensusProblem.hpp:43-46).  bvs/wpts [n][3]; Twc_out 3x4 row-major [R | t]; outlier [n].
int ref_p3p_lmeds(const double* bvs, const double* wpts, int n, int max_iter, float err_px, float fx, float fy,
                  double* Twc_out, uint8_t* outlier) {
    std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> b(n), w(n);
    for (int i = 0; i < n; i++) { b[i] = Eigen::Vector3d(bvs[3 * i], bvs[3 * i + 1], bvs[3 * i + 2]); w[i] = Eigen::Vector3d(wpts[3 * i], wpts[3 * i + 1], wpts[3 * i + 2]); }
    Sophus::SE3d Twc;
    std::vector<int> out;
    bool ok = MultiViewGeometry::p3pRansac(b, w, max_iter, err_px, false, false, fx, fy, Twc, out);
    memset(outlier, 0, n);
    for (int i : out) outlier[i] = 1;
    Eigen::Matrix3d R = Twc.rotationMatrix();
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Twc_out[4 * i + j] = R(i, j); Twc_out[4 * i + 3] = Twc.translation()[i]; }
    return ok ? 1 : 0;
}

// MultiViewGeometry::ceresPnP (src/slam/src/multi_view_geometry.cpp:129-223), unmodified (its 5 ms wall-clock cap stays:
// the test problems solve in well under a millisecond).  pose [t, q(x,y,z,w)] in/out; uv [n][2]; X [n][3]; outlier [n].
int ref_pnp(const double* uv, const double* X, int n, double* pose, int max_iter, float chi2th, int use_robust, int apply_l2,
            float fx, float fy, float cx, float cy, uint8_t* outlier) {
    std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d>> k(n);
    std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> w(n);
    for (int i = 0; i < n; i++) { k[i] = Eigen::Vector2d(uv[2 * i], uv[2 * i + 1]); w[i] = Eigen::Vector3d(X[3 * i], X[3 * i + 1], X[3 * i + 2]); }
    Sophus::SE3d Twc(Eigen::Quaterniond(pose[6], pose[3], pose[4], pose[5]), Eigen::Vector3d(pose[0], pose[1], pose[2]));
    std::vector<int> out;
    bool ok = MultiViewGeometry::ceresPnP(k, w, Twc, max_iter, chi2th, use_robust != 0, apply_l2 != 0, fx, fy, cx, cy, out);
    memset(outlier, 0, n);
    for (int i : out) outlier[i] = 1;
    for (int i = 0; i < 3; i++) pose[i] = Twc.translation()[i];
    const Eigen::Quaterniond& q = Twc.unit_quaternion();
    pose[3] = q.x(); pose[4] = q.y(); pose[5] = q.z(); pose[6] = q.w();
    return ok ? 1 : 0;
}

// FeatureExtractor(max_quality).detectFeaturePoints(image, cell, currKeypoints, roi)  -- unmodified reference code
// (src/slam/src/feature_extractor.cpp:11-158), cv::setNumThreads(1) through ref_config.  The adapted maxQuality_ is private;
// it is recovered by replaying the reference's own rule (:138-145) on the returned count -- n_occ is returned for that.
// cur [ncur][2]; roi {x, y, w, h}; out [cap][2] (sub-pixel positions).  Returns the number of points.
int ref_detect_points(const uint8_t* gray, int w, int h, int cell, const float* cur, int ncur, const int* roi,
                      double max_quality, float* out, int cap) {
    cv::Mat g(h, w, CV_8UC1, (void*)gray);
    std::vector<cv::Point2f> c(ncur);
    for (int i = 0; i < ncur; i++) c[i] = cv::Point2f(cur[2 * i], cur[2 * i + 1]);
    FeatureExtractor fe(max_quality);
    std::vector<cv::Point2f> pts = fe.detectFeaturePoints(g, cell, c, cv::Rect(roi[0], roi[1], roi[2], roi[3]));
    for (int i = 0; i < (int)pts.size() && i < cap; i++) { out[2 * i] = pts[i].x; out[2 * i + 1] = pts[i].y; }
    return (int)pts.size();
}

// the cell pipeline's intermediate, for pinning the float stages: GaussianBlur(3x3) on the ROI of the full image followed by
// cornerMinEigenVal(block 3, Sobel 3), exactly as feature_extractor.cpp:64-70 calls them.  hmap [cell][cell].
void ref_min_eig_cell(const uint8_t* gray, int w, int h, int x0, int y0, int cell, float* hmap, uint8_t* blurred) {
    cv::Mat g(h, w, CV_8UC1, (void*)gray), f, m;
    cv::GaussianBlur(g(cv::Rect(x0, y0, cell, cell)), f, cv::Size(3, 3), 0.);
    cv::cornerMinEigenVal(f, m, 3, 3);
    for (int y = 0; y < cell; y++) { memcpy(hmap + (size_t)y * cell, m.ptr<float>(y), sizeof(float) * cell); if (blurred) memcpy(blurred + (size_t)y * cell, f.ptr(y), cell); }
}

// cv::cornerSubPix(image, pts, Size(win, win), Size(-1, -1), TermCriteria(EPS + MAX_ITER, max_iter, eps))
void ref_corner_subpix(const uint8_t* gray, int w, int h, float* pts, int n, int win, int max_iter, double eps) {
    cv::Mat g(h, w, CV_8UC1, (void*)gray);
    std::vector<cv::Point2f> p(n);
    for (int i = 0; i < n; i++) p[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    cv::cornerSubPix(g, p, cv::Size(win, win), cv::Size(-1, -1), cv::TermCriteria(cv::TermCriteria::EPS + cv::TermCriteria::MAX_ITER, max_iter, eps));
    for (int i = 0; i < n; i++) { pts[2 * i] = p[i].x; pts[2 * i + 1] = p[i].y; }
}

// MultiViewGeometry::compute5ptEssentialMatrix (src/slam/src/multi_view_geometry.cpp:225-318), unmodified, called as
// VisualFrontend::checkReadyForInit does (visual_frontend.cpp:517-528): OpenGV's Ransac over CentralRelativePoseSacProblem
// (NISTER), doRandom = false (seed 12345).  bv1 = keyframe bearing vectors, bv2 = current frame's, [n][3].
// Rt_out: 3x4 row-major [Rwc | twc] (twc NOT yet normalised); outlier [n].  Returns 1 on success.
int ref_essential_5pt(const double* bv1, const double* bv2, int n, int max_iter, float err_px, int optimize, float fx, float fy,
                      double* Rt_out, uint8_t* outlier) {
    std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> a(n), b(n);
    for (int i = 0; i < n; i++) { a[i] = Eigen::Vector3d(bv1[3 * i], bv1[3 * i + 1], bv1[3 * i + 2]); b[i] = Eigen::Vector3d(bv2[3 * i], bv2[3 * i + 1], bv2[3 * i + 2]); }
    Eigen::Matrix3d R = Eigen::Matrix3d::Identity();
    Eigen::Vector3d t = Eigen::Vector3d::Zero();
    std::vector<int> out;
    bool ok = MultiViewGeometry::compute5ptEssentialMatrix(a, b, max_iter, err_px, optimize != 0, false, fx, fy, R, t, out);
    memset(outlier, 0, n);
    for (int i : out) outlier[i] = 1;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Rt_out[4 * i + j] = R(i, j); Rt_out[4 * i + 3] = t[i]; }
    return ok ? 1 : 0;
}

// MultiViewGeometry::triangulate (src/slam/src/multi_view_geometry.cpp:12-22): OpenGV triangulate2 (mid-point) of n pairs.
// Tlr = [t, q(x,y,z,w)] (left <- right); bvl / bvr [n][3]; out [n][3] in the left camera frame.
void ref_triangulate(const double* Tlr, const double* bvl, const double* bvr, int n, double* out) {
    Sophus::SE3d T(Eigen::Quaterniond(Tlr[6], Tlr[3], Tlr[4], Tlr[5]), Eigen::Vector3d(Tlr[0], Tlr[1], Tlr[2]));
    for (int i = 0; i < n; i++) {
        Eigen::Vector3d p = MultiViewGeometry::triangulate(T, Eigen::Vector3d(bvl[3 * i], bvl[3 * i + 1], bvl[3 * i + 2]),
                                                           Eigen::Vector3d(bvr[3 * i], bvr[3 * i + 1], bvr[3 * i + 2]));
        out[3 * i] = p[0]; out[3 * i + 1] = p[1]; out[3 * i + 2] = p[2];
    }
}

}  // extern "C"
