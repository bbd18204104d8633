/* oracle/klt_oracle.c -- CPU restatement of the reference's per-frame keypoint association:
 * pyramidal Lucas-Kanade (cv::calcOpticalFlowPyrLK on prebuilt pyramids) and AlvaAR's forward-backward
 * wrapper around it.
 *
 * TEST INFRASTRUCTURE ONLY (see alva_oracle.c's header): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call this.  The product (alvaar_b200/) never links or executes it.
 *
 * Reference (paths under /root/reference):
 *   FeatureTracker::fbKltTracking            src/slam/src/feature_tracker.cpp:5-111  (inBorder :113-119)
 *   caller VisualFrontend::kltTrackingFromMotionPrior   src/slam/src/visual_frontend.cpp:103-243
 *   cv::calcOpticalFlowPyrLK / SparsePyrLKOpticalFlowImpl::calc   src/libs/opencv/modules/video/src/lkpyramid.cpp:1238-1398
 *   cv::detail::LKTrackerInvoker::operator()                      lkpyramid.cpp:183-722
 *
 * PINNED: tests/test_oracle_klt.py checks it bit-for-bit (positions as float bit patterns, status, err) against
 * golden vectors dumped from the reference's own FeatureTracker + vendored OpenCV 4.5.5 (tools/make_golden_klt.py
 * through oracle/_ref/libalva_ref.so) and, when that library is present, against the live reference.
 *
 * The float sums of the reference are ORDER-sensitive (integers up to 2^25 accumulated in float32), and the
 * reference's order is the one its SSE universal-intrinsics code path produces (CV_SIMD128 is a compile-time
 * switch, so cv::setUseOptimized does not change it):
 *   - window columns are processed in blocks of 8 by four float lanes, the remainder (column 8 of a 9-wide window)
 *     by a scalar accumulator; v_reduce_sum adds lanes as (l0 + l2) + (l1 + l3)          (intrin_sse.hpp:1690-1697)
 *   - the mismatch vector b accumulates, per 8-column block, int32 pair sums (column c and c + 4) converted to float
 *     (v_dotprod + v_cvt_f32, lkpyramid.cpp:536-562).
 * Pyramid levels are passed tightly packed; the 9-px REFLECT_101 image border and the constant-0 derivative border
 * that buildOpticalFlowPyramid stores around each level (lkpyramid.cpp:726-822) are produced by index arithmetic.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define KLT_MAX_WIN 21

static inline int k_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}
static inline int k_floor(float v) { return (int)floorf(v); }
static inline int k_round(float v) { return (int)lrintf(v); }   /* cvRound(float): cvtss2si, half to even */
#define K_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

static inline int img_at(const uint8_t* img, int w, int h, int x, int y)
{
    return img[(size_t)k_reflect101(y, h) * w + k_reflect101(x, w)];
}
static inline int der_at(const int16_t* d, int w, int h, int x, int y, int c)
{
    if (x < 0 || x >= w || y < 0 || y >= h) return 0;   /* derivBorder = BORDER_CONSTANT (tracking.hpp:121-125) */
    return d[((size_t)y * w + x) * 2 + c];
}

/* float accumulators laid out as the SSE path keeps them: four lanes + the scalar remainder */
typedef struct { float q[4]; float s; } acc5;
static inline float acc5_total(const acc5* a) { return a->s + ((a->q[0] + a->q[2]) + (a->q[1] + a->q[3])); }

/* One point on one level: LKTrackerInvoker::operator() body (lkpyramid.cpp:203-720).
 * next[2] in/out (the nextPts entry), status/err in/out. */
static void lk_point_level(const uint8_t* I, const int16_t* dI, const uint8_t* J, int w, int h, int level, int max_level,
                           int use_initial, float px, float py, float* next, uint8_t* status, float* err, int win,
                           int max_count, double eps2, float min_eig_thr)
{
    const float half = (float)(win - 1) * 0.5f;
    const float scale = (float)(1. / (1 << level));
    float prevx = px * scale, prevy = py * scale;
    float nx, ny;
    if (level == max_level) {
        if (use_initial) { nx = next[0] * scale; ny = next[1] * scale; }
        else { nx = prevx; ny = prevy; }
    } else {
        nx = next[0] * 2.f; ny = next[1] * 2.f;
    }
    next[0] = nx; next[1] = ny;

    prevx -= half; prevy -= half;
    const int ipx = k_floor(prevx), ipy = k_floor(prevy);
    if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) {
        if (level == 0) { *status = 0; *err = 0.f; }
        return;
    }
    float a = prevx - (float)ipx, b = prevy - (float)ipy;
    const int W_BITS = 14;
    const float FLT_SCALE = 1.f / (1 << 20);
    int iw00 = k_round((1.f - a) * (1.f - b) * (1 << W_BITS));
    int iw01 = k_round(a * (1.f - b) * (1 << W_BITS));
    int iw10 = k_round((1.f - a) * b * (1 << W_BITS));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;

    short Iw[KLT_MAX_WIN * KLT_MAX_WIN], Dx[KLT_MAX_WIN * KLT_MAX_WIN], Dy[KLT_MAX_WIN * KLT_MAX_WIN];
    acc5 A11 = {{0, 0, 0, 0}, 0}, A12 = A11, A22 = A11;
    const int nblk = win / 8;
    for (int y = 0; y < win; y++) {
        for (int x = 0; x < win; x++) {
            const int X = ipx + x, Y = ipy + y;
            const int ival = K_DESCALE(img_at(I, w, h, X, Y) * iw00 + img_at(I, w, h, X + 1, Y) * iw01 +
                                       img_at(I, w, h, X, Y + 1) * iw10 + img_at(I, w, h, X + 1, Y + 1) * iw11, W_BITS - 5);
            const int ix = K_DESCALE(der_at(dI, w, h, X, Y, 0) * iw00 + der_at(dI, w, h, X + 1, Y, 0) * iw01 +
                                     der_at(dI, w, h, X, Y + 1, 0) * iw10 + der_at(dI, w, h, X + 1, Y + 1, 0) * iw11, W_BITS);
            const int iy = K_DESCALE(der_at(dI, w, h, X, Y, 1) * iw00 + der_at(dI, w, h, X + 1, Y, 1) * iw01 +
                                     der_at(dI, w, h, X, Y + 1, 1) * iw10 + der_at(dI, w, h, X + 1, Y + 1, 1) * iw11, W_BITS);
            Iw[y * win + x] = (short)ival; Dx[y * win + x] = (short)ix; Dy[y * win + x] = (short)iy;
        }
        /* covariance sums in the reference's order: per block of 8 columns, lanes take columns (l) then (l + 4) */
        for (int blk = 0; blk < nblk; blk++)
            for (int half8 = 0; half8 < 2; half8++)
                for (int l = 0; l < 4; l++) {
                    const int x = blk * 8 + half8 * 4 + l;
                    const float fx = (float)Dx[y * win + x], fy = (float)Dy[y * win + x];
                    A22.q[l] = fy * fy + A22.q[l];
                    A12.q[l] = fx * fy + A12.q[l];
                    A11.q[l] = fx * fx + A11.q[l];
                }
        for (int x = nblk * 8; x < win; x++) {
            const int ix = Dx[y * win + x], iy = Dy[y * win + x];
            A11.s += (float)(ix * ix);
            A12.s += (float)(ix * iy);
            A22.s += (float)(iy * iy);
        }
    }
    const float a11 = acc5_total(&A11) * FLT_SCALE, a12 = acc5_total(&A12) * FLT_SCALE, a22 = acc5_total(&A22) * FLT_SCALE;
    float D = a11 * a22 - a12 * a12;
    const float min_eig = (a22 + a11 - sqrtf((a11 - a22) * (a11 - a22) + 4.f * a12 * a12)) / (float)(2 * win * win);
    *err = min_eig;   /* OPTFLOW_LK_GET_MIN_EIGENVALS (the only mode AlvaAR uses, feature_tracker.cpp:35,84) */
    if (min_eig < min_eig_thr || D < FLT_EPSILON) {
        if (level == 0) *status = 0;
        return;
    }
    D = 1.f / D;
    nx -= half; ny -= half;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < max_count; j++) {
        const int inx = k_floor(nx), iny = k_floor(ny);
        if (inx < -win || inx >= w || iny < -win || iny >= h) {
            if (level == 0) *status = 0;
            break;
        }
        a = nx - (float)inx; b = ny - (float)iny;
        iw00 = k_round((1.f - a) * (1.f - b) * (1 << W_BITS));
        iw01 = k_round(a * (1.f - b) * (1 << W_BITS));
        iw10 = k_round((1.f - a) * b * (1 << W_BITS));
        iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        float qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0}, ib1 = 0.f, ib2 = 0.f;
        for (int y = 0; y < win; y++) {
            int diff[KLT_MAX_WIN];
            for (int x = 0; x < win; x++) {
                const int X = inx + x, Y = iny + y;
                diff[x] = K_DESCALE(img_at(J, w, h, X, Y) * iw00 + img_at(J, w, h, X + 1, Y) * iw01 +
                                    img_at(J, w, h, X, Y + 1) * iw10 + img_at(J, w, h, X + 1, Y + 1) * iw11, W_BITS - 5) -
                          Iw[y * win + x];
                diff[x] = (short)diff[x];   /* v_pack(t0, t1) - diff0 is int16 arithmetic; |diff| <= 8160 never wraps */
            }
            for (int blk = 0; blk < nblk; blk++) {
                const int x0 = blk * 8;
                const short* dx = Dx + y * win + x0;
                const short* dy = Dy + y * win + x0;
                const int* d = diff + x0;
                qb0[0] += (float)(d[0] * dx[0] + d[4] * dx[4]);
                qb0[1] += (float)(d[0] * dy[0] + d[4] * dy[4]);
                qb0[2] += (float)(d[1] * dx[1] + d[5] * dx[5]);
                qb0[3] += (float)(d[1] * dy[1] + d[5] * dy[5]);
                qb1[0] += (float)(d[2] * dx[2] + d[6] * dx[6]);
                qb1[1] += (float)(d[2] * dy[2] + d[6] * dy[6]);
                qb1[2] += (float)(d[3] * dx[3] + d[7] * dx[7]);
                qb1[3] += (float)(d[3] * dy[3] + d[7] * dy[7]);
            }
            for (int x = nblk * 8; x < win; x++) {
                ib1 += (float)(diff[x] * Dx[y * win + x]);
                ib2 += (float)(diff[x] * Dy[y * win + x]);
            }
        }
        {
            float s[4];
            for (int l = 0; l < 4; l++) s[l] = qb0[l] + qb1[l];
            ib1 += (s[0] + 0.f) + (s[2] + 0.f);
            ib2 += (s[1] + 0.f) + (s[3] + 0.f);
        }
        const float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
        const float dx = (a12 * b2 - a22 * b1) * D;
        const float dy = (a12 * b1 - a11 * b2) * D;
        nx += dx; ny += dy;
        next[0] = nx + half; next[1] = ny + half;
        if ((double)dx * dx + (double)dy * dy <= eps2) break;
        if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
            next[0] -= dx * 0.5f; next[1] -= dy * 0.5f;
            break;
        }
        pdx = dx; pdy = dy;
    }
}

/* cv::calcOpticalFlowPyrLK(prevPyr, nextPyr, prevPts, nextPts, status, err, Size(win,win), max_level,
 *                          TermCriteria(COUNT+EPS, max_count, epsilon), USE_INITIAL_FLOW | LK_GET_MIN_EIGENVALS, 1e-4)
 * on prebuilt pyramids with derivatives.  imgs/derivs: per level (0..max_level) tightly packed u8 / int16x2 arrays of
 * size w_k x h_k, w_k = (w_{k-1}+1)/2.  next: in (initial flow) / out.  use_initial = 0 starts from prevPts. */
void orc_klt_lk(const uint8_t* const* prev_img, const int16_t* const* prev_deriv, const uint8_t* const* next_img, int w0,
                int h0, int max_level, const float* prev_pts, float* next_pts, uint8_t* status, float* err, int n, int win,
                int max_count, double epsilon, int use_initial, double min_eig_thr)
{
    if (max_count < 0) max_count = 0;
    if (max_count > 100) max_count = 100;
    if (epsilon < 0) epsilon = 0;
    if (epsilon > 10) epsilon = 10;
    const double eps2 = epsilon * epsilon;
    int ws[16], hs[16];
    ws[0] = w0; hs[0] = h0;
    for (int k = 1; k <= max_level; k++) { ws[k] = (ws[k - 1] + 1) / 2; hs[k] = (hs[k - 1] + 1) / 2; }
    for (int i = 0; i < n; i++) status[i] = 1;
    for (int level = max_level; level >= 0; level--)
        for (int i = 0; i < n; i++)
            lk_point_level(prev_img[level], prev_deriv[level], next_img[level], ws[level], hs[level], level, max_level,
                           use_initial, prev_pts[2 * i], prev_pts[2 * i + 1], next_pts + 2 * i, status + i, err + i, win,
                           max_count, eps2, (float)min_eig_thr);
}

/* FeatureTracker::fbKltTracking (src/slam/src/feature_tracker.cpp:5-111) with kltConvCriteria_ = (COUNT+EPS, 30, 0.01)
 * (feature_tracker.hpp:14): forward LK on `levels` pyramid levels from the priors, gates (status, err > error_value,
 * inBorder 1 px), backward LK on level 0 only from the original position, |p - back| > max_fb_dist gate.
 * pts [n][2] (previous-frame positions, unchanged), priors [n][2] in/out (= the tracked positions, written for every
 * point like the reference's priorKeypoints), good [n] out.  npyr_levels = levels the pyramids actually hold - 1. */
void orc_fb_klt(const uint8_t* const* prev_img, const int16_t* const* prev_deriv, const uint8_t* const* cur_img,
                const int16_t* const* cur_deriv, int w0, int h0, int npyr_levels, int levels, int win, float error_value,
                float max_fb_dist, const float* pts, float* priors, uint8_t* good, int n, int max_count, double epsilon)
{
    if (n <= 0) return;
    if (levels > npyr_levels) levels = npyr_levels;   /* feature_tracker.cpp:18-21 */
    uint8_t* st = (uint8_t*)malloc((size_t)n);
    float* er = (float*)malloc(sizeof(float) * (size_t)n);
    orc_klt_lk(prev_img, prev_deriv, cur_img, w0, h0, levels, pts, priors, st, er, n, win, max_count, epsilon, 1, 1e-4);
    for (int i = 0; i < n; i++) {
        const float x = priors[2 * i], y = priors[2 * i + 1];
        good[i] = st[i] && !(er[i] > error_value) &&
                  (1.0f <= x && x < (float)w0 - 1.0f && 1.0f <= y && y < (float)h0 - 1.0f);
    }
    for (int i = 0; i < n; i++) {
        if (!good[i]) continue;
        float back[2] = {pts[2 * i], pts[2 * i + 1]};
        uint8_t s = 1;
        float e = 0.f;
        const double eps = epsilon < 0 ? 0 : (epsilon > 10 ? 10 : epsilon);
        int mc = max_count < 0 ? 0 : (max_count > 100 ? 100 : max_count);
        lk_point_level(cur_img[0], cur_deriv[0], prev_img[0], w0, h0, 0, 0, 1, priors[2 * i], priors[2 * i + 1], back, &s, &e,
                       win, mc, eps * eps, 1e-4f);
        if (!s) { good[i] = 0; continue; }
        /* cv::norm(Point2f): sqrt((double)x*x + (double)y*y)  (core/include/opencv2/core/types.hpp) */
        const float ddx = pts[2 * i] - back[0], ddy = pts[2 * i + 1] - back[1];
        if (sqrt((double)ddx * ddx + (double)ddy * ddy) > max_fb_dist) good[i] = 0;
    }
    free(st); free(er);
}
