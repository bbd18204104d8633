/* oracle/alva_oracle.c -- CPU restatement of the reference's per-frame hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs are the only callers.  Nothing under
 * alvaar_b200/ (the product) links, imports or executes it; the product fails loudly without its
 * CUDA library.
 *
 * Every function restates, in plain scalar C, the arithmetic of one reference function (file:line
 * given per function; paths relative to /root/reference).  The oracle is PINNED: tests/test_oracle.py
 * checks it bit-for-bit against golden vectors dumped from the reference's own vendored OpenCV 4.5.5
 * (tools/make_golden.py through oracle/_ref/libalva_ref.so) and, when that library is present, against
 * the live reference on fresh random inputs.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).  -ffp-contract=off matters: the
 * float paths below are order- and fusion-sensitive.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

/* cv::borderInterpolate(BORDER_REFLECT_101): pattern  d c b | a b c d ... | c b a
 * (src/libs/opencv/modules/core/src/copy.cpp, borderInterpolate) */
static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

/* cvRound(float|double): round-half-to-even (SSE cvtss2si) -- core/include/opencv2/core/fast_math.hpp */
static inline int cv_roundf(float v) { return (int)lrintf(v); }

/* ------------------------------------------------------------------------------------------------
 * RGBA -> gray.   cv::cvtColor(COLOR_RGBA2GRAY) as called by System::findCameraPose
 * (src/slam/src/system.cpp:111-112); 8-bit fixed point RGB2Gray<uchar>
 * (src/libs/opencv/modules/imgproc/src/color_rgb.simd.hpp:650-680, constants
 * color.simd_helpers.hpp:16-24: RY15 9798, GY15 19235, BY15 3735, shift 15).
 * ---------------------------------------------------------------------------------------------- */
void orc_gray(const uint8_t* rgba, int w, int h, uint8_t* gray)
{
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) {
        int r = rgba[4 * i], g = rgba[4 * i + 1], b = rgba[4 * i + 2];
        gray[i] = (uint8_t)((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15);
    }
}

/* ------------------------------------------------------------------------------------------------
 * One Gaussian-pyramid level.  cv::pyrDown 8U (src/libs/opencv/modules/imgproc/src/pyramids.cpp:
 * 1260-1303 -> pyrDown_<FixPtCast<uchar,8>> :745-781, PyrDownInvoker :783-900, cast :53-58).
 * Separable [1 4 6 4 1] on even columns then even rows, integer, no intermediate rounding,
 * dst = (sum + 128) >> 8, BORDER_REFLECT_101, dst size ((w+1)/2, (h+1)/2).
 * This is also the level step of cv::buildOpticalFlowPyramid (video/src/lkpyramid.cpp:726-822),
 * which VisualFrontend::preprocessImage calls (src/slam/src/visual_frontend.cpp:696).
 * ---------------------------------------------------------------------------------------------- */
void orc_pyrdown(const uint8_t* src, int w, int h, uint8_t* dst)
{
    static const int k[5] = {1, 4, 6, 4, 1};
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    for (int y = 0; y < dh; y++) {
        for (int x = 0; x < dw; x++) {
            int s = 0;
            for (int j = -2; j <= 2; j++) {
                const uint8_t* row = src + (size_t)reflect101(2 * y + j, h) * w;
                int rs = 0;
                for (int i = -2; i <= 2; i++) rs += k[i + 2] * row[reflect101(2 * x + i, w)];
                s += k[j + 2] * rs;
            }
            dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
        }
    }
}

/* Number of pyramid levels cv::buildOpticalFlowPyramid produces beyond level 0
 * (lkpyramid.cpp:811-816: stop when the level is not larger than the LK window). */
int orc_pyramid_levels(int w, int h, int win, int max_level)
{
    int lv = 0;
    while (lv < max_level) {
        int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win || nh <= win) break;
        w = nw; h = nh; lv++;
    }
    return lv;
}

/* ------------------------------------------------------------------------------------------------
 * FAST-9/16 + score + 3x3 NMS.  cv::FAST(img, kps, thr, nms, TYPE_9_16)
 * (src/libs/opencv/modules/features2d/src/fast.cpp:57-292 FAST_t<16>; ring offsets
 * fast_score.cpp:50-81; score cornerScore<16> fast_score.cpp:119-210).
 *   corner  : >= 9 contiguous ring pixels all > v+t or all < v-t, rows/cols 3..n-4
 *   score   : max threshold for which it stays a corner = max(max_arc min(v-ring), max_arc min(ring-v)) - 1
 *   nms     : strict > against the 8 neighbours' scores (0 where not a corner)
 *   order   : row-major (y, then x).   xys[3*i] = {x, y, score}.  Returns the true count.
 * ---------------------------------------------------------------------------------------------- */
static const int fast_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int fast_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* returns score+1 (i.e. max(A,B)) if corner at threshold thr, else 0 */
static int fast_corner_strength(const uint8_t* p, int stride, int thr)
{
    int v = p[0], d[25];
    for (int k = 0; k < 16; k++) d[k] = v - p[fast_dy[k] * stride + fast_dx[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    /* any 9-arc holds one pixel of every antipodal pair: cheap exact reject (fast.cpp:204-221) */
    for (int k = 0; k < 8; k++)
        if (abs(d[k]) <= thr && abs(d[k + 8]) <= thr) return 0;
    int best = -1000;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int j = 1; j < 9; j++) {
            if (d[s + j] < mn) mn = d[s + j];
            if (d[s + j] > mx) mx = d[s + j];
        }
        if (mn > best) best = mn;      /* dark arc: all ring < v - t  <=> min(v - ring) > t */
        if (-mx > best) best = -mx;    /* bright arc */
    }
    return best > thr ? best : 0;
}

int orc_fast9(const uint8_t* img, int w, int h, int thr, int nms, int32_t* xys, int cap)
{
    if (thr < 0) thr = 0;
    if (thr > 255) thr = 255;
    if (w < 7 || h < 7) return 0;
    uint8_t* score = (uint8_t*)calloc((size_t)w * h, 1);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = fast_corner_strength(img + (size_t)y * w + x, w, thr);
            if (s) score[(size_t)y * w + x] = (uint8_t)(s - 1); /* uchar score row, fast.cpp:198,241 */
        }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = fast_corner_strength(img + (size_t)y * w + x, w, thr);
            if (!s) continue;
            int sc = s - 1;
            if (nms) {
                const uint8_t* c = score + (size_t)y * w + x;
                if (!(sc > c[-1] && sc > c[1] && sc > c[-w - 1] && sc > c[-w] && sc > c[-w + 1] &&
                      sc > c[w - 1] && sc > c[w] && sc > c[w + 1]))
                    continue;
            }
            if (n < cap) { xys[3 * n] = x; xys[3 * n + 1] = y; xys[3 * n + 2] = nms ? sc : 0; }
            n++;
        }
    free(score);
    return n;
}

/* dense variant used by the GPU parity tests: score map (0 = not a corner, else score) */
void orc_fast9_scoremap(const uint8_t* img, int w, int h, int thr, uint8_t* score)
{
    memset(score, 0, (size_t)w * h);
    if (w < 7 || h < 7) return;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = fast_corner_strength(img + (size_t)y * w + x, w, thr);
            if (s) score[(size_t)y * w + x] = (uint8_t)(s - 1);
        }
}

/* ------------------------------------------------------------------------------------------------
 * ORB pre-blur.  ORB_Impl::detectAndCompute copies the image into a 32-px BORDER_REFLECT_101 buffer
 * (features2d/src/orb.cpp:985-989,1102) and runs GaussianBlur(7x7, sigma 2) in place over the image
 * ROI (orb.cpp:1188).  The ROI is a sub-matrix, so the bit-exact fixed-point branch is skipped
 * (imgproc/src/smooth.dispatch.cpp:654) and the generic separable engine runs with a float kernel
 * (filter.dispatch.cpp:290-345 rejects the 8-bit kernel): RowFilter<uchar,float> then
 * SymmColumnFilter<Cast<float,uchar>> (filter.simd.hpp:2431-2490, 2697-2790; vector bodies :468-510,
 * :1163-1215), saturate_cast<uchar> = round-half-even.
 *   row    : s = k0*S[0]; s += k[i]*S[i]  (left to right)
 *   column : s = k3*R[0];  s += k[3+j]*(R[+j] + R[-j]), j = 1..3
 * fused = 0 : separate multiply and add (SSE3 baseline dispatch, and the shipped WASM simd128 build)
 * fused = 1 : v_muladd -> FMA (AVX2/AVX-512 dispatch on a native x86 host)
 * kernel = cv::getGaussianKernel(7, 2, CV_32F) (smooth.dispatch.cpp:76-190), values pinned from the
 * reference (tests/test_oracle.py checks them).
 * ---------------------------------------------------------------------------------------------- */
static const uint32_t orb_gauss7_bits[4] = {0x3e5d4ae0u, 0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u}; /* k[3..6] */

void orc_gauss7_kernel(float* k7)
{
    float c[4];
    memcpy(c, orb_gauss7_bits, sizeof c);
    k7[3] = c[0]; k7[2] = k7[4] = c[1]; k7[1] = k7[5] = c[2]; k7[0] = k7[6] = c[3];
}

void orc_orb_blur(const uint8_t* img, int w, int h, int fused, uint8_t* out)
{
    float k[7];
    orc_gauss7_kernel(k);
    float* rows = (float*)malloc((size_t)w * h * sizeof(float));
    for (int y = 0; y < h; y++) {
        const uint8_t* S = img + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float s;
            if (fused) {
                s = 0.f;
                for (int i = 0; i < 7; i++) s = fmaf((float)S[reflect101(x + i - 3, w)], k[i], s);
            } else {
                s = k[0] * (float)S[reflect101(x - 3, w)];
                for (int i = 1; i < 7; i++) s += k[i] * (float)S[reflect101(x + i - 3, w)];
            }
            rows[(size_t)y * w + x] = s;
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float* R0 = rows + (size_t)y * w + x;
            float s = fused ? fmaf(k[3], R0[0], 0.f) : k[3] * R0[0];
            for (int j = 1; j <= 3; j++) {
                float a = rows[(size_t)reflect101(y + j, h) * w + x];
                float b = rows[(size_t)reflect101(y - j, h) * w + x];
                if (fused) s = fmaf(k[3 + j], a + b, s);
                else s += k[3 + j] * (a + b);
            }
            int v = cv_roundf(s);
            out[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    free(rows);
}

/* ------------------------------------------------------------------------------------------------
 * rBRIEF-256 descriptors at given points.
 * FeatureExtractor::describeFeaturePoints (src/slam/src/feature_extractor.cpp:160-214):
 *   KeyPoint::convert -> angle = -1, octave 0 (core/src/types.cpp:93-101);
 *   ORB::create(500, 1., 0)->compute -> ORB_Impl::detectAndCompute(useProvidedKeypoints)
 *   (features2d/src/orb.cpp:970-1218): drop keypoints whose ROUNDED position is within 31 px of the
 *   border (orb.cpp:1130, keypoint.cpp:92-117), blur, then computeOrbDescriptors (orb.cpp:219-350):
 *     a = (float)cos(angle_rad), b = (float)sin(angle_rad), centre = (cvRound(x), cvRound(y)),
 *     sample(p) = img[cy + cvRound(p.x*b + p.y*a)][cx + cvRound(p.x*a - p.y*b)],
 *     bit i of byte j = sample(pattern[2*(8j+i)]) < sample(pattern[2*(8j+i)+1]).
 * angles == NULL reproduces AlvaAR (constant -1 degree); otherwise angles[i] in degrees (ORB's own
 * detect path feeds the intensity-centroid angle here).
 * `blurred` is the output of orc_orb_blur.  kept[i] = 1 if described.
 * ---------------------------------------------------------------------------------------------- */
static const int8_t orb_pattern[1024] = {
#include "../alvaar_b200/csrc/orb_pattern.inc"
};

int orc_orb_keep(float x, float y, int w, int h)
{
    if (h <= 62 || w <= 62) return 0;
    int ix = cv_roundf(x), iy = cv_roundf(y);
    return ix >= 31 && ix < w - 31 && iy >= 31 && iy < h - 31;
}

void orc_orb_trig(float angle_deg, float* a, float* b)
{
    float ang = angle_deg * (float)(3.1415926535897932384626433832795 / 180.f);
    *a = (float)cos(ang);
    *b = (float)sin(ang);
}

int orc_orb_describe(const uint8_t* blurred, int w, int h, const float* pts, const float* angles, int n,
                     uint8_t* desc, uint8_t* kept)
{
    int nk = 0;
    for (int i = 0; i < n; i++) {
        float px = pts[2 * i], py = pts[2 * i + 1];
        kept[i] = (uint8_t)orc_orb_keep(px, py, w, h);
        if (!kept[i]) { memset(desc + (size_t)i * 32, 0, 32); continue; }
        nk++;
        float a, b;
        orc_orb_trig(angles ? angles[i] : -1.f, &a, &b);
        const uint8_t* center = blurred + (size_t)cv_roundf(py) * w + cv_roundf(px);
        for (int j = 0; j < 32; j++) {
            int val = 0;
            for (int t = 0; t < 8; t++) {
                const int8_t* p = orb_pattern + 4 * (8 * j + t);
                float x0 = (float)p[0] * a - (float)p[1] * b, y0 = (float)p[0] * b + (float)p[1] * a;
                float x1 = (float)p[2] * a - (float)p[3] * b, y1 = (float)p[2] * b + (float)p[3] * a;
                int t0 = center[cv_roundf(y0) * w + cv_roundf(x0)];
                int t1 = center[cv_roundf(y1) * w + cv_roundf(x1)];
                val |= (t0 < t1) << t;
            }
            desc[(size_t)i * 32 + j] = (uint8_t)val;
        }
    }
    return nk;
}

/* ------------------------------------------------------------------------------------------------
 * Intensity-centroid orientation.  ICAngles (features2d/src/orb.cpp:181-215) with the umax table of
 * computeKeyPoints (orb.cpp:819-834, halfPatchSize 15) and cv::fastAtan2
 * (core/src/mathfuncs_core.simd.hpp:34-71: degree-scaled odd polynomial, baseline build, no FMA).
 * ---------------------------------------------------------------------------------------------- */
void orc_umax(int* umax /* [17] */)
{
    const int hp = 15;
    int v, v0, vmax = (int)floor(hp * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceil(hp * sqrtf(2.f) / 2);
    for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt((double)hp * hp - v * v));
    for (v = hp, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

float orc_fast_atan2(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / 3.1415926535897932384626433832795);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.1415926535897932384626433832795);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.1415926535897932384626433832795);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.1415926535897932384626433832795);
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* moments over the circular patch r=15 around (cvRound(x), cvRound(y)); caller guarantees >=16 px margin */
void orc_ic_moments(const uint8_t* img, int w, int cx, int cy, int* m01, int* m10)
{
    int umax[17];
    orc_umax(umax);
    const uint8_t* center = img + (size_t)cy * w + cx;
    int m_01 = 0, m_10 = 0;
    for (int u = -15; u <= 15; ++u) m_10 += u * center[u];
    for (int v = 1; v <= 15; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = center[u + v * w], vm = center[u - v * w];
            v_sum += (vp - vm);
            m_10 += u * (vp + vm);
        }
        m_01 += v * v_sum;
    }
    *m01 = m_01; *m10 = m_10;
}

void orc_ic_angles(const uint8_t* img, int w, int h, const float* pts, int n, float* angles)
{
    for (int i = 0; i < n; i++) {
        int cx = cv_roundf(pts[2 * i]), cy = cv_roundf(pts[2 * i + 1]);
        if (cx < 15 || cy < 15 || cx >= w - 15 || cy >= h - 15) { angles[i] = -1.f; continue; }
        int m01, m10;
        orc_ic_moments(img, w, cx, cy, &m01, &m10);
        angles[i] = orc_fast_atan2((float)m01, (float)m10);
    }
}

/* Harris response, HarrisResponses (features2d/src/orb.cpp:130-177), blockSize 7, k = 0.04 */
void orc_harris(const uint8_t* img, int w, int h, const float* pts, int n, float* resp)
{
    const int bs = 7, r = bs / 2;
    float scale = 1.f / ((1 << 2) * bs * 255.f);
    float scale_sq_sq = scale * scale * scale * scale;
    for (int i = 0; i < n; i++) {
        int x0 = cv_roundf(pts[2 * i]), y0 = cv_roundf(pts[2 * i + 1]);
        if (x0 < r + 1 || y0 < r + 1 || x0 >= w - r - 1 || y0 >= h - r - 1) { resp[i] = 0.f; continue; }
        int a = 0, b = 0, c = 0;
        for (int dy = 0; dy < bs; dy++)
            for (int dx = 0; dx < bs; dx++) {
                const uint8_t* p = img + (size_t)(y0 - r + dy) * w + (x0 - r + dx);
                int Ix = (p[1] - p[-1]) * 2 + (p[-w + 1] - p[-w - 1]) + (p[w + 1] - p[w - 1]);
                int Iy = (p[w] - p[-w]) * 2 + (p[w - 1] - p[-w - 1]) + (p[w + 1] - p[-w + 1]);
                a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
            }
        resp[i] = ((float)a * b - (float)c * c - 0.04f * ((float)a + b) * ((float)a + b)) * scale_sq_sq;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Hamming distance + brute-force 2-NN.
 *   cv::norm(a, b, NORM_HAMMING) (core/src/norm.cpp:99-140 -> hal::normHamming stat.simd.hpp:81-128),
 *   called by MapPoint::computeMinDescDist (src/slam/src/map_point.cpp:204-222);
 *   BFMatcher(NORM_HAMMING).knnMatch(k=2) (features2d/src/matchers.cpp:757 -> batchDistance,
 *   core/src/batch_distance.cpp:103-123, k-NN insertion :235-248: strict '<' so ties keep the
 *   LOWEST train index).   out[4*i] = {idx0, dist0, idx1, dist1}, -1 when missing.
 * ---------------------------------------------------------------------------------------------- */
int orc_hamming(const uint8_t* a, const uint8_t* b, int nbytes)
{
    int d = 0;
    for (int i = 0; i < nbytes; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

void orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* out)
{
    for (int i = 0; i < nq; i++) {
        int d0 = INT32_MAX, d1 = INT32_MAX, i0 = -1, i1 = -1;
        const uint64_t* a = (const uint64_t*)(q + (size_t)i * 32);
        uint64_t a0, a1, a2, a3;
        memcpy(&a0, a, 8); memcpy(&a1, a + 1, 8); memcpy(&a2, a + 2, 8); memcpy(&a3, a + 3, 8);
        for (int j = 0; j < nt; j++) {
            uint64_t b[4];
            memcpy(b, t + (size_t)j * 32, 32);
            int d = __builtin_popcountll(a0 ^ b[0]) + __builtin_popcountll(a1 ^ b[1]) +
                    __builtin_popcountll(a2 ^ b[2]) + __builtin_popcountll(a3 ^ b[3]);
            if (d < d1) {
                if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
                else { d1 = d; i1 = j; }
            }
        }
        out[4 * i] = i0; out[4 * i + 1] = i0 < 0 ? -1 : d0;
        out[4 * i + 2] = i1; out[4 * i + 3] = i1 < 0 ? -1 : d1;
    }
}

/* MapPoint::computeMinDescDist (src/slam/src/map_point.cpp:204-222): min over all descriptor pairs,
 * initial value 1000. */
int orc_min_desc_dist(const uint8_t* A, int na, const uint8_t* B, int nb)
{
    int best = 1000;
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++) {
            int d = orc_hamming(A + (size_t)i * 32, B + (size_t)j * 32, 32);
            if (d < best) best = d;
        }
    return best;
}

/* KeyPointsFilter::retainBest semantics on integer FAST responses (features2d/src/keypoint.cpp:69-90):
 * returns the response threshold such that every keypoint with response >= threshold is kept
 * (can keep more than n on ties); returns 0 when count <= n (keep all). */
int orc_retain_best_threshold(const int32_t* xys, int count, int n)
{
    if (n < 0 || count <= n) return 0;
    if (n == 0) return 256;
    int hist[256];
    memset(hist, 0, sizeof hist);
    for (int i = 0; i < count; i++) hist[xys[3 * i + 2] & 255]++;
    int acc = 0;
    for (int s = 255; s >= 0; s--) {
        acc += hist[s];
        if (acc >= n) return s;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * ORB::detectAndCompute with nlevels = 1, HARRIS_SCORE, edgeThreshold = patchSize = 31
 * (features2d/src/orb.cpp:970-1218; computeKeyPoints :785-958):
 *   FAST(thr, nms) -> runByImageBorder(31) -> retainBest(2n) on the FAST score -> HarrisResponses ->
 *   retainBest(n) on the Harris response -> ICAngles -> GaussianBlur -> steered rBRIEF.
 * The reference's output order is whatever nth_element/partition leave; here: row-major (y, x).
 * kp[4*i] = {x, y, response, angle}; returns the count (only the first cap are written).
 * ---------------------------------------------------------------------------------------------- */
static int cmp_float_desc(const void* a, const void* b)
{
    float x = *(const float*)a, y = *(const float*)b;
    return (x < y) - (x > y);
}

int orc_orb_detect(const uint8_t* img, int w, int h, int nfeatures, int fast_thr, int fused, float* kp, uint8_t* desc,
                   int cap)
{
    int cap0 = w * h / 8 + 16;
    int32_t* xys = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)cap0);
    int n0 = orc_fast9(img, w, h, fast_thr, 1, xys, cap0);
    /* border 31: Rect(31, 31, w-31, h-31).contains(pt) */
    int n1 = 0;
    if (!(h <= 62 || w <= 62))
        for (int i = 0; i < n0; i++) {
            int x = xys[3 * i], y = xys[3 * i + 1];
            if (x >= 31 && x < w - 31 && y >= 31 && y < h - 31) { memmove(xys + 3 * n1, xys + 3 * i, 12); n1++; }
        }
    int thr = orc_retain_best_threshold(xys, n1, 2 * nfeatures);
    int n2 = 0;
    for (int i = 0; i < n1; i++)
        if (xys[3 * i + 2] >= thr) { memmove(xys + 3 * n2, xys + 3 * i, 12); n2++; }
    float* pts = (float*)malloc(sizeof(float) * 2 * (size_t)(n2 + 1));
    float* hr = (float*)malloc(sizeof(float) * (size_t)(n2 + 1));
    float* tmp = (float*)malloc(sizeof(float) * (size_t)(n2 + 1));
    for (int i = 0; i < n2; i++) { pts[2 * i] = (float)xys[3 * i]; pts[2 * i + 1] = (float)xys[3 * i + 1]; }
    orc_harris(img, w, h, pts, n2, hr);
    float fthr = -INFINITY;
    if (n2 > nfeatures) {
        memcpy(tmp, hr, sizeof(float) * n2);
        qsort(tmp, n2, sizeof(float), cmp_float_desc);
        fthr = tmp[nfeatures - 1];
    }
    int n3 = 0;
    for (int i = 0; i < n2; i++)
        if (hr[i] >= fthr) { pts[2 * n3] = pts[2 * i]; pts[2 * n3 + 1] = pts[2 * i + 1]; hr[n3] = hr[i]; n3++; }
    int nw = n3 < cap ? n3 : cap;
    float* ang = (float*)malloc(sizeof(float) * (size_t)(n3 + 1));
    uint8_t* blur = (uint8_t*)malloc((size_t)w * h);
    uint8_t* keptv = (uint8_t*)malloc((size_t)n3 + 1);
    orc_ic_angles(img, w, h, pts, n3, ang);
    orc_orb_blur(img, w, h, fused, blur);
    orc_orb_describe(blur, w, h, pts, ang, nw, desc, keptv);
    for (int i = 0; i < nw; i++) { kp[4 * i] = pts[2 * i]; kp[4 * i + 1] = pts[2 * i + 1]; kp[4 * i + 2] = hr[i]; kp[4 * i + 3] = ang[i]; }
    free(xys); free(pts); free(hr); free(tmp); free(ang); free(blur); free(keptv);
    return n3;
}

/* ------------------------------------------------------------------------------------------------
 * Scharr derivative image of one pyramid level, as cv::buildOpticalFlowPyramid(withDerivatives = true)
 * stores it (video/src/lkpyramid.cpp:57-150, ScharrDerivInvoker): with reflect-101 neighbours in both
 * directions (row -1 -> row 1, column -1 -> column 1; a single row/column reflects onto itself),
 *   t0 = 3 (s[y-1] + s[y+1]) + 10 s[y]      t1 = s[y+1] - s[y-1]                (vertical, per column)
 *   dx = t0[x+1] - t0[x-1]                  dy = 3 (t1[x-1] + t1[x+1]) + 10 t1[x]
 * out: int16 [h][w][2] = (dx, dy) interleaved.  |dx|, |dy| <= 16 * 255: no int16 wrap.
 * ---------------------------------------------------------------------------------------------- */
void orc_scharr(const uint8_t* img, int w, int h, int16_t* out)
{
    int* t0 = (int*)malloc(sizeof(int) * (size_t)(w + 2));
    int* t1 = (int*)malloc(sizeof(int) * (size_t)(w + 2));
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = img + (size_t)(y > 0 ? y - 1 : (h > 1 ? 1 : 0)) * w;
        const uint8_t* r1 = img + (size_t)y * w;
        const uint8_t* r2 = img + (size_t)(y < h - 1 ? y + 1 : (h > 1 ? h - 2 : 0)) * w;
        for (int x = 0; x < w; x++) {
            t0[x + 1] = (r0[x] + r2[x]) * 3 + r1[x] * 10;
            t1[x + 1] = r2[x] - r0[x];
        }
        int x0 = w > 1 ? 1 : 0, x1 = w > 1 ? w - 2 : 0;
        t0[0] = t0[x0 + 1]; t0[w + 1] = t0[x1 + 1];
        t1[0] = t1[x0 + 1]; t1[w + 1] = t1[x1 + 1];
        for (int x = 0; x < w; x++) {
            out[((size_t)y * w + x) * 2] = (int16_t)(t0[x + 2] - t0[x]);
            out[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
    free(t0); free(t1);
}
