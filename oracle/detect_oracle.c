/* oracle/detect_oracle.c -- CPU restatement of the reference's keyframe corner detector:
 * per-grid-cell Shi-Tomasi (minimum eigenvalue) maxima with a shared suppression mask, then cornerSubPix.
 *
 * TEST INFRASTRUCTURE ONLY (see alva_oracle.c's header): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call this.  The product (alvaar_b200/) never links or executes it.
 *
 * Reference (paths under /root/reference):
 *   FeatureExtractor::detectFeaturePoints     src/slam/src/feature_extractor.cpp:11-158   (caller map_manager.cpp:193-222)
 *   cv::GaussianBlur(3x3, sigma 0) on a cell ROI   src/libs/opencv/modules/imgproc/src/smooth.dispatch.cpp:611-755: the ROI is a
 *       non-isolated submatrix, so the bit-exact branch (:653) is skipped and sepFilter2D runs its 8-bit fixed-point path whose
 *       SIMD body rounds half-to-even and whose scalar tail rounds half-up (see orc_blur3_cell); pixels outside the ROI come
 *       from the parent image, REFLECT_101 at the image border
 *   cv::cornerMinEigenVal(cell, block 3, Sobel 3)  imgproc/src/corner.cpp:237-330,550-572: Sobel with scale 1/(4*3*255) folded into
 *       the smoothing kernel (deriv.cpp Sobel), float row filter RowFilter<uchar,float> (left-to-right products) and
 *       SymmColumnSmallVec_32f (filter.simd.hpp:2094-2165: (S0 + S2) * k1 + (S1 * k0 + delta)), products, unnormalised 3x3 box
 *       sums in double (box_filter), lambda_min = (a + c) - sqrt((a - c)^2 + b^2) in float (corner.cpp:52-103); all borders
 *       REFLECT_101 on the 40x40 cell
 *   cv::circle (filled, radius cell/4)             imgproc/src/drawing.cpp:1476-1610 (Bresenham disk; centre = cvRound(px))
 *   cv::minMaxLoc                                  first maximum in row-major order
 *   cv::cornerSubPix(3x3 half-window, 30 it, 0.01) imgproc/src/cornersubpix.cpp:44-160, getRectSubPix 8U->32F (samplers.cpp)
 *
 * PINNED: tests/test_oracle_detect.py -- corner list bit-for-bit (integer cell maxima, order, count, adapted quality) and
 * sub-pixel positions as float bit patterns against golden vectors dumped from the reference's own FeatureExtractor
 * (tools/make_golden_detect.py) and the live reference.  cv::setNumThreads(1): the reference's parallel_for_ body is racy
 * (shared mask / counter); the serial cell order is the defined behaviour.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

static inline int d_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

/* GaussianBlur(image(cell ROI), 3x3, sigma 0) as the reference computes it (measured against the reference, then read off
 * filter.simd.hpp): sepFilter2D takes its 8-bit fixed-point path (kernel [64 128 64] per axis, 32-bit row sums).  The
 * column stage's SIMD body (SymmColumnVec_32s8u, :1010-1099) converts the sums to float, scales by 1/256 -- every value is a
 * multiple of 1/16, exact -- and rounds half-to-EVEN (v_round); it covers the first (cs & ~3) columns of the ROI.  The
 * remaining (cs & 3) columns go through the scalar FixedPtCastEx: (sum + 2^15) >> 16, i.e. half-UP.  Pixels outside the ROI
 * come from the parent image (non-isolated submatrix), REFLECT_101 at the image border. */
void orc_blur3_cell(const uint8_t* img, int w, int h, int x0, int y0, int cs, uint8_t* out)
{
    static const int k[3] = {1, 2, 1};
    for (int y = 0; y < cs; y++)
        for (int x = 0; x < cs; x++) {
            int s = 0;
            for (int j = -1; j <= 1; j++) {
                const uint8_t* row = img + (size_t)d_reflect101(y0 + y + j, h) * w;
                int hs = 0;
                for (int i = -1; i <= 1; i++) hs += k[i + 1] * row[d_reflect101(x0 + x + i, w)];
                s += k[j + 1] * hs;
            }
            int q;
            if (x < (cs & ~3)) {
                q = s >> 4;
                const int r = s & 15;
                if (r > 8 || (r == 8 && (q & 1))) q++;
            } else
                q = (s + 8) >> 4;
            out[y * cs + x] = (uint8_t)q;
        }
}

/* cornerMinEigenVal(block 3, ksize 3) of the blurred cs x cs cell at (x0, y0) (borders REFLECT_101 on the cell) */
void orc_min_eig_cell(const uint8_t* img, int w, int h, int x0, int y0, int cs, float* hmap)
{
    uint8_t* blur = (uint8_t*)malloc((size_t)cs * cs);
    orc_blur3_cell(img, w, h, x0, y0, cs, blur);
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0));   /* (float)scale: Sobel scales the smoothing kernel, deriv.cpp */
    const float s2 = 2.0f * s;
    float* dx = (float*)malloc(sizeof(float) * cs * cs);
    float* dy = (float*)malloc(sizeof(float) * cs * cs);
    float* rd = (float*)malloc(sizeof(float) * cs * (cs + 2));   /* row-filtered, rows -1..cs */
    float* rs = (float*)malloc(sizeof(float) * cs * (cs + 2));
#define PIX(xx, yy) ((float)blur[d_reflect101((yy), cs) * cs + d_reflect101((xx), cs)])
    for (int y = -1; y <= cs; y++)
        for (int x = 0; x < cs; x++) {
            const float a = PIX(x - 1, y), b = PIX(x, y), c = PIX(x + 1, y);
            /* RowFilter<uchar,float>: s = k0*S0; s += k1*S1; s += k2*S2 */
            float d = -1.0f * a; d += 0.0f * b; d += 1.0f * c;
            float m = s * a; m += s2 * b; m += s * c;
            rd[(y + 1) * cs + x] = d; rs[(y + 1) * cs + x] = m;
        }
#undef PIX
    for (int y = 0; y < cs; y++)
        for (int x = 0; x < cs; x++) {
            const float d0 = rd[y * cs + x], d1 = rd[(y + 1) * cs + x], d2 = rd[(y + 2) * cs + x];
            const float t1 = d1 * s2 + 0.0f;             /* v_muladd(S1, k0, delta) */
            dx[y * cs + x] = (d0 + d2) * s + t1;         /* v_muladd(S0 + S2, k1, .) */
            dy[y * cs + x] = rs[(y + 2) * cs + x] - rs[y * cs + x] + 0.0f;
        }
    for (int y = 0; y < cs; y++)
        for (int x = 0; x < cs; x++) {
            double A = 0, B = 0, C = 0;
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    const int yy = d_reflect101(y + j, cs), xx = d_reflect101(x + i, cs);
                    const float gx = dx[yy * cs + xx], gy = dy[yy * cs + xx];
                    A += (double)(gx * gx); B += (double)(gx * gy); C += (double)(gy * gy);
                }
            const float a = (float)A * 0.5f, b = (float)B, c = (float)C * 0.5f;
            const float t = a - c;
            hmap[y * cs + x] = (a + c) - sqrtf(b * b + t * t);
        }
    free(dx); free(dy); free(rd); free(rs); free(blur);
}

/* half-widths of cv::circle's filled disk per |row offset| (drawing.cpp:1483-1610) */
void orc_circle_halfwidths(int radius, int* hw /* [radius + 1] */)
{
    for (int i = 0; i <= radius; i++) hw[i] = -1;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        if (dx > hw[dy]) hw[dy] = dx;
        if (dy > hw[dx]) hw[dx] = dy;
        dy++;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

static void draw_disk(uint8_t* mask, int w, int h, int cx, int cy, int radius, const int* hw)
{
    for (int r = -radius; r <= radius; r++) {
        const int y = cy + r, half = hw[r < 0 ? -r : r];
        if (y < 0 || y >= h || half < 0) continue;
        int xa = cx - half, xb = cx + half;
        if (xa < 0) xa = 0;
        if (xb > w - 1) xb = w - 1;
        for (int x = xa; x <= xb; x++) mask[(size_t)y * w + x] = 0;
    }
}

/* cv::getRectSubPix(8U -> 32F): getRectSubPix_8u32f (imgproc/src/samplers.cpp:219-268).  Interior: a running recurrence
 * along each row (prev + t, prev = (float)(t * (1 - a) / a), a clamped to >= 1e-4); otherwise the generic replicate-border
 * template (samplers.cpp:127-215): plain bilinear weights, border columns blended vertically only. */
static void rect_subpix(const uint8_t* img, int w, int h, float cx, float cy, int pw, int ph, float* out)
{
    cx -= (pw - 1) * 0.5f; cy -= (ph - 1) * 0.5f;
    const int ipx = (int)floorf(cx), ipy = (int)floorf(cy);
    if (0 <= ipx && ipx + pw < w && 0 <= ipy && ipy + ph < h) {
        float a = cx - ipx;
        const float b = cy - ipy;
        a = a > 0.0001f ? a : 0.0001f;
        const float a12 = a * (1.f - b), a22 = a * b, b1 = 1.f - b, b2 = b;
        const double s = (1. - a) / a;
        for (int i = 0; i < ph; i++) {
            const uint8_t* r0 = img + (size_t)(ipy + i) * w + ipx;
            const uint8_t* r1 = r0 + w;
            float prev = (1 - a) * (b1 * r0[0] + b2 * r1[0]);
            for (int j = 0; j < pw; j++) {
                const float t = a12 * r0[j + 1] + a22 * r1[j + 1];
                out[i * pw + j] = prev + t;
                prev = (float)(t * s);
            }
        }
    } else {
        const float a = cx - ipx, b = cy - ipy;
        const float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b, b1 = 1.f - b, b2 = b;
        for (int i = 0; i < ph; i++) {
            int y0 = ipy + i, y1 = y0 + 1;
            y0 = y0 < 0 ? 0 : (y0 > h - 1 ? h - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > h - 1 ? h - 1 : y1);
            const uint8_t* r0 = img + (size_t)y0 * w;
            const uint8_t* r1 = img + (size_t)y1 * w;
            for (int j = 0; j < pw; j++) {
                const int x0 = ipx + j, x1 = x0 + 1;
                if (x0 < 0 || x1 > w - 1) {
                    const int xc = x0 < 0 ? 0 : w - 1;
                    out[i * pw + j] = r0[xc] * b1 + r1[xc] * b2;
                } else
                    out[i * pw + j] = r0[x0] * a11 + r0[x1] * a12 + r1[x0] * a21 + r1[x1] * a22;
            }
        }
    }
}

/* cv::cornerSubPix(image, pts, Size(win, win), Size(-1,-1), TermCriteria(EPS + MAX_ITER, max_iter, eps)) */
void orc_corner_subpix(const uint8_t* img, int w, int h, float* pts, int n, int win, int max_iter, double eps)
{
    const int ww = 2 * win + 1;
    float* mask = (float*)malloc(sizeof(float) * ww * ww);
    float* buf = (float*)malloc(sizeof(float) * (ww + 2) * (ww + 2));
    for (int i = 0; i < ww; i++) {
        const float y = (float)(i - win) / win;
        const float vy = expf(-y * y);
        for (int j = 0; j < ww; j++) {
            const float x = (float)(j - win) / win;
            mask[i * ww + j] = (float)(vy * expf(-x * x));
        }
    }
    if (max_iter < 1) max_iter = 1;
    if (max_iter > 100) max_iter = 100;
    eps = eps > 0 ? eps : 0;
    eps *= eps;
    for (int p = 0; p < n; p++) {
        const float cTx = pts[2 * p], cTy = pts[2 * p + 1];
        float cx = cTx, cy = cTy;
        int iter = 0;
        double err = 0;
        do {
            double a = 0, b = 0, c = 0, bb1 = 0, bb2 = 0;
            rect_subpix(img, w, h, cx, cy, ww + 2, ww + 2, buf);
            const int st = ww + 2;
            for (int i = 0, k = 0; i < ww; i++) {
                const float* sp = buf + (i + 1) * st + 1;
                const double py = i - win;
                for (int j = 0; j < ww; j++, k++) {
                    const double m = mask[k];
                    const double tgx = sp[j + 1] - sp[j - 1];
                    const double tgy = sp[j + st] - sp[j - st];
                    const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                    const double px = j - win;
                    a += gxx; b += gxy; c += gyy;
                    bb1 += gxx * px + gxy * py;
                    bb2 += gxy * px + gyy * py;
                }
            }
            const double det = a * c - b * b;
            if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
            const double scale = 1.0 / det;
            const float nx = (float)(cx + c * scale * bb1 - b * scale * bb2);
            const float ny = (float)(cy - b * scale * bb1 + a * scale * bb2);
            err = (nx - cx) * (nx - cx) + (ny - cy) * (ny - cy);
            cx = nx; cy = ny;
            if (cx < 0 || cx >= w || cy < 0 || cy >= h) break;
        } while (++iter < max_iter && err > eps);
        if (fabs(cx - cTx) > win || fabs(cy - cTy) > win) { cx = cTx; cy = cTy; }
        pts[2 * p] = cx; pts[2 * p + 1] = cy;
    }
    free(mask); free(buf);
}

/* FeatureExtractor::detectFeaturePoints(image, cell, currKeypoints, roi) with cv::setNumThreads(1).
 * cur [ncur][2]: pixel positions of the keypoints the frame already has.  roi = {x, y, width, height}.
 * *quality: maxQuality_ in/out (adapted as feature_extractor.cpp:138-145).  out [cap][2]; out_int (optional) [cap][2]: the
 * integer cell maxima before cornerSubPix.  Returns the number of points (primaries in cell order, then secondaries). */
int orc_detect_points(const uint8_t* img, int w, int h, int cs, const float* cur, int ncur, const int* roi, double* quality,
                      float* out, int32_t* out_int, int cap)
{
    const int rad = cs / 4;
    const int nch = h / cs, ncw = w / cs, ncells = nch * ncw;
    int* hw = (int*)malloc(sizeof(int) * (rad + 1));
    orc_circle_halfwidths(rad, hw);
    uint8_t* mask = (uint8_t*)malloc((size_t)w * h);
    memset(mask, 1, (size_t)w * h);
    uint8_t* occ = (uint8_t*)calloc((size_t)(nch + 1) * (ncw + 1), 1);
    for (int i = 0; i < ncur; i++) {
        const float px = cur[2 * i], py = cur[2 * i + 1];
        const size_t r = (size_t)(py / cs), c = (size_t)(px / cs);
        if (r <= (size_t)nch && c <= (size_t)ncw) occ[r * (ncw + 1) + c] = 1;
        draw_disk(mask, w, h, (int)lrintf(px), (int)lrintf(py), rad, hw);
    }
    float* hmap = (float*)malloc(sizeof(float) * cs * cs);
    int32_t* prim = (int32_t*)malloc(sizeof(int32_t) * 2 * (ncells + 1));
    int32_t* sec = (int32_t*)malloc(sizeof(int32_t) * 2 * (ncells + 1));
    uint8_t* has_p = (uint8_t*)calloc(ncells + 1, 1);
    uint8_t* has_s = (uint8_t*)calloc(ncells + 1, 1);
    int n_occ = 0;
    const double q = *quality;
    for (int i = 0; i < ncells; i++) {
        const int r = i / ncw, c = i % ncw;
        if (occ[r * (ncw + 1) + c]) { n_occ++; continue; }
        const int x0 = c * cs, y0 = r * cs;
        if (!(x0 + cs < w - 1 && y0 + cs < h - 1)) continue;
        orc_min_eig_cell(img, w, h, x0, y0, cs, hmap);
        for (int pass = 0; pass < 2; pass++) {
            float best = -FLT_MAX;
            int bx = 0, by = 0;
            for (int y = 0; y < cs; y++)
                for (int x = 0; x < cs; x++) {
                    const float v = hmap[y * cs + x] * (float)mask[(size_t)(y0 + y) * w + x0 + x];
                    if (v > best) { best = v; bx = x; by = y; }
                }
            const int X = bx + x0, Y = by + y0;
            if (X < roi[0] || Y < roi[1] || X >= roi[0] + roi[2] || Y >= roi[1] + roi[3]) break;   /* `continue` to the next cell */
            if ((double)best >= q) {
                if (pass == 0) { prim[2 * i] = X; prim[2 * i + 1] = Y; has_p[i] = 1; }
                else { sec[2 * i] = X; sec[2 * i + 1] = Y; has_s[i] = 1; }
                draw_disk(mask, w, h, X, Y, rad, hw);
            }
        }
    }
    int n = 0;
    for (int i = 0; i < ncells; i++)
        if (has_p[i] && n < cap) { out[2 * n] = (float)prim[2 * i]; out[2 * n + 1] = (float)prim[2 * i + 1]; n++; }
    const int nk = n;
    if (nk + n_occ < ncells) {
        const int nsec = ncells - (nk + n_occ);
        int k = 0;
        for (int i = 0; i < ncells && k < nsec; i++)
            if (has_s[i] && n < cap) { out[2 * n] = (float)sec[2 * i]; out[2 * n + 1] = (float)sec[2 * i + 1]; n++; k++; }
    }
    if (n < 0.33 * (ncells - n_occ)) *quality = q * 0.5;
    else if (n > 0.9 * (ncells - n_occ)) *quality = q * 1.5;
    if (out_int)
        for (int i = 0; i < n; i++) { out_int[2 * i] = (int32_t)out[2 * i]; out_int[2 * i + 1] = (int32_t)out[2 * i + 1]; }
    if (n > 0) orc_corner_subpix(img, w, h, out, n, 3, 30, 0.01);
    free(hw); free(mask); free(occ); free(hmap); free(prim); free(sec); free(has_p); free(has_s);
    return n;
}
