/* oracle/pose_oracle.c -- CPU restatement of the reference's per-frame pose estimation:
 * P3P (Kneip) inside a Least-Median-of-Squares loop, then the Ceres motion-only refinement (PnP).
 *
 * TEST INFRASTRUCTURE ONLY (see alva_oracle.c's header): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call this.  The product (alvaar_b200/) never links or executes it.
 *
 * Reference (paths under /root/reference):
 *   MultiViewGeometry::p3pRansac       src/slam/src/multi_view_geometry.cpp:24-127   (caller visual_frontend.cpp:245-417)
 *   MultiViewGeometry::ceresPnP        src/slam/src/multi_view_geometry.cpp:129-223
 *   DirectSE3::ReprojectionErrorSE3    src/slam/src/ceres_parametrization.cpp:96-155
 *   opengv::sac::Lmeds::computeModel   src/libs/opengv/include/opengv/sac/implementation/Lmeds.hpp:40-190
 *   SampleConsensusProblem (sampling)  src/libs/opengv/include/opengv/sac/implementation/SampleConsensusProblem.hpp:36-120
 *   AbsolutePoseSacProblem (KNEIP)     src/libs/opengv/src/sac_problems/absolute_pose/AbsolutePoseSacProblem.cpp:40-210
 *   p3p_kneip_main                     src/libs/opengv/src/absolute_pose/modules/main.cpp:50-220
 *   math::o4_roots                     src/libs/opengv/src/math/roots.cpp:88-136
 *   Ceres trust-region loop            as restated in oracle/ba_oracle.c (same LM strategy; DENSE_QR solves the same
 *                                      damped least-squares problem, here through its normal equations)
 *
 * Floating point (fp64): tolerance 1e-4 relative on the pose (BASELINE.json north_star); inlier / outlier sets exact on
 * the seeded test inputs.  PINNED against the reference's own MultiViewGeometry (compiled unmodified with the vendored
 * OpenGV / Ceres into oracle/_ref/libalva_ref.so): tests/test_oracle_pose.py + tests/golden/pose.npz.
 *
 * Determinism: the reference seeds its sampler from the clock unless State::multiViewRandomEnabled_ is false
 * (state.hpp:67 -> SampleConsensusProblem.hpp:43-46: mt19937 seeded 12345).  With the pin, rnd() is
 * std::uniform_int_distribution<int>(0, INT_MAX) over std::mt19937, which in libstdc++ (GCC >= 11, 32-bit engine range)
 * is Lemire's multiply-shift: (x * 2^31) >> 32 = x >> 1, no rejection.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <complex.h>

/* ------------------------------------------------------------------ std::mt19937 */
typedef struct { uint32_t s[624]; int i; } mt_t;
static void mt_seed(mt_t* m, uint32_t seed)
{
    m->s[0] = seed;
    for (int i = 1; i < 624; i++) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
    m->i = 624;
}
static uint32_t mt_next(mt_t* m)
{
    if (m->i >= 624) {
        for (int k = 0; k < 624; k++) {
            uint32_t y = (m->s[k] & 0x80000000u) | (m->s[(k + 1) % 624] & 0x7fffffffu);
            m->s[k] = m->s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->i = 0;
    }
    uint32_t y = m->s[m->i++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* first n values of SampleConsensusProblem::rnd() for the given seed (test helper; the CUDA library gets the same table) */
void orc_sac_rnd(uint32_t seed, int n, int32_t* out)
{
    mt_t m;
    mt_seed(&m, seed);
    for (int i = 0; i < n; i++) out[i] = (int32_t)(mt_next(&m) >> 1);
}

/* ------------------------------------------------------------------ small vector helpers */
static void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* std::pow(std::complex<double>, double) as libstdc++ computes it (complex:1019-1033) */
static double complex cpow_real(double complex x, double y)
{
    if (cimag(x) == 0.0 && creal(x) > 0.0) return pow(creal(x), y);
    double complex t = clog(x);
    const double rho = exp(y * creal(t)), th = y * cimag(t);
    return CMPLX(rho * cos(th), rho * sin(th));
}

/* opengv::math::o4_roots (roots.cpp:88-136): Ferrari in complex arithmetic; the real parts of the four roots are returned
 * whether or not the roots are real (complex ones turn into NaN poses downstream and never win) */
static void o4_roots(const double* p, double* roots)
{
    const double A = p[0], B = p[1], C = p[2], D = p[3], E = p[4];
    const double A2 = A * A, B2 = B * B, A3 = A2 * A, B3 = B2 * B, A4 = A3 * A, B4 = B3 * B;
    const double alpha = -3 * B2 / (8 * A2) + C / A;
    const double beta = B3 / (8 * A3) - B * C / (2 * A2) + D / A;
    const double gamma = -3 * B4 / (256 * A4) + B2 * C / (16 * A3) - B * D / (4 * A2) + E / A;
    const double alpha2 = alpha * alpha, alpha3 = alpha2 * alpha;
    const double complex P = CMPLX(-alpha2 / 12 - gamma, 0);
    const double complex Q = CMPLX(-alpha3 / 108 + alpha * gamma / 3 - beta * beta / 8, 0);
    const double complex R = -Q / 2.0 + csqrt(cpow_real(Q, 2.0) / 4.0 + cpow_real(P, 3.0) / 27.0);
    const double complex U = cpow_real(R, 1.0 / 3.0);
    double complex y;
    if (creal(U) == 0) y = -5.0 * alpha / 6.0 - cpow_real(Q, 1.0 / 3.0);
    else y = -5.0 * alpha / 6.0 - P / (3.0 * U) + U;
    const double complex w = csqrt(alpha + 2.0 * y);
    const double complex bw = CMPLX(2.0 * beta, 0) / w;
    const double complex s1 = csqrt(-(3.0 * alpha + 2.0 * y + bw)), s2 = csqrt(-(3.0 * alpha + 2.0 * y - bw));
    const double sh = -B / (4.0 * A);
    roots[0] = creal(sh + 0.5 * (w + s1));
    roots[1] = creal(sh + 0.5 * (w - s1));
    roots[2] = creal(sh + 0.5 * (-w + s2));
    roots[3] = creal(sh + 0.5 * (-w - s2));
}

/* p3p_kneip_main: f[3][3] unit bearings, p[3][3] world points -> up to 4 camera-to-world transforms sol[k] = [R | C] (3x4
 * row-major).  Returns the number of solutions (0 for collinear world points, else 4). */
int orc_p3p_kneip(const double* f, const double* p, double* sol)
{
    const double *P1 = p, *P2 = p + 3, *P3 = p + 6;
    double t1[3], t2[3], c[3];
    for (int i = 0; i < 3; i++) { t1[i] = P2[i] - P1[i]; t2[i] = P3[i] - P1[i]; }
    cross3(t1, t2, c);
    if (norm3(c) == 0) return 0;
    const double *f1 = f, *f2 = f + 3, *f3 = f + 6;
    double T[9], f3t[3];
    for (int pass = 0; pass < 2; pass++) {
        double e3[3], e2[3];
        cross3(f1, f2, e3);
        const double n = norm3(e3);
        for (int i = 0; i < 3; i++) e3[i] /= n;
        cross3(e3, f1, e2);
        for (int i = 0; i < 3; i++) { T[i] = f1[i]; T[3 + i] = e2[i]; T[6 + i] = e3[i]; }
        for (int i = 0; i < 3; i++) f3t[i] = T[3 * i] * f3[0] + T[3 * i + 1] * f3[1] + T[3 * i + 2] * f3[2];
        if (pass == 0 && f3t[2] > 0) { f1 = f + 3; f2 = f; P1 = p + 3; P2 = p; }
        else break;
    }
    double n1[3], n2[3], n3[3], d[3], N[9];
    for (int i = 0; i < 3; i++) n1[i] = P2[i] - P1[i];
    { const double n = norm3(n1); for (int i = 0; i < 3; i++) n1[i] /= n; }
    for (int i = 0; i < 3; i++) d[i] = P3[i] - P1[i];
    cross3(n1, d, n3);
    { const double n = norm3(n3); for (int i = 0; i < 3; i++) n3[i] /= n; }
    cross3(n3, n1, n2);
    for (int i = 0; i < 3; i++) { N[i] = n1[i]; N[3 + i] = n2[i]; N[6 + i] = n3[i]; }
    double P3n[3];
    for (int i = 0; i < 3; i++) P3n[i] = N[3 * i] * d[0] + N[3 * i + 1] * d[1] + N[3 * i + 2] * d[2];
    const double d_12 = norm3(t1);
    const double f_1 = f3t[0] / f3t[2], f_2 = f3t[1] / f3t[2], p_1 = P3n[0], p_2 = P3n[1];
    const double cos_beta = dot3(f1, f2);
    double b = 1 / (1 - cos_beta * cos_beta) - 1;
    b = cos_beta < 0 ? -sqrt(b) : sqrt(b);
    const double f_1_pw2 = f_1 * f_1, f_2_pw2 = f_2 * f_2, p_1_pw2 = p_1 * p_1, p_1_pw3 = p_1_pw2 * p_1, p_1_pw4 = p_1_pw3 * p_1;
    const double p_2_pw2 = p_2 * p_2, p_2_pw3 = p_2_pw2 * p_2, p_2_pw4 = p_2_pw3 * p_2, d_12_pw2 = d_12 * d_12, b_pw2 = b * b;
    double fac[5];
    fac[0] = -f_2_pw2 * p_2_pw4 - p_2_pw4 * f_1_pw2 - p_2_pw4;
    fac[1] = 2 * p_2_pw3 * d_12 * b + 2 * f_2_pw2 * p_2_pw3 * d_12 * b - 2 * f_2 * p_2_pw3 * f_1 * d_12;
    fac[2] = -f_2_pw2 * p_2_pw2 * p_1_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2 - f_2_pw2 * p_2_pw2 * d_12_pw2 + f_2_pw2 * p_2_pw4 +
             p_2_pw4 * f_1_pw2 + 2 * p_1 * p_2_pw2 * d_12 + 2 * f_1 * f_2 * p_1 * p_2_pw2 * d_12 * b - p_2_pw2 * p_1_pw2 * f_1_pw2 +
             2 * p_1 * p_2_pw2 * f_2_pw2 * d_12 - p_2_pw2 * d_12_pw2 * b_pw2 - 2 * p_1_pw2 * p_2_pw2;
    fac[3] = 2 * p_1_pw2 * p_2 * d_12 * b + 2 * f_2 * p_2_pw3 * f_1 * d_12 - 2 * f_2_pw2 * p_2_pw3 * d_12 * b - 2 * p_1 * p_2 * d_12_pw2 * b;
    fac[4] = -2 * f_2 * p_2_pw2 * f_1 * p_1 * d_12 * b + f_2_pw2 * p_2_pw2 * d_12_pw2 + 2 * p_1_pw3 * d_12 - p_1_pw2 * d_12_pw2 +
             f_2_pw2 * p_2_pw2 * p_1_pw2 - p_1_pw4 - 2 * f_2_pw2 * p_2_pw2 * p_1 * d_12 + p_2_pw2 * f_1_pw2 * p_1_pw2 +
             f_2_pw2 * p_2_pw2 * d_12_pw2 * b_pw2;
    double roots[4];
    o4_roots(fac, roots);
    for (int k = 0; k < 4; k++) {
        const double r = roots[k];
        const double cot_alpha = (-f_1 * p_1 / f_2 - r * p_2 + d_12 * b) / (-f_1 * r * p_2 / f_2 + p_1 - d_12);
        const double cos_theta = r, sin_theta = sqrt(1 - r * r);
        const double sin_alpha = sqrt(1 / (cot_alpha * cot_alpha + 1));
        double cos_alpha = sqrt(1 - sin_alpha * sin_alpha);
        if (cot_alpha < 0) cos_alpha = -cos_alpha;
        const double k0 = sin_alpha * b + cos_alpha;
        const double Cc[3] = {d_12 * cos_alpha * k0, cos_theta * d_12 * sin_alpha * k0, sin_theta * d_12 * sin_alpha * k0};
        const double Rm[9] = {-cos_alpha, -sin_alpha * cos_theta, -sin_alpha * sin_theta,
                              sin_alpha,  -cos_alpha * cos_theta, -cos_alpha * sin_theta,
                              0.0,        -sin_theta,             cos_theta};
        double* S = sol + 12 * k;
        /* C = P1 + N^T Cc ;  R = N^T Rm^T T */
        for (int i = 0; i < 3; i++) S[4 * i + 3] = P1[i] + (N[i] * Cc[0] + N[3 + i] * Cc[1] + N[6 + i] * Cc[2]);
        double NR[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) NR[3 * i + j] = N[i] * Rm[3 * j] + N[3 + i] * Rm[3 * j + 1] + N[6 + i] * Rm[3 * j + 2];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) S[4 * i + j] = NR[3 * i] * T[j] + NR[3 * i + 1] * T[3 + j] + NR[3 * i + 2] * T[6 + j];
    }
    return 4;
}

/* 1 - f . normalise(R^T (X - t))   (AbsolutePoseSacProblem::getSelectedDistancesToModel, central camera) */
static double bearing_dist(const double* T, const double* X, const double* f)
{
    double q[3];
    for (int i = 0; i < 3; i++) {
        const double tr = -(T[i] * T[3] + T[4 + i] * T[7] + T[8 + i] * T[11]);   /* (-R^T t)_i */
        q[i] = T[i] * X[0] + T[4 + i] * X[1] + T[8 + i] * X[2] + tr;
    }
    const double n = norm3(q);
    return 1.0 - (q[0] / n * f[0] + q[1] / n * f[1] + q[2] / n * f[2]);
}

/* AbsolutePoseSacProblem::computeModelCoefficients (KNEIP): P3P on the first three sample indices, the fourth
 * disambiguates.  Returns 0 when no model (degenerate sample, or no solution with a finite score). */
int orc_p3p_sample_model(const double* bvs, const double* wpts, const int* idx, double* T)
{
    double f[9], p[9], sol[48];
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) { f[3 * k + i] = bvs[3 * idx[k] + i]; p[3 * k + i] = wpts[3 * idx[k] + i]; }
    if (orc_p3p_kneip(f, p, sol) != 4) return 0;
    double best = 1000000.0;
    int bi = -1;
    for (int k = 0; k < 4; k++) {
        const double s = bearing_dist(sol + 12 * k, wpts + 3 * idx[3], bvs + 3 * idx[3]);
        if (s < best) { best = s; bi = k; }
    }
    if (bi < 0) return 0;
    memcpy(T, sol + 12 * bi, sizeof(double) * 12);
    return 1;
}

static int cmp_double(const void* a, const void* b)
{
    const double x = *(const double*)a, y = *(const double*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* MultiViewGeometry::p3pRansac with optimize = false (the only way the reference calls it, visual_frontend.cpp:299-312):
 * LMedS over max_iter successful draws; threshold = 1 - cos(atan(err_px / focal)) in the reference's float arithmetic.
 * bvs [n][3] unit bearing vectors, wpts [n][3].  Twc_out: 3x4 row-major [R | t]; outlier [n] (1 = outlier).
 * Returns 1 on success (>= 5 inliers and an orthogonal R), else 0.  info[0] = #inliers, [1] = best median, [2] = #draws. */
int orc_p3p_lmeds(const double* bvs, const double* wpts, int n, int max_iter, float err_px, float fx, float fy, uint32_t seed,
                  double* Twc_out, uint8_t* outlier, double* info)
{
    if (n < 4) return 0;
    float focal = fx + fy;
    focal = (float)(focal / 2.);
    const double threshold = 1.0 - cosf(atanf(err_px / focal));
    mt_t m;
    mt_seed(&m, seed);
    int* sh = (int*)malloc(sizeof(int) * n);
    double* dist = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) sh[i] = i;
    double best = DBL_MAX, bestT[12];
    int have = 0, iterations = 0, draws = 0;
    unsigned skipped = 0;
    const unsigned max_skip = (unsigned)max_iter * 10;
    while (iterations < max_iter && skipped < max_skip) {
        int idx[4];
        for (int i = 0; i < 4; i++) {
            const int j = i + (int)((mt_next(&m) >> 1) % (uint32_t)(n - i));
            const int t = sh[i]; sh[i] = sh[j]; sh[j] = t;
        }
        for (int i = 0; i < 4; i++) idx[i] = sh[i];
        draws++;
        double T[12];
        if (!orc_p3p_sample_model(bvs, wpts, idx, T)) { skipped++; continue; }
        for (int i = 0; i < n; i++) {
            double d = bearing_dist(T, wpts + 3 * i, bvs + 3 * i);
            if (d < 0) d = 0;
            dist[i] = d * d;
        }
        qsort(dist, n, sizeof(double), cmp_double);   /* NaNs (never produced by a selected model) would break any sort */
        const int mid = n / 2;
        const double pen = (n % 2 == 0) ? (dist[mid - 1] + dist[mid]) / 2 : dist[mid];
        if (pen < best) { best = pen; memcpy(bestT, T, sizeof bestT); have = 1; }
        iterations++;
    }
    int ninl = 0, ok = 0;
    if (have) {
        for (int i = 0; i < n; i++) {
            const double d = bearing_dist(bestT, wpts + 3 * i, bvs + 3 * i);
            outlier[i] = !(d <= threshold);
            ninl += !outlier[i];
        }
        /* Sophus::isOrthogonal: |R R^T - I|_F < 1e-10  (sophus/rotation_matrix.hpp) */
        double e = 0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                const double v = bestT[4 * i] * bestT[4 * j] + bestT[4 * i + 1] * bestT[4 * j + 1] + bestT[4 * i + 2] * bestT[4 * j + 2] - (i == j);
                e += v * v;
            }
        ok = ninl >= 5 && sqrt(e) < 1e-10;
        memcpy(Twc_out, bestT, sizeof bestT);
    } else {
        for (int i = 0; i < n; i++) outlier[i] = 1;
    }
    if (info) { info[0] = ninl; info[1] = best; info[2] = draws; }
    free(sh); free(dist);
    return ok;
}

/* ================================================================== ceresPnP */
static void q_normalize(const double* q, double* o)
{
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) o[i] = q[i] / n;
}
static void q_to_R(const double* q, double* R)
{
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
void orc_se3_plus(const double* x, const double* delta, double* out);   /* ba_oracle.c: SE3Parameterization::Plus */

/* DirectSE3::ReprojectionErrorSE3::Evaluate: pose = [t_wc, q_wc(x,y,z,w)]; res[2]; J[2][6] local (first 6 of the 2x7);
 * returns the depth-positive flag; *chi2 = |res|^2 */
int orc_pnp_evaluate(const double* K, const double* pose, const double* X, const double* uv, double* res, double* J, double* chi2)
{
    double q[4], R[9];
    q_normalize(pose + 3, q);
    q_to_R(q, R);
    const double d[3] = {X[0] - pose[0], X[1] - pose[1], X[2] - pose[2]};
    double c[3];
    for (int i = 0; i < 3; i++) c[i] = R[i] * d[0] + R[3 + i] * d[1] + R[6 + i] * d[2];   /* Rcw (X - t) */
    const double iz = 1. / c[2];
    res[0] = K[0] * c[0] * iz + K[2] - uv[0];
    res[1] = K[1] * c[1] * iz + K[3] - uv[1];
    if (chi2) *chi2 = res[0] * res[0] + res[1] * res[1];
    if (J) {
        const double iz2 = iz * iz;
        const double Jc[6] = {iz * K[0], 0, -c[0] * iz2 * K[0], 0, iz * K[1], -c[1] * iz2 * K[1]};
        double JR[6];
        for (int r = 0; r < 2; r++)
            for (int k = 0; k < 3; k++) JR[3 * r + k] = Jc[3 * r] * R[3 * k] + Jc[3 * r + 1] * R[3 * k + 1] + Jc[3 * r + 2] * R[3 * k + 2];
        const double Sk[9] = {0, -X[2], X[1], X[2], 0, -X[0], -X[1], X[0], 0};
        for (int r = 0; r < 2; r++)
            for (int k = 0; k < 3; k++) {
                J[6 * r + k] = -JR[3 * r + k];
                J[6 * r + 3 + k] = JR[3 * r] * Sk[k] + JR[3 * r + 1] * Sk[3 + k] + JR[3 * r + 2] * Sk[6 + k];
            }
    }
    return c[2] > 0;
}

static void pnp_huber(double s, double delta, double* rho0, double* rho1)
{
    if (delta > 0 && s > delta * delta) {
        const double r = sqrt(s);
        *rho0 = 2 * delta * r - delta * delta;
        const double v = delta / r;
        *rho1 = v > DBL_MIN ? v : DBL_MIN;
    } else { *rho0 = s; *rho1 = 1.0; }
}

typedef struct { const double *K, *X, *uv; const uint8_t* removed; int n; double huber; } pnp_view;

static double pnp_cost(const pnp_view* v, const double* pose)
{
    double cost = 0;
    for (int i = 0; i < v->n; i++) {
        if (v->removed && v->removed[i]) continue;
        double r[2], s, r0, r1;
        orc_pnp_evaluate(v->K, pose, v->X + 3 * i, v->uv + 2 * i, r, 0, &s);
        pnp_huber(s, v->huber, &r0, &r1);
        cost += 0.5 * r0;
    }
    return cost;
}
/* normal equations of the corrected problem: H = J^T J (6x6), g = J^T r, column norms = diag(H); returns the cost */
static double pnp_linearize(const pnp_view* v, const double* pose, double* H, double* g)
{
    double cost = 0;
    memset(H, 0, sizeof(double) * 36);
    memset(g, 0, sizeof(double) * 6);
    for (int i = 0; i < v->n; i++) {
        if (v->removed && v->removed[i]) continue;
        double r[2], J[12], s, r0, r1;
        orc_pnp_evaluate(v->K, pose, v->X + 3 * i, v->uv + 2 * i, r, J, &s);
        pnp_huber(s, v->huber, &r0, &r1);
        cost += 0.5 * r0;
        const double sc = sqrt(r1);
        r[0] *= sc; r[1] *= sc;
        for (int k = 0; k < 12; k++) J[k] *= sc;
        for (int a = 0; a < 6; a++) {
            g[a] += J[a] * r[0] + J[6 + a] * r[1];
            for (int b = 0; b < 6; b++) H[6 * a + b] += J[a] * J[b] + J[6 + a] * J[6 + b];
        }
    }
    return cost;
}
static int chol6(const double* S, const double* b, double* x)
{
    double L[36];
    memcpy(L, S, sizeof L);
    for (int j = 0; j < 6; j++) {
        double d = L[6 * j + j];
        for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k];
        if (!(d > 0)) return 0;
        d = sqrt(d);
        L[6 * j + j] = d;
        for (int i = j + 1; i < 6; i++) {
            double s = L[6 * i + j];
            for (int k = 0; k < j; k++) s -= L[6 * i + k] * L[6 * j + k];
            L[6 * i + j] = s / d;
        }
    }
    for (int i = 0; i < 6; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[6 * i + k] * x[k]; x[i] = s / L[6 * i + i]; }
    for (int i = 5; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < 6; k++) s -= L[6 * k + i] * x[k]; x[i] = s / L[6 * i + i]; }
    return 1;
}
static double v7norm(const double* p) { double s = 0; for (int i = 0; i < 7; i++) s += p[i] * p[i]; return sqrt(s); }
static double v7diff(const double* a, const double* b, int inf)
{
    double s = 0;
    for (int i = 0; i < 7; i++) { const double e = fabs(a[i] - b[i]); if (inf) { if (e > s) s = e; } else s += e * e; }
    return inf ? s : sqrt(s);
}

/* one ceres::Solve of the PnP problem (trust-region LM exactly as ba_oracle.c's ba_solve_impl, one 6-dof block).
 * last_pose: the point the functors were evaluated at last.  summary: initial cost, final cost, #successful, #iterations,
 * termination (0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE). */
static int pnp_solve(const pnp_view* v, double* pose, int max_iter, double* summary, double* last_pose)
{
    double H[36], g[6], nf[6], sc[6], diag[6], S[36], rhs[6], y[6], df[6], cand[7], gneg[6];
    double radius = 1e4, decrease_factor = 2.0;
    int reuse_diagonal = 0, invalid_steps = 0;
    double x_cost = pnp_linearize(v, pose, H, g);
    memcpy(last_pose, pose, sizeof(double) * 7);
    for (int i = 0; i < 6; i++) { nf[i] = H[7 * i]; sc[i] = 1.0 / (1.0 + sqrt(nf[i])); }
    double xn = v7norm(pose);
    double se_min = x_cost, se_cur = x_cost, se_ref = x_cost, se_cand = x_cost, se_acc_ref = 0, se_acc_cand = 0;
    int n_success = 0, n_iter = 0, term = 1;
    summary[0] = x_cost;
    double gmax;
    for (int i = 0; i < 6; i++) gneg[i] = -g[i];
    orc_se3_plus(pose, gneg, cand);
    gmax = v7diff(pose, cand, 1);
    int iteration = 0, last_success = 1;
    for (;;) {
        if (last_success) n_success++;
        n_iter++;
        if (iteration >= max_iter) { term = 1; break; }
        if (last_success && gmax <= 1e-10) { term = 0; break; }
        if (radius <= 1e-32) { term = 0; break; }
        iteration++;
        last_success = 0;
        if (!reuse_diagonal)
            for (int i = 0; i < 6; i++) diag[i] = fmin(fmax(nf[i] * sc[i] * sc[i], 1e-6), 1e32);
        reuse_diagonal = 1;
        for (int a = 0; a < 6; a++) {
            rhs[a] = g[a] * sc[a];
            for (int b = 0; b < 6; b++) S[6 * a + b] = H[6 * a + b] * sc[a] * sc[b];
            S[7 * a] += diag[a] / radius;
        }
        const int ok = chol6(S, rhs, y);
        double model_change = -1;
        if (ok) {
            /* step = -y (scaled space); model cost change = -(step^T g_s + 0.5 step^T H_s step), H_s without the damping */
            double lin = 0, quad = 0;
            for (int a = 0; a < 6; a++) {
                lin += -y[a] * rhs[a];
                for (int b = 0; b < 6; b++) quad += y[a] * (H[6 * a + b] * sc[a] * sc[b]) * y[b];
            }
            model_change = -(lin + 0.5 * quad);
        }
        if (!ok || !(model_change > 0.0)) {
            if (++invalid_steps >= 5) { term = 2; break; }
            radius *= 0.5;
            reuse_diagonal = 1;
            continue;
        }
        invalid_steps = 0;
        for (int i = 0; i < 6; i++) df[i] = -y[i] * sc[i];
        orc_se3_plus(pose, df, cand);
        const double cand_cost = pnp_cost(v, cand);
        memcpy(last_pose, cand, sizeof(double) * 7);
        const double step_norm = v7diff(pose, cand, 0);
        if (step_norm <= 1e-8 * (xn + 1e-8)) { term = 0; break; }
        if (fabs(x_cost - cand_cost) <= 1e-3 * x_cost) { term = 0; break; }
        const double rel = (se_cur - cand_cost) / model_change;
        const double hist = (se_ref - cand_cost) / (se_acc_ref + model_change);
        const double quality = rel > hist ? rel : hist;
        if (quality > 1e-3) {
            memcpy(pose, cand, sizeof(double) * 7);
            xn = v7norm(pose);
            x_cost = pnp_linearize(v, pose, H, g);
            for (int i = 0; i < 6; i++) nf[i] = H[7 * i];
            for (int i = 0; i < 6; i++) gneg[i] = -g[i];
            orc_se3_plus(pose, gneg, cand);
            gmax = v7diff(pose, cand, 1);
            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * quality - 1.0, 3));
            radius = fmin(1e16, radius);
            decrease_factor = 2.0;
            reuse_diagonal = 0;
            se_cur = cand_cost; se_acc_cand += model_change; se_acc_ref += model_change;
            int nonmono = 0;
            if (se_cur < se_min) { se_min = se_cur; se_cand = se_cur; se_acc_cand = 0; }
            else { nonmono = 1; if (se_cur > se_cand) { se_cand = se_cur; se_acc_cand = 0; } }
            if (!nonmono) { se_ref = se_cand; se_acc_ref = se_acc_cand; }
            last_success = 1;
        } else {
            radius = radius / decrease_factor;
            decrease_factor *= 2.0;
            reuse_diagonal = 1;
        }
    }
    summary[1] = x_cost; summary[2] = n_success; summary[3] = n_iter; summary[4] = term;
    return term != 2;
}

/* MultiViewGeometry::ceresPnP (multi_view_geometry.cpp:129-223), wall-clock cap lifted: robust solve (Huber huber_delta),
 * flag residuals with chi2 > chi2_thr or non-positive depth at their last evaluation, optionally remove them and re-solve
 * without the loss.  K = {fx, fy, cx, cy}; pose [t, q(x,y,z,w)] in/out; outlier [n] out.  Returns the reference's bool
 * (0 when every point is an outlier -- the pose is then left at its input value, as in the reference -- or the last solve
 * failed).  summary [10]: the two solves as in pnp_solve (zeros if the second is skipped). */
int orc_pnp(const double* K, const double* uv, const double* X, int n, double* pose, double huber_delta, double chi2_thr,
            int max_iter, int use_robust, int apply_l2, uint8_t* outlier, double* summary)
{
    pnp_view v = {K, X, uv, 0, n, use_robust ? huber_delta : 0.0};
    double work[7], last[7];
    memcpy(work, pose, sizeof work);
    for (int i = 0; i < 10; i++) summary[i] = 0;
    int usable = pnp_solve(&v, work, max_iter, summary, last);
    int nbad = 0;
    for (int i = 0; i < n; i++) {
        double r[2], s;
        const int dp = orc_pnp_evaluate(K, last, X + 3 * i, uv + 2 * i, r, 0, &s);
        outlier[i] = (s > chi2_thr || !dp) ? 1 : 0;
        nbad += outlier[i];
    }
    if (nbad == n) return 0;
    if (apply_l2 && nbad > 0) {
        v.removed = outlier;
        v.huber = 0.0;
        usable = pnp_solve(&v, work, max_iter, summary + 5, last);
    }
    memcpy(pose, work, sizeof work);
    return usable;
}
