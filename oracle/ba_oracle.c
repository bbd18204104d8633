/* oracle/ba_oracle.c -- CPU restatement of the reference's local bundle adjustment numerics.
 *
 * TEST INFRASTRUCTURE ONLY (see alva_oracle.c header).  Restates, in plain C doubles:
 *   - the BA cost functor  DirectSE3::ReprojectionErrorKSE3AnchInvDepth::Evaluate
 *       src/slam/src/ceres_parametrization.cpp:157-269  (anchored inverse depth, poses T_wc = [t, q(x,y,z,w)])
 *   - the pose plus-op     SE3Parameterization::Plus / ComputeJacobian
 *       src/slam/src/ceres_parametrization.hpp:224-248 (T+ = exp(delta) * T, delta = [upsilon, omega], J_local = [I6; 0])
 *       with Sophus SE3::exp / SO3::expAndTheta (src/libs/Sophus/sophus/se3.hpp:763-784, so3.hpp:585-621)
 *   - Huber loss + corrector as Ceres applies them (ceres-solver/internal/ceres/loss_function.cc:48-62,
 *       corrector.cc:36-134, residual_block.cc:133-195): in both Huber regions rho'' <= 0, so residual and
 *       Jacobian are scaled by sqrt(rho')
 *   - the problem Optimizer::localBA builds (src/slam/src/optimizer.cpp:20-262: calib constant, inverse depths in
 *       elimination group 0, poses in group 1, >= 2 constant keyframes, HuberLoss(sqrt(5.9915)), SPARSE_SCHUR,
 *       LEVENBERG_MARQUARDT, <= 5 iterations, function_tolerance 1e-3; the wall-clock cap is lifted)
 *   - Ceres' trust-region loop (trust_region_minimizer.cc:60-130, 232-300, 399-450, 700-830): Jacobi column
 *       scaling 1/(1+sqrt(|col|^2)) fixed at iteration 0, LM diagonal sqrt(clamp(|col|^2,1e-6,1e32)/radius)
 *       (levenberg_marquardt_strategy.cc:66-100), Schur elimination of the 1-dim inverse-depth blocks
 *       (schur_eliminator_impl.h:176-375), step quality / radius update (:147-160, trust_region_step_evaluator.cc),
 *       parameter / function / gradient tolerance tests, defaults from include/ceres/solver.h:263-322.
 * Pinned against ceres::Solve itself through oracle/_ref (tests/test_oracle_ba.py) and golden vectors.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

/* ------------------------------------------------------------------ small SE(3) helpers */
static void quat_normalize(const double* q, double* o) /* (x,y,z,w) */
{
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) o[i] = q[i] / n;
}
static void quat_to_R(const double* q, double* R) /* Eigen::Quaternion::toRotationMatrix */
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
    double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
static void quat_mul(const double* a, const double* b, double* o) /* a*b, (x,y,z,w) */
{
    double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
}
static void mat3_vec(const double* R, const double* v, double* o)
{
    for (int i = 0; i < 3; i++) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}

/* SE3Parameterization::Plus: x = [t, q], delta = [upsilon, omega] -> exp(delta) * T */
void orc_se3_plus(const double* x, const double* delta, double* out)
{
    const double* ups = delta;
    const double* om = delta + 3;
    double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    double theta, imag, real;
    const double eps = 1e-10;
    if (theta_sq < eps * eps) {
        theta = 0;
        double t4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
        real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * t4;
    } else {
        theta = sqrt(theta_sq);
        double half = 0.5 * theta;
        imag = sin(half) / theta;
        real = cos(half);
    }
    double dq[4] = {imag * om[0], imag * om[1], imag * om[2], real};
    double Rd[9];
    quat_to_R(dq, Rd);
    double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0}, O2[9], V[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += O[3 * i + k] * O[3 * k + j];
            O2[3 * i + j] = s;
        }
    if (theta < eps) memcpy(V, Rd, sizeof V);
    else {
        double a = (1 - cos(theta)) / theta_sq, b = (theta - sin(theta)) / (theta_sq * theta);
        for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
    }
    double td[3], qn[4], Rt[3];
    mat3_vec(V, ups, td);
    quat_normalize(x + 3, qn);
    mat3_vec(Rd, x, Rt);
    for (int i = 0; i < 3; i++) out[i] = td[i] + Rt[i];
    double qo[4];
    quat_mul(dq, qn, qo);
    /* Sophus SO3 product renormalises when needed; dq and qn are unit to 1 ulp */
    quat_normalize(qo, out + 3);
}

/* ------------------------------------------------------------------ cost functor */
/* res[2]; Ja, Jp: 2x6 row-major LOCAL Jacobians (first six columns of the 2x7 global ones); Jd[2].
 * returns depth-positive flag; *chi2 = |res|^2. */
int orc_ba_evaluate(const double* calib, const double* anch, const double* pose, double invd, const double* obs,
                    double* res, double* Ja, double* Jp, double* Jd, double* chi2)
{
    const double fx = calib[0], fy = calib[1], cx = calib[2], cy = calib[3];
    double qa[4], qc[4], Rwa[9], Rwc[9];
    quat_normalize(anch + 3, qa);
    quat_normalize(pose + 3, qc);
    quat_to_R(qa, Rwa);
    quat_to_R(qc, Rwc);
    const double zanch = 1.0 / invd;
    /* anchpt = zanch * invK * [ua, va, 1] */
    double ap[3] = {zanch * (obs[2] - cx) / fx, zanch * (obs[3] - cy) / fy, zanch};
    double wpt[3], tmp[3];
    mat3_vec(Rwa, ap, tmp);
    for (int i = 0; i < 3; i++) wpt[i] = tmp[i] + anch[i];
    /* lcampt = Rcw * (wpt - twc) */
    double d[3] = {wpt[0] - pose[0], wpt[1] - pose[1], wpt[2] - pose[2]}, cp[3];
    for (int i = 0; i < 3; i++) cp[i] = Rwc[i] * d[0] + Rwc[3 + i] * d[1] + Rwc[6 + i] * d[2];
    const double iz = 1.0 / cp[2];
    res[0] = fx * cp[0] * iz + cx - obs[0];
    res[1] = fy * cp[1] * iz + cy - obs[1];
    if (chi2) *chi2 = res[0] * res[0] + res[1] * res[1];
    if (Ja || Jp || Jd) {
        const double iz2 = iz * iz;
        double Jc[6] = {iz * fx, 0, -cp[0] * iz2 * fx, 0, iz * fy, -cp[1] * iz2 * fy};
        double JR[6]; /* J_lRcw = Jc * Rcw, Rcw = Rwc^T */
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++)
                JR[3 * r + c] = Jc[3 * r] * Rwc[3 * c] + Jc[3 * r + 1] * Rwc[3 * c + 1] + Jc[3 * r + 2] * Rwc[3 * c + 2];
        /* skew(wpt) */
        double Sk[9] = {0, -wpt[2], wpt[1], wpt[2], 0, -wpt[0], -wpt[1], wpt[0], 0};
        double JS[6];
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++)
                JS[3 * r + c] = JR[3 * r] * Sk[c] + JR[3 * r + 1] * Sk[3 + c] + JR[3 * r + 2] * Sk[6 + c];
        if (Ja)
            for (int r = 0; r < 2; r++)
                for (int c = 0; c < 3; c++) { Ja[6 * r + c] = JR[3 * r + c]; Ja[6 * r + 3 + c] = -JS[3 * r + c]; }
        if (Jp)
            for (int r = 0; r < 2; r++)
                for (int c = 0; c < 3; c++) { Jp[6 * r + c] = -JR[3 * r + c]; Jp[6 * r + 3 + c] = JS[3 * r + c]; }
        if (Jd) {
            double Ra[3];
            mat3_vec(Rwa, ap, Ra);
            for (int r = 0; r < 2; r++)
                Jd[r] = -zanch * (JR[3 * r] * Ra[0] + JR[3 * r + 1] * Ra[1] + JR[3 * r + 2] * Ra[2]);
        }
    }
    return cp[2] > 0;
}

/* Huber as Ceres: rho[0], rho[1] for s = |r|^2; delta <= 0 -> trivial loss */
static void huber(double s, double delta, double* rho0, double* rho1)
{
    if (delta > 0 && s > delta * delta) {
        double r = sqrt(s);
        *rho0 = 2 * delta * r - delta * delta;
        double v = delta / r;
        *rho1 = v > DBL_MIN ? v : DBL_MIN;
    } else { *rho0 = s; *rho1 = 1.0; }
}

/* ------------------------------------------------------------------ problem view */
typedef struct {
    const double* calib;
    int nkf, nlm, nobs;
    const uint8_t* pose_const;
    const int32_t *anch_kf, *obs_kf, *obs_lm;
    const double *anch_uv, *obs_uv;
    double huber;
    /* reduced program */
    int* pose_col;  /* nkf: column offset of a free, referenced pose (else -1) */
    int* lm_used;   /* nlm: has >= 1 residual */
    int npose_free, ncols_f;
} ba_view;

static double ba_cost_only(const ba_view* v, const double* poses, const double* invd)
{
    double cost = 0;
    for (int o = 0; o < v->nobs; o++) {
        int l = v->obs_lm[o];
        if (l < 0) continue; /* removed residual block */
        double obs[4] = {v->obs_uv[2 * o], v->obs_uv[2 * o + 1], v->anch_uv[2 * l], v->anch_uv[2 * l + 1]}, r[2], s;
        orc_ba_evaluate(v->calib, poses + 7 * v->anch_kf[l], poses + 7 * v->obs_kf[o], invd[l], obs, r, 0, 0, 0, &s);
        double r0, r1;
        huber(s, v->huber, &r0, &r1);
        cost += 0.5 * r0;
    }
    return cost;
}

/* residuals (corrected), local Jacobians (corrected, UNSCALED); returns cost */
static double ba_linearize(const ba_view* v, const double* poses, const double* invd, double* res, double* Ja,
                           double* Jp, double* Jd)
{
    double cost = 0;
    for (int o = 0; o < v->nobs; o++) {
        int l = v->obs_lm[o];
        if (l < 0) {
            res[2 * o] = res[2 * o + 1] = Jd[2 * o] = Jd[2 * o + 1] = 0;
            for (int i = 0; i < 12; i++) Ja[12 * o + i] = Jp[12 * o + i] = 0;
            continue;
        }
        double obs[4] = {v->obs_uv[2 * o], v->obs_uv[2 * o + 1], v->anch_uv[2 * l], v->anch_uv[2 * l + 1]}, s;
        orc_ba_evaluate(v->calib, poses + 7 * v->anch_kf[l], poses + 7 * v->obs_kf[o], invd[l], obs, res + 2 * o,
                        Ja + 12 * o, Jp + 12 * o, Jd + 2 * o, &s);
        double r0, r1;
        huber(s, v->huber, &r0, &r1);
        cost += 0.5 * r0;
        double sc = sqrt(r1);
        res[2 * o] *= sc; res[2 * o + 1] *= sc;
        for (int i = 0; i < 12; i++) { Ja[12 * o + i] *= sc; Jp[12 * o + i] *= sc; }
        Jd[2 * o] *= sc; Jd[2 * o + 1] *= sc;
    }
    return cost;
}

/* public: linearisation dump for kernel-level parity (corrected residuals / local Jacobians) */
double orc_ba_linearize(const double* calib, const double* poses, int nkf, const double* invd, const int32_t* anch_kf,
                        const double* anch_uv, int nlm, const int32_t* obs_kf, const int32_t* obs_lm,
                        const double* obs_uv, int nobs, double huber_delta, double* res, double* Ja, double* Jp,
                        double* Jd)
{
    ba_view v = {calib, nkf, nlm, nobs, 0, anch_kf, obs_kf, obs_lm, anch_uv, obs_uv, huber_delta, 0, 0, 0, 0};
    return ba_linearize(&v, poses, invd, res, Ja, Jp, Jd);
}

/* squared column norms of the (unscaled) Jacobian: out_f[ncols_f], out_e[nlm] */
static void ba_colnorms(const ba_view* v, const double* Ja, const double* Jp, const double* Jd, double* nf, double* ne)
{
    memset(nf, 0, sizeof(double) * v->ncols_f);
    memset(ne, 0, sizeof(double) * v->nlm);
    for (int o = 0; o < v->nobs; o++) {
        if (v->obs_lm[o] < 0) continue;
        int l = v->obs_lm[o], ca = v->pose_col[v->anch_kf[l]], cp = v->pose_col[v->obs_kf[o]];
        ne[l] += Jd[2 * o] * Jd[2 * o] + Jd[2 * o + 1] * Jd[2 * o + 1];
        for (int c = 0; c < 6; c++) {
            if (ca >= 0) nf[ca + c] += Ja[12 * o + c] * Ja[12 * o + c] + Ja[12 * o + 6 + c] * Ja[12 * o + 6 + c];
            if (cp >= 0) nf[cp + c] += Jp[12 * o + c] * Jp[12 * o + c] + Jp[12 * o + 6 + c] * Jp[12 * o + 6 + c];
        }
    }
    /* NOTE: when the anchor and the observing keyframe coincide (never produced by localBA: the anchor is the first
     * observer and adds no residual) the two cells would share columns; not supported here. */
}

/* Build the reduced camera system for scaled Jacobian J*diag(sc) and LM diagonal D (both in reduced column order:
 * sc_f/D_f[ncols_f], sc_e/D_e[nlm]).  S[n*n] row-major, rhs[n];  ete[nlm], etb[nlm] kept for back-substitution. */
static void ba_schur(const ba_view* v, const double* res, const double* Ja, const double* Jp, const double* Jd,
                     const double* sc_f, const double* sc_e, const double* D_f, const double* D_e, double* S,
                     double* rhs, double* ete, double* etb, double* W /* nlm x n : E^T F */)
{
    const int n = v->ncols_f;
    memset(S, 0, sizeof(double) * n * n);
    memset(rhs, 0, sizeof(double) * n);
    memset(W, 0, sizeof(double) * (size_t)v->nlm * n);
    for (int l = 0; l < v->nlm; l++) { ete[l] = D_e[l] * D_e[l]; etb[l] = 0; }
    for (int i = 0; i < n; i++) S[i * n + i] = D_f[i] * D_f[i];
    for (int o = 0; o < v->nobs; o++) {
        if (v->obs_lm[o] < 0) continue;
        int l = v->obs_lm[o], ca = v->pose_col[v->anch_kf[l]], cp = v->pose_col[v->obs_kf[o]];
        double e[2] = {Jd[2 * o] * sc_e[l], Jd[2 * o + 1] * sc_e[l]};
        double F[2][12];
        int cols[12], nc = 0;
        for (int c = 0; c < 6; c++)
            if (ca >= 0) { cols[nc] = ca + c; F[0][nc] = Ja[12 * o + c] * sc_f[ca + c]; F[1][nc] = Ja[12 * o + 6 + c] * sc_f[ca + c]; nc++; }
        for (int c = 0; c < 6; c++)
            if (cp >= 0) { cols[nc] = cp + c; F[0][nc] = Jp[12 * o + c] * sc_f[cp + c]; F[1][nc] = Jp[12 * o + 6 + c] * sc_f[cp + c]; nc++; }
        ete[l] += e[0] * e[0] + e[1] * e[1];
        etb[l] += e[0] * res[2 * o] + e[1] * res[2 * o + 1];
        for (int a = 0; a < nc; a++) {
            W[(size_t)l * n + cols[a]] += e[0] * F[0][a] + e[1] * F[1][a];
            rhs[cols[a]] += F[0][a] * res[2 * o] + F[1][a] * res[2 * o + 1];
            for (int b = 0; b < nc; b++) S[cols[a] * n + cols[b]] += F[0][a] * F[0][b] + F[1][a] * F[1][b];
        }
    }
    for (int l = 0; l < v->nlm; l++) {
        if (!v->lm_used[l]) continue;
        const double inv = 1.0 / ete[l];
        const double* w = W + (size_t)l * n;
        for (int a = 0; a < n; a++) {
            if (w[a] == 0) continue;
            rhs[a] -= w[a] * inv * etb[l];
            for (int b = 0; b < n; b++)
                if (w[b] != 0) S[a * n + b] -= w[a] * inv * w[b];
        }
    }
}

/* dense Cholesky solve (S symmetric positive definite), in place on copies */
static int chol_solve(int n, const double* S, const double* b, double* x)
{
    double* L = (double*)malloc(sizeof(double) * n * n);
    memcpy(L, S, sizeof(double) * n * n);
    for (int j = 0; j < n; j++) {
        double d = L[j * n + j];
        for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0)) { free(L); return 0; }
        d = sqrt(d);
        L[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = L[i * n + j];
            for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
        x[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
    free(L);
    return 1;
}

/* public: reduced system of the FIRST LM iteration (radius 1e4) for kernel-level parity.
 * S[n*n], rhs[n], n = 6 * (#free referenced poses); pose_col_out[nkf]. Returns n. */
int orc_ba_first_schur(const double* calib, const double* poses, const uint8_t* pose_const, int nkf, const double* invd,
                       const int32_t* anch_kf, const double* anch_uv, int nlm, const int32_t* obs_kf,
                       const int32_t* obs_lm, const double* obs_uv, int nobs, double huber_delta, double* S,
                       double* rhs, int32_t* pose_col_out);

/* ------------------------------------------------------------------ the LM loop */
static void apply_step(const ba_view* v, const double* poses, const double* invd, const double* delta_f,
                       const double* delta_e, double* poses_o, double* invd_o)
{
    memcpy(poses_o, poses, sizeof(double) * 7 * v->nkf);
    memcpy(invd_o, invd, sizeof(double) * v->nlm);
    for (int k = 0; k < v->nkf; k++)
        if (v->pose_col[k] >= 0) orc_se3_plus(poses + 7 * k, delta_f + v->pose_col[k], poses_o + 7 * k);
    for (int l = 0; l < v->nlm; l++)
        if (v->lm_used[l]) invd_o[l] = invd[l] + delta_e[l];
}

static double x_norm(const ba_view* v, const double* poses, const double* invd)
{
    double s = 0;
    for (int k = 0; k < v->nkf; k++)
        if (v->pose_col[k] >= 0)
            for (int i = 0; i < 7; i++) s += poses[7 * k + i] * poses[7 * k + i];
    for (int l = 0; l < v->nlm; l++)
        if (v->lm_used[l]) s += invd[l] * invd[l];
    return sqrt(s);
}

static double x_diff(const ba_view* v, const double* p0, const double* d0, const double* p1, const double* d1, int inf)
{
    double s = 0;
    for (int k = 0; k < v->nkf; k++)
        if (v->pose_col[k] >= 0)
            for (int i = 0; i < 7; i++) {
                double e = fabs(p0[7 * k + i] - p1[7 * k + i]);
                if (inf) { if (e > s) s = e; } else s += e * e;
            }
    for (int l = 0; l < v->nlm; l++)
        if (v->lm_used[l]) {
            double e = fabs(d0[l] - d1[l]);
            if (inf) { if (e > s) s = e; } else s += e * e;
        }
    return inf ? s : sqrt(s);
}

static void setup_view(ba_view* v, int* pose_col, int* lm_used)
{
    int* ref = (int*)calloc(v->nkf, sizeof(int));
    memset(lm_used, 0, sizeof(int) * v->nlm);
    for (int o = 0; o < v->nobs; o++) {
        int l = v->obs_lm[o];
        if (l < 0) continue;
        lm_used[l] = 1;
        ref[v->anch_kf[l]] = 1;
        ref[v->obs_kf[o]] = 1;
    }
    int c = 0, np = 0;
    for (int k = 0; k < v->nkf; k++) {
        if (!v->pose_const[k] && ref[k]) { pose_col[k] = c; c += 6; np++; }
        else pose_col[k] = -1;
    }
    v->pose_col = pose_col; v->lm_used = lm_used; v->npose_free = np; v->ncols_f = c;
    free(ref);
}

int orc_ba_first_schur(const double* calib, const double* poses, const uint8_t* pose_const, int nkf, const double* invd,
                       const int32_t* anch_kf, const double* anch_uv, int nlm, const int32_t* obs_kf,
                       const int32_t* obs_lm, const double* obs_uv, int nobs, double huber_delta, double* S,
                       double* rhs, int32_t* pose_col_out)
{
    ba_view v = {calib, nkf, nlm, nobs, pose_const, anch_kf, obs_kf, obs_lm, anch_uv, obs_uv, huber_delta, 0, 0, 0, 0};
    int* pose_col = (int*)malloc(sizeof(int) * nkf);
    int* lm_used = (int*)malloc(sizeof(int) * nlm);
    setup_view(&v, pose_col, lm_used);
    const int n = v.ncols_f;
    double *res = malloc(sizeof(double) * 2 * nobs), *Ja = malloc(sizeof(double) * 12 * nobs),
           *Jp = malloc(sizeof(double) * 12 * nobs), *Jd = malloc(sizeof(double) * 2 * nobs);
    double *nf = malloc(sizeof(double) * (n + 1)), *ne = malloc(sizeof(double) * nlm), *scf = malloc(sizeof(double) * (n + 1)),
           *sce = malloc(sizeof(double) * nlm), *Df = malloc(sizeof(double) * (n + 1)), *De = malloc(sizeof(double) * nlm);
    double *ete = malloc(sizeof(double) * nlm), *etb = malloc(sizeof(double) * nlm), *W = malloc(sizeof(double) * (size_t)nlm * (n + 1));
    ba_linearize(&v, poses, invd, res, Ja, Jp, Jd);
    ba_colnorms(&v, Ja, Jp, Jd, nf, ne);
    for (int i = 0; i < n; i++) scf[i] = 1.0 / (1.0 + sqrt(nf[i]));
    for (int l = 0; l < nlm; l++) sce[l] = 1.0 / (1.0 + sqrt(ne[l]));
    const double radius = 1e4;
    for (int i = 0; i < n; i++) { double d = nf[i] * scf[i] * scf[i]; d = fmin(fmax(d, 1e-6), 1e32); Df[i] = sqrt(d / radius); }
    for (int l = 0; l < nlm; l++) { double d = ne[l] * sce[l] * sce[l]; d = fmin(fmax(d, 1e-6), 1e32); De[l] = sqrt(d / radius); }
    ba_schur(&v, res, Ja, Jp, Jd, scf, sce, Df, De, S, rhs, ete, etb, W);
    for (int k = 0; k < nkf; k++) pose_col_out[k] = pose_col[k];
    free(res); free(Ja); free(Jp); free(Jd); free(nf); free(ne); free(scf); free(sce); free(Df); free(De);
    free(ete); free(etb); free(W); free(pose_col); free(lm_used);
    return n;
}

/* Same contract as ref_ba_solve in oracle/ref_harness.cpp.  summary[0..4] = initial cost, final cost,
 * #successful steps, #iterations (including iteration 0), termination (0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE). */
/* last_p / last_d (optional): the point at which the cost functors were evaluated LAST -- the candidate of the last
 * iteration that got as far as evaluating one (accepted or not), else the start point.  The reference reads chi2err_ /
 * isDepthPositive_ out of the functors after Solve (optimizer.cpp:273-299), i.e. at exactly this point. */
static int ba_solve_impl(const double* calib, double* poses, const uint8_t* pose_const, int nkf, double* invd,
                         const int32_t* anch_kf, const double* anch_uv, int nlm, const int32_t* obs_kf,
                         const int32_t* obs_lm, const double* obs_uv, int nobs, double huber_delta, int max_iter,
                         double* summary, double* iter_costs, double* last_p, double* last_d)
{
    ba_view v = {calib, nkf, nlm, nobs, pose_const, anch_kf, obs_kf, obs_lm, anch_uv, obs_uv, huber_delta, 0, 0, 0, 0};
    int* pose_col = (int*)malloc(sizeof(int) * nkf);
    int* lm_used = (int*)malloc(sizeof(int) * nlm);
    setup_view(&v, pose_col, lm_used);
    const int n = v.ncols_f;
    double *res = malloc(sizeof(double) * 2 * nobs), *Ja = malloc(sizeof(double) * 12 * nobs),
           *Jp = malloc(sizeof(double) * 12 * nobs), *Jd = malloc(sizeof(double) * 2 * nobs);
    double *nf = malloc(sizeof(double) * (n + 1)), *ne = malloc(sizeof(double) * nlm), *scf = malloc(sizeof(double) * (n + 1)),
           *sce = malloc(sizeof(double) * nlm), *Df = malloc(sizeof(double) * (n + 1)), *De = malloc(sizeof(double) * nlm),
           *diag_f = malloc(sizeof(double) * (n + 1)), *diag_e = malloc(sizeof(double) * nlm);
    double *ete = malloc(sizeof(double) * nlm), *etb = malloc(sizeof(double) * nlm), *W = malloc(sizeof(double) * (size_t)nlm * (n + 1));
    double *S = malloc(sizeof(double) * (n + 1) * (n + 1)), *rhs = malloc(sizeof(double) * (n + 1)), *yf = malloc(sizeof(double) * (n + 1)),
           *ye = malloc(sizeof(double) * nlm), *df = malloc(sizeof(double) * (n + 1)), *de = malloc(sizeof(double) * nlm);
    double *cand_p = malloc(sizeof(double) * 7 * nkf), *cand_d = malloc(sizeof(double) * nlm), *gf = malloc(sizeof(double) * (n + 1)),
           *ge = malloc(sizeof(double) * nlm);

    double radius = 1e4, decrease_factor = 2.0;
    int reuse_diagonal = 0, invalid_steps = 0;
    double x_cost = ba_linearize(&v, poses, invd, res, Ja, Jp, Jd);
    if (last_p) { memcpy(last_p, poses, sizeof(double) * 7 * nkf); memcpy(last_d, invd, sizeof(double) * nlm); }
    ba_colnorms(&v, Ja, Jp, Jd, nf, ne);
    for (int i = 0; i < n; i++) scf[i] = 1.0 / (1.0 + sqrt(nf[i]));
    for (int l = 0; l < nlm; l++) sce[l] = 1.0 / (1.0 + sqrt(ne[l]));
    double xn = x_norm(&v, poses, invd);
    /* step evaluator state (monotonic) */
    double se_min = x_cost, se_cur = x_cost, se_ref = x_cost, se_cand = x_cost, se_acc_ref = 0, se_acc_cand = 0;
    int n_success = 0, n_iter = 0, term = 1;
    summary[0] = x_cost;

    /* gradient max norm at x (for the gradient-tolerance test) */
#define GRAD_MAXNORM(out)                                                                                             \
    do {                                                                                                              \
        memset(gf, 0, sizeof(double) * n);                                                                            \
        memset(ge, 0, sizeof(double) * nlm);                                                                          \
        for (int o = 0; o < nobs; o++) {                                                                              \
            if (obs_lm[o] < 0) continue;                                                                              \
            int l = obs_lm[o], ca = pose_col[anch_kf[l]], cp = pose_col[obs_kf[o]];                                   \
            ge[l] += Jd[2 * o] * res[2 * o] + Jd[2 * o + 1] * res[2 * o + 1];                                         \
            for (int c = 0; c < 6; c++) {                                                                             \
                if (ca >= 0) gf[ca + c] += Ja[12 * o + c] * res[2 * o] + Ja[12 * o + 6 + c] * res[2 * o + 1];         \
                if (cp >= 0) gf[cp + c] += Jp[12 * o + c] * res[2 * o] + Jp[12 * o + 6 + c] * res[2 * o + 1];         \
            }                                                                                                         \
        }                                                                                                             \
        for (int i = 0; i < n; i++) gf[i] = -gf[i];                                                                   \
        for (int l = 0; l < nlm; l++) ge[l] = -ge[l];                                                                 \
        apply_step(&v, poses, invd, gf, ge, cand_p, cand_d);                                                          \
        out = x_diff(&v, poses, invd, cand_p, cand_d, 1);                                                             \
    } while (0)

    double gmax;
    GRAD_MAXNORM(gmax);
    int iteration = 0, last_success = 1;
    double push_cost = x_cost;
    for (;;) {
        /* FinalizeIterationAndCheckIfMinimizerCanContinue: push the iteration summary, then the stop tests */
        if (last_success) n_success++;
        if (iter_costs && n_iter < 64) iter_costs[n_iter] = push_cost;
        n_iter++;
        if (iteration >= max_iter) { term = 1; break; }
        if (last_success && gmax <= 1e-10) { term = 0; break; }
        if (radius <= 1e-32) { term = 0; break; }
        iteration++;
        last_success = 0;
        /* LevenbergMarquardtStrategy::ComputeStep */
        if (!reuse_diagonal) {
            for (int i = 0; i < n; i++) diag_f[i] = fmin(fmax(nf[i] * scf[i] * scf[i], 1e-6), 1e32);
            for (int l = 0; l < nlm; l++) diag_e[l] = fmin(fmax(ne[l] * sce[l] * sce[l], 1e-6), 1e32);
        }
        for (int i = 0; i < n; i++) Df[i] = sqrt(diag_f[i] / radius);
        for (int l = 0; l < nlm; l++) De[l] = sqrt(diag_e[l] / radius);
        reuse_diagonal = 1;
        ba_schur(&v, res, Ja, Jp, Jd, scf, sce, Df, De, S, rhs, ete, etb, W);
        int ok = chol_solve(n, S, rhs, yf);
        double model_change = -1;
        if (ok) {
            for (int l = 0; l < nlm; l++) {
                if (!lm_used[l]) { ye[l] = 0; continue; }
                double s = etb[l];
                const double* w = W + (size_t)l * n;
                for (int a = 0; a < n; a++) s -= w[a] * yf[a];
                ye[l] = s / ete[l];
            }
            /* step = -y ; model residuals = J_scaled * step */
            double acc = 0;
            for (int o = 0; o < nobs; o++) {
                if (obs_lm[o] < 0) continue;
                int l = obs_lm[o], ca = pose_col[anch_kf[l]], cp = pose_col[obs_kf[o]];
                double m[2] = {Jd[2 * o] * sce[l] * -ye[l], Jd[2 * o + 1] * sce[l] * -ye[l]};
                for (int c = 0; c < 6; c++) {
                    if (ca >= 0) { m[0] += Ja[12 * o + c] * scf[ca + c] * -yf[ca + c]; m[1] += Ja[12 * o + 6 + c] * scf[ca + c] * -yf[ca + c]; }
                    if (cp >= 0) { m[0] += Jp[12 * o + c] * scf[cp + c] * -yf[cp + c]; m[1] += Jp[12 * o + 6 + c] * scf[cp + c] * -yf[cp + c]; }
                }
                acc += m[0] * (res[2 * o] + m[0] / 2.0) + m[1] * (res[2 * o + 1] + m[1] / 2.0);
            }
            model_change = -acc;
        }
        if (!ok || !(model_change > 0.0)) {
            /* HandleInvalidStep (trust_region_minimizer.cc:452-480) + LM StepIsInvalid */
            if (++invalid_steps >= 5) { term = 2; break; }
            radius *= 0.5;
            reuse_diagonal = 1;
            push_cost = x_cost;
            continue;
        }
        invalid_steps = 0;
        for (int i = 0; i < n; i++) df[i] = -yf[i] * scf[i];
        for (int l = 0; l < nlm; l++) de[l] = -ye[l] * sce[l];
        apply_step(&v, poses, invd, df, de, cand_p, cand_d);
        double cand_cost = ba_cost_only(&v, cand_p, cand_d);
        if (last_p) { memcpy(last_p, cand_p, sizeof(double) * 7 * nkf); memcpy(last_d, cand_d, sizeof(double) * nlm); }
        /* ParameterToleranceReached / FunctionToleranceReached: return WITHOUT taking the step or pushing a summary */
        double step_norm = x_diff(&v, poses, invd, cand_p, cand_d, 0);
        if (step_norm <= 1e-8 * (xn + 1e-8)) { term = 0; break; }
        if (fabs(x_cost - cand_cost) <= 1e-3 * x_cost) { term = 0; break; }
        /* IsStepSuccessful */
        double rel = (se_cur - cand_cost) / model_change;
        double hist = (se_ref - cand_cost) / (se_acc_ref + model_change);
        double quality = rel > hist ? rel : hist;
        if (quality > 1e-3) {
            memcpy(poses, cand_p, sizeof(double) * 7 * nkf);
            memcpy(invd, cand_d, sizeof(double) * nlm);
            xn = x_norm(&v, poses, invd);
            x_cost = ba_linearize(&v, poses, invd, res, Ja, Jp, Jd);
            ba_colnorms(&v, Ja, Jp, Jd, nf, ne);
            GRAD_MAXNORM(gmax);
            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * quality - 1.0, 3));
            radius = fmin(1e16, radius);
            decrease_factor = 2.0;
            reuse_diagonal = 0;
            /* TrustRegionStepEvaluator::StepAccepted, max_consecutive_nonmonotonic_steps = 0 */
            se_cur = cand_cost; se_acc_cand += model_change; se_acc_ref += model_change;
            int nonmono = 0;
            if (se_cur < se_min) { se_min = se_cur; se_cand = se_cur; se_acc_cand = 0; }
            else { nonmono = 1; if (se_cur > se_cand) { se_cand = se_cur; se_acc_cand = 0; } }
            if (!nonmono) { se_ref = se_cand; se_acc_ref = se_acc_cand; }
            last_success = 1;
            push_cost = x_cost;
        } else {
            radius = radius / decrease_factor;
            decrease_factor *= 2.0;
            reuse_diagonal = 1;
            push_cost = cand_cost;
        }
    }
#undef GRAD_MAXNORM
    summary[1] = x_cost; summary[2] = n_success; summary[3] = n_iter; summary[4] = term;
    free(res); free(Ja); free(Jp); free(Jd); free(nf); free(ne); free(scf); free(sce); free(Df); free(De); free(diag_f);
    free(diag_e); free(ete); free(etb); free(W); free(S); free(rhs); free(yf); free(ye); free(df); free(de); free(cand_p);
    free(cand_d); free(gf); free(ge); free(pose_col); free(lm_used);
    return term != 2;
}

int orc_ba_solve(const double* calib, double* poses, const uint8_t* pose_const, int nkf, double* invd,
                 const int32_t* anch_kf, const double* anch_uv, int nlm, const int32_t* obs_kf, const int32_t* obs_lm,
                 const double* obs_uv, int nobs, double huber_delta, int max_iter, double* summary, double* iter_costs)
{
    return ba_solve_impl(calib, poses, pose_const, nkf, invd, anch_kf, anch_uv, nlm, obs_kf, obs_lm, obs_uv, nobs,
                         huber_delta, max_iter, summary, iter_costs, 0, 0);
}

/* chi2 = |r|^2 of the RAW residual (sigma = 1) and the depth sign, per observation, at (poses, invd):
 * ReprojectionErrorKSE3AnchInvDepth::Evaluate's side outputs (ceres_parametrization.hpp:136-139). */
static int ba_flag_outliers(const double* calib, const double* poses, const double* invd, const int32_t* anch_kf,
                            const double* anch_uv, const int32_t* obs_kf, const int32_t* obs_lm, const double* obs_uv,
                            int nobs, double chi2_thr, int32_t* flags, int mark)
{
    int n = 0;
    for (int o = 0; o < nobs; o++) {
        int l = obs_lm[o];
        if (l < 0) continue;
        double obs[4] = {obs_uv[2 * o], obs_uv[2 * o + 1], anch_uv[2 * l], anch_uv[2 * l + 1]}, r[2], s;
        int depth_pos = orc_ba_evaluate(calib, poses + 7 * anch_kf[l], poses + 7 * obs_kf[o], invd[l], obs, r, 0, 0, 0, &s);
        if (s > chi2_thr || !depth_pos) { flags[o] = mark; n++; }
    }
    return n;
}

/* Optimizer::localBA steps 2-4 (optimizer.cpp:251-359): solve, drop the residuals flagged at the last evaluated point,
 * and -- only if something was dropped (and the robust loss is on) -- solve again with <= 5 iterations and flag once
 * more.  flags[o]: 0 kept, 1 removed after the first solve, 2 flagged after the second.  Returns #removed (pass 1).
 * summary[0..4] / [5..9] as orc_ba_solve for the two solves (zeros when the second is skipped). */
int orc_ba_local(const double* calib, double* poses, const uint8_t* pose_const, int nkf, double* invd,
                 const int32_t* anch_kf, const double* anch_uv, int nlm, const int32_t* obs_kf, const int32_t* obs_lm,
                 const double* obs_uv, int nobs, double huber_delta, double chi2_thr, int max_iter, int32_t* flags,
                 double* summary)
{
    double* last_p = malloc(sizeof(double) * 7 * nkf);
    double* last_d = malloc(sizeof(double) * nlm);
    int32_t* lm = malloc(sizeof(int32_t) * nobs);
    memcpy(lm, obs_lm, sizeof(int32_t) * nobs);
    memset(flags, 0, sizeof(int32_t) * nobs);
    for (int i = 0; i < 10; i++) summary[i] = 0;
    ba_solve_impl(calib, poses, pose_const, nkf, invd, anch_kf, anch_uv, nlm, obs_kf, lm, obs_uv, nobs, huber_delta,
                  max_iter, summary, 0, last_p, last_d);
    int nbad = ba_flag_outliers(calib, last_p, last_d, anch_kf, anch_uv, obs_kf, lm, obs_uv, nobs, chi2_thr, flags, 1);
    if (huber_delta > 0 && nbad > 0) {
        for (int o = 0; o < nobs; o++)
            if (flags[o]) lm[o] = -1;
        ba_solve_impl(calib, poses, pose_const, nkf, invd, anch_kf, anch_uv, nlm, obs_kf, lm, obs_uv, nobs, huber_delta, 5,
                      summary + 5, 0, last_p, last_d);
        ba_flag_outliers(calib, last_p, last_d, anch_kf, anch_uv, obs_kf, lm, obs_uv, nobs, chi2_thr, flags, 2);
    }
    free(last_p); free(last_d); free(lm);
    return nbad;
}
