// lmdif_core.h -- the minimiser the reference's 5-point refinement runs, restated: MINPACK's Levenberg-Marquardt on a
// forward-difference Jacobian, in the form Eigen's unsupported NonLinearOptimization module drives it.
//
// Reference behaviour (file:line under /root/reference/src/libs):
//   opengv/src/relative_pose/methods.cpp:1152-1177     optimize_nonlinear: NumericalDiff<Functor> (forward, step sqrt(eps) |x_j|)
//                                                      + LevenbergMarquardt, ftol = xtol = 10 eps, maxfev = 1000, lm.minimize(x)
//   eigen/unsupported/Eigen/src/NumericalDiff/NumericalDiff.h:63-127          df(): f(x) once, then one f per column
//   eigen/unsupported/Eigen/src/NonLinearOptimization/LevenbergMarquardt.h    minimizeInit / minimizeOneStep (MINPACK lmder):
//        column norms, QR with column pivoting, scaling diag = max(diag, column norm), trust region delta = 100 |D x|,
//        lmpar (Newton iteration on the LM parameter, <= 10 steps), gain ratio, the delta / par updates and the stopping tests
//   eigen/unsupported/Eigen/src/NonLinearOptimization/lmpar.h (lmpar2), qrsolv.h
// Why it exists: the cost is (1 - cos) residuals of ~1e-7 whose forward differences at sqrt(eps) are dominated by rounding
// noise, so where this iteration stalls is a property of the ALGORITHM, not of the cost's minimum.  A cleaner minimiser
// (central differences, as round 1 used) converges somewhere else in the flat valley: 2e-3 away in translation on the golden
// trace, 7x the reference's own spread under a 1-ulp input change.  Following the reference's iteration puts the result inside
// that spread.  Not bit-identical by construction (Eigen's blueNorm / Householder / Givens kernels round differently), and it
// need not be: the reference itself moves by the same amount when one input bit changes.
//
// Plain scalar code behind ALVA_HD: the device runs it in one thread (the stage runs once per session), the CPU suite compiles
// the same source for the host (tests/host/init_core_host.cpp).
#pragma once
#include <float.h>
#include <math.h>

#ifndef ALVA_HD
#ifdef __CUDACC__
#define ALVA_HD __host__ __device__
#else
#define ALVA_HD
#endif
#endif

namespace alva_lm {

constexpr int N = 6;

ALVA_HD inline double norm_n(const double* v, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) s += v[i] * v[i];
    return sqrt(s);
}

// QR of the m x N matrix a (column-major, a[j * m + i]) with column pivoting, Householder vectors stored below the diagonal,
// R in the upper triangle; tau[k] the Householder coefficients; perm[k] = original index of the column now in position k.
ALVA_HD inline void qr_colpiv(double* a, int m, double* tau, int* perm) {
    double cn[N];
    for (int j = 0; j < N; j++) { perm[j] = j; cn[j] = 0; for (int i = 0; i < m; i++) cn[j] += a[j * m + i] * a[j * m + i]; }
    for (int k = 0; k < N; k++) {
        int big = k;
        for (int j = k + 1; j < N; j++) if (cn[j] > cn[big]) big = j;
        if (big != k) {
            for (int i = 0; i < m; i++) { const double t = a[k * m + i]; a[k * m + i] = a[big * m + i]; a[big * m + i] = t; }
            const double t = cn[k]; cn[k] = cn[big]; cn[big] = t;
            const int p = perm[k]; perm[k] = perm[big]; perm[big] = p;
        }
        // Householder of column k, rows k..m-1: H = I - tau v v^T, v = [1; essential]
        double* c = a + k * m;
        double tail = 0;
        for (int i = k + 1; i < m; i++) tail += c[i] * c[i];
        const double c0 = c[k];
        double beta;
        if (tail <= DBL_MIN) { tau[k] = 0; beta = c0; for (int i = k + 1; i < m; i++) c[i] = 0; }
        else {
            beta = sqrt(c0 * c0 + tail);
            if (c0 >= 0) beta = -beta;
            const double d = c0 - beta;
            for (int i = k + 1; i < m; i++) c[i] /= d;
            tau[k] = (beta - c0) / beta;
        }
        c[k] = beta;
        for (int j = k + 1; j < N; j++) {
            double* cj = a + j * m;
            double s = cj[k];
            for (int i = k + 1; i < m; i++) s += c[i] * cj[i];
            s *= tau[k];
            cj[k] -= s;
            for (int i = k + 1; i < m; i++) cj[i] -= s * c[i];
            // remaining squared norm of column j below row k (recomputed: m is a few hundred, N is 6)
            double r = 0;
            for (int i = k + 1; i < m; i++) r += cj[i] * cj[i];
            cn[j] = r;
        }
    }
}
#define ALVA_LM_R(i, j) a[(j) * m + (i)]

// qrsolv: given R (upper triangle of a), the permutation, a diagonal D (indexed by ORIGINAL column) and Q^T b, solve
// min |[R P^T; D] x - [Q^T b; 0]|.  s (N x N, row-major) receives the lower-triangular factor, sdiag its diagonal.
ALVA_HD inline void qrsolv(const double* a, int m, const int* perm, const double* diag, const double* qtb, double* x, double* sdiag,
                           double* s) {
    double wa[N];
    for (int j = 0; j < N; j++) {
        for (int i = j; i < N; i++) s[i * N + j] = ALVA_LM_R(j, i);   // lower triangle := R^T
        x[j] = s[j * N + j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < N; j++) {
        const int l = perm[j];
        if (diag[l] != 0.) {
            for (int k = j; k < N; k++) sdiag[k] = 0.;
            sdiag[j] = diag[l];
            double qtbpj = 0.;
            for (int k = j; k < N; k++) {
                if (sdiag[k] == 0.) continue;
                double cs, sn;
                if (fabs(s[k * N + k]) < fabs(sdiag[k])) {
                    const double cot = s[k * N + k] / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * cot * cot);
                    cs = sn * cot;
                } else {
                    const double tn = sdiag[k] / s[k * N + k];
                    cs = 0.5 / sqrt(0.25 + 0.25 * tn * tn);
                    sn = cs * tn;
                }
                s[k * N + k] = cs * s[k * N + k] + sn * sdiag[k];
                const double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < N; i++) {
                    const double t2 = cs * s[i * N + k] + sn * sdiag[i];
                    sdiag[i] = -sn * s[i * N + k] + cs * sdiag[i];
                    s[i * N + k] = t2;
                }
            }
        }
        sdiag[j] = s[j * N + j];
        s[j * N + j] = x[j];
    }
    int nsing = N;
    for (int j = 0; j < N; j++) {
        if (sdiag[j] == 0. && nsing == N) nsing = j;
        if (nsing < N) wa[j] = 0.;
    }
    for (int j = nsing - 1; j >= 0; j--) {
        double sum = 0.;
        for (int i = j + 1; i < nsing; i++) sum += s[i * N + j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < N; j++) x[perm[j]] = wa[j];
}

// lmpar: the Levenberg-Marquardt parameter par and the step x with |D x| ~ delta
ALVA_HD inline void lmpar(const double* a, int m, const int* perm, const double* diag, const double* qtb, double delta, double& par,
                          double* x) {
    const double dwarf = DBL_MIN;
    double wa1[N], wa2[N], sdiag[N], s[N * N];
    // Gauss-Newton direction; rank: leading diagonal entries of R that are not negligible
    int rank = N;
    {
        double maxd = 0;
        for (int j = 0; j < N; j++) maxd = fmax(maxd, fabs(ALVA_LM_R(j, j)));
        const double thr = maxd * DBL_EPSILON * (double)(m < N ? m : N);
        rank = 0;
        for (int j = 0; j < N; j++) if (fabs(ALVA_LM_R(j, j)) > thr) rank++; else break;
    }
    for (int j = 0; j < N; j++) wa1[j] = j < rank ? qtb[j] : 0.;
    for (int j = rank - 1; j >= 0; j--) {
        wa1[j] /= ALVA_LM_R(j, j);
        for (int i = 0; i < j; i++) wa1[i] -= ALVA_LM_R(i, j) * wa1[j];
    }
    for (int j = 0; j < N; j++) x[perm[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < N; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = norm_n(wa2, N);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { par = 0; return; }
    double parl = 0.;
    if (rank == N) {
        for (int j = 0; j < N; j++) { const int l = perm[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < N; j++) {   // solve R^T w = wa1
            double sum = 0.;
            for (int i = 0; i < j; i++) sum += ALVA_LM_R(i, j) * wa1[i];
            wa1[j] = (wa1[j] - sum) / ALVA_LM_R(j, j);
        }
        const double temp = norm_n(wa1, N);
        parl = fp / delta / temp / temp;
    }
    for (int j = 0; j < N; j++) {
        double sum = 0.;
        for (int i = 0; i <= j; i++) sum += ALVA_LM_R(i, j) * qtb[i];
        wa1[j] = sum / diag[perm[j]];
    }
    const double gnorm = norm_n(wa1, N);
    double paru = gnorm / delta;
    if (paru == 0.) paru = dwarf / fmin(delta, 0.1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.) par = gnorm / dxnorm;
    while (true) {
        ++iter;
        if (par == 0.) par = fmax(dwarf, 0.001 * paru);
        double dsc[N];
        const double sp = sqrt(par);
        for (int j = 0; j < N; j++) dsc[j] = sp * diag[j];
        qrsolv(a, m, perm, dsc, qtb, x, sdiag, s);
        for (int j = 0; j < N; j++) wa2[j] = diag[j] * x[j];
        dxnorm = norm_n(wa2, N);
        const double temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0. && fp <= temp && temp < 0.) || iter == 10) break;
        for (int j = 0; j < N; j++) { const int l = perm[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < N; j++) {
            wa1[j] /= sdiag[j];
            const double t = wa1[j];
            for (int i = j + 1; i < N; i++) wa1[i] -= s[i * N + j] * t;
        }
        const double t2 = norm_n(wa1, N);
        const double parc = fp / delta / t2 / t2;
        if (fp > 0.) parl = fmax(parl, par);
        if (fp < 0.) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
    if (iter == 0) par = 0.;
}

// F: void operator()(const double* x, double* fvec) -- m residuals.  Work arrays: fvec[m], fjac[N * m], wa4[m].
// Returns Eigen's LevenbergMarquardtSpace::Status code (1..8; 5 = maxfev reached).
template <class F>
ALVA_HD inline int lmdif(F& fun, int m, double* x, double* fvec, double* fjac, double* wa4, double ftol, double xtol, int maxfev,
                         int* nfev_out = nullptr) {
    const double eps = DBL_EPSILON, factor = 100., gtol = 0.;
    const double dstep = sqrt(eps);   // NumericalDiff: eps = sqrt(max(epsfcn = 0, epsilon))
    double diag[N], qtf[N], wa1[N], wa2[N], wa3[N], tau[N];
    int perm[N];
    double* a = fjac;
    int nfev = 1, iter = 1, status = 0;
    fun(x, fvec);
    double fnorm = norm_n(fvec, m), par = 0., delta = 0., xnorm = 0.;
    while (status == 0) {
        // forward-difference Jacobian: NumericalDiff::df evaluates f(x) again, then one column at a time
        fun(x, wa4);
        for (int j = 0; j < N; j++) {
            double h = dstep * fabs(x[j]);
            if (h == 0.) h = dstep;
            const double keep = x[j];
            x[j] = keep + h;
            fun(x, a + j * m);
            x[j] = keep;
            for (int i = 0; i < m; i++) a[j * m + i] = (a[j * m + i] - wa4[i]) / h;
        }
        nfev += N + 1;
        for (int j = 0; j < N; j++) wa2[j] = norm_n(a + j * m, m);
        qr_colpiv(a, m, tau, perm);
        if (iter == 1) {
            for (int j = 0; j < N; j++) diag[j] = wa2[j] == 0. ? 1. : wa2[j];
            for (int j = 0; j < N; j++) wa3[j] = diag[j] * x[j];
            xnorm = norm_n(wa3, N);
            delta = factor * xnorm;
            if (delta == 0.) delta = factor;
        }
        // qtf = first N components of Q^T fvec
        for (int i = 0; i < m; i++) wa4[i] = fvec[i];
        for (int k = 0; k < N; k++) {
            const double* c = a + k * m;
            double s = wa4[k];
            for (int i = k + 1; i < m; i++) s += c[i] * wa4[i];
            s *= tau[k];
            wa4[k] -= s;
            for (int i = k + 1; i < m; i++) wa4[i] -= s * c[i];
        }
        for (int j = 0; j < N; j++) qtf[j] = wa4[j];
        double gnorm = 0.;
        if (fnorm != 0.)
            for (int j = 0; j < N; j++)
                if (wa2[perm[j]] != 0.) {
                    double sum = 0.;
                    for (int i = 0; i <= j; i++) sum += ALVA_LM_R(i, j) * (qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / wa2[perm[j]]));
                }
        if (gnorm <= gtol) { status = 4; break; }
        for (int j = 0; j < N; j++) diag[j] = fmax(diag[j], wa2[j]);
        double ratio;
        do {
            lmpar(a, m, perm, diag, qtf, delta, par, wa1);
            for (int j = 0; j < N; j++) { wa1[j] = -wa1[j]; wa2[j] = x[j] + wa1[j]; wa3[j] = diag[j] * wa1[j]; }
            const double pnorm = norm_n(wa3, N);
            if (iter == 1) delta = fmin(delta, pnorm);
            fun(wa2, wa4);
            ++nfev;
            const double fnorm1 = norm_n(wa4, m);
            double actred = -1.;
            if (0.1 * fnorm1 < fnorm) actred = 1. - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            for (int i = 0; i < N; i++) {   // R * (P^-1 step)
                double sum = 0.;
                for (int j = i; j < N; j++) sum += ALVA_LM_R(i, j) * wa1[perm[j]];
                wa3[i] = sum;
            }
            const double t1 = norm_n(wa3, N) / fnorm, t2 = sqrt(par) * pnorm / fnorm;
            const double temp1 = t1 * t1, temp2 = t2 * t2;
            const double prered = temp1 + temp2 / 0.5;
            const double dirder = -(temp1 + temp2);
            ratio = prered != 0. ? actred / prered : 0.;
            if (ratio <= 0.25) {
                double temp = actred >= 0. ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par /= temp;
            } else if (!(par != 0. && ratio < 0.75)) {
                delta = pnorm / 0.5;
                par = 0.5 * par;
            }
            if (ratio >= 1e-4) {
                for (int j = 0; j < N; j++) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                for (int i = 0; i < m; i++) fvec[i] = wa4[i];
                xnorm = norm_n(wa2, N);
                fnorm = fnorm1;
                ++iter;
            }
            const bool fsmall = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.;
            if (fsmall && delta <= xtol * xnorm) { status = 3; break; }
            if (fsmall) { status = 1; break; }
            if (delta <= xtol * xnorm) { status = 2; break; }
            if (nfev >= maxfev) { status = 5; break; }
            if (fabs(actred) <= eps && prered <= eps && 0.5 * ratio <= 1.) { status = 6; break; }
            if (delta <= eps * xnorm) { status = 7; break; }
            if (gnorm <= eps) { status = 8; break; }
        } while (ratio < 1e-4);
    }
    if (nfev_out) *nfev_out = nfev;
    return status;
}
#undef ALVA_LM_R

}  // namespace alva_lm
