// init_core.h -- per-thread arithmetic of the map initialisation kernels (init.cu): Nister's five-point solver, the
// decomposition / disambiguation of an essential matrix, mid-point triangulation, the RANSAC bookkeeping and the pieces of the
// final non-linear refinement.  Everything is plain scalar C++ behind ALVA_HD so that the very same source the device runs is
// also compiled for the host by tests/init_core_host.cpp and checked against the oracle in the CPU suite (the product only
// ever calls it from CUDA kernels).
//
// Reference behaviour (file:line under /root/reference/src):
//   slam/src/multi_view_geometry.cpp:225-318    MultiViewGeometry::compute5ptEssentialMatrix
//   libs/opengv/include/opengv/sac/implementation/Ransac.hpp:44-143               Ransac::computeModel
//   libs/opengv/src/sac_problems/relative_pose/CentralRelativePoseSacProblem.cpp  computeModelCoefficients (NISTER, :38-255),
//                                                                                 getSelectedDistancesToModel (:257-294)
//   libs/opengv/src/relative_pose/methods.cpp:239-268, modules/main.cpp:135-276   fivept_nister
//   libs/opengv/src/relative_pose/methods.cpp:1085-1177                           optimize_nonlinear
//   libs/opengv/src/triangulation/methods.cpp:65-88                               triangulate2
#pragma once
#include <math.h>
#include <stdint.h>
#include <float.h>

#ifndef ALVA_HD
#ifdef __CUDACC__
#define ALVA_HD __host__ __device__
#else
#define ALVA_HD
#endif
#endif

namespace alva_init {

ALVA_HD inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
ALVA_HD inline void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
ALVA_HD inline void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
ALVA_HD inline double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// triangulate2: mid-point of the two rays; (R12, t12) = pose of camera 2 in camera 1, result in camera 1
ALVA_HD inline void triangulate2(const double* R12, const double* t12, const double* f1, const double* f2, double* out) {
    double f2u[3];
    for (int i = 0; i < 3; i++) f2u[i] = R12[3 * i] * f2[0] + R12[3 * i + 1] * f2[1] + R12[3 * i + 2] * f2[2];
    const double b0 = dot3(t12, f1), b1 = dot3(t12, f2u);
    const double a00 = dot3(f1, f1), a10 = dot3(f1, f2u), a01 = -a10, a11 = -dot3(f2u, f2u);
    const double invdet = 1.0 / (a00 * a11 - a01 * a10);
    const double l0 = (a11 * b0 - a01 * b1) * invdet, l1 = (a00 * b1 - a10 * b0) * invdet;
    for (int i = 0; i < 3; i++) out[i] = (l0 * f1[i] + (t12[i] + l1 * f2u[i])) / 2;
}

// (1 - cos) reprojection error in both views of one correspondence under (R, t)
ALVA_HD inline double relpose_dist(const double* R, const double* t, const double* f1, const double* f2) {
    double p[3], q[3], r2[3];
    triangulate2(R, t, f1, f2, p);
    for (int i = 0; i < 3; i++) q[i] = p[i] - t[i];
    for (int i = 0; i < 3; i++) r2[i] = R[i] * q[0] + R[3 + i] * q[1] + R[6 + i] * q[2];
    const double n1 = sqrt(dot3(p, p)), n2 = sqrt(dot3(r2, r2));
    const double e1 = 1.0 - (f1[0] * (p[0] / n1) + f1[1] * (p[1] / n1) + f1[2] * (p[2] / n1));
    const double e2 = 1.0 - (f2[0] * (r2[0] / n2) + f2[1] * (r2[1] / n2) + f2[2] * (r2[2] / n2));
    return e1 + e2;
}

// ---------------------------------------------------------------- trivariate polynomials of total degree <= 3
// monomial m = x^a y^b z^c packed as a | b << 2 | c << 4; Nister's order: ten eliminated monomials, then
// [x z^2, x z, x, y z^2, y z, y, z^3, z^2, z, 1]
ALVA_HD inline int mono_exp(int m) {
    const int T[20] = {3, 3 << 2, 2 | 1 << 2, 1 | 2 << 2, 2 | 1 << 4, 2, 2 << 2 | 1 << 4, 2 << 2, 1 | 1 << 2 | 1 << 4, 1 | 1 << 2,
                       1 | 2 << 4, 1 | 1 << 4, 1, 1 << 2 | 2 << 4, 1 << 2 | 1 << 4, 1 << 2, 3 << 4, 2 << 4, 1 << 4, 0};
    return T[m];
}
ALVA_HD inline int mono_index(int packed) {
    for (int i = 0; i < 20; i++)
        if (mono_exp(i) == packed) return i;
    return -1;
}
// out += s * p * q, p and q given as sparse lists of monomial indices (np / nq entries)
ALVA_HD inline void poly_mul_acc(const double* p, const double* q, double s, double* out) {
    for (int i = 0; i < 20; i++) {
        const double pi = p[i];
        if (pi == 0.0) continue;
        const int ei = mono_exp(i);
        for (int j = 0; j < 20; j++) {
            const double qj = q[j];
            if (qj == 0.0) continue;
            const int ej = mono_exp(j);
            const int a = (ei & 3) + (ej & 3), b = ((ei >> 2) & 3) + ((ej >> 2) & 3), c = (ei >> 4) + (ej >> 4);
            if (a + b + c > 3) continue;
            out[mono_index(a | b << 2 | c << 4)] += s * pi * qj;
        }
    }
}
ALVA_HD inline void upoly_mul(const double* a, int da, const double* b, int db, double* o) {
    for (int i = 0; i <= da + db; i++) o[i] = 0;
    for (int i = 0; i <= da; i++)
        for (int j = 0; j <= db; j++) o[i + j] += a[i] * b[j];
}
ALVA_HD inline double upoly_val(const double* a, int d, double z) {
    double v = a[d];
    for (int i = d - 1; i >= 0; i--) v = v * z + a[i];
    return v;
}

// Sturm chain (each member scaled by a positive factor) and the number of sign changes at z
struct Sturm { double c[11][11]; int deg[11]; int n; };
ALVA_HD inline void sturm_build(const double* p, int d, Sturm& S) {
    for (int k = 0; k < 11; k++) { S.deg[k] = 0; for (int i = 0; i < 11; i++) S.c[k][i] = 0; }
    for (int i = 0; i <= d; i++) S.c[0][i] = p[i];
    S.deg[0] = d;
    for (int i = 1; i <= d; i++) S.c[1][i - 1] = i * p[i];
    S.deg[1] = d - 1;
    S.n = d >= 1 ? 2 : 1;
    for (int k = 2; k <= d && S.n == k; k++) {
        double r[11];
        int dr = S.deg[k - 2];
        const int dq = S.deg[k - 1];
        for (int i = 0; i < 11; i++) r[i] = S.c[k - 2][i];
        const double* q = S.c[k - 1];
        while (dr >= dq) {
            const double f = r[dr] / q[dq];
            for (int i = 0; i <= dq; i++) r[dr - dq + i] -= f * q[i];
            r[dr] = 0;
            dr--;
        }
        double mx = 0;
        for (int i = 0; i <= dr; i++) mx = fmax(mx, fabs(r[i]));
        if (dr < 0 || !(mx >= 1e-300)) break;
        while (dr > 0 && fabs(r[dr]) < 1e-14 * mx) dr--;
        for (int i = 0; i <= dr; i++) S.c[k][i] = -r[i] / mx;
        S.deg[k] = dr;
        S.n = k + 1;
        if (dr == 0) break;
    }
}
ALVA_HD inline int sturm_changes(const Sturm& S, double z) {
    int changes = 0, last = 0;
    for (int k = 0; k < S.n; k++) {
        const double v = upoly_val(S.c[k], S.deg[k], z);
        const int s = v > 0 ? 1 : (v < 0 ? -1 : 0);
        if (s != 0) { if (last != 0 && s != last) changes++; last = s; }
    }
    return changes;
}
// all real roots of p (degree d <= 10), ascending; isolation by Sturm counts, then safeguarded Newton to convergence
ALVA_HD inline int real_roots(const double* p, int d, double* roots) {
    while (d > 0 && p[d] == 0.0) d--;
    if (d < 1) return 0;
    Sturm S;
    sturm_build(p, d, S);
    double bound = 0;
    for (int i = 0; i < d; i++) bound = fmax(bound, fabs(p[i] / p[d]));
    bound += 1.0;
    double slo[48], shi[48];
    int sclo[48], schi[48];
    int sp = 0, nr = 0;
    slo[0] = -bound; shi[0] = bound; sclo[0] = sturm_changes(S, -bound); schi[0] = sturm_changes(S, bound); sp = 1;
    double dp[11];
    for (int i = 1; i <= d; i++) dp[i - 1] = i * p[i];
    while (sp > 0) {
        sp--;
        double lo = slo[sp], hi = shi[sp];
        const int clo = sclo[sp], chi = schi[sp], n = clo - chi;
        if (n <= 0) continue;
        const double mid = 0.5 * (lo + hi);
        if (n > 1 && mid > lo && mid < hi && (hi - lo) > 1e-13 * fmax(1.0, fabs(mid)) && sp + 2 <= 48) {
            const int cm = sturm_changes(S, mid);
            slo[sp] = mid; shi[sp] = hi; sclo[sp] = cm; schi[sp] = chi; sp++;
            slo[sp] = lo; shi[sp] = mid; sclo[sp] = clo; schi[sp] = cm; sp++;
            continue;
        }
        if (n > 1) { for (int k = 0; k < n && nr < 10; k++) roots[nr++] = mid; continue; }
        double flo = upoly_val(p, d, lo), fhi = upoly_val(p, d, hi);
        double x = mid;
        if (flo == 0.0) x = lo;
        else if (fhi == 0.0) x = hi;
        else if ((flo < 0) != (fhi < 0)) {
            for (int it = 0; it < 200; it++) {
                const double f = upoly_val(p, d, x), df = upoly_val(dp, d - 1, x);
                if (f == 0.0) break;
                if ((f < 0) == (flo < 0)) { lo = x; flo = f; } else { hi = x; fhi = f; }
                double xn = df != 0.0 ? x - f / df : 0.5 * (lo + hi);
                if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
                if (fabs(xn - x) <= 4e-16 * fmax(1.0, fabs(x))) { x = xn; break; }
                x = xn;
            }
        }
        if (nr < 10) roots[nr++] = x;
    }
    for (int i = 1; i < nr; i++) { const double v = roots[i]; int j = i - 1; while (j >= 0 && roots[j] > v) { roots[j + 1] = roots[j]; j--; } roots[j + 1] = v; }
    return nr;
}

// fivept_nister: f1_i^T E f2_i = 0, i < 5; Es = up to 10 row-major E of unit Frobenius norm; returns their number
ALVA_HD inline int fivept_nister(const double* f1, const double* f2, double* Es) {
    double Q[5][9];
    for (int i = 0; i < 5; i++)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) Q[i][3 * r + c] = f1[3 * i + r] * f2[3 * i + c];
    for (int i = 0; i < 5; i++) {   // orthonormal basis of the row space
        for (int pass = 0; pass < 2; pass++)
            for (int j = 0; j < i; j++) {
                double d = 0;
                for (int k = 0; k < 9; k++) d += Q[i][k] * Q[j][k];
                for (int k = 0; k < 9; k++) Q[i][k] -= d * Q[j][k];
            }
        double n = 0;
        for (int k = 0; k < 9; k++) n += Q[i][k] * Q[i][k];
        n = sqrt(n);
        if (!(n > 1e-12)) return 0;
        for (int k = 0; k < 9; k++) Q[i][k] /= n;
    }
    double EE[4][9];   // orthonormal basis of the null space
    for (int b = 0; b < 4; b++) {
        double best = -1, bv[9];
        for (int e = 0; e < 9; e++) {
            double v[9];
            for (int k = 0; k < 9; k++) v[k] = (k == e);
            for (int pass = 0; pass < 2; pass++) {
                for (int j = 0; j < 5; j++) { double d = 0; for (int k = 0; k < 9; k++) d += v[k] * Q[j][k]; for (int k = 0; k < 9; k++) v[k] -= d * Q[j][k]; }
                for (int j = 0; j < b; j++) { double d = 0; for (int k = 0; k < 9; k++) d += v[k] * EE[j][k]; for (int k = 0; k < 9; k++) v[k] -= d * EE[j][k]; }
            }
            double n = 0;
            for (int k = 0; k < 9; k++) n += v[k] * v[k];
            if (n > best) { best = n; for (int k = 0; k < 9; k++) bv[k] = v[k]; }
        }
        best = sqrt(best);
        for (int k = 0; k < 9; k++) EE[b][k] = bv[k] / best;
    }
    const int LIN[4] = {12, 15, 18, 19};   // x, y, z, 1
    double Ep[9][20];
    for (int e = 0; e < 9; e++) {
        for (int k = 0; k < 20; k++) Ep[e][k] = 0;
        for (int b = 0; b < 4; b++) Ep[e][LIN[b]] = EE[b][e];
    }
    double A[10][20], A0[10][20];
    for (int r = 0; r < 10; r++)
        for (int k = 0; k < 20; k++) A[r][k] = 0;
    {   // det E = 0
        double m[3][20];
        for (int r = 0; r < 3; r++)
            for (int k = 0; k < 20; k++) m[r][k] = 0;
        poly_mul_acc(Ep[4], Ep[8], 1, m[0]); poly_mul_acc(Ep[5], Ep[7], -1, m[0]);
        poly_mul_acc(Ep[3], Ep[8], 1, m[1]); poly_mul_acc(Ep[5], Ep[6], -1, m[1]);
        poly_mul_acc(Ep[3], Ep[7], 1, m[2]); poly_mul_acc(Ep[4], Ep[6], -1, m[2]);
        poly_mul_acc(Ep[0], m[0], 1, A[0]); poly_mul_acc(Ep[1], m[1], -1, A[0]); poly_mul_acc(Ep[2], m[2], 1, A[0]);
    }
    {   // (E E^T - 1/2 trace(E E^T) I) E = 0
        double G[9][20];
        for (int r = 0; r < 9; r++)
            for (int k = 0; k < 20; k++) G[r][k] = 0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) poly_mul_acc(Ep[3 * i + k], Ep[3 * j + k], 1, G[3 * i + j]);
        for (int k = 0; k < 20; k++) {
            const double tr = G[0][k] + G[4][k] + G[8][k];
            G[0][k] -= 0.5 * tr; G[4][k] -= 0.5 * tr; G[8][k] -= 0.5 * tr;
        }
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) poly_mul_acc(G[3 * i + k], Ep[3 * k + j], 1, A[1 + 3 * i + j]);
    }
    for (int r = 0; r < 10; r++)
        for (int k = 0; k < 20; k++) A0[r][k] = A[r][k];
    for (int c = 0; c < 10; c++) {   // Gauss-Jordan, partial pivoting
        int piv = c;
        for (int r = c + 1; r < 10; r++)
            if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (!(fabs(A[piv][c]) > 1e-300)) return 0;
        if (piv != c) for (int k = 0; k < 20; k++) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        const double inv = 1.0 / A[c][c];
        for (int k = 0; k < 20; k++) A[c][k] *= inv;
        for (int r = 0; r < 10; r++) {
            if (r == c) continue;
            const double f = A[r][c];
            if (f == 0.0) continue;
            for (int k = 0; k < 20; k++) A[r][k] -= f * A[c][k];
        }
    }
    double bx[3][4], by[3][4], b1[3][5];   // B(z), ascending coefficients
    for (int r = 0; r < 3; r++) {
        const double* e = &A[4 + 2 * r][10];
        const double* f = &A[5 + 2 * r][10];
        bx[r][0] = e[2]; bx[r][1] = e[1] - f[2]; bx[r][2] = e[0] - f[1]; bx[r][3] = -f[0];
        by[r][0] = e[5]; by[r][1] = e[4] - f[5]; by[r][2] = e[3] - f[4]; by[r][3] = -f[3];
        b1[r][0] = e[9]; b1[r][1] = e[8] - f[9]; b1[r][2] = e[7] - f[8]; b1[r][3] = e[6] - f[7]; b1[r][4] = -f[6];
    }
    double p1[8], p2[8], p3[7], t1[8], t2[8];
    upoly_mul(by[0], 3, b1[1], 4, t1); upoly_mul(b1[0], 4, by[1], 3, t2);
    for (int i = 0; i < 8; i++) p1[i] = t1[i] - t2[i];
    upoly_mul(b1[0], 4, bx[1], 3, t1); upoly_mul(bx[0], 3, b1[1], 4, t2);
    for (int i = 0; i < 8; i++) p2[i] = t1[i] - t2[i];
    upoly_mul(bx[0], 3, by[1], 3, t1); upoly_mul(by[0], 3, bx[1], 3, t2);
    for (int i = 0; i < 7; i++) p3[i] = t1[i] - t2[i];
    double P[11], u1[11], u2[11], u3[11];
    upoly_mul(p1, 7, bx[2], 3, u1); upoly_mul(p2, 7, by[2], 3, u2); upoly_mul(p3, 6, b1[2], 4, u3);
    for (int i = 0; i < 11; i++) P[i] = u1[i] + u2[i] + u3[i];
    double roots[10];
    const int nr = real_roots(P, 10, roots);
    int ne = 0;
    for (int k = 0; k < nr; k++) {
        double z = roots[k];
        const double d3 = upoly_val(p3, 6, z);
        double x = upoly_val(p1, 7, z) / d3, y = upoly_val(p2, 7, z) / d3;
        if (!isfinite(x) || !isfinite(y)) continue;
        {   // one Gauss-Newton step on the ten cubics
            double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
            const double px[4] = {1, x, x * x, x * x * x}, py[4] = {1, y, y * y, y * y * y}, pz[4] = {1, z, z * z, z * z * z};
            for (int q = 0; q < 10; q++) {
                double r = 0, J[3] = {0, 0, 0};
                for (int m = 0; m < 20; m++) {
                    const double c = A0[q][m];
                    if (c == 0.0) continue;
                    const int e = mono_exp(m), a = e & 3, b = (e >> 2) & 3, cc = e >> 4;
                    r += c * px[a] * py[b] * pz[cc];
                    if (a) J[0] += c * a * px[a - 1] * py[b] * pz[cc];
                    if (b) J[1] += c * b * px[a] * py[b - 1] * pz[cc];
                    if (cc) J[2] += c * cc * px[a] * py[b] * pz[cc - 1];
                }
                for (int i = 0; i < 3; i++) { g[i] += J[i] * r; for (int j = 0; j < 3; j++) H[3 * i + j] += J[i] * J[j]; }
            }
            const double dH = det3(H);
            if (fabs(dH) > 1e-300) {
                double dd[3];
                for (int c = 0; c < 3; c++) {
                    double Hi[9];
                    for (int i = 0; i < 9; i++) Hi[i] = H[i];
                    for (int rr = 0; rr < 3; rr++) Hi[3 * rr + c] = g[rr];
                    dd[c] = det3(Hi) / dH;
                }
                const double sc = fabs(x) + fabs(y) + fabs(z) + 1.0;
                if (isfinite(dd[0]) && isfinite(dd[1]) && isfinite(dd[2]) && fabs(dd[0]) + fabs(dd[1]) + fabs(dd[2]) < 1e-6 * sc) { x -= dd[0]; y -= dd[1]; z -= dd[2]; }
            }
        }
        double E[9], n = 0;
        for (int e = 0; e < 9; e++) { E[e] = x * EE[0][e] + y * EE[1][e] + z * EE[2][e] + EE[3][e]; n += E[e] * E[e]; }
        n = sqrt(n);
        if (!(n > 0) || !isfinite(n)) continue;
        for (int e = 0; e < 9; e++) Es[9 * ne + e] = E[e] / n;
        ne++;
    }
    return ne;
}

// symmetric 3x3 eigen-decomposition (cyclic Jacobi), eigenvalues descending, eigenvectors in the columns of V
ALVA_HD inline void eig3(const double* Ain, double* w, double* V) {
    double A[9];
    for (int i = 0; i < 9; i++) { A[i] = Ain[i]; V[i] = (i % 4 == 0); }
    for (int sweep = 0; sweep < 60; sweep++) {
        const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                const double apq = A[3 * p + q];
                if (apq == 0.0) continue;
                const double theta = (A[3 * q + q] - A[3 * p + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 3; k++) { const double akp = A[3 * k + p], akq = A[3 * k + q]; A[3 * k + p] = c * akp - s * akq; A[3 * k + q] = s * akp + c * akq; }
                for (int k = 0; k < 3; k++) { const double apk = A[3 * p + k], aqk = A[3 * q + k]; A[3 * p + k] = c * apk - s * aqk; A[3 * q + k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; k++) { const double vkp = V[3 * k + p], vkq = V[3 * k + q]; V[3 * k + p] = c * vkp - s * vkq; V[3 * k + q] = s * vkp + c * vkq; }
            }
    }
    w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
    for (int i = 0; i < 2; i++)
        for (int j = i + 1; j < 3; j++)
            if (w[j] > w[i]) {
                const double t = w[i]; w[i] = w[j]; w[j] = t;
                for (int k = 0; k < 3; k++) { const double v = V[3 * k + i]; V[3 * k + i] = V[3 * k + j]; V[3 * k + j] = v; }
            }
}

// E = U diag(s) V^T -> Ra = U W V^T, Rb = U W^T V^T (negated when det < 0), ta = s0 U.col(2)
ALVA_HD inline void decompose_essential(const double* E, double* Ra, double* Rb, double* ta) {
    double EtE[9], w[3], V[9], U[9], u[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) EtE[3 * i + j] = E[i] * E[j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    eig3(EtE, w, V);
    const double s0 = sqrt(fmax(w[0], 0.0));
    for (int c = 0; c < 2; c++) {
        for (int r = 0; r < 3; r++) u[c][r] = E[3 * r] * V[c] + E[3 * r + 1] * V[3 + c] + E[3 * r + 2] * V[6 + c];
        if (c == 1) { const double d = dot3(u[1], u[0]); for (int r = 0; r < 3; r++) u[1][r] -= d * u[0][r]; }
        const double n = sqrt(dot3(u[c], u[c]));
        for (int r = 0; r < 3; r++) u[c][r] /= n;
    }
    cross3(u[0], u[1], u[2]);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) U[3 * r + c] = u[c][r];
    const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    double Vt[9], T[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Vt[3 * r + c] = V[3 * c + r];
    mat3_mul(U, W, T); mat3_mul(T, Vt, Ra);
    mat3_mul(U, Wt, T); mat3_mul(T, Vt, Rb);
    if (det3(Ra) < 0) for (int i = 0; i < 9; i++) Ra[i] = -Ra[i];
    if (det3(Rb) < 0) for (int i = 0; i < 9; i++) Rb[i] = -Rb[i];
    for (int r = 0; r < 3; r++) ta[r] = s0 * u[2][r];
}

// computeModelCoefficients (NISTER): 5 sample points -> E candidates -> the (R, t) with the smallest summed error over the 8
// sample points (strict <, candidates in the order (ta,Ra) (ta,Rb) (tb,Ra) (tb,Rb)); model = [R (9), t (3)]
ALVA_HD inline bool relpose_sample_model(const double* bv1, const double* bv2, const int* idx, double* model) {
    double f1[15], f2[15], Es[90];
    for (int i = 0; i < 5; i++)
        for (int k = 0; k < 3; k++) { f1[3 * i + k] = bv1[3 * idx[i] + k]; f2[3 * i + k] = bv2[3 * idx[i] + k]; }
    const int ne = fivept_nister(f1, f2, Es);
    double bestq = 1000000.0;
    bool have = false;
    for (int e = 0; e < ne; e++) {
        double Ra[9], Rb[9], ta[3];
        decompose_essential(Es + 9 * e, Ra, Rb, ta);
        for (int j = 0; j < 4; j++) {
            const double* R = (j & 1) ? Rb : Ra;
            const double sg = j >= 2 ? -1.0 : 1.0;
            const double t[3] = {sg * ta[0], sg * ta[1], sg * ta[2]};
            double q = 0;
            for (int k = 0; k < 8; k++) q += relpose_dist(R, t, bv1 + 3 * idx[k], bv2 + 3 * idx[k]);
            if (q < bestq) {
                bestq = q; have = true;
                for (int i = 0; i < 9; i++) model[i] = R[i];
                for (int i = 0; i < 3; i++) model[9 + i] = t[i];
            }
        }
    }
    return have;
}

// Ransac::computeModel's bookkeeping, fed one drawn sample at a time in draw order
struct RansacState {
    int iterations, best, skipped, max_iter, max_skip, draws, have;
    double k;
    ALVA_HD void init(int max_iterations) {
        iterations = 0; best = -2147483647; skipped = 0; max_iter = max_iterations; max_skip = max_iterations * 10; draws = 0;
        have = 0; k = 1.0;
    }
    ALVA_HD bool running() const { return (double)iterations < k && skipped < max_skip; }
    // returns true when the sample's model became the best one; `stop` is raised by the iterations > max_iterations break
    ALVA_HD bool consume(bool valid, int inliers, int npoints, bool& stop) {
        stop = false;
        draws++;
        if (!valid) { skipped++; return false; }
        bool better = false;
        if (inliers > best) {
            best = inliers; better = true; have = 1;
            const double w = (double)best / (double)npoints;
            double p = 1.0 - pow(w, 8.0);
            p = fmax(DBL_EPSILON, p);
            p = fmin(1.0 - DBL_EPSILON, p);
            k = log(1.0 - 0.99) / log(p);
        }
        ++iterations;
        if (iterations > max_iter) stop = true;
        return better;
    }
};

ALVA_HD inline void cayley2rot(const double* c, double* R) {
    const double s = 1 + c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    R[0] = 1 + c[0] * c[0] - c[1] * c[1] - c[2] * c[2]; R[1] = 2 * (c[0] * c[1] - c[2]); R[2] = 2 * (c[0] * c[2] + c[1]);
    R[3] = 2 * (c[0] * c[1] + c[2]); R[4] = 1 - c[0] * c[0] + c[1] * c[1] - c[2] * c[2]; R[5] = 2 * (c[1] * c[2] - c[0]);
    R[6] = 2 * (c[0] * c[2] - c[1]); R[7] = 2 * (c[1] * c[2] + c[0]); R[8] = 1 - c[0] * c[0] - c[1] * c[1] + c[2] * c[2];
    for (int i = 0; i < 9; i++) R[i] *= 1 / s;
}
ALVA_HD inline void rot2cayley(const double* R, double* c) {
    double A[9], B[9], Bi[9], Cm[9];
    for (int i = 0; i < 9; i++) { A[i] = R[i] - (i % 4 == 0); B[i] = R[i] + (i % 4 == 0); }
    const double d = det3(B);
    Bi[0] = (B[4] * B[8] - B[5] * B[7]) / d; Bi[1] = (B[2] * B[7] - B[1] * B[8]) / d; Bi[2] = (B[1] * B[5] - B[2] * B[4]) / d;
    Bi[3] = (B[5] * B[6] - B[3] * B[8]) / d; Bi[4] = (B[0] * B[8] - B[2] * B[6]) / d; Bi[5] = (B[2] * B[3] - B[0] * B[5]) / d;
    Bi[6] = (B[3] * B[7] - B[4] * B[6]) / d; Bi[7] = (B[1] * B[6] - B[0] * B[7]) / d; Bi[8] = (B[0] * B[4] - B[1] * B[3]) / d;
    mat3_mul(A, Bi, Cm);
    c[0] = -Cm[5]; c[1] = Cm[2]; c[2] = -Cm[1];
}

// one point's residual and Jacobian row of optimize_nonlinear's cost at x = [t, cayley].  Central differences with a 1e-6
// relative step: the reference differentiates forward with sqrt(eps) steps, which drowns in the rounding noise of the
// (1 - cos) residuals (~1e-16 absolute on values of ~1e-7) and leaves its end point noise-limited; the wider central stencil
// keeps both truncation (h^2) and noise (eps / h) three orders below that, so this minimiser lands inside the reference's band
ALVA_HD inline double nl_point(const double* x, const double* f1, const double* f2, double* Jrow) {
    double R[9];
    cayley2rot(x + 3, R);
    const double f = relpose_dist(R, x, f1, f2);
    if (Jrow) {
        for (int c = 0; c < 6; c++) {
            double xp[6];
            for (int i = 0; i < 6; i++) xp[i] = x[i];
            const double h = 1e-6 * fmax(fabs(x[c]), 1e-2);
            xp[c] = x[c] + h;
            cayley2rot(xp + 3, R);
            const double fp = relpose_dist(R, xp, f1, f2);
            xp[c] = x[c] - h;
            cayley2rot(xp + 3, R);
            const double fm = relpose_dist(R, xp, f1, f2);
            Jrow[c] = (fp - fm) / (2 * h);
        }
    }
    return f;
}
// (H + lambda diag(H)) dx = g by Cholesky; H symmetric 6x6 row-major
ALVA_HD inline bool solve6_damped(const double* H, const double* g, double lambda, double* dx) {
    double L[36], y[6];
    for (int i = 0; i < 36; i++) L[i] = H[i];
    for (int a = 0; a < 6; a++) L[7 * a] += lambda * fmax(H[7 * a], 1e-30);
    for (int j = 0; j < 6; j++) {
        double d = L[6 * j + j];
        for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k];
        if (!(d > 0)) return false;
        d = sqrt(d);
        L[6 * j + j] = d;
        for (int i = j + 1; i < 6; i++) { double v = L[6 * i + j]; for (int k = 0; k < j; k++) v -= L[6 * i + k] * L[6 * j + k]; L[6 * i + j] = v / d; }
    }
    for (int i = 0; i < 6; i++) { double v = g[i]; for (int k = 0; k < i; k++) v -= L[6 * i + k] * y[k]; y[i] = v / L[6 * i + i]; }
    for (int i = 5; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < 6; k++) v -= L[6 * k + i] * dx[k]; dx[i] = v / L[6 * i + i]; }
    return true;
}

}  // namespace alva_init
