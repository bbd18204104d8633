/* oracle/init_oracle.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * CPU restatement of the reference's map initialisation arithmetic:
 *   MultiViewGeometry::compute5ptEssentialMatrix   src/slam/src/multi_view_geometry.cpp:225-318
 *     -> opengv::sac::Ransac<CentralRelativePoseSacProblem>::computeModel     opengv/sac/implementation/Ransac.hpp:44-143
 *     -> CentralRelativePoseSacProblem::computeModelCoefficients (NISTER)     src/sac_problems/relative_pose/CentralRelativePoseSacProblem.cpp:38-255
 *          fivept_nister (null space of the 5x9 epipolar system)              src/relative_pose/methods.cpp:239-268
 *          fivept_nister_main (10 cubic constraints, Gauss-Jordan, det B(z))  src/relative_pose/modules/main.cpp:135-276
 *          math::Sturm (real roots of the 10th-order polynomial)              src/math/Sturm.cpp:141-330
 *          decomposition of E, four (R, t) candidates, 8-point disambiguation CentralRelativePoseSacProblem.cpp:89-252
 *     -> getSelectedDistancesToModel (mid-point triangulation, 1 - cos)       CentralRelativePoseSacProblem.cpp:257-294
 *     -> optimizeModelCoefficients -> relative_pose::optimize_nonlinear        src/relative_pose/methods.cpp:1085-1177
 *   MultiViewGeometry::triangulate -> opengv::triangulation::triangulate2     src/triangulation/methods.cpp:65-88
 *
 * What is restated EXACTLY (decides discrete outcomes): the sampler (mt19937 seeded 12345, partial Fisher-Yates over a
 * persistent shuffle, 8 indices per draw), Ransac's bookkeeping (adaptive k from the best inlier count with probability 0.99,
 * skip accounting, `iterations > max_iterations` stop), the strict comparisons, the float arithmetic of the threshold.
 * What is restated MATHEMATICALLY (same solutions, different floating-point route; agreement ~1e-9, far inside the 1e-4 pose
 * tolerance): the null space (projector + pivoted Gram-Schmidt instead of a Jacobi SVD -- any basis spans the same E's), the
 * constraint expansion (generic trivariate polynomial products instead of OpenGV's generated composeA), root isolation
 * (Sturm counts + bisection/Newton to convergence; the reference stops after 5 Newton steps and one LM polishing step), the
 * E decomposition (eigenvectors of E^T E), and the final refinement (Levenberg-Marquardt with central differences on the same
 * 6-parameter cost, run to convergence; the reference uses Eigen's MINPACK port on forward differences with ftol = xtol =
 * 10 eps, whose end point is noise-limited: it moves by 1e-6 .. 1e-3 under a 1-ulp change of its input, so the refined pose
 * is compared within that band and by cost).
 * PINNED against the reference's own MultiViewGeometry (oracle/_ref/libalva_ref.so): tests/test_oracle_init.py +
 * tests/golden/init.npz.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

void orc_sac_rnd(uint32_t seed, int n, int32_t* out);   /* pose_oracle.c: SampleConsensusProblem::rnd() table */

/* ------------------------------------------------------------------ small helpers */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static void mat3_mul(const double* A, const double* B, double* C)   /* row-major */
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static double det3(const double* M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* opengv::triangulation::triangulate2 (triangulation/methods.cpp:65-88): mid-point of the two rays; R12/t12 = pose of camera 2
 * in camera 1; the point is expressed in camera 1 */
void orc_triangulate2(const double* R12, const double* t12, const double* f1, const double* f2, double* out)
{
    double f2u[3];
    for (int i = 0; i < 3; i++) f2u[i] = R12[3 * i] * f2[0] + R12[3 * i + 1] * f2[1] + R12[3 * i + 2] * f2[2];
    const double b0 = dot3(t12, f1), b1 = dot3(t12, f2u);
    const double a00 = dot3(f1, f1), a10 = dot3(f1, f2u), a01 = -a10, a11 = -dot3(f2u, f2u);
    /* Eigen's 2x2 inverse: adjugate times 1/det */
    const double invdet = 1.0 / (a00 * a11 - a01 * a10);
    const double i00 = a11 * invdet, i01 = -a01 * invdet, i10 = -a10 * invdet, i11 = a00 * invdet;
    const double l0 = i00 * b0 + i01 * b1, l1 = i10 * b0 + i11 * b1;
    for (int i = 0; i < 3; i++) out[i] = (l0 * f1[i] + (t12[i] + l1 * f2u[i])) / 2;
}

/* MultiViewGeometry::triangulate for n pairs; Tlr = [t, q(x,y,z,w)] */
void orc_triangulate(const double* Tlr, const double* bvl, const double* bvr, int n, double* out)
{
    const double x = Tlr[3], y = Tlr[4], z = Tlr[5], w = Tlr[6];
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    for (int i = 0; i < n; i++) orc_triangulate2(R, Tlr, bvl + 3 * i, bvr + 3 * i, out + 3 * i);
}

/* reprojection distance of one correspondence under (R, t) (CentralRelativePoseSacProblem.cpp:271-293) */
static double relpose_dist(const double* R, const double* t, const double* f1, const double* f2)
{
    double p[3], q[3], r2[3];
    orc_triangulate2(R, t, f1, f2, p);
    for (int i = 0; i < 3; i++) q[i] = p[i] - t[i];
    for (int i = 0; i < 3; i++) r2[i] = R[i] * q[0] + R[3 + i] * q[1] + R[6 + i] * q[2];   /* R^T (p - t) */
    const double n1 = sqrt(dot3(p, p)), n2 = sqrt(dot3(r2, r2));
    const double e1 = 1.0 - (f1[0] * (p[0] / n1) + f1[1] * (p[1] / n1) + f1[2] * (p[2] / n1));
    const double e2 = 1.0 - (f2[0] * (r2[0] / n2) + f2[1] * (r2[1] / n2) + f2[2] * (r2[2] / n2));
    return e1 + e2;
}

/* ------------------------------------------------------------------ trivariate polynomials, total degree <= 3
 * 20 monomials in Nister's order: the first ten are eliminated, the last ten are [x z^2, x z, x, y z^2, y z, y, z^3, z^2, z, 1] */
static const int8_t MONO[20][3] = {{3, 0, 0}, {0, 3, 0}, {2, 1, 0}, {1, 2, 0}, {2, 0, 1}, {2, 0, 0}, {0, 2, 1}, {0, 2, 0}, {1, 1, 1}, {1, 1, 0},
                                   {1, 0, 2}, {1, 0, 1}, {1, 0, 0}, {0, 1, 2}, {0, 1, 1}, {0, 1, 0}, {0, 0, 3}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};
static int mono_index(int a, int b, int c)
{
    for (int i = 0; i < 20; i++)
        if (MONO[i][0] == a && MONO[i][1] == b && MONO[i][2] == c) return i;
    return -1;
}
static void poly_mul_acc(const double* p, const double* q, double s, double* out)   /* out += s * p * q */
{
    for (int i = 0; i < 20; i++) {
        if (p[i] == 0.0) continue;
        for (int j = 0; j < 20; j++) {
            if (q[j] == 0.0) continue;
            const int a = MONO[i][0] + MONO[j][0], b = MONO[i][1] + MONO[j][1], c = MONO[i][2] + MONO[j][2];
            if (a + b + c > 3) continue;
            out[mono_index(a, b, c)] += s * p[i] * q[j];
        }
    }
}

/* univariate helpers (ascending coefficients) */
static void upoly_mul(const double* a, int da, const double* b, int db, double* o)
{
    for (int i = 0; i <= da + db; i++) o[i] = 0;
    for (int i = 0; i <= da; i++)
        for (int j = 0; j <= db; j++) o[i + j] += a[i] * b[j];
}
static double upoly_val(const double* a, int d, double z)
{
    double v = a[d];
    for (int i = d - 1; i >= 0; i--) v = v * z + a[i];
    return v;
}

/* Sturm chain of p (degree d <= 10): chain[0] = p, chain[1] = p', chain[k] = -rem(chain[k-2], chain[k-1]) (each scaled to
 * unit max-magnitude, a positive factor that leaves every sign unchanged).  Returns the chain length. */
typedef struct { double c[12][11]; int deg[12]; int n; } sturm_t;
static void sturm_build(const double* p, int d, sturm_t* S)
{
    while (d > 0 && p[d] == 0.0) d--;
    memset(S, 0, sizeof *S);
    for (int i = 0; i <= d; i++) S->c[0][i] = p[i];
    S->deg[0] = d;
    for (int i = 1; i <= d; i++) S->c[1][i - 1] = i * p[i];
    S->deg[1] = d - 1;
    S->n = d >= 1 ? 2 : 1;
    for (int k = 2; k <= d && S->n == k; k++) {
        double r[11];
        int dr = S->deg[k - 2];
        const int dq = S->deg[k - 1];
        memcpy(r, S->c[k - 2], sizeof r);
        const double* q = S->c[k - 1];
        while (dr >= dq) {
            const double f = r[dr] / q[dq];
            for (int i = 0; i <= dq; i++) r[dr - dq + i] -= f * q[i];
            r[dr] = 0;
            dr--;
        }
        double mx = 0;
        for (int i = 0; i <= dr; i++) mx = fmax(mx, fabs(r[i]));
        if (dr < 0 || mx == 0.0 || mx < 1e-300) break;
        /* drop numerically vanished leading terms */
        while (dr > 0 && fabs(r[dr]) < 1e-14 * mx) dr--;
        for (int i = 0; i <= dr; i++) S->c[k][i] = -r[i] / mx;
        S->deg[k] = dr;
        S->n = k + 1;
        if (dr == 0) break;
    }
}
static int sturm_changes(const sturm_t* S, double z)
{
    int changes = 0, last = 0;
    for (int k = 0; k < S->n; k++) {
        const double v = upoly_val(S->c[k], S->deg[k], z);
        const int s = v > 0 ? 1 : (v < 0 ? -1 : 0);
        if (s != 0) { if (last != 0 && s != last) changes++; last = s; }
    }
    return changes;
}
/* all real roots of p (degree d), ascending */
static int real_roots(const double* p, int d, double* roots)
{
    while (d > 0 && p[d] == 0.0) d--;
    if (d < 1) return 0;
    sturm_t S;
    sturm_build(p, d, &S);
    double bound = 0;   /* Cauchy: 1 + max |a_i / a_d| */
    for (int i = 0; i < d; i++) bound = fmax(bound, fabs(p[i] / p[d]));
    bound += 1.0;
    struct { double lo, hi; int clo, chi; } st[64];
    int sp = 0, nr = 0;
    st[sp].lo = -bound; st[sp].hi = bound; st[sp].clo = sturm_changes(&S, -bound); st[sp].chi = sturm_changes(&S, bound); sp++;
    double dp[11];
    for (int i = 1; i <= d; i++) dp[i - 1] = i * p[i];
    while (sp > 0) {
        sp--;
        double lo = st[sp].lo, hi = st[sp].hi;
        const int clo = st[sp].clo, chi = st[sp].chi, n = clo - chi;
        if (n <= 0) continue;
        const double mid = 0.5 * (lo + hi);
        if (n > 1 && mid > lo && mid < hi && (hi - lo) > 1e-13 * fmax(1.0, fabs(mid))) {
            const int cm = sturm_changes(&S, mid);
            if (sp + 2 > 64) continue;
            /* push the upper half first so that roots come out ascending */
            st[sp].lo = mid; st[sp].hi = hi; st[sp].clo = cm; st[sp].chi = chi; sp++;
            st[sp].lo = lo; st[sp].hi = mid; st[sp].clo = clo; st[sp].chi = cm; sp++;
            continue;
        }
        if (n > 1) { for (int k = 0; k < n && nr < 10; k++) roots[nr++] = mid; continue; }   /* (near-)multiple root */
        /* one root in (lo, hi]: safeguarded Newton on p (sign change across the bracket when the root is simple) */
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
    /* ascending order (the stack discipline already yields it; keep it explicit) */
    for (int i = 1; i < nr; i++) { const double v = roots[i]; int j = i - 1; while (j >= 0 && roots[j] > v) { roots[j + 1] = roots[j]; j--; } roots[j + 1] = v; }
    return nr;
}

/* fivept_nister: essential matrices E (row-major, unit Frobenius norm) with f1_i^T E f2_i = 0 for the five pairs.
 * f1 = bearing vectors in camera 1 (the keyframe), f2 in camera 2.  Returns the number of real solutions (<= 10). */
int orc_fivept_nister(const double* f1, const double* f2, double* Es)
{
    /* rows of the epipolar system (relative_pose/methods.cpp:253-259: f = bv2, f' = bv1): q = [f'_r * f_c] */
    double Q[5][9];
    for (int i = 0; i < 5; i++)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) Q[i][3 * r + c] = f1[3 * i + r] * f2[3 * i + c];
    /* orthonormal basis of the row space (modified Gram-Schmidt, twice) */
    for (int i = 0; i < 5; i++) {
        for (int pass = 0; pass < 2; pass++)
            for (int j = 0; j < i; j++) {
                double d = 0;
                for (int k = 0; k < 9; k++) d += Q[i][k] * Q[j][k];
                for (int k = 0; k < 9; k++) Q[i][k] -= d * Q[j][k];
            }
        double n = 0;
        for (int k = 0; k < 9; k++) n += Q[i][k] * Q[i][k];
        n = sqrt(n);
        if (!(n > 1e-12)) return 0;   /* degenerate sample */
        for (int k = 0; k < 9; k++) Q[i][k] /= n;
    }
    /* null space: project the unit vectors, pick the largest residual, orthonormalise, four times */
    double EE[4][9];
    for (int b = 0; b < 4; b++) {
        double best = -1, bv[9];
        for (int e = 0; e < 9; e++) {
            double v[9] = {0};
            v[e] = 1;
            for (int pass = 0; pass < 2; pass++) {
                for (int j = 0; j < 5; j++) { double d = 0; for (int k = 0; k < 9; k++) d += v[k] * Q[j][k]; for (int k = 0; k < 9; k++) v[k] -= d * Q[j][k]; }
                for (int j = 0; j < b; j++) { double d = 0; for (int k = 0; k < 9; k++) d += v[k] * EE[j][k]; for (int k = 0; k < 9; k++) v[k] -= d * EE[j][k]; }
            }
            double n = 0;
            for (int k = 0; k < 9; k++) n += v[k] * v[k];
            if (n > best) { best = n; memcpy(bv, v, sizeof bv); }
        }
        best = sqrt(best);
        for (int k = 0; k < 9; k++) EE[b][k] = bv[k] / best;
    }
    /* E(x, y, z) = x EE0 + y EE1 + z EE2 + EE3 as entry polynomials */
    static const int LIN[4] = {12, 15, 18, 19};
    double Ep[9][20];
    memset(Ep, 0, sizeof Ep);
    for (int e = 0; e < 9; e++)
        for (int b = 0; b < 4; b++) Ep[e][LIN[b]] = EE[b][e];
    double A[10][20];
    memset(A, 0, sizeof A);
    /* det E = 0 */
    {
        double m[3][20];
        memset(m, 0, sizeof m);
        poly_mul_acc(Ep[4], Ep[8], 1, m[0]); poly_mul_acc(Ep[5], Ep[7], -1, m[0]);
        poly_mul_acc(Ep[3], Ep[8], 1, m[1]); poly_mul_acc(Ep[5], Ep[6], -1, m[1]);
        poly_mul_acc(Ep[3], Ep[7], 1, m[2]); poly_mul_acc(Ep[4], Ep[6], -1, m[2]);
        poly_mul_acc(Ep[0], m[0], 1, A[0]); poly_mul_acc(Ep[1], m[1], -1, A[0]); poly_mul_acc(Ep[2], m[2], 1, A[0]);
    }
    /* (E E^T - 1/2 trace(E E^T) I) E = 0 */
    {
        double G[9][20], tr[20];
        memset(G, 0, sizeof G);
        memset(tr, 0, sizeof tr);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) poly_mul_acc(Ep[3 * i + k], Ep[3 * j + k], 1, G[3 * i + j]);
        for (int k = 0; k < 20; k++) tr[k] = G[0][k] + G[4][k] + G[8][k];
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 20; k++) G[4 * i][k] -= 0.5 * tr[k];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) poly_mul_acc(G[3 * i + k], Ep[3 * k + j], 1, A[1 + 3 * i + j]);
    }
    double A0[10][20];
    memcpy(A0, A, sizeof A0);
    /* Gauss-Jordan on the first ten columns (partial pivoting) */
    for (int c = 0; c < 10; c++) {
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
    /* B(z): rows <e> - z <f>, <g> - z <h>, <i> - z <j>; columns x, y, 1 (ascending coefficients in z) */
    double bx[3][4], by[3][4], b1[3][5];
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
        /* one Gauss-Newton step on the ten original cubics (the reference's pollishCoefficients is a single LM step) */
        {
            double r[10], J[10][3], H[9] = {0}, g[3] = {0};
            for (int q = 0; q < 10; q++) {
                r[q] = 0; J[q][0] = J[q][1] = J[q][2] = 0;
                for (int m = 0; m < 20; m++) {
                    const double c = A0[q][m];
                    if (c == 0.0) continue;
                    const int a = MONO[m][0], b = MONO[m][1], cc = MONO[m][2];
                    const double xa = pow(x, a), yb = pow(y, b), zc = pow(z, cc);
                    r[q] += c * xa * yb * zc;
                    if (a) J[q][0] += c * a * pow(x, a - 1) * yb * zc;
                    if (b) J[q][1] += c * b * xa * pow(y, b - 1) * zc;
                    if (cc) J[q][2] += c * cc * xa * yb * pow(z, cc - 1);
                }
                for (int i = 0; i < 3; i++) { g[i] += J[q][i] * r[q]; for (int j = 0; j < 3; j++) H[3 * i + j] += J[q][i] * J[q][j]; }
            }
            const double dH = det3(H);
            if (fabs(dH) > 1e-300) {
                double Hi[9], dd[3];
                for (int c = 0; c < 3; c++) { memcpy(Hi, H, sizeof Hi); for (int rr = 0; rr < 3; rr++) Hi[3 * rr + c] = g[rr]; dd[c] = det3(Hi) / dH; }
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

/* symmetric 3x3 eigen-decomposition (cyclic Jacobi): A = V diag(w) V^T, eigenvalues descending, V columns */
static void eig3(const double* Ain, double* w, double* V)
{
    double A[9];
    memcpy(A, Ain, sizeof A);
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0);
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

/* the four (R, t) candidates of an essential matrix (CentralRelativePoseSacProblem.cpp:89-141): E = U diag(s) V^T,
 * Ra = U W V^T, Rb = U W^T V^T (negated when det < 0), ta = s0 U.col(2), tb = -ta; order (ta,Ra) (ta,Rb) (tb,Ra) (tb,Rb) */
static void decompose_essential(const double* E, double Rs[2][9], double* ta)
{
    double EtE[9], w[3], V[9], U[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) EtE[3 * i + j] = E[i] * E[j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    eig3(EtE, w, V);
    const double s0 = sqrt(fmax(w[0], 0.0));
    double u[3][3];
    for (int c = 0; c < 2; c++) {
        for (int r = 0; r < 3; r++) u[c][r] = E[3 * r] * V[c] + E[3 * r + 1] * V[3 + c] + E[3 * r + 2] * V[6 + c];
        if (c == 1) { const double d = dot3(u[1], u[0]); for (int r = 0; r < 3; r++) u[1][r] -= d * u[0][r]; }
        const double n = sqrt(dot3(u[c], u[c]));
        for (int r = 0; r < 3; r++) u[c][r] /= n;
    }
    cross3(u[0], u[1], u[2]);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) U[3 * r + c] = u[c][r];
    static const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    double Vt[9], T[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Vt[3 * r + c] = V[3 * c + r];
    mat3_mul(U, W, T); mat3_mul(T, Vt, Rs[0]);
    mat3_mul(U, Wt, T); mat3_mul(T, Vt, Rs[1]);
    for (int k = 0; k < 2; k++)
        if (det3(Rs[k]) < 0) for (int i = 0; i < 9; i++) Rs[k][i] = -Rs[k][i];
    for (int r = 0; r < 3; r++) ta[r] = s0 * u[2][r];
}

/* CentralRelativePoseSacProblem::computeModelCoefficients (NISTER): idx = 8 sample indices; model = [R (9, row-major), t (3)] */
int orc_relpose_sample_model(const double* bv1, const double* bv2, const int* idx, double* model)
{
    double f1[15], f2[15], Es[90];
    for (int i = 0; i < 5; i++) { memcpy(f1 + 3 * i, bv1 + 3 * idx[i], 24); memcpy(f2 + 3 * i, bv2 + 3 * idx[i], 24); }
    const int ne = orc_fivept_nister(f1, f2, Es);
    double bestq = 1000000.0;
    int have = 0;
    for (int e = 0; e < ne; e++) {
        double Rs[2][9], ta[3];
        decompose_essential(Es + 9 * e, Rs, ta);
        for (int j = 0; j < 4; j++) {
            const double* R = Rs[j & 1];
            double t[3] = {ta[0], ta[1], ta[2]};
            if (j >= 2) { t[0] = -t[0]; t[1] = -t[1]; t[2] = -t[2]; }
            double q = 0;
            for (int k = 0; k < 8; k++) q += relpose_dist(R, t, bv1 + 3 * idx[k], bv2 + 3 * idx[k]);
            if (q < bestq) { bestq = q; memcpy(model, R, 72); memcpy(model + 9, t, 24); have = 1; }
        }
    }
    return have;
}

/* cayley <-> rotation (src/math/cayley.cpp) */
static void cayley2rot(const double* c, double* R)
{
    const double s = 1 + c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    R[0] = 1 + c[0] * c[0] - c[1] * c[1] - c[2] * c[2]; R[1] = 2 * (c[0] * c[1] - c[2]); R[2] = 2 * (c[0] * c[2] + c[1]);
    R[3] = 2 * (c[0] * c[1] + c[2]); R[4] = 1 - c[0] * c[0] + c[1] * c[1] - c[2] * c[2]; R[5] = 2 * (c[1] * c[2] - c[0]);
    R[6] = 2 * (c[0] * c[2] - c[1]); R[7] = 2 * (c[1] * c[2] + c[0]); R[8] = 1 - c[0] * c[0] - c[1] * c[1] + c[2] * c[2];
    for (int i = 0; i < 9; i++) R[i] *= 1 / s;
}
static void rot2cayley(const double* R, double* c)
{
    /* C = (R - I)(R + I)^-1; cayley = (-C12, C02, -C01) */
    double A[9], B[9], Bi[9], Cm[9];
    for (int i = 0; i < 9; i++) { A[i] = R[i] - (i % 4 == 0); B[i] = R[i] + (i % 4 == 0); }
    const double d = det3(B);
    Bi[0] = (B[4] * B[8] - B[5] * B[7]) / d; Bi[1] = (B[2] * B[7] - B[1] * B[8]) / d; Bi[2] = (B[1] * B[5] - B[2] * B[4]) / d;
    Bi[3] = (B[5] * B[6] - B[3] * B[8]) / d; Bi[4] = (B[0] * B[8] - B[2] * B[6]) / d; Bi[5] = (B[2] * B[3] - B[0] * B[5]) / d;
    Bi[6] = (B[3] * B[7] - B[4] * B[6]) / d; Bi[7] = (B[1] * B[6] - B[0] * B[7]) / d; Bi[8] = (B[0] * B[4] - B[1] * B[3]) / d;
    mat3_mul(A, Bi, Cm);
    c[0] = -Cm[5]; c[1] = Cm[2]; c[2] = -Cm[1];
}

/* residuals of relative_pose::OptimizeNonlinearFunctor1 for x = [t, cayley] over the inliers; returns the sum of squares */
static double nl_residuals(const double* x, const double* bv1, const double* bv2, const int* inl, int m, double* f)
{
    double R[9], s = 0;
    cayley2rot(x + 3, R);
    for (int i = 0; i < m; i++) { f[i] = relpose_dist(R, x, bv1 + 3 * inl[i], bv2 + 3 * inl[i]); s += f[i] * f[i]; }
    return s;
}
static int solve6(double* H, double* g, double* dx)   /* Cholesky, in place; returns 0 when not positive definite */
{
    for (int j = 0; j < 6; j++) {
        double d = H[6 * j + j];
        for (int k = 0; k < j; k++) d -= H[6 * j + k] * H[6 * j + k];
        if (!(d > 0)) return 0;
        d = sqrt(d);
        H[6 * j + j] = d;
        for (int i = j + 1; i < 6; i++) { double v = H[6 * i + j]; for (int k = 0; k < j; k++) v -= H[6 * i + k] * H[6 * j + k]; H[6 * i + j] = v / d; }
    }
    double y[6];
    for (int i = 0; i < 6; i++) { double v = g[i]; for (int k = 0; k < i; k++) v -= H[6 * i + k] * y[k]; y[i] = v / H[6 * i + i]; }
    for (int i = 5; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < 6; k++) v -= H[6 * k + i] * dx[k]; dx[i] = v / H[6 * i + i]; }
    return 1;
}
/* relative_pose::optimize_nonlinear: minimise sum_i (e1_i + e2_i)^2 over [t, cayley] from the RANSAC model */
static void optimize_nonlinear(const double* bv1, const double* bv2, const int* inl, int m, double* model)
{
    double x[6], *f = (double*)malloc(sizeof(double) * m * 8), *J = f + m, *ft = f + 7 * m;
    memcpy(x, model + 9, 24);
    rot2cayley(model, x + 3);
    double cost = nl_residuals(x, bv1, bv2, inl, m, f), lambda = 1e-3;
    for (int it = 0; it < 200; it++) {
        /* central differences, 1e-6 relative step (the reference: forward, sqrt(eps) -- noise-limited, see the header) */
        for (int c = 0; c < 6; c++) {
            const double h = 1e-6 * fmax(fabs(x[c]), 1e-2);
            const double keep = x[c];
            x[c] = keep + h;
            nl_residuals(x, bv1, bv2, inl, m, ft);
            for (int i = 0; i < m; i++) J[c * m + i] = ft[i];
            x[c] = keep - h;
            nl_residuals(x, bv1, bv2, inl, m, ft);
            x[c] = keep;
            for (int i = 0; i < m; i++) J[c * m + i] = (J[c * m + i] - ft[i]) / (2 * h);
        }
        double H[36], g[6];
        for (int a = 0; a < 6; a++) {
            g[a] = 0;
            for (int i = 0; i < m; i++) g[a] -= J[a * m + i] * f[i];
            for (int b = 0; b <= a; b++) { double s = 0; for (int i = 0; i < m; i++) s += J[a * m + i] * J[b * m + i]; H[6 * a + b] = H[6 * b + a] = s; }
        }
        int improved = 0;
        double step_rel = 0;
        for (int tries = 0; tries < 40 && !improved; tries++) {
            double Hd[36], gd[6], dx[6], xn[6];
            memcpy(Hd, H, sizeof Hd); memcpy(gd, g, sizeof gd);
            for (int a = 0; a < 6; a++) Hd[7 * a] += lambda * fmax(H[7 * a], 1e-30);
            if (!solve6(Hd, gd, dx)) { lambda *= 10; continue; }
            double nx = 0, nd = 0;
            for (int a = 0; a < 6; a++) { xn[a] = x[a] + dx[a]; nx += x[a] * x[a]; nd += dx[a] * dx[a]; }
            const double c2 = nl_residuals(xn, bv1, bv2, inl, m, ft);
            if (c2 < cost) {
                step_rel = sqrt(nd) / fmax(sqrt(nx), 1e-300);
                memcpy(x, xn, sizeof x); memcpy(f, ft, sizeof(double) * m);
                const double rel = (cost - c2) / cost;
                cost = c2; lambda = fmax(lambda * 0.1, 1e-12); improved = 1;
                if (rel < 1e-15) step_rel = 0;
            } else lambda *= 10;
        }
        if (!improved || step_rel < 1e-13) break;
    }
    memcpy(model + 9, x, 24);
    cayley2rot(x + 3, model);
    free(f);
}

/* MultiViewGeometry::compute5ptEssentialMatrix.  bv1 (keyframe) / bv2 (current frame) [n][3] unit bearing vectors.
 * Rt_out: 3x4 row-major [Rwc | twc] (not normalised, as the reference returns it); outlier [n].
 * info[0] = #inliers, [1] = #RANSAC iterations, [2] = #draws, [3] = best inlier count during RANSAC.  Returns 1 on success. */
int orc_essential_5pt(const double* bv1, const double* bv2, int n, int max_iter, float err_px, int optimize, float fx, float fy,
                      uint32_t seed, double* Rt_out, uint8_t* outlier, double* info)
{
    if (info) info[0] = info[1] = info[2] = info[3] = 0;
    for (int i = 0; i < n; i++) outlier[i] = 0;
    if (n < 8) return 0;
    float focal = fx + fy;
    focal = (float)(focal / 2.);
    const double threshold = 2.0 * (1.0 - cosf(atanf(err_px / focal)));
    const int max_draws = 11 * max_iter + 16;
    int32_t* rnd = (int32_t*)malloc(sizeof(int32_t) * 8 * (size_t)max_draws);
    orc_sac_rnd(seed, 8 * max_draws, rnd);
    int* sh = (int*)malloc(sizeof(int) * n);
    int* inl = (int*)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) sh[i] = i;
    int iterations = 0, best = -INT32_MAX, draws = 0, have = 0;
    unsigned skipped = 0;
    const unsigned max_skip = (unsigned)max_iter * 10;
    double k = 1.0, bestm[12], model[12];
    while (iterations < k && skipped < max_skip) {
        int idx[8];
        for (int i = 0; i < 8; i++) {
            const int j = i + (int)((uint32_t)rnd[8 * draws + i] % (uint32_t)(n - i));
            const int t = sh[i]; sh[i] = sh[j]; sh[j] = t;
        }
        for (int i = 0; i < 8; i++) idx[i] = sh[i];
        draws++;
        if (!orc_relpose_sample_model(bv1, bv2, idx, model)) { skipped++; continue; }
        int cnt = 0;
        for (int i = 0; i < n; i++) cnt += relpose_dist(model, model + 9, bv1 + 3 * i, bv2 + 3 * i) < threshold;
        if (cnt > best) {
            best = cnt; memcpy(bestm, model, sizeof bestm); have = 1;
            const double w = (double)best / (double)n;
            double p = 1.0 - pow(w, 8.0);
            p = fmax(DBL_EPSILON, p);
            p = fmin(1.0 - DBL_EPSILON, p);
            k = log(1.0 - 0.99) / log(p);
        }
        ++iterations;
        if (iterations > max_iter) break;
    }
    int ok = 0, m = 0;
    if (have) {
        for (int i = 0; i < n; i++)
            if (relpose_dist(bestm, bestm + 9, bv1 + 3 * i, bv2 + 3 * i) < threshold) inl[m++] = i;
        if (m >= 10) {
            ok = 1;
            if (optimize) optimize_nonlinear(bv1, bv2, inl, m, bestm);
            for (int i = 0; i < n; i++) outlier[i] = 1;
            for (int i = 0; i < m; i++) outlier[inl[i]] = 0;
            for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Rt_out[4 * i + j] = bestm[3 * i + j]; Rt_out[4 * i + 3] = bestm[9 + i]; }
        }
    }
    if (info) { info[0] = m; info[1] = iterations; info[2] = draws; info[3] = best; }
    free(rnd); free(sh); free(inl);
    return ok;
}
