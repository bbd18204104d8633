// tests/host/init_core_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles alvaar_b200/csrc/init_core.h -- the per-thread arithmetic the CUDA kernels in init.cu execute -- for the HOST, so
// that the CPU suite (no GPU in the build container) can check the very source the device runs against the oracle and the
// reference.  The serial driver below mirrors essential_kernel's control flow (chunk-free: one hypothesis at a time, which is
// what the kernel's replay of RansacState amounts to).  Never linked into libalva_b200.so.
#include "../../alvaar_b200/csrc/init_core.h"
#include "../../alvaar_b200/csrc/lmdif_core.h"
#include <cstdint>
#include <cstring>
#include <vector>
using namespace alva_init;

extern "C" {

int host_fivept_nister(const double* f1, const double* f2, double* Es) { return fivept_nister(f1, f2, Es); }
int host_relpose_sample_model(const double* bv1, const double* bv2, const int* idx, double* model) {
    return relpose_sample_model(bv1, bv2, idx, model) ? 1 : 0;
}
void host_triangulate2(const double* R, const double* t, const double* f1, const double* f2, double* out) { triangulate2(R, t, f1, f2, out); }

// rnd: the sampler table (SampleConsensusProblem::rnd() values), 8 per draw
int host_essential_5pt(const double* bv1, const double* bv2, int n, int max_iter, double threshold, int optimize, const int32_t* rnd,
                       int table_len, double* Rt_out, uint8_t* outlier, double* info) {
    for (int i = 0; i < n; i++) outlier[i] = 0;
    info[0] = info[1] = info[2] = info[3] = 0;
    if (n < 8) return 0;
    std::vector<int> sh(n);
    for (int i = 0; i < n; i++) sh[i] = i;
    RansacState rs;
    rs.init(max_iter);
    double bestm[12], model[12];
    int d = 0;
    while (rs.running() && 8 * (d + 1) <= table_len) {
        int idx[8];
        for (int i = 0; i < 8; i++) { const int j = i + (int)((uint32_t)rnd[8 * d + i] % (uint32_t)(n - i)); std::swap(sh[i], sh[j]); }
        for (int i = 0; i < 8; i++) idx[i] = sh[i];
        d++;
        const bool ok = relpose_sample_model(bv1, bv2, idx, model);
        int cnt = 0;
        if (ok) for (int i = 0; i < n; i++) cnt += relpose_dist(model, model + 9, bv1 + 3 * i, bv2 + 3 * i) < threshold;
        bool stop;
        if (rs.consume(ok, cnt, n, stop)) memcpy(bestm, model, sizeof bestm);
        if (stop) break;
    }
    info[2] = rs.iterations; info[3] = rs.draws;
    if (!rs.have) { for (int i = 0; i < n; i++) outlier[i] = 1; return 0; }
    int m = 0;
    for (int i = 0; i < n; i++) { const bool in = relpose_dist(bestm, bestm + 9, bv1 + 3 * i, bv2 + 3 * i) < threshold; outlier[i] = !in; m += in; }
    info[0] = m >= 10; info[1] = m;
    if (m < 10) return 0;
    if (optimize == 2) {   // the reference's minimiser restated (lmdif_core.h): MINPACK LM on forward differences, ftol = xtol = 10 eps
        double x[6];
        for (int i = 0; i < 3; i++) x[i] = bestm[9 + i];
        rot2cayley(bestm, x + 3);
        std::vector<int> inl;
        for (int i = 0; i < n; i++) if (!outlier[i]) inl.push_back(i);
        const int mm = (int)inl.size();
        std::vector<double> work((size_t)mm * 8);
        auto fun = [&](const double* xx, double* f) {
            double R[9];
            cayley2rot(xx + 3, R);
            for (int i = 0; i < mm; i++) f[i] = relpose_dist(R, xx, bv1 + 3 * inl[i], bv2 + 3 * inl[i]);
        };
        int nfev = 0;
        const int st = alva_lm::lmdif(fun, mm, x, work.data(), work.data() + mm, work.data() + 7 * (size_t)mm, 10 * DBL_EPSILON, 10 * DBL_EPSILON, 1000, &nfev);
        (void)st; (void)nfev;
        for (int i = 0; i < 3; i++) bestm[9 + i] = x[i];
        cayley2rot(x + 3, bestm);
    } else if (optimize) {
        double x[6], xn[6];
        for (int i = 0; i < 3; i++) x[i] = bestm[9 + i];
        rot2cayley(bestm, x + 3);
        auto cost_at = [&](const double* xx) { double s = 0; for (int i = 0; i < n; i++) if (!outlier[i]) { const double f = nl_point(xx, bv1 + 3 * i, bv2 + 3 * i, nullptr); s += f * f; } return s; };
        double cost = cost_at(x), lambda = 1e-3;
        for (int it = 0; it < 200; it++) {
            double H[36] = {0}, g[6] = {0};
            for (int i = 0; i < n; i++) {
                if (outlier[i]) continue;
                double J[6];
                const double f = nl_point(x, bv1 + 3 * i, bv2 + 3 * i, J);
                for (int a = 0; a < 6; a++) { g[a] -= J[a] * f; for (int b = 0; b < 6; b++) H[6 * a + b] += J[a] * J[b]; }
            }
            bool improved = false;
            double step_rel = 0;
            for (int tries = 0; tries < 40 && !improved; tries++) {
                double dx[6];
                if (!solve6_damped(H, g, lambda, dx)) { lambda *= 10; continue; }
                double nx = 0, nd = 0;
                for (int a = 0; a < 6; a++) { xn[a] = x[a] + dx[a]; nx += x[a] * x[a]; nd += dx[a] * dx[a]; }
                const double c2 = cost_at(xn);
                if (c2 < cost) {
                    const double rel = (cost - c2) / cost;
                    step_rel = rel < 1e-15 ? 0.0 : sqrt(nd) / fmax(sqrt(nx), 1e-300);
                    cost = c2; lambda = fmax(lambda * 0.1, 1e-12); improved = true;
                    memcpy(x, xn, sizeof x);
                } else lambda *= 10;
            }
            if (!improved || step_rel < 1e-13) break;
        }
        for (int i = 0; i < 3; i++) bestm[9 + i] = x[i];
        cayley2rot(x + 3, bestm);
    }
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Rt_out[4 * r + c] = bestm[3 * r + c]; Rt_out[4 * r + 3] = bestm[9 + r]; }
    return 1;
}

}  // extern "C"
