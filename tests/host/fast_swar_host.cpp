// tests/host/fast_swar_host.cpp -- TEST INFRASTRUCTURE ONLY.  Host emulation of one warp of the front end's FAST-9 candidate
// pre-test: alvaar_b200/csrc/fast_swar.h (the antipodal-sharing variant, single source for device and host) against (a) the
// baseline formulation -- all 16 ring flags computed directly, what frontend.cu's fast_candidates8 does -- and (b) the scalar
// definition of the pre-test (9 contiguous ring pixels differ from the centre by more than t, either sign mixed), on a gray
// tile laid out as the kernel's shared tile (pitch 144 bytes, 14 rows per warp).
#include "../../alvaar_b200/csrc/fast_swar.h"
#include <cstdint>
#include <cstring>
#include <vector>
using namespace fast_swar;

namespace {
constexpr int GP = 144, GPW = GP / 4;

template <bool HI>
void baseline_lane(const uint32_t* g0, uint32_t Kc, const uint32_t* mul, uint32_t* acc) {   // every element directly
    uint32_t Lw[7], Mw[7], Rw[7];
    for (int k = 0; k < 16; k++) acc[k] = 0;
    for (int r = 0; r < 6; r++) { Lw[r] = g0[r * GPW - 1]; Mw[r] = g0[r * GPW]; Rw[r] = g0[r * GPW + 1]; }
    for (int i = 0; i < 8; i++) {
        const int r = i + 6;
        Lw[r % 7] = g0[r * GPW - 1]; Mw[r % 7] = g0[r * GPW]; Rw[r % 7] = g0[r * GPW + 1];
        const uint32_t c = Mw[(i + 3) % 7];
        for (int k = 0; k < 16; k++) {
            const int rr = ((i + 3) % 7 + dy_of(k) + 7) % 7;
            uint32_t ring;
            switch (dx_of(k)) {
                case 0: ring = Mw[rr]; break;
                case 1: ring = FSW_PRMT(Mw[rr], Rw[rr], 0x4321); break;
                case 2: ring = FSW_PRMT(Mw[rr], Rw[rr], 0x5432); break;
                case 3: ring = FSW_PRMT(Mw[rr], Rw[rr], 0x6543); break;
                case -1: ring = FSW_PRMT(Lw[rr], Mw[rr], 0x6543); break;
                case -2: ring = FSW_PRMT(Lw[rr], Mw[rr], 0x5432); break;
                default: ring = FSW_PRMT(Lw[rr], Mw[rr], 0x4321); break;
            }
            const uint32_t U = absdiff_gt<HI>(ring, c, Kc);
            acc[k] = i < 7 ? fsw_madhi(U, mul[i], acc[k]) : acc[k] + U;
        }
    }
}

template <int M>
void assemble_all(uint32_t acc[32][16], int lane) {
    constexpr int o = source_of(M);
    const uint32_t left = acc[lane > 0 ? lane - 1 : lane][o], right = acc[lane < 31 ? lane + 1 : lane][o];   // __shfl_up / __shfl_down by 1
    acc[lane][M] = assemble<M>(acc[lane][o], left, right, acc[lane][M]);
}
}  // namespace

extern "C" {

// gray: [14][GP] bytes (the 14 rows a warp needs; image column of byte b is arbitrary), thr: FAST threshold.
// cand_anti / cand_base [32]: candidate masks per lane (bit 8 j + i: pixel j of the lane's word, row i; centre rows 3..10);
// lane l owns the word at byte offset 4 * (first_word + l).
void fsw_warp(const uint8_t* gray, int first_word, int thr, uint32_t* cand_anti, uint32_t* cand_base) {
    const bool hi = thr >= 128;
    const uint32_t Kc = (uint32_t)(hi ? 255 - thr : 127 - thr) * 0x01010101u;
    uint32_t mul[7];
    for (int i = 0; i < 7; i++) mul[i] = 1u << (25 + i);
    const uint32_t* G = reinterpret_cast<const uint32_t*>(gray);
    static uint32_t acc[32][16];
    for (int lane = 0; lane < 32; lane++) {
        uint32_t b[16];
        if (hi) baseline_lane<true>(G + first_word + lane, Kc, mul, b); else baseline_lane<false>(G + first_word + lane, Kc, mul, b);
        cand_base[lane] = contiguous9(b);
        if (hi) phase1<true, GPW>(G + first_word + lane, Kc, mul, acc[lane]); else phase1<false, GPW>(G + first_word + lane, Kc, mul, acc[lane]);
    }
    // phase 2 reads the neighbours' DIRECT words only, which phase 2 never writes: the order of lanes / elements is free
    for (int lane = 0; lane < 32; lane++) {
        assemble_all<5>(acc, lane); assemble_all<6>(acc, lane); assemble_all<7>(acc, lane); assemble_all<8>(acc, lane);
        assemble_all<9>(acc, lane); assemble_all<10>(acc, lane); assemble_all<11>(acc, lane); assemble_all<12>(acc, lane);
    }
    for (int lane = 0; lane < 32; lane++) cand_anti[lane] = contiguous9(acc[lane]);
}

// scalar definition for pixel (x = byte column, y = row 3..10): 9 contiguous ring pixels with |ring - c| > t
int fsw_scalar(const uint8_t* gray, int x, int y, int thr) {
    int f[16];
    const int c = gray[y * GP + x];
    for (int k = 0; k < 16; k++) { const int v = gray[(y + dy_of(k)) * GP + x + dx_of(k)]; f[k] = (v > c ? v - c : c - v) > thr; }
    for (int s = 0; s < 16; s++) { int all = 1; for (int q = 0; q < 9; q++) all &= f[(s + q) & 15]; if (all) return 1; }
    return 0;
}
}
