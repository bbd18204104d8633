// hamming.cu -- brute-force Hamming 2-NN over 256-bit descriptors, sm_100a.
//
// Reference behaviour (exact): cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2)
//   opencv features2d/src/matchers.cpp:757 -> core/src/batch_distance.cpp:103-123 (batchDistHamming),
//   k-NN insertion :235-248 (strict '<': ties keep the lowest train index), popcount core/src/stat.simd.hpp:81-128.
// The same primitive serves MapPoint::computeMinDescDist (src/slam/src/map_point.cpp:204-222).
//
// Integer-ALU bound (xor + popc), not HBM: 8 queries live in registers per warp, every lane streams its own train
// descriptors with two 128-bit loads, candidates are packed as (dist << 22 | index) so the lexicographic
// (distance, index) order OpenCV's insertion produces is a plain unsigned min; the 32 per-lane top-2 lists are merged
// with warp shuffles, chunk partials with a tiny second kernel.
#include "alva_common.cuh"
#include "../../include/alva_b200.h"

namespace {

constexpr int QPW = 8;          // queries per warp
constexpr int WPC = 8;          // warps per CTA
constexpr int CHUNK = 1024;     // train descriptors per partial
constexpr uint32_t NONE = 0xffffffffu;

__device__ __forceinline__ void top2_insert(uint32_t& k0, uint32_t& k1, uint32_t key) {
    if (key < k1) {
        if (key < k0) { k1 = k0; k0 = key; }
        else k1 = key;
    }
}

// counts != nullptr: the queries are [nbatch][qcap] slots of which only the first counts[b] are live; warps that hold
// no live query leave immediately (their partials are never read).
__global__ void __launch_bounds__(WPC * 32) knn2_partial_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t,
                                                                int nt, uint2* __restrict__ partial, int nchunks,
                                                                const int32_t* __restrict__ counts, int qcap) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = (blockIdx.x * WPC + warp) * QPW;
    if (q0 >= nq) return;
    if (counts) { const int b = q0 / qcap; if (q0 - b * qcap >= counts[b]) return; }
    const int chunk = blockIdx.y;
    const int tb = chunk * CHUNK, te = min(nt, tb + CHUNK);
    uint4 qa[QPW], qb[QPW];
#pragma unroll
    for (int i = 0; i < QPW; i++) {
        const int qi = min(q0 + i, nq - 1);
        qa[i] = __ldg(q + 2 * qi);
        qb[i] = __ldg(q + 2 * qi + 1);
    }
    uint32_t k0[QPW], k1[QPW];
#pragma unroll
    for (int i = 0; i < QPW; i++) k0[i] = k1[i] = NONE;
    for (int j = tb + lane; j < te; j += 32) {
        const uint4 a = __ldg(t + 2 * j), b = __ldg(t + 2 * j + 1);
#pragma unroll
        for (int i = 0; i < QPW; i++) {
            // 256-bit popcount with 4 POPC instead of 8: the POPC pipe (16 lanes/clk/SM) is the bound, the LOP3 pipe has
            // slack, so three carry-save adders (sum = a^b^c, carry = maj(a,b,c): 2 LOP3 each) first compress the eight
            // xor words into ones / twos / fours planes:  d = popc(ones) + popc(w7) + 2 popc(twos) + 4 popc(fours).
            const uint32_t x0 = qa[i].x ^ a.x, x1 = qa[i].y ^ a.y, x2 = qa[i].z ^ a.z, x3 = qa[i].w ^ a.w;
            const uint32_t x4 = qb[i].x ^ b.x, x5 = qb[i].y ^ b.y, x6 = qb[i].z ^ b.z, x7 = qb[i].w ^ b.w;
            const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 | x1));
            const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 | x4));
            const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 | s2));
            const uint32_t s4 = c1 ^ c2 ^ c3, c4 = (c1 & c2) | (c3 & (c1 | c2));
            const int d = __popc(s3) + __popc(x7) + 2 * __popc(s4) + 4 * __popc(c4);
            top2_insert(k0[i], k1[i], ((uint32_t)d << 22) | (uint32_t)j);
        }
    }
    // warp-shuffle merge of the 32 per-lane (best, second) pairs
#pragma unroll
    for (int i = 0; i < QPW; i++) {
        uint32_t a0 = k0[i], a1 = k1[i];
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const uint32_t b0 = __shfl_xor_sync(0xffffffffu, a0, off), b1 = __shfl_xor_sync(0xffffffffu, a1, off);
            const uint32_t lo = min(a0, b0), hi = max(a0, b0);
            a1 = min(hi, min(a1, b1));
            a0 = lo;
        }
        if (lane == 0 && q0 + i < nq) partial[(size_t)(q0 + i) * nchunks + chunk] = make_uint2(a0, a1);
    }
}

__global__ void knn2_merge_kernel(const uint2* __restrict__ partial, int nq, int nchunks, int32_t* __restrict__ out,
                                  const int32_t* __restrict__ counts, int qcap) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    uint32_t k0 = NONE, k1 = NONE;
    bool live = true;
    if (counts) { const int b = qi / qcap; live = (qi - b * qcap) < counts[b]; }
    for (int c = 0; live && c < nchunks; c++) {
        const uint2 p = partial[(size_t)qi * nchunks + c];
        top2_insert(k0, k1, p.x);
        top2_insert(k0, k1, p.y);
    }
    int4 r;
    r.x = k0 == NONE ? -1 : (int)(k0 & 0x3fffff);
    r.y = k0 == NONE ? -1 : (int)(k0 >> 22);
    r.z = k1 == NONE ? -1 : (int)(k1 & 0x3fffff);
    r.w = k1 == NONE ? -1 : (int)(k1 >> 22);
    reinterpret_cast<int4*>(out)[qi] = r;
}

}  // namespace

extern "C" int alva_k_hamming_knn2(alva_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* out) {
    if (!ctx || !q || !t || !out || nq < 1 || nt < 1 || nt >= (1 << 22) || ((uintptr_t)q & 15) || ((uintptr_t)t & 15) ||
        ((uintptr_t)out & 15)) {
        alva_set_error("alva_k_hamming_knn2: bad argument (need 16-byte aligned buffers, 1 <= nt < 2^22)");
        return ALVA_E_INVALID;
    }
    const int nchunks = (nt + CHUNK - 1) / CHUNK;
    uint2* partial = (uint2*)alva_scratch(ctx, (size_t)nq * nchunks * sizeof(uint2));
    if (!partial) return ALVA_E_CUDA;
    dim3 grid((nq + QPW * WPC - 1) / (QPW * WPC), nchunks);
    knn2_partial_kernel<<<grid, WPC * 32, 0, ctx->stream>>>((const uint4*)q, nq, (const uint4*)t, nt, partial, nchunks, nullptr, 0);
    ALVA_LAUNCH_CHECK(ctx);
    knn2_merge_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(partial, nq, nchunks, out, nullptr, 0);
    ALVA_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int alva_k_hamming_knn2_batch(alva_ctx* ctx, const uint8_t* q, const int32_t* counts, int nbatch, int qcap,
                                         const uint8_t* t, int nt, int32_t* out) {
    if (!ctx || !q || !counts || !t || !out || nbatch < 1 || qcap < 1 || (qcap % QPW) != 0 || nt < 1 || nt >= (1 << 22) ||
        ((uintptr_t)q & 15) || ((uintptr_t)t & 15) || ((uintptr_t)out & 15)) {
        alva_set_error("alva_k_hamming_knn2_batch: bad argument (qcap must be a multiple of %d)", QPW);
        return ALVA_E_INVALID;
    }
    const int nq = nbatch * qcap;
    const int nchunks = (nt + CHUNK - 1) / CHUNK;
    uint2* partial = (uint2*)alva_scratch(ctx, (size_t)nq * nchunks * sizeof(uint2));
    if (!partial) return ALVA_E_CUDA;
    dim3 grid((nq + QPW * WPC - 1) / (QPW * WPC), nchunks);
    knn2_partial_kernel<<<grid, WPC * 32, 0, ctx->stream>>>((const uint4*)q, nq, (const uint4*)t, nt, partial, nchunks, counts, qcap);
    ALVA_LAUNCH_CHECK(ctx);
    knn2_merge_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(partial, nq, nchunks, out, counts, qcap);
    ALVA_LAUNCH_CHECK(ctx);
    return 0;
}
