// hamming_mma.cu -- brute-force Hamming 2-NN on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Reference behaviour (exact, same as hamming.cu): cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2)
//   opencv features2d/src/matchers.cpp:757 -> core/src/batch_distance.cpp:103-123 (batchDistHamming),
//   k-NN insertion :235-248 (strict '<': ties keep the lowest train index).
//
// Why tensor cores: brute-force Hamming IS a contraction.  With every descriptor bit b expanded to the 8-bit value
// 2b - 1 (+1 / -1), the dot product of two 256-element rows is  256 - 2 * hamming,  exactly (an int32, or a float whose
// partial sums are all integers of magnitude <= 256).  So the N_q x N_t distance matrix is one 8-bit GEMM with K = 256, and the
// integer pipes (which bound the LOP3/POPC kernel of hamming.cu at ~7.5e11 distances/s) are left with nothing but the top-2
// selection.  Two operand kinds are built: kind::i8 (+-1 as int8, int32 accumulators; the default) and kind::f8f6f4 (+-1.0 as
// E4M3, fp32 accumulators); the results are bit-identical (every value involved is an exactly representable integer).
// Measured on B200 (profiles/r02a_knn_mma_i8_full.txt, tools/gpu_knn_mma_prof.py): once the epilogue was out of the way both
// kinds run at the same rate, ~4 030 MAC/clk/SM (M = N = 128, cta_group::1: ~130 clk per instruction, tensor pipe ~49 % busy) --
// the limit at this tile shape is the rate the MMA unit streams its two 32 KB operands from shared memory, not the arithmetic
// kind.  N = 256 tiles (half the A re-reads per MAC), A kept in TMEM, or cta_group::2 are the next steps.
//
// Shape of the kernel (one CTA per 256 query rows, 1 CTA / SM, 20 warps):
//   warp 0   producer : bulk async copies (cp.async.bulk, SASS UBLKCP -- the TMA engine's 1-D mode) of pre-tiled 32 KB
//                       operand blobs into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1   issuer   : one elected thread issues tcgen05.mma.kind::f8f6f4 / kind::i8 (SASS UTCQMMA / UTCIMMA), M = 128, N = 128, K = 32 per
//                       instruction, 8 per K = 256, for TWO 128-row query tiles per train tile (every train byte that
//                       leaves L2 feeds 256 query rows); accumulators live in TMEM, double-buffered (4 x 128 columns
//                       = all 512), tcgen05.commit signals "smem stage free" and "accumulator ready"
//   warp 2   TMEM allocator
//   warps 4-19 epilogue: 2 query tiles x 4 TMEM lane quarters x 2 column halves.  tcgen05.ld (SASS LDTM) 2 x 32 columns in
//                       flight; a row (= query) lives in one thread per column half, which reduces every 8 dot products to
//                       their maximum (3-input max) and compares it with the dot product of its current second best; only
//                       the groups that hit run the top-2 insertion, on packed keys hamming << 22 | index (one multiply-add
//                       builds a key, three min / max insert it: OpenCV's strict '<' with ties to the lowest train index is
//                       the unsigned order of those keys).  The two halves of a row are merged through shared memory.
//                       (With 8 epilogue warps and a branchy insertion the MMAs waited on the epilogue 3/4 of the time and
//                       int8 and FP8 operands ran equally slowly; 16 warps + the gated insertion moved the bound to the MMAs.)
// The operands are expanded from the packed 256-bit descriptors by knn2_expand_kernel straight into the shared-memory
// image of a tile (UMMA canonical K-major layout), so the producer needs no tensor map and no swizzle pattern has to be
// matched by hand anywhere else.  Live queries of a batch are compacted on the way ([nbatch][qcap] slots, counts[b] live).
#include "alva_common.cuh"
#include "../../include/alva_b200.h"

namespace {

constexpr int TM = 128;                 // query rows per M tile (UMMA M)
constexpr int TN = 128;                 // train rows per N tile (UMMA N)
constexpr int KBYTES = 256;             // int8 elements per expanded descriptor
constexpr int BLOB = TM * KBYTES;       // one operand tile in shared memory: 32 KB
constexpr int NSTAGE = 4;               // train-tile ring
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 16;            // 2 query tiles x 4 TMEM lane quarters x 2 column halves
constexpr int NTHREADS = (EPI_WARP0 + EPI_WARPS) * 32;
constexpr uint32_t NONE = 0xffffffffu;
constexpr int NEG = -(1 << 20);         // "no candidate yet" dot product

// UMMA shared-memory matrix descriptor pieces, per layout mode (host-filled; kernel parameters so that a debugging run can
// try a variant without a rebuild)
struct MmaLayout {
    uint32_t desc_hi;     // bits 32..63: stride byte offset >> 4 | version 1 (bit 46) | layout type (bits 61..63)
    uint32_t lbo16;       // leading byte offset >> 4 (bits 16..29)
    uint32_t koff[8];     // byte offset of K step j (32 int8 each) inside a tile
    int mode;             // 0: no swizzle ("interleaved" 8 x 16 B core matrices)   1: 128-byte swizzle
                          // 2: as 0 with the two offsets exchanged (debugging aid for the descriptor convention)
};

// byte offset of (row r, 16-byte K chunk c) inside a 128-row tile
__host__ __device__ __forceinline__ uint32_t tile_offset(int mode, int r, int c) {
    if (mode != 1) return (uint32_t)c * (TM * 16) + (uint32_t)r * 16;   // core matrix = 8 rows x 16 B contiguous
    return (uint32_t)(c >> 3) * (TM * 128) + (uint32_t)r * 128 + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
}

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------------------------------------- operand expansion
// offsets[b] = number of live queries in batches < b  (counts clamped to [0, qcap]); one small CTA
__global__ void __launch_bounds__(256) knn2_offsets_kernel(const int32_t* __restrict__ counts, int nbatch, int qcap,
                                                           int32_t* __restrict__ offsets) {
    __shared__ int part[256];
    const int tid = threadIdx.x;
    const int per = (nbatch + 255) / 256;
    const int b0 = tid * per, b1 = min(nbatch, b0 + per);
    int s = 0;
    for (int b = b0; b < b1; b++) s += max(0, min(counts[b], qcap));
    part[tid] = s;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int v = tid >= off ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = tid ? part[tid - 1] : 0;
    for (int b = b0; b < b1; b++) { offsets[b] = run; run += max(0, min(counts[b], qcap)); }
    if (tid == 255) offsets[nbatch] = part[255];
}

// 16 descriptor bits -> 16 bytes (+1 for a set bit, -1 for a clear one), LSB first.  int8: 0x01 / 0xFF; E4M3: 0x38 / 0xB8
__device__ __forceinline__ uint4 expand16(uint32_t v, int kind) {
    uint32_t w[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t nib = (v >> (4 * g)) & 15u;
        const uint32_t b01 = (nib * 0x00204081u) & 0x01010101u;   // bit i of the nibble -> byte i
        w[g] = kind ? (0xB8B8B8B8u ^ (b01 * 0x80u)) : ~(b01 * 0xFEu);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// One thread per (tile row, 16-byte K chunk).  blockIdx.x < a_blocks: query side (compacting live slots through
// `offsets`), else train side.  Rows past the end of either side are written as zeros (their dot products are masked
// / never stored).  rowmap[compact row] = query slot, or -1.
__global__ void __launch_bounds__(256) knn2_expand_kernel(const uint8_t* __restrict__ q, const int32_t* __restrict__ offsets,
                                                          int nbatch, int qcap, int nq_slots, uint8_t* __restrict__ Aexp,
                                                          int32_t* __restrict__ rowmap, int a_blocks, const uint8_t* __restrict__ t,
                                                          int nt, uint8_t* __restrict__ Bexp, int mode, int kind) {
    const bool is_a = (int)blockIdx.x < a_blocks;
    const int idx = (is_a ? blockIdx.x : blockIdx.x - a_blocks) * 256 + threadIdx.x;
    const int tile = idx >> 11, wi = idx & 2047, c = wi >> 7, r = wi & 127;
    const int R = tile * TM + r;
    const uint8_t* src = nullptr;
    if (is_a) {
        const int total = offsets ? offsets[nbatch] : nq_slots;
        if (R >= ((total + 255) & ~255)) return;   // beyond the last 256-row group any CTA will touch
        int slot = -1;
        if (R < total) {
            if (offsets) {
                int lo = 0, hi = nbatch;
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offsets[mid] <= R) lo = mid; else hi = mid; }
                slot = lo * qcap + (R - offsets[lo]);
            } else slot = R;
            src = q + (size_t)slot * 32;
        }
        if (c == 0) rowmap[R] = slot;
    } else {
        if (R < nt) src = t + (size_t)R * 32;
    }
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (src) o = expand16(*reinterpret_cast<const uint16_t*>(src + 2 * c), kind);
    uint8_t* dst = (is_a ? Aexp : Bexp) + (size_t)tile * BLOB + tile_offset(mode, r, c);
    *reinterpret_cast<uint4*>(dst) = o;
}

// ---------------------------------------------------------------------------------------------- the matcher
struct SmemBars {
    uint64_t full[NSTAGE], empty[NSTAGE], tfull[2], tempty[2], afull;
    uint32_t tmem_base;
    uint32_t pad[3];
    uint2 xchg[2 * TM];     // the two best keys of the upper column half of every row
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, const MmaLayout& L) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((L.lbo16 & 0x3FFFu) << 16);
    return (uint64_t)lo | ((uint64_t)L.desc_hi << 32);
}

// accumulator value type per operand kind, and the bit pattern <-> value conversions of the epilogue
template <int KIND> struct Acc;
template <> struct Acc<0> {
    using T = int;
    static __device__ __forceinline__ int from_bits(int b) { return b; }
    static __device__ __forceinline__ int neg() { return NEG; }
    static __device__ __forceinline__ int to_int(int v) { return v; }
    // (256 - dot) << 21 | idx as one multiply-add: dot * -(2^21) + ((256 << 21) + idx)  (idx < 2^22, (256 - dot) even)
    static __device__ __forceinline__ uint32_t key(int v, int idx) { return (uint32_t)v * 0xFFE00000u + ((256u << 21) + (uint32_t)idx); }
    static __device__ __forceinline__ int dot_of_key(uint32_t k) { return 256 - (int)(k >> 21); }   // NONE -> -1791: below every dot product
};
template <> struct Acc<1> {
    using T = float;
    static __device__ __forceinline__ float from_bits(int b) { return __int_as_float(b); }
    static __device__ __forceinline__ float neg() { return (float)NEG; }
    // the accumulators are integer-valued floats of magnitude <= 256: adding 1.5 * 2^23 leaves the integer in the low mantissa bits
    static __device__ __forceinline__ int to_int(float v) { return __float_as_int(v + 12582912.0f) - 0x4B400000; }
    static __device__ __forceinline__ uint32_t key(float v, int idx) { return (uint32_t)to_int(v) * 0xFFE00000u + ((256u << 21) + (uint32_t)idx); }
    static __device__ __forceinline__ float dot_of_key(uint32_t k) { return (float)(256 - (int)(k >> 21)); }
};

template <int KIND>
__global__ void __launch_bounds__(NTHREADS, 1)
knn2_mma_kernel(const uint8_t* __restrict__ Aexp, const uint8_t* __restrict__ Bexp, const int32_t* __restrict__ rowmap,
                const int32_t* __restrict__ total_ptr, int nrows, int nt, int ntiles, int tiles_per_chunk, int nchunks,
                uint2* __restrict__ partial, const MmaLayout L, uint32_t idesc, int32_t* __restrict__ dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = sm;
    uint8_t* sB = sm + 2 * BLOB;
    SmemBars& B = *reinterpret_cast<SmemBars*>(sm + (2 + NSTAGE) * BLOB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = total_ptr ? *total_ptr : nrows;
    const int pair = blockIdx.x, chunk = blockIdx.y;
    if (pair * 2 * TM >= total) return;                       // no live query in this group (uniform: before any barrier)
    const int tile0 = chunk * tiles_per_chunk;
    const int nit = min(tiles_per_chunk, ntiles - tile0);     // >= 1 by construction of nchunks

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&B.full[s], 1); mbar_init(&B.empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&B.tfull[s], 1); mbar_init(&B.tempty[s], EPI_WARPS); }
        mbar_init(&B.afull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&B.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = B.tmem_base;

    if (warp == 0) {
        // ---------------------------------------------------------------- producer
        if (lane == 0) {
            mbar_arrive_expect_tx(&B.afull, 2 * BLOB);
            bulk_g2s(sA, Aexp + (size_t)pair * 2 * BLOB, BLOB, &B.afull);
            bulk_g2s(sA + BLOB, Aexp + (size_t)pair * 2 * BLOB + BLOB, BLOB, &B.afull);
        }
        for (int it = 0; it < nit; it++) {
            const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
            mbar_wait(&B.empty[s], ph ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&B.full[s], BLOB);
                bulk_g2s(sB + s * BLOB, Bexp + (size_t)(tile0 + it) * BLOB, BLOB, &B.full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        mbar_wait(&B.afull, 0);
        for (int it = 0; it < nit; it++) {
            const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
            const int as = it & 1, aph = (it >> 1) & 1;
            mbar_wait(&B.tempty[as], aph ^ 1);     // the epilogue has drained this accumulator stage
            mbar_wait(&B.full[s], ph);             // the train tile has landed
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + s * BLOB);
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const uint32_t d = tmem + (uint32_t)((as * 2 + t) * TN);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint64_t da = make_desc(a0 + t * BLOB + L.koff[j], L), db = make_desc(b0 + L.koff[j], L);
                        if (KIND) umma_f8(d, da, db, idesc, j > 0 ? 1u : 0u); else umma_i8(d, da, db, idesc, j > 0 ? 1u : 0u);
                    }
                }
                umma_commit(&B.empty[s]);          // fires when the MMAs above have finished reading shared memory
                umma_commit(&B.tfull[as]);         // ... and their results are in TMEM
            }
            __syncwarp();
        }
    } else if (warp >= EPI_WARP0) {
        // ---------------------------------------------------------------- epilogue: top-2 per query row and column half
        const int e = warp - EPI_WARP0;
        const int lq = e & 3, t = (e >> 2) & 1, half = e >> 3;   // a warp may only touch TMEM lanes 32 * (warp % 4) ..; EPI_WARP0 % 4 == 0
        const int trow = lq * 32 + lane;                         // row inside the 128-row tile
        const int row = pair * 2 * TM + t * TM + trow;
        using A = Acc<KIND>;
        using T = typename A::T;
        // Per thread: the two smallest keys seen, key = hamming << 22 | index = (256 - dot) << 21 | index -- the lexicographic
        // (distance, index) order of OpenCV's insertion as ONE unsigned compare, so the insertion itself is three min / max and
        // needs no order of arrival.  thr = the dot product of the current second best: a candidate can only matter if its dot
        // product is strictly above it (its index is larger than everything this thread has seen).
        uint32_t k0 = NONE, k1 = NONE;
        T thr = A::neg();
        auto max3 = [](T x, T y, T z) { return max(max(x, y), z); };
        auto scan32 = [&](int (&vb)[32], int n0) {
            T v[32];
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = A::from_bits(vb[j]);
            const bool tail = n0 + 32 > nt;   // only the last train tile: columns past nt hold the dot products of zero rows
            if (tail) {
#pragma unroll
                for (int j = 0; j < 32; j++) if (n0 + j >= nt) v[j] = A::neg();   // below every threshold: never opens a group
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const T g = max3(max3(v[8 * k], v[8 * k + 1], v[8 * k + 2]), max3(v[8 * k + 3], v[8 * k + 4], v[8 * k + 5]), max(v[8 * k + 6], v[8 * k + 7]));
                if (g > thr) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        uint32_t key = A::key(v[8 * k + j], n0 + 8 * k + j);
                        if (tail && n0 + 8 * k + j >= nt) key = NONE;
                        const uint32_t hi = max(k0, key);
                        k0 = min(k0, key);
                        k1 = min(k1, hi);
                    }
                    thr = A::dot_of_key(k1);
                }
            }
        };
        for (int it = 0; it < nit; it++) {
            const int as = it & 1, aph = (it >> 1) & 1;
            mbar_wait(&B.tfull[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)((as * 2 + t) * TN + half * (TN / 2));
            const int n0 = (tile0 + it) * TN + half * (TN / 2);
            int vb0[32], vb1[32];
            tmem_ld32_issue(taddr, vb0);
            tmem_ld32_issue(taddr + 32, vb1);
            tmem_ld_wait();
            // the accumulators are in registers: hand the stage back to the MMA issuer before the selection work
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&B.tempty[as]);
            if (dbg && pair == 0 && chunk == 0 && it == 0 && t == 0) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    dbg[trow * TN + half * (TN / 2) + j] = A::to_int(A::from_bits(vb0[j]));
                    dbg[trow * TN + half * (TN / 2) + 32 + j] = A::to_int(A::from_bits(vb1[j]));
                }
            }
            scan32(vb0, n0);
            scan32(vb1, n0 + 32);
        }
        // merge the two column halves of a row: the upper half hands its pair over through shared memory
        if (half == 1) B.xchg[t * TM + trow] = make_uint2(k0, k1);
        bar_sync_named(1, EPI_WARPS * 32);
        if (half == 0 && row < total) {
            const int slot = rowmap ? rowmap[row] : row;
            if (slot >= 0) {
                const uint2 o = B.xchg[t * TM + trow];
                const uint32_t hi = max(k0, o.x);
                const uint32_t b0 = min(k0, o.x);
                const uint32_t b1 = min(min(k1, o.y), hi);
                partial[(size_t)slot * nchunks + chunk] = make_uint2(b0, b1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host side
int alva_g_knn_mma = 1;        // alva_set_option("knn_mma", 0 never | 1 automatic (large query sets) | 2 always)
int alva_g_knn_mma_kind = 0;   // alva_set_option("knn_mma_kind", 0 int8 operands / int32 accumulators | 1 E4M3 operands / fp32 accumulators)
int alva_g_knn_mma_mode = 0;   // alva_set_option("knn_mma_mode", 0 no swizzle | 1 128-byte swizzle | 2 debugging variant of 0)

int alva_knn2_merge_launch(alva_ctx* ctx, const uint2* partial, int nq, int nchunks, int32_t* out, const int32_t* counts, int qcap);

static MmaLayout make_layout(int mode) {
    MmaLayout L{};
    L.mode = mode;
    if (mode != 1) {
        // K-major, no swizzle: ((8, n), 2) : ((1, SBO), LBO) in 16-byte units -- core matrices of 8 rows x 16 B
        L.lbo16 = (TM * 16) >> 4;                 // next 16-byte K chunk
        L.desc_hi = (128u >> 4) | (1u << 14);     // SBO = next 8-row group; version 1; layout 0
        for (int j = 0; j < 8; j++) L.koff[j] = (uint32_t)j * 2 * TM * 16;
        if (mode == 2) { L.lbo16 = 128u >> 4; L.desc_hi = ((TM * 16u) >> 4) | (1u << 14); }
    } else {
        // K-major, 128-byte swizzle: rows of 128 B, 8-row groups of 1024 B, two 128-byte K atoms per tile
        L.lbo16 = 1;
        L.desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        for (int j = 0; j < 8; j++) L.koff[j] = (uint32_t)(j >> 2) * TM * 128 + (uint32_t)(j & 3) * 32;
    }
    return L;
}

// true if the tensor-core path should serve this problem
bool alva_knn2_mma_wanted(int nq, int nt) {
    if (alva_g_knn_mma == 2) return true;
    if (alva_g_knn_mma == 0) return false;
    return nq >= 8192 && nt >= 1024;
}

// q: [nq][32] (or [nbatch][qcap][32] with counts), t: [nt][32]; out: [nq][4] int32 as alva_k_hamming_knn2
int alva_knn2_mma_launch(alva_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* out, const int32_t* counts,
                         int nbatch, int qcap, int32_t* dbg_out) {
    const int a_tiles = 2 * ((nq + 2 * TM - 1) / (2 * TM));
    const int b_tiles = (nt + TN - 1) / TN;
    const int npairs = a_tiles / 2;
    int nchunks = 1;
    if (npairs < ctx->num_sms) {
        nchunks = (2 * ctx->num_sms + npairs - 1) / npairs;
        if (nchunks > b_tiles) nchunks = b_tiles;
    }
    const int tiles_per_chunk = (b_tiles + nchunks - 1) / nchunks;
    nchunks = (b_tiles + tiles_per_chunk - 1) / tiles_per_chunk;

    // workspace: offsets | rowmap | A tiles | B tiles | dbg
    const size_t off_bytes = (((size_t)(counts ? nbatch + 1 : 1) * 4) + 255) & ~(size_t)255;
    const size_t map_bytes = (((size_t)a_tiles * TM * 4) + 255) & ~(size_t)255;
    const size_t a_bytes = (size_t)a_tiles * BLOB, b_bytes = (size_t)b_tiles * BLOB;
    const size_t need = off_bytes + map_bytes + a_bytes + b_bytes;
    if (need > ctx->knn_ws_bytes) {
        if (ctx->knn_ws) { ALVA_CUDA(cudaStreamSynchronize(ctx->stream)); ALVA_CUDA(cudaFree(ctx->knn_ws)); ctx->knn_ws = nullptr; ctx->knn_ws_bytes = 0; }
        ALVA_CUDA(cudaMalloc(&ctx->knn_ws, need + need / 8));
        ctx->knn_ws_bytes = need + need / 8;
    }
    uint8_t* ws = (uint8_t*)ctx->knn_ws;
    int32_t* offsets = counts ? (int32_t*)ws : nullptr;
    int32_t* rowmap = (int32_t*)(ws + off_bytes);
    uint8_t* Aexp = ws + off_bytes + map_bytes;
    uint8_t* Bexp = Aexp + a_bytes;
    uint2* partial = (uint2*)alva_scratch(ctx, (size_t)nq * nchunks * sizeof(uint2));
    if (!partial) return ALVA_E_CUDA;

    const MmaLayout L = make_layout(alva_g_knn_mma_mode);
    if (counts) {
        knn2_offsets_kernel<<<1, 256, 0, ctx->stream>>>(counts, nbatch, qcap, offsets);
        ALVA_LAUNCH_CHECK(ctx);
    }
    const int a_blocks = a_tiles * 8, b_blocks = b_tiles * 8;
    const int kind = alva_g_knn_mma_kind ? 1 : 0;
    knn2_expand_kernel<<<a_blocks + b_blocks, 256, 0, ctx->stream>>>(q, offsets, nbatch, qcap, nq, Aexp, rowmap, a_blocks, t, nt, Bexp, L.mode, kind);
    ALVA_LAUNCH_CHECK(ctx);

    // instruction descriptor (both operands K-major): N >> 3 at bit 17, M >> 4 at bit 24;
    //   kind::i8     : D = S32 (bits 4-5 = 2), A = B = signed int8 (bits 7-9, 10-12 = 1)
    //   kind::f8f6f4 : D = F32 (bits 4-5 = 1), A = B = E4M3 (bits 7-9, 10-12 = 0)
    const uint32_t shape = ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    const uint32_t idesc = kind ? ((1u << 4) | shape) : ((2u << 4) | (1u << 7) | (1u << 10) | shape);
    const size_t smem = (size_t)(2 + NSTAGE) * BLOB + sizeof(SmemBars) + 1024;
    const dim3 grid(npairs, nchunks);
    const int32_t* total_ptr = counts ? offsets + nbatch : nullptr;
    if (kind) {
        ALVA_CUDA(cudaFuncSetAttribute(knn2_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn2_mma_kernel<1><<<grid, NTHREADS, smem, ctx->stream>>>(Aexp, Bexp, rowmap, total_ptr, nq, nt, b_tiles, tiles_per_chunk, nchunks, partial, L, idesc, dbg_out);
    } else {
        ALVA_CUDA(cudaFuncSetAttribute(knn2_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn2_mma_kernel<0><<<grid, NTHREADS, smem, ctx->stream>>>(Aexp, Bexp, rowmap, total_ptr, nq, nt, b_tiles, tiles_per_chunk, nchunks, partial, L, idesc, dbg_out);
    }
    ALVA_LAUNCH_CHECK(ctx);
    return alva_knn2_merge_launch(ctx, partial, nq, nchunks, out, counts, qcap);
}

// Debugging aid (tools/gpu_knn_mma_check.py): run the tensor-core path unconditionally and also return the raw dot
// products of the first 128 x 128 tile (dbg_dev: 16384 int32, device memory).
extern "C" int alva_debug_knn2_mma(alva_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* out, int32_t* dbg_dev) { AlvaDeviceGuard guard__(ctx);
    if (!ctx || !q || !t || !out || nq < 1 || nt < 1 || nt >= (1 << 22)) { alva_set_error("alva_debug_knn2_mma: bad argument"); return ALVA_E_INVALID; }
    return alva_knn2_mma_launch(ctx, q, nq, t, nt, out, nullptr, 0, 0, dbg_dev);
}
