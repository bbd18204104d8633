// orb.cu -- ORB pre-blur and steered-BRIEF (rBRIEF-256) descriptors at given points, sm_100a.
//
// Reference behaviour (bit-exact; CPU restatement in oracle/alva_oracle.c):
//   FeatureExtractor::describeFeaturePoints   src/slam/src/feature_extractor.cpp:160-214
//     -> ORB::create(500, 1., 0)->compute      opencv features2d/src/orb.cpp:970-1218
//        border rule (31 px, rounded)          orb.cpp:1130, keypoint.cpp:92-117
//        GaussianBlur(7x7, sigma 2) float path orb.cpp:1188 -> imgproc/src/filter.simd.hpp:468-510, 1163-1215
//        computeOrbDescriptors                 orb.cpp:219-350 (pattern orb.cpp:380-638)
//        ICAngles + fastAtan2 (detect mode)    orb.cpp:181-215, core/src/mathfuncs_core.simd.hpp:34-71
//
// The blur is order- and fusion-sensitive float arithmetic: every product and sum below is an explicit
// __fmul_rn/__fadd_rn (or one __fmaf_rn in ALVA_ORB_FMA mode) so nvcc cannot re-associate or contract it.
#include "alva_common.cuh"
#include "../../include/alva_b200.h"
#include <math.h>

namespace {

__constant__ int8_t c_pattern[1024] = {
#include "orb_pattern.inc"
};
// getGaussianKernel(7, 2, CV_32F) (imgproc/src/smooth.dispatch.cpp:76-190): k[3], k[2]=k[4], k[1]=k[5], k[0]=k[6]
__constant__ uint32_t c_gauss_bits[4] = {0x3e5d4ae0u, 0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u};

constexpr int BTW = 128, BTH = 32;          // blur tile
constexpr int BIP = BTW + 8;                // input smem pitch (halo 3, padded to 4)
constexpr int BIR = BTH + 6;

template <bool FMA>
__global__ void __launch_bounds__(256) orb_blur_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int w,
                                                       int h) {
    __shared__ uint8_t in_s[BIR][BIP];
    __shared__ float row_s[BIR][BTW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * BTW, y0 = blockIdx.y * BTH;
    const size_t fo = (size_t)blockIdx.z * w * h;
    const uint8_t* s = src + fo;
    float k[4];
#pragma unroll
    for (int i = 0; i < 4; i++) k[i] = __uint_as_float(c_gauss_bits[i]);   // k[j] = weight at distance j

    for (int i = tid; i < BIR * BIP; i += 256) {
        const int r = i / BIP, c = i - r * BIP;
        const int x = reflect101(min(x0 + c - 4, w + 3), w), y = reflect101(min(y0 + r - 3, h + 3), h);
        in_s[r][c] = __ldg(s + (size_t)y * w + x);
    }
    __syncthreads();
    // row filter: s = k0*S[0]; s += k[i]*S[i], left to right (RowFilter<uchar,float>, filter.simd.hpp:2477-2487)
    for (int i = tid; i < BIR * BTW; i += 256) {
        const int r = i / BTW, c = i - r * BTW;
        const uint8_t* S = &in_s[r][c + 1];   // S[0] = pixel x-3
        float acc;
        if (FMA) {
            acc = 0.f;
#pragma unroll
            for (int t = 0; t < 7; t++) acc = __fmaf_rn((float)S[t], k[t < 3 ? 3 - t : t - 3], acc);
        } else {
            acc = __fmul_rn(k[3], (float)S[0]);
#pragma unroll
            for (int t = 1; t < 7; t++) acc = __fadd_rn(acc, __fmul_rn(k[t < 3 ? 3 - t : t - 3], (float)S[t]));
        }
        row_s[r][c] = acc;
    }
    __syncthreads();
    // column filter: s = k3*R[0]; s += k[3+j]*(R[+j] + R[-j]) (SymmColumnFilter, filter.simd.hpp:2697-2790), cvRound
    for (int i = tid; i < BTH * (BTW / 4); i += 256) {
        const int r = i / (BTW / 4), c4 = (i - r * (BTW / 4)) * 4;
        const int y = y0 + r;
        if (y >= h) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c4 + j;
            float acc = FMA ? __fmaf_rn(k[0], row_s[r + 3][c], 0.f) : __fmul_rn(k[0], row_s[r + 3][c]);
#pragma unroll
            for (int d = 1; d <= 3; d++) {
                const float ab = __fadd_rn(row_s[r + 3 + d][c], row_s[r + 3 - d][c]);
                acc = FMA ? __fmaf_rn(k[d], ab, acc) : __fadd_rn(acc, __fmul_rn(k[d], ab));
            }
            int v = __float2int_rn(acc);
            v = max(0, min(255, v));
            packed |= (uint32_t)v << (8 * j);
        }
        const int x = x0 + c4;
        uint8_t* d = dst + fo + (size_t)y * w + x;
        if (x + 3 < w && ((w & 3) == 0)) *reinterpret_cast<uint32_t*>(d) = packed;
        else
            for (int j = 0; j < 4 && x + j < w; j++) d[j] = (uint8_t)(packed >> (8 * j));
    }
}

// cv::fastAtan2 (degrees), baseline arithmetic: no contraction
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / 3.1415926535897932384626433832795);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.1415926535897932384626433832795);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.1415926535897932384626433832795);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.1415926535897932384626433832795);
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// umax of computeKeyPoints (orb.cpp:819-834) for halfPatchSize 15
__constant__ int8_t c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// one warp per keypoint; lane j produces descriptor byte j
__global__ void __launch_bounds__(256) orb_describe_kernel(const uint8_t* __restrict__ gray, const uint8_t* __restrict__ blurred,
                                                           int w, int h, const float* __restrict__ pts,
                                                           const int32_t* __restrict__ npts_per_frame, int npts, int flags,
                                                           uint8_t* __restrict__ desc, uint8_t* __restrict__ kept,
                                                           float* __restrict__ angles_out) {
    __shared__ int8_t pat_s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) pat_s[i] = c_pattern[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const int kp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n = npts_per_frame ? min(npts_per_frame[f], npts) : npts;
    if (kp >= npts) return;
    const size_t o = (size_t)f * npts + kp;
    if (kp >= n) {   // unused slot
        if (lane == 0) kept[o] = 0;
        desc[o * 32 + lane] = 0;
        if (angles_out && lane == 0) angles_out[o] = -1.f;
        return;
    }
    const float px = pts[o * 2], py = pts[o * 2 + 1];
    const int cx = __float2int_rn(px), cy = __float2int_rn(py);
    const bool keep = (w > 62 && h > 62) && cx >= 31 && cx < w - 31 && cy >= 31 && cy < h - 31;
    if (!keep) {
        if (lane == 0) kept[o] = 0;
        desc[o * 32 + lane] = 0;
        if (angles_out && lane == 0) angles_out[o] = -1.f;
        return;
    }
    const size_t fo = (size_t)f * w * h;
    float angle = -1.f;   // KeyPoint::convert default (core/src/types.cpp:93-101)
    if (flags & ALVA_ORB_IC_ANGLE) {
        // intensity-centroid moments over the circular patch r = 15 (exact integers), warp-reduced
        const uint8_t* c = gray + fo + (size_t)cy * w + cx;
        int m01 = 0, m10 = 0;
        // lanes split the 31 columns u = -15..15 (lane 31 idle)
        const int u = lane - 15;
        if (lane < 31) {
            m10 += u * c[u];   // v = 0 row
            for (int v = 1; v <= 15; v++) {
                if (abs(u) <= c_umax[v]) {
                    const int vp = c[u + v * w], vm = c[u - v * w];
                    m01 += v * (vp - vm);
                    m10 += u * (vp + vm);
                }
            }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            m01 += __shfl_xor_sync(0xffffffffu, m01, off);
            m10 += __shfl_xor_sync(0xffffffffu, m10, off);
        }
        angle = fast_atan2_deg((float)m01, (float)m10);
    }
    // orb.cpp:232-235: angle *= (float)(CV_PI/180.f); a = (float)cos(angle), b = (float)sin(angle)
    const float ang = __fmul_rn(angle, (float)(3.1415926535897932384626433832795 / 180.f));
    const float a = (float)cos((double)ang), b = (float)sin((double)ang);
    const uint8_t* center = blurred + fo + (size_t)cy * w + cx;
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int8_t* p = pat_s + 4 * (8 * lane + t);
        const float p0x = (float)p[0], p0y = (float)p[1], p1x = (float)p[2], p1y = (float)p[3];
        const int x0 = __float2int_rn(__fsub_rn(__fmul_rn(p0x, a), __fmul_rn(p0y, b)));
        const int y0 = __float2int_rn(__fadd_rn(__fmul_rn(p0x, b), __fmul_rn(p0y, a)));
        const int x1 = __float2int_rn(__fsub_rn(__fmul_rn(p1x, a), __fmul_rn(p1y, b)));
        const int y1 = __float2int_rn(__fadd_rn(__fmul_rn(p1x, b), __fmul_rn(p1y, a)));
        const int t0 = center[y0 * w + x0], t1 = center[y1 * w + x1];
        val |= (t0 < t1) << t;
    }
    desc[o * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        kept[o] = 1;
        if (angles_out) angles_out[o] = angle;
    }
}

}  // namespace

extern "C" int alva_k_orb_blur(alva_ctx* ctx, const uint8_t* gray, uint8_t* blurred, int w, int h, int nframes, int flags) {
    if (!ctx || !gray || !blurred || w < 8 || h < 8 || nframes < 1 || gray == blurred) {
        alva_set_error("alva_k_orb_blur: bad argument (in-place not supported)");
        return ALVA_E_INVALID;
    }
    dim3 grid((w + BTW - 1) / BTW, (h + BTH - 1) / BTH, nframes);
    if (flags & ALVA_ORB_FMA) orb_blur_kernel<true><<<grid, 256, 0, ctx->stream>>>(gray, blurred, w, h);
    else orb_blur_kernel<false><<<grid, 256, 0, ctx->stream>>>(gray, blurred, w, h);
    ALVA_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int alva_k_orb_describe(alva_ctx* ctx, const uint8_t* gray, const uint8_t* blurred, int w, int h, int nframes,
                                   const float* pts, const int32_t* npts_per_frame, int npts, int flags, uint8_t* desc,
                                   uint8_t* kept, float* angles_out) {
    if (!ctx || !blurred || !pts || !desc || !kept || w < 1 || h < 1 || nframes < 1 || npts < 1 ||
        ((flags & ALVA_ORB_IC_ANGLE) && !gray)) {
        alva_set_error("alva_k_orb_describe: bad argument");
        return ALVA_E_INVALID;
    }
    dim3 grid((npts + 7) / 8, nframes);
    orb_describe_kernel<<<grid, 256, 0, ctx->stream>>>(gray, blurred, w, h, pts, npts_per_frame, npts, flags, desc, kept,
                                                       angles_out);
    ALVA_LAUNCH_CHECK(ctx);
    return 0;
}
