// Weight gradient of the pointwise convs on the tcgen05 tensor cores (sm_100a), persistent and warp-specialised.
//
//   dw[n,k] += sum_{b,r} pro_dy(dy[b,r,n], dy2[b,r,n]) * pro_x(x[b,r,k])          (x3d_fine.py:100-105 backward)
//
// The reduction runs over ROWS (millions of them) and the result is a tiny [N,K] matrix: a tall-skinny "transposed"
// GEMM.  The CUDA-core kernel it replaces (pw_wgrad_kernel, x3d_pw.cu) took 28 % of the training step.  Here both
// operands are MN-major for the tensor core: a block of RB rows of dy (resp. x) is stored as [row][32 channels] =
// 128-byte rows (32-byte-unit swizzle) and read by tcgen05.mma as A = dy^T
// (M = dy channels, K = rows) and B = x (N = x channels), 8 rows per instruction, 3xTF32 (hi/lo split, fp32
// accumulation in TMEM).  One CTA per SM walks a contiguous range of row blocks and keeps its [N,K] partial in TMEM
// for the whole launch; at the end the partials are added to dw with fp32 atomics (148 x N x K per launch).
//
//   warps 0-15  producers: batches of 4 (tensor, 32-channel chunk) units per thread: loads issued first, then the
//               BatchNorm-backward / Swish prologue, the hi/lo split and the swizzled stores into a ring of stages;
//   warp  16    MMA issuer (convergent loop, one elected lane issues); commits stage-free barriers;
//   all warps   final TMEM -> global atomics.
// Problems whose [N,K] partial exceeds the 512 TMEM columns are split over blockIdx.y (dy channel tiles or x channel
// halves).  Strided / windowed inputs, bias gradients, odd channel counts and N or K > 512 stay on the CUDA-core kernel.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include "tc_ptx.cuh"
#include "tma_host.cuh"
#include <stdlib.h>
#include <string.h>

#define WG_PROD_WARPS 16
#define WG_PROD_THREADS (WG_PROD_WARPS * 32)
#define WG_THREADS (WG_PROD_THREADS + 64)        /* + MMA issuer warp + TMA loader warp */
#define WG_MAX_RAW 8
#define WG_MAX_STAGES 4
#define WG_UB 2                                  /* units per batch; two batches (this one and the next) are in registers */
#define WG_SMEM_MAX (220 * 1024)

struct WgParams {
    int B, R, N, K;                              // N = dy channels, K = x channels
    int RB, rbps, nstages;                       // rows per stage, row blocks per sample
    int nchA_pad, nchB;                          // padded M-side chunks (4 per 128-channel tile) and N-side chunks of ONE CTA
    // The tensor with MORE channels is the M side (padded to 128 per tile), the other one the N side (padded to 32):
    // swap = 0: M = dy channels, N = x channels;  swap = 1: M = x channels, N = dy channels (D is written transposed)
    int swap, ndyc, nxc;                         // dy / x chunks this CTA loads
    uint32_t dy_off, dy_lo, x_off, x_lo;         // stage offsets of the dy / x hi regions and hi -> lo distances
    int mt_per, npad_per;                        // dy channel tiles (128) and padded x channels of one CTA
    int msplit, nsplit;                          // blockIdx.y = ms * nsplit + ns
    int items_per_cta;
    long long total_items;
    uint32_t chunk_bytes, stage_bytes, tmem_cols;
    int av_dy, av_x;
    // strided 1x1x1 conv (downsample branch): the x rows are gathered from a [Ti,Hi,Wi] volume at (t*st, h*sh, w*sw)
    int gmode, gH, gW, gHi, gWi, gst, gsh, gsw;
    long long g_sample_stride;
    // TMA-fed producers: per row block the loader lane lands dy (+ dy2) and x in a raw stage [dy | dy2 | x]; a tensor whose
    // rows are not 16-byte multiples (54 channels) is presented with `fold` rows per TMA row (tma_host.cuh)
    int tma, fold_dy, fold_x, nraw;
    uint32_t raw_dy_bytes, raw_x_bytes, raw_stage_bytes, raw_off;
};

// MN-major descriptor.  32-bit (tf32) MN-major operands only exist in the SWIZZLE_128B_BASE32B layout (layout type 1;
// cute::UMMA::Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over [4 rows][128 B]): rows of 32 channels are 128 B apart, the
// four 32-byte units of a row are XORed with (row & 3), the pattern repeats every 4 rows (512 B).  LBO = stride between
// 32-channel blocks, SBO = stride between 4-row groups.
__device__ __forceinline__ uint64_t wg_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           (1ull << 61);
}
// byte offset of (row, 16-byte chunk q8) inside a [rows][32 floats] chunk in that layout
__device__ __forceinline__ uint32_t wg_sw_off(int row, int q8) {
    return (uint32_t)(row * 128 + ((((q8 >> 1) ^ (row & 3)) << 5) | ((q8 & 1) << 4)));
}
__device__ __forceinline__ bool wg_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float wg_sigmoid(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
template <int MODE>
__device__ __forceinline__ float wg_pro(float x, float x2, float a, float b, float c) {
    if (MODE == CF_PRO_AFFINE) return fmaf(a, x, b);
    if (MODE == CF_PRO_AFFINE_RELU) return fmaxf(fmaf(a, x, b), 0.f);
    if (MODE == CF_PRO_AFFINE_SWISH) { float z = fmaf(a, x, b); return z * wg_sigmoid(z); }
    if (MODE == CF_PRO_AFFINE2) return fmaf(a, x, fmaf(b, x2, c));
    return x;
}
// 4 floats at p (channel c .. c+3 of a row with C channels), zero beyond C; AV = 4 or 2
__device__ __forceinline__ void wg_ld4(const float* p, int c, int C, int av, float* v) {
    if (av == 4) {
        if (c < C) {
            float4 t = __ldg(reinterpret_cast<const float4*>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            v[0] = v[1] = v[2] = v[3] = 0.f;
        }
    } else {
        if (c < C) {
            float2 t = __ldg(reinterpret_cast<const float2*>(p));
            v[0] = t.x; v[1] = t.y;
        } else {
            v[0] = v[1] = 0.f;
        }
        if (c + 2 < C) {
            float2 t = __ldg(reinterpret_cast<const float2*>(p + 2));
            v[2] = t.x; v[3] = t.y;
        } else {
            v[2] = v[3] = 0.f;
        }
    }
}

template <int DYM, int XM>
__global__ void __launch_bounds__(WG_THREADS, 1) pw_wgrad_tc_kernel(const cf_pw_wgrad_args a, const WgParams p,
                                                                    const __grid_constant__ CUtensorMap tm_dy,
                                                                    const __grid_constant__ CUtensorMap tm_dy2,
                                                                    const __grid_constant__ CUtensorMap tm_x) {
    cf_pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[WG_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty[WG_MAX_STAGES];
    __shared__ __align__(8) uint64_t rfull[WG_MAX_RAW];
    __shared__ __align__(8) uint64_t rempty[WG_MAX_RAW];
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_addr_s;

    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stages = base;
    float* tabA = reinterpret_cast<float*>(stages + (size_t)p.nstages * p.stage_bytes);      // dy tables [3][ndyc*32]
    float* tabB = tabA + 3 * p.ndyc * 32;                                                      // x tables  [2][nxc*32]
    uint8_t* raw = base + p.raw_off;                                                           // TMA landing ring

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ms = blockIdx.y / p.nsplit, ns = blockIdx.y - ms * p.nsplit;
    const int m_base = ms * p.mt_per * 128;                  // first M-side / N-side channel of this CTA
    const int c_base = ns * p.npad_per;
    const int n_base = p.swap ? c_base : m_base;             // first dy channel of this CTA
    const int k_base = p.swap ? m_base : c_base;             // first x channel of this CTA
    const long long item0 = (long long)blockIdx.x * p.items_per_cta;
    const long long item1 = min(item0 + p.items_per_cta, p.total_items);

    if (warp == WG_PROD_WARPS) {
        tmem_alloc(&tmem_addr_s, p.tmem_cols);
        tmem_relinquish();
    }
    if (tid == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            mbar_init(&full[s], WG_PROD_WARPS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&done_bar, 1);
        for (int s = 0; s < p.nraw; ++s) {
            mbar_init(&rfull[s], 1);
            mbar_init(&rempty[s], WG_PROD_WARPS);
        }
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_addr_s;
    const uint32_t a_lo_off = (uint32_t)p.nchA_pad * p.chunk_bytes;              // A lo region follows A hi
    const uint32_t b_hi_off = 2u * a_lo_off;
    const uint32_t b_lo_off = b_hi_off + (uint32_t)p.nchB * p.chunk_bytes;

    if (warp < WG_PROD_WARPS) {
        // ================= producers =================
        // A thread's work is a sequence of batches (item, u0): up to WG_UB (tensor, chunk) units of one row block.  The
        // loads of batch i+1 -- usually the next row block -- are issued before batch i is transformed and stored, so a
        // full memory latency is never exposed per stage (it was: 2400 cycles per 32-row stage in the first version).
        const int TPC = p.RB * 8;                            // threads per chunk
        const int groups = WG_PROD_THREADS / TPC;
        const int grp = tid / TPC, lt = tid - grp * TPC;
        const int row = lt >> 3, q8 = lt & 7;
        const uint32_t soff = wg_sw_off(row, q8);
        const int nunits = p.ndyc + p.nxc;
        const int N = p.N, K = p.K;
        const int ustep = groups * WG_UB;
        constexpr bool aff2 = DYM == CF_PRO_AFFINE2;
        float vA[WG_UB][4], wA[WG_UB][4], vB[WG_UB][4], wB[WG_UB][4];

        // (b, rb) of a row block advance incrementally: no divisions in the per-stage path
        auto load_batch = [&](int b, int rb, int u0, float (&v)[WG_UB][4], float (&v2)[WG_UB][4]) {
            const int r0 = rb * p.RB;
            const bool rv = row < min(p.RB, p.R - r0);
            const size_t grow = (size_t)b * p.R + r0 + row;
            const float* dyp = a.dy + grow * N + n_base + q8 * 4;
            const float* dy2p = aff2 ? a.dy2 + grow * N + n_base + q8 * 4 : nullptr;
            const float* xp;
            if (p.gmode) {
                const int r = rv ? r0 + row : 0;
                const int w = r % p.gW, q = r / p.gW;
                const int h = q % p.gH, t = q / p.gH;
                xp = a.x + (size_t)b * p.g_sample_stride + (((size_t)(t * p.gst) * p.gHi + h * p.gsh) * p.gWi + w * p.gsw) * K + k_base + q8 * 4;
            } else {
                xp = a.x + grow * K + k_base + q8 * 4;
            }
#pragma unroll
            for (int i = 0; i < WG_UB; ++i) {
                const int u = u0 + i * groups;
                if (u < nunits && rv) {
                    if (u < p.ndyc) {
                        const int c = n_base + u * 32 + q8 * 4;
                        wg_ld4(dyp + u * 32, c, N, p.av_dy, v[i]);
                        if (aff2) wg_ld4(dy2p + u * 32, c, N, p.av_dy, v2[i]);
                    } else {
                        const int c = k_base + (u - p.ndyc) * 32 + q8 * 4;
                        wg_ld4(xp + (u - p.ndyc) * 32, c, K, p.av_x, v[i]);
                    }
                } else {
                    v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
                    v2[i][0] = v2[i][1] = v2[i][2] = v2[i][3] = 0.f;
                }
            }
        };
        auto store_batch = [&](uint8_t* stage, int u0, bool rv, const float (&v)[WG_UB][4], const float (&v2)[WG_UB][4]) {
#pragma unroll
            for (int i = 0; i < WG_UB; ++i) {
                const int u = u0 + i * groups;
                if (u >= nunits) break;
                float hi[4], lo[4];
                if (u < p.ndyc) {
                    const int tl = u * 32 + q8 * 4;
                    const float4 ta = *reinterpret_cast<const float4*>(tabA + tl);
                    const float4 tb = *reinterpret_cast<const float4*>(tabA + p.ndyc * 32 + tl);
                    const float4 tc = *reinterpret_cast<const float4*>(tabA + 2 * p.ndyc * 32 + tl);
                    const float pa[4] = {ta.x, ta.y, ta.z, ta.w}, pb[4] = {tb.x, tb.y, tb.z, tb.w}, pc[4] = {tc.x, tc.y, tc.z, tc.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t = v[i][e];
                        if (DYM != CF_PRO_NONE) t = rv ? wg_pro<DYM>(t, v2[i][e], pa[e], pb[e], pc[e]) : 0.f;
                        tf32_split(t, hi[e], lo[e]);
                    }
                    uint8_t* dst = stage + p.dy_off + (uint32_t)u * p.chunk_bytes + soff;
                    *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(dst + p.dy_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                } else {
                    const int j = u - p.ndyc;
                    const int tl = j * 32 + q8 * 4;
                    const float4 ta = *reinterpret_cast<const float4*>(tabB + tl);
                    const float4 tb = *reinterpret_cast<const float4*>(tabB + p.nxc * 32 + tl);
                    const float pa[4] = {ta.x, ta.y, ta.z, ta.w}, pb[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t = v[i][e];
                        if (XM != CF_PRO_NONE) t = rv ? wg_pro<XM>(t, 0.f, pa[e], pb[e], 0.f) : 0.f;
                        tf32_split(t, hi[e], lo[e]);
                    }
                    uint8_t* dst = stage + p.x_off + (uint32_t)j * p.chunk_bytes + soff;
                    *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(dst + p.x_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        };

        int s = 0, cur_b = -1;
        uint32_t ph = 0;
        int b = (int)(item0 / p.rbps);
        int rb = (int)(item0 - (long long)b * p.rbps);
        long long item = item0;
        if (p.tma) {
            // ---- TMA-fed: the row block's dy (+ dy2) and x tiles are already in shared memory (raw ring, filled by the
            // loader lane with cp.async.bulk.tensor); read -> prologue -> hi/lo split -> swizzled operand stage.
            // Written for a SHORT instruction stream: the ncu source view of round 1's producers showed ~510 issue slots per
            // thread and 64-row stage of which ~10 % were loads / FMAs / stores -- the rest address arithmetic, predicates
            // and reconvergence of the generic batch / unit bookkeeping.  Here: one loop per tensor, all offsets of a
            // thread (row, 16-byte slot) computed once, nothing per unit but a constant stride.
            int rs = 0;
            uint32_t rph = 0;
            const uint32_t roff1 = (uint32_t)row * 128u + (uint32_t)q8 * 16u;                  // fold == 1: [RB][32] chunks
            const uint32_t ustride = (uint32_t)groups * p.chunk_bytes;
            const int cq = q8 * 4;
            const float* tA = tabA + cq;
            const float* tB = tabB + cq;
            const int nA = p.ndyc * 32, nB = p.nxc * 32;
            for (; item < item1; ++item) {
                const bool rv = row < min(p.RB, p.R - rb * p.RB);
                if (b != cur_b) {
                    named_bar_sync(1, WG_PROD_THREADS);
                    for (int t = tid; t < nA; t += WG_PROD_THREADS) {
                        const int n = n_base + t;
                        const bool v = n < N && DYM != CF_PRO_NONE;
                        tabA[t] = v ? a.dy_a[(size_t)b * N + n] : 0.f;
                        tabA[nA + t] = (v && a.dy_b) ? a.dy_b[(size_t)b * N + n] : 0.f;
                        tabA[2 * nA + t] = (v && a.dy_c) ? a.dy_c[(size_t)b * N + n] : 0.f;
                    }
                    for (int t = tid; t < nB; t += WG_PROD_THREADS) {
                        const int k = k_base + t;
                        const bool v = k < K && XM != CF_PRO_NONE;
                        tabB[t] = v ? a.x_a[(size_t)b * K + k] : 0.f;
                        tabB[nB + t] = (v && a.x_b) ? a.x_b[(size_t)b * K + k] : 0.f;
                    }
                    named_bar_sync(1, WG_PROD_THREADS);
                    cur_b = b;
                }
                mbar_wait_b(&rfull[rs], rph);
                mbar_wait_b(&empty[s], ph ^ 1u);
                const uint8_t* rdy = raw + (size_t)rs * p.raw_stage_bytes;
                const uint8_t* rx = rdy + (aff2 ? 2u : 1u) * p.raw_dy_bytes;
                uint8_t* stage = stages + (size_t)s * p.stage_bytes;
                // ---- dy units (this thread's group takes every `groups`-th 32-channel chunk)
                {
                    uint8_t* dst = stage + p.dy_off + soff + (uint32_t)grp * p.chunk_bytes;
                    for (int u = grp; u < p.ndyc; u += groups, dst += ustride) {
                        float v[4], v2[4];
                        if (p.fold_dy == 1) {
                            const float4 t = *reinterpret_cast<const float4*>(rdy + (uint32_t)u * p.chunk_bytes + roff1);
                            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                            if (aff2) {
                                const float4 t2 = *reinterpret_cast<const float4*>(rdy + p.raw_dy_bytes + (uint32_t)u * p.chunk_bytes + roff1);
                                v2[0] = t2.x; v2[1] = t2.y; v2[2] = t2.z; v2[3] = t2.w;
                            }
                        } else {                                 // whole [RB][N] tile, 8-byte aligned rows (N even)
                            const int c = u * 32 + cq;
                            const float* src = reinterpret_cast<const float*>(rdy) + row * N + c;
#pragma unroll
                            for (int e = 0; e < 4; e += 2) {
                                float2 t = make_float2(0.f, 0.f), t2 = make_float2(0.f, 0.f);
                                if (c + e < N) {
                                    t = *reinterpret_cast<const float2*>(src + e);
                                    if (aff2) t2 = *reinterpret_cast<const float2*>(src + (p.raw_dy_bytes >> 2) + e);
                                }
                                v[e] = t.x; v[e + 1] = t.y; v2[e] = t2.x; v2[e + 1] = t2.y;
                            }
                        }
                        float hi[4], lo[4];
                        if (DYM != CF_PRO_NONE) {
                            const float4 ta = *reinterpret_cast<const float4*>(tA + u * 32);
                            const float4 tb = *reinterpret_cast<const float4*>(tA + nA + u * 32);
                            const float4 tc = *reinterpret_cast<const float4*>(tA + 2 * nA + u * 32);
                            const float pa[4] = {ta.x, ta.y, ta.z, ta.w}, pb[4] = {tb.x, tb.y, tb.z, tb.w}, pc[4] = {tc.x, tc.y, tc.z, tc.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = rv ? wg_pro<DYM>(v[e], aff2 ? v2[e] : 0.f, pa[e], pb[e], pc[e]) : 0.f;
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) tf32_split(v[e], hi[e], lo[e]);
                        *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4*>(dst + p.dy_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                // ---- x units
                {
                    uint8_t* dst = stage + p.x_off + soff + (uint32_t)grp * p.chunk_bytes;
                    for (int u = grp; u < p.nxc; u += groups, dst += ustride) {
                        float v[4];
                        if (p.fold_x == 1) {
                            const float4 t = *reinterpret_cast<const float4*>(rx + (uint32_t)u * p.chunk_bytes + roff1);
                            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        } else {
                            const int c = u * 32 + cq;
                            const float* src = reinterpret_cast<const float*>(rx) + row * K + c;
#pragma unroll
                            for (int e = 0; e < 4; e += 2) {
                                float2 t = make_float2(0.f, 0.f);
                                if (c + e < K) t = *reinterpret_cast<const float2*>(src + e);
                                v[e] = t.x; v[e + 1] = t.y;
                            }
                        }
                        float hi[4], lo[4];
                        if (XM != CF_PRO_NONE) {
                            const float4 ta = *reinterpret_cast<const float4*>(tB + u * 32);
                            const float4 tb = *reinterpret_cast<const float4*>(tB + nB + u * 32);
                            const float pa[4] = {ta.x, ta.y, ta.z, ta.w}, pb[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = rv ? wg_pro<XM>(v[e], 0.f, pa[e], pb[e], 0.f) : 0.f;
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) tf32_split(v[e], hi[e], lo[e]);
                        *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4*>(dst + p.x_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&rempty[rs]);                 // raw stage read: the loader may refill it
                if (++rs == p.nraw) { rs = 0; rph ^= 1u; }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                if (++s == p.nstages) { s = 0; ph ^= 1u; }
                if (++rb == p.rbps) { rb = 0; ++b; }
            }
        }
        if (!p.tma && item0 < item1 && grp < nunits) load_batch(b, rb, grp, vA, wA);

        // One row block.  Its batches alternate between the two register sets starting with (cv, cw); which set a batch
        // uses is fixed in the code (a run-time "current set" flag made the compiler copy registers right after the
        // loads, i.e. wait for them: no prefetch at all -- 34 % of all stall samples sat on those moves).
        auto item_body = [&](float (&cv)[WG_UB][4], float (&cw)[WG_UB][4], float (&nv)[WG_UB][4], float (&nw)[WG_UB][4]) {
            const bool rv = row < min(p.RB, p.R - rb * p.RB);
            int nb = b, nrb = rb + 1;                        // the next row block
            if (nrb == p.rbps) { nrb = 0; ++nb; }
            if (b != cur_b) {                                // per-sample prologue tables (zero beyond the real channels)
                named_bar_sync(1, WG_PROD_THREADS);
                for (int t = tid; t < p.ndyc * 32; t += WG_PROD_THREADS) {
                    const int n = n_base + t;
                    const bool v = n < N && DYM != CF_PRO_NONE;
                    tabA[t] = v ? a.dy_a[(size_t)b * N + n] : 0.f;
                    tabA[p.ndyc * 32 + t] = (v && a.dy_b) ? a.dy_b[(size_t)b * N + n] : 0.f;
                    tabA[2 * p.ndyc * 32 + t] = (v && a.dy_c) ? a.dy_c[(size_t)b * N + n] : 0.f;
                }
                for (int t = tid; t < p.nxc * 32; t += WG_PROD_THREADS) {
                    const int k = k_base + t;
                    const bool v = k < K && XM != CF_PRO_NONE;
                    tabB[t] = v ? a.x_a[(size_t)b * K + k] : 0.f;
                    tabB[p.nxc * 32 + t] = (v && a.x_b) ? a.x_b[(size_t)b * K + k] : 0.f;
                }
                named_bar_sync(1, WG_PROD_THREADS);
                cur_b = b;
            }
            mbar_wait_b(&empty[s], ph ^ 1u);
            uint8_t* stage = stages + (size_t)s * p.stage_bytes;
            const bool more_items = item + 1 < item1;
            for (int u0 = grp; u0 < nunits;) {
                {
                    const int nu0 = u0 + ustep;
                    if (nu0 < nunits) load_batch(b, rb, nu0, nv, nw);
                    else if (more_items) load_batch(nb, nrb, grp, nv, nw);
                    store_batch(stage, u0, rv, cv, cw);
                    u0 = nu0;
                }
                if (u0 >= nunits) break;
                {
                    const int nu0 = u0 + ustep;
                    if (nu0 < nunits) load_batch(b, rb, nu0, cv, cw);
                    else if (more_items) load_batch(nb, nrb, grp, cv, cw);
                    store_batch(stage, u0, rv, nv, nw);
                    u0 = nu0;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (++s == p.nstages) { s = 0; ph ^= 1u; }
            b = nb;
            rb = nrb;
            ++item;
        };
        const int nbatches = grp < nunits ? (nunits - grp + ustep - 1) / ustep : 0;
        if (nbatches & 1) {                                  // odd: consecutive row blocks start on alternating sets
            while (item + 1 < item1) {
                item_body(vA, wA, vB, wB);
                item_body(vB, wB, vA, wA);
            }
            if (item < item1) item_body(vA, wA, vB, wB);
        } else {
            while (item < item1) item_body(vA, wA, vB, wB);
        }
    } else if (warp == WG_PROD_WARPS + 1) {
        // ================= TMA loader (one lane) =================
        if (p.tma && lane == 0) {
            constexpr bool aff2 = DYM == CF_PRO_AFFINE2;
            tma_prefetch_desc(&tm_dy);
            tma_prefetch_desc(&tm_x);
            if (aff2) tma_prefetch_desc(&tm_dy2);
            int rs = 0;
            uint32_t rph = 0;
            int b = (int)(item0 / p.rbps);
            int rb = (int)(item0 - (long long)b * p.rbps);
            for (long long item = item0; item < item1; ++item) {
                mbar_wait_b(&rempty[rs], rph ^ 1u);
                uint8_t* dst = raw + (size_t)rs * p.raw_stage_bytes;
                mbar_expect_tx(&rfull[rs], p.raw_stage_bytes);
                const int r0 = rb * p.RB;
                if (p.fold_dy == 1) {
                    for (int u = 0; u < p.ndyc; ++u) {
                        tma_load_3d(dst + (size_t)u * p.chunk_bytes, &tm_dy, n_base + u * 32, r0, b, &rfull[rs]);
                        if (aff2) tma_load_3d(dst + p.raw_dy_bytes + (size_t)u * p.chunk_bytes, &tm_dy2, n_base + u * 32, r0, b, &rfull[rs]);
                    }
                } else {
                    tma_load_3d(dst, &tm_dy, 0, r0 / p.fold_dy, b, &rfull[rs]);
                    if (aff2) tma_load_3d(dst + p.raw_dy_bytes, &tm_dy2, 0, r0 / p.fold_dy, b, &rfull[rs]);
                }
                uint8_t* xd = dst + (aff2 ? 2u : 1u) * p.raw_dy_bytes;
                if (p.fold_x == 1) {
                    for (int u = 0; u < p.nxc; ++u) tma_load_3d(xd + (size_t)u * p.chunk_bytes, &tm_x, k_base + u * 32, r0, b, &rfull[rs]);
                } else {
                    tma_load_3d(xd, &tm_x, 0, r0 / p.fold_x, b, &rfull[rs]);
                }
                if (++rs == p.nraw) { rs = 0; rph ^= 1u; }
                if (++rb == p.rbps) { rb = 0; ++b; }
            }
        }
    } else {
        // ================= MMA issuer =================
        // instruction descriptor: D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7, 10), A and B MN-major (bits 15, 16),
        // N >> 3 @ bit 17, M >> 4 @ bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.npad_per >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        const uint32_t stages_s = smem_u32(stages);
        const int ksteps = p.RB >> 3;
        int s = 0;
        uint32_t ph = 0;
        bool first = true;
        for (long long item = item0; item < item1; ++item) {
            mbar_wait_b(&full[s], ph);
            tc_fence_after();
            const uint32_t st = stages_s + (uint32_t)s * p.stage_bytes;
            const uint64_t a_hi = wg_desc_mn(st, p.chunk_bytes, 512u);
            const uint64_t a_lo = wg_desc_mn(st + a_lo_off, p.chunk_bytes, 512u);
            const uint64_t b_hi = wg_desc_mn(st + b_hi_off, p.chunk_bytes, 512u);
            const uint64_t b_lo = wg_desc_mn(st + b_lo_off, p.chunk_bytes, 512u);
            if (wg_elect_one()) {
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t ko = (uint64_t)(ks * (1024 >> 4));                  // next group of 8 rows
                    for (int mt = 0; mt < p.mt_per; ++mt) {
                        const uint64_t mo = (uint64_t)(((uint32_t)mt * 4u * p.chunk_bytes) >> 4) + ko;   // 128 dy channels = 4 chunks
                        const uint32_t d = tmem + (uint32_t)(mt * p.npad_per);
                        umma_tf32(d, a_lo + mo, b_hi + ko, idesc, (uint32_t)(!(first && ks == 0)));
                        umma_tf32(d, a_hi + mo, b_lo + ko, idesc, 1u);
                        umma_tf32(d, a_hi + mo, b_hi + ko, idesc, 1u);
                    }
                }
                umma_commit(&empty[s]);
                if (item == item1 - 1) umma_commit(&done_bar);
            }
            __syncwarp();
            first = false;
            if (++s == p.nstages) { s = 0; ph ^= 1u; }
        }
    }

    // ================= all warps: TMEM partial -> dw (fp32 atomics) =================
    __syncthreads();
    if (item1 > item0) {
        mbar_wait_b(&done_bar, 0u);
        tc_fence_after();
        const int qd = warp & 3;                             // TMEM lane quadrant of this warp
        const int nslab = p.npad_per >> 5;
        const int njobs = p.mt_per * nslab;                  // (channel tile, 32-column slab) jobs, spread over warps / 4
        for (int job = warp >> 2; warp < WG_PROD_WARPS && job < njobs; job += WG_PROD_WARPS / 4) {
            const int mt = job / nslab, sl = job - mt * nslab;
            const int m = m_base + mt * 128 + qd * 32 + lane;    // M-side channel of this TMEM lane
            float r32[32];
            tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(mt * p.npad_per + sl * 32), r32);
            const int msize = p.swap ? p.K : p.N, csize = p.swap ? p.N : p.K;
            if (m < msize) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = c_base + sl * 32 + i;           // N-side channel of this column
                    if (c < csize && sl * 32 + i < p.npad_per)
                        atomicAdd(a.dw + (p.swap ? (size_t)c * p.K + m : (size_t)m * p.K + c), r32[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WG_PROD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem, p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------

static int wg_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// returns CF_OK when launched, -1 when the problem is not eligible (the caller runs the CUDA-core kernel)
int cf_pw_wgrad_tc(const cf_pw_wgrad_args* a, cudaStream_t stream) {
    if (cf_env("CFNET_PW_WGRAD_SIMT", 0)) return -1;
    const int N = a->N, K = a->K;
    if (a->dbias || (N & 1) || (K & 1) || N > 512 || K > 512) return -1;
    if (a->gather_in && !(a->g.kt == 1 && a->g.kh == 1 && a->g.kw == 1 && a->g.pt == 0 && a->g.ph == 0 && a->g.pw == 0 && a->g.ch_stride == 1 &&
                          a->g.pos_stride == K && a->x_mode == CF_PRO_NONE && ((a->g.sample_stride * 4) & 15) == 0))
        return -1;
    if (a->dy_mode != CF_PRO_NONE && a->dy_mode != CF_PRO_AFFINE2) return -1;
    if (a->x_mode == CF_PRO_AFFINE2) return -1;
    const long long R = (long long)a->g.T * a->g.H * a->g.W;
    if (R * a->B < 4096) return -1;                          // tiny problems: launch overhead dominates, keep the simple kernel
    WgParams p;
    p.B = a->B; p.R = (int)R; p.N = N; p.K = K;
    p.gmode = a->gather_in ? 1 : 0;
    p.gH = a->g.H; p.gW = a->g.W; p.gHi = a->g.Hi; p.gWi = a->g.Wi; p.gst = a->g.st; p.gsh = a->g.sh; p.gsw = a->g.sw;
    p.g_sample_stride = a->g.sample_stride;
    uintptr_t da = (uintptr_t)a->dy | (uintptr_t)(a->dy2 ? a->dy2 : a->dy);
    p.av_dy = ((N & 3) == 0 && (da & 15) == 0) ? 4 : 2;
    p.av_x = ((K & 3) == 0 && (((uintptr_t)a->x) & 15) == 0) ? 4 : 2;
    if ((da & 7) || (((uintptr_t)a->x) & 7)) return -1;
    // The tensor with more channels is the M side (128 per tile), the other the N side (32 per chunk): the tensor work of
    // one 8-row step is ceil(M/128)*128 x ceil32(N).  Then split so that the partial fits 512 TMEM columns and one MMA
    // covers <= 256 N-side channels.
    {
        const long long cost0 = (long long)((N + 127) / 128) * 128 * ((K + 31) / 32 * 32);
        const long long cost1 = (long long)((K + 127) / 128) * 128 * ((N + 31) / 32 * 32);
        p.swap = (cost1 < cost0 && !cf_env("CFNET_WG_NOSWAP", 0)) ? 1 : 0;
    }
    const int Msz = p.swap ? K : N, Csz = p.swap ? N : K;
    const int mtiles = (Msz + 127) / 128;
    const int cpad = (Csz + 31) / 32 * 32;
    p.nsplit = cpad > 256 ? 2 : 1;
    p.npad_per = ((cpad / 32 + p.nsplit - 1) / p.nsplit) * 32;
    p.msplit = 1;
    while (((mtiles + p.msplit - 1) / p.msplit) * p.npad_per > 512) ++p.msplit;
    p.mt_per = (mtiles + p.msplit - 1) / p.msplit;
    p.nchB = p.npad_per / 32;
    p.nchA_pad = 4 * p.mt_per;
    int nchA;                                                // M-side chunks a CTA actually loads
    {
        int nreal = (Msz + 31) / 32;
        nchA = nreal < p.nchA_pad ? nreal : p.nchA_pad;
        if (p.msplit > 1) nchA = p.nchA_pad;                 // (padded channels of the last tile load as zeros)
    }
    p.ndyc = p.swap ? p.nchB : nchA;
    p.nxc = p.swap ? nchA : p.nchB;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < p.mt_per * p.npad_per) p.tmem_cols <<= 1;
    const size_t tab_bytes = (size_t)(3 * p.ndyc + 2 * p.nxc) * 32 * 4;
    // ---- TMA-fed producers (dense rows; tensors describable by a tensor map): 2 operand stages + a ring of raw stages
    CUtensorMap tm_dy, tm_dy2, tm_x;
    memset(&tm_dy, 0, sizeof(tm_dy));
    memset(&tm_dy2, 0, sizeof(tm_dy2));
    memset(&tm_x, 0, sizeof(tm_x));
    const bool aff2 = a->dy_mode == CF_PRO_AFFINE2;
    p.tma = 0; p.fold_dy = p.fold_x = 1; p.nraw = 0; p.raw_dy_bytes = p.raw_x_bytes = p.raw_stage_bytes = p.raw_off = 0;
    bool planned = false;
    if (!p.gmode && cf_env("CFNET_WG_TMA", 1)) {
        const int fdy = (N % 4 == 0) ? 1 : 2, fx = (K % 4 == 0) ? 1 : 2;
        // 54-channel tensors (rows that are not 16-byte multiples) travel as folded whole-tile copies (no channel split then)
        const bool split = p.msplit * p.nsplit > 1;
        if ((fdy == 1 && fx == 1) || (!split && cf_env("CFNET_WG_TMA_FOLD", 1))) {
            for (int rb = 64; rb >= 16 && !planned; rb >>= 1) {
                const uint32_t chunk = (uint32_t)rb * 128u;
                const uint32_t stage = (uint32_t)(2 * (p.nchA_pad + p.nchB)) * chunk;
                const uint32_t rdy = fdy == 1 ? (uint32_t)p.ndyc * chunk : (uint32_t)rb * N * 4u;
                const uint32_t rx = fx == 1 ? (uint32_t)p.nxc * chunk : (uint32_t)rb * K * 4u;
                const uint32_t rstage = (aff2 ? 2u : 1u) * rdy + rx;
                const size_t used = 1024 + 2 * (size_t)stage + tab_bytes + 128;
                if (used + 3 * (size_t)rstage > WG_SMEM_MAX) continue;
                if (!cf_make_row_tmap(&tm_dy, a->dy, a->B, R, N, fdy, rb) || (aff2 && !cf_make_row_tmap(&tm_dy2, a->dy2, a->B, R, N, fdy, rb)) ||
                    !cf_make_row_tmap(&tm_x, a->x, a->B, R, K, fx, rb))
                    break;
                p.tma = 1; p.fold_dy = fdy; p.fold_x = fx;
                p.RB = rb; p.chunk_bytes = chunk; p.stage_bytes = stage; p.nstages = 2;
                p.raw_dy_bytes = rdy; p.raw_x_bytes = rx; p.raw_stage_bytes = rstage;
                p.nraw = (int)((WG_SMEM_MAX - used) / rstage);
                if (p.nraw > WG_MAX_RAW) p.nraw = WG_MAX_RAW;
                p.raw_off = (uint32_t)((2 * (size_t)stage + tab_bytes + 127) / 128 * 128);
                planned = true;
            }
        }
    }
    if (!planned) {
        int rb0 = cf_env("CFNET_WG_RB", 64);
        if (rb0 != 16 && rb0 != 32) rb0 = 64;
        p.RB = rb0;                                          // rows per stage: as many as leave >= 2 stages (per-stage overhead amortised)
        for (;;) {
            p.chunk_bytes = (uint32_t)p.RB * 128u;
            p.stage_bytes = (uint32_t)(2 * (p.nchA_pad + p.nchB)) * p.chunk_bytes;
            p.nstages = (int)((WG_SMEM_MAX - 1024 - tab_bytes) / p.stage_bytes);
            if (p.nstages >= 2 || p.RB == 16) break;
            p.RB >>= 1;
        }
        if (p.nstages < 2) return -1;
        if (p.nstages > WG_MAX_STAGES) p.nstages = WG_MAX_STAGES;
        // keep the shared-memory carve-out at <= 196 KB (some L1 left for the 8-byte loads of the 54-channel tensors: see x3d_pw_tc2.cu)
        while (p.nstages > 2 && 1024 + (size_t)p.nstages * p.stage_bytes + tab_bytes > 193 * 1024) --p.nstages;
    }
    {
        const uint32_t m_lo = (uint32_t)p.nchA_pad * p.chunk_bytes, c_off = 2u * m_lo, c_lo = (uint32_t)p.nchB * p.chunk_bytes;
        if (p.swap) { p.x_off = 0; p.x_lo = m_lo; p.dy_off = c_off; p.dy_lo = c_lo; }
        else { p.dy_off = 0; p.dy_lo = m_lo; p.x_off = c_off; p.x_lo = c_lo; }
    }
    p.rbps = (int)((R + p.RB - 1) / p.RB);
    p.total_items = (long long)a->B * p.rbps;
    int gx = wg_sm_count() / (p.msplit * p.nsplit);
    if (gx < 1) gx = 1;
    if (gx > p.total_items) gx = (int)p.total_items;
    p.items_per_cta = (int)((p.total_items + gx - 1) / gx);
    gx = (int)((p.total_items + p.items_per_cta - 1) / p.items_per_cta);
    const size_t smem = p.tma ? 1024 + (size_t)p.raw_off + (size_t)p.nraw * p.raw_stage_bytes
                              : 1024 + (size_t)p.nstages * p.stage_bytes + tab_bytes;
    dim3 grid((unsigned)gx, (unsigned)(p.msplit * p.nsplit));
    cudaError_t e = cudaSuccess;
#define WG_LAUNCH(DYM_, XM_)                                                                                                  \
    do {                                                                                                                      \
        static CfOncePerDevice attr_done;                                                                                        \
        if (attr_done.need()) {                                                                                                     \
            e = cudaFuncSetAttribute(pw_wgrad_tc_kernel<DYM_, XM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_MAX); \
            if (e == cudaSuccess) attr_done.mark();                                                                                     \
        }                                                                                                                     \
        if (e == cudaSuccess) cf_launch(pw_wgrad_tc_kernel<DYM_, XM_>, grid, WG_THREADS, smem, stream, *a, p, tm_dy, tm_dy2, tm_x);                     \
    } while (0)
#define WG_LAUNCH_X(DYM_)                                                              \
    switch (a->x_mode) {                                                               \
        case CF_PRO_NONE: WG_LAUNCH(DYM_, CF_PRO_NONE); break;                         \
        case CF_PRO_AFFINE: WG_LAUNCH(DYM_, CF_PRO_AFFINE); break;                     \
        case CF_PRO_AFFINE_RELU: WG_LAUNCH(DYM_, CF_PRO_AFFINE_RELU); break;           \
        default: WG_LAUNCH(DYM_, CF_PRO_AFFINE_SWISH); break;                          \
    }
    if (a->dy_mode == CF_PRO_AFFINE2) { WG_LAUNCH_X(CF_PRO_AFFINE2) } else { WG_LAUNCH_X(CF_PRO_NONE) }
#undef WG_LAUNCH_X
#undef WG_LAUNCH
    if (e != cudaSuccess) {
        cf_set_error("cf_pw_wgrad_tc: cannot opt in to %d B of shared memory: %s", WG_SMEM_MAX, cudaGetErrorString(e));
        return CF_ERR_CUDA;
    }
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
