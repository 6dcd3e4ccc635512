// Temporal depthwise 5x1x1 convolution of the stem (conv1_t, x3d_fine.py:216-222), channels-last:
// forward, data gradient and weight gradient as register-window streaming kernels.
//
// A thread owns one float4 (4 channels of one spatial position) of EVERY frame and marches along T with the five
// frames of its window in registers: every element is loaded exactly once with a perfectly coalesced 16-byte access
// (a frame is one contiguous [H*W*C] block, consecutive threads = consecutive float4s), no shared memory, no halo.
// The general direct kernels of x3d_dw.cu spent 12 ms of the 156 ms step here (tap loops with runtime geometry,
// 5 loads per output).
//
//   forward : y[t]  = sum_dt pro(x[t+dt-2]) * w[dt]                               (+ sum y, sum y^2 per sample/channel)
//   dgrad   : dx[t] = sum_dt d'[t-dt+2] * w[dt],   d' = P*dz + Q*y + R            (= forward with the flipped stencil)
//   wgrad   : dw[dt] += sum_t d'[t] * act(x[t+dt-2]),  act = relu(a*x+b) when tables are given
//   fused   : the data-gradient march also accumulates dw[dt] = sum_q act(x[q]) * d'[q-dt+2] from the d' window it holds
//             (one more load per step -- the forward input at the output position -- instead of a second pass over d', y)
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include <stdlib.h>

enum { T5_FWD = 0, T5_DGRAD = 1, T5_WGRAD = 2, T5_FUSED = 3 };   // FUSED: data gradient + weight gradient in one pass

struct T5Params {
    int B, C, T;
    long long F4;          // float4s per frame = H*W*C/4
    int tseg, ntseg;
};

__device__ __forceinline__ float4 t5_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 t5_fma(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

template <int MODE>
__global__ void __launch_bounds__(256) dwt5_kernel(const cf_dw_args a, const T5Params p) {
    cf_pdl_enter();
    __shared__ float red[20 * 256];
    const int tid = threadIdx.x;
    const long long e = (long long)blockIdx.x * 256 + tid;
    const bool live = e < p.F4;
    const int b = blockIdx.z;
    const int C = p.C, T = p.T, CQ = C >> 2;
    const int cq = (int)(e % CQ), c0 = cq * 4;
    const int t0 = blockIdx.y * p.tseg, t1 = min(T, t0 + p.tseg);
    const size_t tc = (size_t)b * C + c0;

    // per-channel constants
    float4 w[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int jj = (MODE == T5_DGRAD || MODE == T5_FUSED) ? 4 - j : j;    // dgrad: flipped stencil
        w[j] = (MODE == T5_WGRAD || !live) ? t5_zero()
                                           : make_float4(a.w[(size_t)(c0 + 0) * 5 + jj], a.w[(size_t)(c0 + 1) * 5 + jj],
                                                         a.w[(size_t)(c0 + 2) * 5 + jj], a.w[(size_t)(c0 + 3) * 5 + jj]);
    }
    auto tab4 = [&](const float* t, float dflt) {
        return (t && live) ? *reinterpret_cast<const float4*>(t + tc) : make_float4(dflt, dflt, dflt, dflt);
    };
    // window tensor: FWD a.x (pro NONE / AFFINE / AFFINE_RELU); DGRAD a.x (+ a.x2, AFFINE2); WGRAD a.aux (relu(a*x+b) if epi tables)
    const float* win = MODE == T5_WGRAD ? a.aux : a.x;
    const int win_mode = MODE == T5_WGRAD ? (a.epi_a ? CF_PRO_AFFINE_RELU : CF_PRO_NONE) : a.pro_mode;
    const float4 ea4 = tab4(a.epi_a, 1.f), eb4 = tab4(a.epi_b, 0.f);     // FUSED: activation tables of the forward input
    const float4 wa = MODE == T5_WGRAD ? tab4(a.epi_a, 1.f) : tab4(a.pro_a, 1.f);
    const float4 wb = MODE == T5_WGRAD ? tab4(a.epi_b, 0.f) : tab4(a.pro_b, 0.f);
    const float4 wc = MODE == T5_WGRAD ? t5_zero() : tab4(a.pro_c, 0.f);
    // WGRAD: d' tables
    const float4 da = tab4(a.pro_a, 1.f), db = tab4(a.pro_b, 0.f), dc = tab4(a.pro_c, 0.f);

    const size_t base = (size_t)b * T * p.F4 + e;                    // float4 index of frame 0
    auto load_win = [&](int t) {
        if (!live || t < 0 || t >= T) return t5_zero();              // zero padding applies AFTER the prologue
        float4 x = __ldg(reinterpret_cast<const float4*>(win) + base + (size_t)t * p.F4);
        if (win_mode == CF_PRO_AFFINE_RELU) {
            x.x = fmaxf(fmaf(wa.x, x.x, wb.x), 0.f); x.y = fmaxf(fmaf(wa.y, x.y, wb.y), 0.f);
            x.z = fmaxf(fmaf(wa.z, x.z, wb.z), 0.f); x.w = fmaxf(fmaf(wa.w, x.w, wb.w), 0.f);
        } else if (win_mode == CF_PRO_AFFINE) {
            x = t5_fma(wa, x, wb);
        } else if (win_mode == CF_PRO_AFFINE2) {
            const float4 x2 = __ldg(reinterpret_cast<const float4*>(a.x2) + base + (size_t)t * p.F4);
            x = t5_fma(wa, x, t5_fma(wb, x2, wc));
        }
        return x;
    };

    float4 xm2 = load_win(t0 - 2), xm1 = load_win(t0 - 1), x0 = load_win(t0), xp1 = load_win(t0 + 1);
    float4 s1 = t5_zero(), s2 = t5_zero();
    float4 acc[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) acc[j] = t5_zero();

    for (int t = t0; t < t1; ++t) {
        const float4 xp2 = load_win(t + 2);
        if (MODE != T5_WGRAD) {
            float4 y = make_float4(xm2.x * w[0].x, xm2.y * w[0].y, xm2.z * w[0].z, xm2.w * w[0].w);
            y = t5_fma(xm1, w[1], y);
            y = t5_fma(x0, w[2], y);
            y = t5_fma(xp1, w[3], y);
            y = t5_fma(xp2, w[4], y);
            if (live) {
                __stcs(reinterpret_cast<float4*>(a.y) + base + (size_t)t * p.F4, y);
                s1.x += y.x; s1.y += y.y; s1.z += y.z; s1.w += y.w;
                s2 = t5_fma(y, y, s2);
                if (MODE == T5_FUSED) {                            // dw[dt] += act(x[t]) * d'[t - dt + 2]
                    float4 xin = __ldg(reinterpret_cast<const float4*>(a.aux) + base + (size_t)t * p.F4);
                    if (a.epi_a) {
                        xin.x = fmaxf(fmaf(ea4.x, xin.x, eb4.x), 0.f); xin.y = fmaxf(fmaf(ea4.y, xin.y, eb4.y), 0.f);
                        xin.z = fmaxf(fmaf(ea4.z, xin.z, eb4.z), 0.f); xin.w = fmaxf(fmaf(ea4.w, xin.w, eb4.w), 0.f);
                    }
                    acc[0] = t5_fma(xin, xp2, acc[0]);
                    acc[1] = t5_fma(xin, xp1, acc[1]);
                    acc[2] = t5_fma(xin, x0, acc[2]);
                    acc[3] = t5_fma(xin, xm1, acc[3]);
                    acc[4] = t5_fma(xin, xm2, acc[4]);
                }
            }
        } else if (live) {
            float4 d = __ldg(reinterpret_cast<const float4*>(a.x) + base + (size_t)t * p.F4);
            if (a.pro_mode == CF_PRO_AFFINE2) {
                const float4 d2 = __ldg(reinterpret_cast<const float4*>(a.x2) + base + (size_t)t * p.F4);
                d = t5_fma(da, d, t5_fma(db, d2, dc));
            } else if (a.pro_mode != CF_PRO_NONE) {
                d = t5_fma(da, d, db);
            }
            acc[0] = t5_fma(d, xm2, acc[0]);
            acc[1] = t5_fma(d, xm1, acc[1]);
            acc[2] = t5_fma(d, x0, acc[2]);
            acc[3] = t5_fma(d, xp1, acc[3]);
            acc[4] = t5_fma(d, xp2, acc[4]);
        }
        xm2 = xm1; xm1 = x0; x0 = xp1; xp1 = xp2;
    }

    // ---- block reductions: threads with the same channel quad (e % CQ) share channels
    if (MODE == T5_WGRAD || MODE == T5_FUSED) {
        float* dwp = MODE == T5_FUSED ? a.dw_out : a.y;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            red[(j * 4 + 0) * 256 + tid] = acc[j].x; red[(j * 4 + 1) * 256 + tid] = acc[j].y;
            red[(j * 4 + 2) * 256 + tid] = acc[j].z; red[(j * 4 + 3) * 256 + tid] = acc[j].w;
        }
        __syncthreads();
        // slot v = j*4 + i (tap j, channel i of the quad); first thread of each quad residue sums its residue class
        for (int job = tid; job < 20 * CQ; job += 256) {
            const int v = job / CQ, q = job - v * CQ;                // q = channel quad
            // threads with (blockIdx.x*256 + tid') % CQ == q
            int start = (int)(((long long)q - ((long long)blockIdx.x * 256) % CQ + CQ) % CQ);
            float s = 0.f;
            for (int u = start; u < 256; u += CQ) s += red[v * 256 + u];
            const int j = v >> 2, i = v & 3;
            atomicAdd(dwp + (size_t)(q * 4 + i) * 5 + j, s);
        }
    } else if (a.stats_mode != CF_STATS_NONE) {
        red[0 * 256 + tid] = s1.x; red[1 * 256 + tid] = s1.y; red[2 * 256 + tid] = s1.z; red[3 * 256 + tid] = s1.w;
        red[4 * 256 + tid] = s2.x; red[5 * 256 + tid] = s2.y; red[6 * 256 + tid] = s2.z; red[7 * 256 + tid] = s2.w;
        __syncthreads();
        for (int job = tid; job < 8 * CQ; job += 256) {
            const int v = job / CQ, q = job - v * CQ;
            int start = (int)(((long long)q - ((long long)blockIdx.x * 256) % CQ + CQ) % CQ);
            float s = 0.f;
            for (int u = start; u < 256; u += CQ) s += red[v * 256 + u];
            const int which = v >> 2, i = v & 3;
            atomicAdd(a.stats + ((size_t)b * C + q * 4 + i) * 2 + which, (double)s);
        }
    }
}

// mode: 0 forward, 1 data gradient, 2 weight gradient, 3 data + weight gradient (a->dw_out).  Returns CF_OK when launched,
// -1 when not eligible.
int cf_dwt5_try(int mode, const cf_dw_args* a, cudaStream_t stream) {
    if (cf_env("CFNET_DWT5_OFF", 0)) return -1;
    const cf_geom& g = a->g;
    if (!(g.kt == 5 && g.kh == 1 && g.kw == 1 && g.pt == 2 && g.ph == 0 && g.pw == 0 && g.st == 1 && g.sh == 1 && g.sw == 1)) return -1;
    if (g.T != g.Ti || g.H != g.Hi || g.W != g.Wi) return -1;
    if (a->C % 4 != 0 || a->C > 1024) return -1;
    uintptr_t al = (uintptr_t)a->x | (uintptr_t)(a->x2 ? a->x2 : a->x) | (uintptr_t)(a->aux ? a->aux : a->x) | (uintptr_t)a->pro_a |
                   (uintptr_t)a->pro_b | (uintptr_t)a->pro_c | (uintptr_t)a->epi_a | (uintptr_t)a->epi_b;
    if (mode != T5_WGRAD) al |= (uintptr_t)a->y;
    if (al & 15) return -1;
    if (mode == T5_FWD && !(a->pro_mode == CF_PRO_NONE || a->pro_mode == CF_PRO_AFFINE || a->pro_mode == CF_PRO_AFFINE_RELU)) return -1;
    if (mode == T5_FWD && a->stats_mode == CF_STATS_SUM_AUX) return -1;
    if ((mode == T5_DGRAD || mode == T5_FUSED) && (a->epi_mode != CF_EPI_NONE || a->stats_mode != CF_STATS_NONE)) return -1;
    if (mode == T5_FUSED && !(a->dw_out && a->aux)) return -1;
    if (mode != T5_FWD && !(a->pro_mode == CF_PRO_NONE || a->pro_mode == CF_PRO_AFFINE || a->pro_mode == CF_PRO_AFFINE2)) return -1;
    T5Params p;
    p.B = a->B; p.C = a->C; p.T = g.T;
    p.F4 = (long long)g.H * g.W * a->C / 4;
    const long long blocks = (p.F4 + 255) / 256;
    long long want = (2LL * 148 * 8 + blocks * a->B - 1) / (blocks * a->B);      // >= 2 waves of 8 blocks per SM
    if (want < 1) want = 1;
    int tseg = (int)((g.T + want - 1) / want);
    if (tseg < 16) tseg = g.T < 16 ? g.T : 16;                     // four halo frames per segment
    p.tseg = tseg;
    p.ntseg = (g.T + tseg - 1) / tseg;
    if (blocks > 2147483647LL || p.ntseg > 65535 || a->B > 65535) return -1;
    dim3 grid((unsigned)blocks, (unsigned)p.ntseg, (unsigned)a->B);
    if (mode == T5_FWD) cf_launch(dwt5_kernel<T5_FWD>, grid, 256, 0, stream, *a, p);
    else if (mode == T5_DGRAD) cf_launch(dwt5_kernel<T5_DGRAD>, grid, 256, 0, stream, *a, p);
    else if (mode == T5_FUSED) cf_launch(dwt5_kernel<T5_FUSED>, grid, 256, 0, stream, *a, p);
    else cf_launch(dwt5_kernel<T5_WGRAD>, grid, 256, 0, stream, *a, p);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
