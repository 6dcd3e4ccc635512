// Depthwise 3x3x3 convolution with spatial stride 2 (conv2 of the first Bottleneck of every stage,
// x3d_fine.py:89-97 with stride (1,2,2), :277-288): forward and weight gradient as plane-marching kernels.
//
// A strided window needs 4 input positions per output, so a ring of three haloed INPUT planes (the stride-1 scheme of
// x3d_dw3.cu) does not fit shared memory at a useful tile size.  These kernels are input-stationary instead: one
// haloed input plane (9 x 29 positions x 54 channels for a 4 x 14 output tile) is resident at a time, the next one
// arrives with cp.async, and every thread (1 x 7 output patch, 2 channels) keeps THREE partial output planes in
// registers: input frame ti adds its dt = 0/1/2 taps to output frames ti+1 / ti / ti-1; frame ti-1 is then complete
// and stored.  Each input element is read from global memory once (+ 16 % halo) and from shared memory once.
//
//   forward : y[t,h,w,c] = sum_taps relu(a*x+b)[t+dt-1, 2h+dh-1, 2w+dw-1, c] * w[c,dt,dh,dw]       (+ sum y, sum y^2)
//   wgrad   : dw[c,dt,dh,dw] += sum_out d'[t,h,w,c] * relu(a1*y1+b1)[t+dt-1, 2h+dh-1, 2w+dw-1, c],  d' = P*dU + Q*y2 + R
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include <stdlib.h>

#define S2_LANES 27
#define S2_CS 54
#define S2_PW 7
#define S2_OTH 4                      /* output rows per tile (one per patch row) */
#define S2_IH (2 * S2_OTH + 1)

enum { S2_FWD = 0, S2_WGRAD = 2 };

struct S2Params {
    int B, C, T, Hi, Wi, Ho, Wo;
    int htiles, wtiles, slabs;
    long long total_steps;
    int steps_per_cta;
};

__device__ __forceinline__ void s2_cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void s2_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int MODE, int NPW>
__global__ void __launch_bounds__(S2_LANES * S2_OTH * NPW) dw3s2_kernel(const cf_dw_args a, const S2Params p) {
    cf_pdl_enter();
    constexpr int OTW = S2_PW * NPW, IW = 2 * OTW + 1;
    constexpr int NPATCH = S2_OTH * NPW, NT = S2_LANES * NPATCH;
    constexpr int NPOS = S2_IH * IW;
    constexpr int PLANE = NPOS * S2_CS;
    constexpr int KPT = (NPOS + NPATCH - 1) / NPATCH;
    extern __shared__ __align__(16) float sm[];
    float* cur = sm;                                               // activated plane being consumed
    float* stg = cur + PLANE;                                      // raw plane in flight
    float* tabs = stg + PLANE;                                     // [5][54]
    float* ws = tabs + 5 * S2_CS;                                  // [27][54] (forward)

    const int tid = threadIdx.x;
    const int patch = tid / S2_LANES, lane = tid - patch * S2_LANES;
    const int pr = patch / NPW, pc = patch - pr * NPW;             // output row / 7-wide column group inside the tile
    const int C = p.C, T = p.T, Hi = p.Hi, Wi = p.Wi, Ho = p.Ho, Wo = p.Wo;
    const float* src = MODE == S2_WGRAD ? a.aux : a.x;             // input-resolution tensor behind the plane
    const int plane_mode = MODE == S2_WGRAD ? (a.epi_a ? CF_PRO_AFFINE_RELU : CF_PRO_NONE) : a.pro_mode;

    const long long step0 = (long long)blockIdx.x * p.steps_per_cta;
    const long long step1 = min(step0 + p.steps_per_cta, p.total_steps);
    for (long long step = step0; step < step1;) {
        const int col = (int)(step / T);
        const int t0 = (int)(step - (long long)col * T);
        const int t1 = (int)min((long long)T, t0 + (step1 - step));
        step += t1 - t0;
        int bx = col;
        const int slab = bx % p.slabs; bx /= p.slabs;
        const int tw_i = bx % p.wtiles; bx /= p.wtiles;
        const int th_i = bx % p.htiles;
        const int b = bx / p.htiles;
        const int ho0 = th_i * S2_OTH, wo0 = tw_i * OTW;
        const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 1;            // image coordinates of haloed position (0,0)
        const int cs0 = slab * S2_CS, c0 = cs0 + lane * 2;
        __syncthreads();

        for (int i = tid; i < S2_CS; i += NT) {
            const size_t tc = (size_t)b * C + cs0 + i;
            float ra = 1.f, rb = 0.f, rc = 0.f, ea = 1.f, eb = 0.f;
            if (MODE == S2_WGRAD) {
                if (a.epi_a) { ra = a.epi_a[tc]; rb = a.epi_b[tc]; }
                if (a.pro_mode != CF_PRO_NONE) {
                    ea = a.pro_a[tc];
                    eb = a.pro_b ? a.pro_b[tc] : 0.f;
                    rc = a.pro_c ? a.pro_c[tc] : 0.f;
                }
            } else if (a.pro_mode != CF_PRO_NONE) {
                ra = a.pro_a[tc];
                rb = a.pro_b ? a.pro_b[tc] : 0.f;
            }
            tabs[i] = ra; tabs[S2_CS + i] = rb; tabs[2 * S2_CS + i] = rc; tabs[3 * S2_CS + i] = ea; tabs[4 * S2_CS + i] = eb;
        }
        if (MODE == S2_FWD)
            for (int i = tid; i < 27 * S2_CS; i += NT) {
                const int tap = i / S2_CS, c = i - tap * S2_CS;
                ws[i] = a.w[(size_t)(cs0 + c) * 27 + tap];
            }
        __syncthreads();
        const float2 ra = *reinterpret_cast<const float2*>(tabs + lane * 2);
        const float2 rb = *reinterpret_cast<const float2*>(tabs + S2_CS + lane * 2);
        const float2 rc = *reinterpret_cast<const float2*>(tabs + 2 * S2_CS + lane * 2);
        const float2 ea = *reinterpret_cast<const float2*>(tabs + 3 * S2_CS + lane * 2);
        const float2 eb = *reinterpret_cast<const float2*>(tabs + 4 * S2_CS + lane * 2);

        // plane movement: positions patch + k * NPATCH, own channel pair; validity / offsets are frame-invariant
        uint64_t vmask = 0;
        int goff[KPT];
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            const int pos = patch + k * NPATCH;
            const int hh = pos / IW, ww = pos - hh * IW;
            const int h = hi0 + hh, w = wi0 + ww;
            const bool v = pos < NPOS && (unsigned)h < (unsigned)Hi && (unsigned)w < (unsigned)Wi;
            goff[k] = v ? (h * Wi + w) * C + c0 : 0;
            vmask |= v ? (1ull << k) : 0ull;
        }
        const int soff0 = patch * S2_CS + lane * 2;
        const size_t frame = (size_t)Hi * Wi * C;
        auto plane_issue = [&](int t) {
            if (t < 0 || t >= T) return;
            const float* f0 = src + ((size_t)b * T + t) * frame;
#pragma unroll
            for (int k = 0; k < KPT; ++k)
                if (vmask & (1ull << k)) s2_cp_async8(stg + soff0 + k * (NPATCH * S2_CS), f0 + goff[k]);
        };
        auto plane_land = [&](int t) {
            const uint64_t m = (t >= 0 && t < T) ? vmask : 0ull;
#pragma unroll
            for (int k = 0; k < KPT; ++k) {
                if (patch + k * NPATCH >= NPOS) break;
                float2 v = make_float2(0.f, 0.f);
                if (m & (1ull << k)) {
                    const float2 x = *reinterpret_cast<const float2*>(stg + soff0 + k * (NPATCH * S2_CS));
                    if (plane_mode == CF_PRO_AFFINE_RELU) {
                        v.x = fmaxf(fmaf(ra.x, x.x, rb.x), 0.f);
                        v.y = fmaxf(fmaf(ra.y, x.y, rb.y), 0.f);
                    } else if (plane_mode == CF_PRO_AFFINE) {
                        v.x = fmaf(ra.x, x.x, rb.x);
                        v.y = fmaf(ra.y, x.y, rb.y);
                    } else {
                        v = x;
                    }
                }
                *reinterpret_cast<float2*>(cur + soff0 + k * (NPATCH * S2_CS)) = v;
            }
        };

        // per-thread state
        float2 wreg[27];                                           // forward: weights; wgrad: accumulators
#pragma unroll
        for (int i = 0; i < 27; ++i)
            wreg[i] = MODE == S2_WGRAD ? make_float2(0.f, 0.f) : *reinterpret_cast<const float2*>(ws + i * S2_CS + lane * 2);
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
        // forward: partial output frames  accA = out(ti-1), accB = out(ti), accC = out(ti+1)
        // wgrad  : d' of the patch at      dA  = out(ti-1), dB   = out(ti), dC   = out(ti+1), dN = prefetched out(ti+2)
        float2 pA[S2_PW], pB[S2_PW], pC[S2_PW], dN[S2_PW], dN2[S2_PW];
#pragma unroll
        for (int j = 0; j < S2_PW; ++j) pA[j] = pB[j] = pC[j] = dN[j] = dN2[j] = make_float2(0.f, 0.f);
        const int ho = ho0 + pr;                                   // this thread's output row
        const bool row_ok = ho < Ho;
        const float* pthr = cur + ((2 * pr) * IW + 2 * pc * S2_PW) * S2_CS + lane * 2;   // haloed origin of the patch's window
        auto out_base = [&](int t) { return ((((size_t)b * T + t) * Ho + ho) * Wo + wo0 + pc * S2_PW) * C + c0; };
        auto d_fetch = [&](int t) {                                // raw output gradient of frame t (zeros outside the piece)
            const bool v = row_ok && t >= t0 && t < t1;
            const size_t g0 = v ? out_base(t) : 0;
#pragma unroll
            for (int j = 0; j < S2_PW; ++j) {
                dN[j] = v ? __ldg(reinterpret_cast<const float2*>(a.x + g0 + (size_t)j * C)) : make_float2(0.f, 0.f);
                if (a.pro_mode == CF_PRO_AFFINE2) dN2[j] = v ? __ldg(reinterpret_cast<const float2*>(a.x2 + g0 + (size_t)j * C)) : make_float2(0.f, 0.f);
            }
        };
        auto d_apply = [&](int t, float2 (&d)[S2_PW]) {            // d' = P*dU + Q*y2 + R, zero outside the piece / image
            const bool v = row_ok && t >= t0 && t < t1;
#pragma unroll
            for (int j = 0; j < S2_PW; ++j) {
                float2 x = dN[j];
                if (a.pro_mode == CF_PRO_AFFINE2) {
                    x.x = fmaf(ea.x, x.x, fmaf(eb.x, dN2[j].x, rc.x));
                    x.y = fmaf(ea.y, x.y, fmaf(eb.y, dN2[j].y, rc.y));
                } else if (a.pro_mode != CF_PRO_NONE) {
                    x.x = fmaf(ea.x, x.x, eb.x);
                    x.y = fmaf(ea.y, x.y, eb.y);
                }
                d[j] = v ? x : make_float2(0.f, 0.f);
            }
        };

        // prologue: input frame t0-1 resident, t0 in flight; wgrad: d'(t0) ready in pC, d'(t0+1) prefetched
        plane_issue(t0 - 1);
        s2_cp_async_wait_all();
        plane_land(t0 - 1);
        plane_issue(t0);
        if (MODE == S2_WGRAD) {
            d_fetch(t0);
            d_apply(t0, pC);
            d_fetch(t0 + 1);
        }
        __syncthreads();

        for (int ti = t0 - 1; ti <= t1; ++ti) {
            // ---- consume input frame ti (zero frames contribute nothing)
            if (ti >= 0 && ti < T) {
#pragma unroll
                for (int dh = 0; dh < 3; ++dh) {
                    float2 in[2 * S2_PW + 1];
#pragma unroll
                    for (int j = 0; j < 2 * S2_PW + 1; ++j) in[j] = *reinterpret_cast<const float2*>(pthr + (dh * IW + j) * S2_CS);
#pragma unroll
                    for (int dw = 0; dw < 3; ++dw) {
                        if (MODE == S2_FWD) {
                            const float2 w0 = wreg[(0 * 3 + dh) * 3 + dw], w1 = wreg[(1 * 3 + dh) * 3 + dw], w2 = wreg[(2 * 3 + dh) * 3 + dw];
#pragma unroll
                            for (int j = 0; j < S2_PW; ++j) {
                                const float2 x = in[2 * j + dw];
                                ffma2(pC[j], x, w0);      // dt = 0 -> out(ti+1)
                                ffma2(pB[j], x, w1);      // dt = 1 -> out(ti)
                                ffma2(pA[j], x, w2);      // dt = 2 -> out(ti-1)
                            }
                        } else {
                            float2 g0 = wreg[(0 * 3 + dh) * 3 + dw], g1 = wreg[(1 * 3 + dh) * 3 + dw], g2 = wreg[(2 * 3 + dh) * 3 + dw];
#pragma unroll
                            for (int j = 0; j < S2_PW; ++j) {
                                const float2 x = in[2 * j + dw];
                                ffma2(g0, pC[j], x);
                                ffma2(g1, pB[j], x);
                                ffma2(g2, pA[j], x);
                            }
                            wreg[(0 * 3 + dh) * 3 + dw] = g0; wreg[(1 * 3 + dh) * 3 + dw] = g1; wreg[(2 * 3 + dh) * 3 + dw] = g2;
                        }
                    }
                }
            }
            // ---- forward: output frame ti-1 is complete
            if (MODE == S2_FWD && row_ok && ti - 1 >= t0 && ti - 1 < t1) {
                const size_t g0 = out_base(ti - 1);
#pragma unroll
                for (int j = 0; j < S2_PW; ++j) {
                    const float2 v = pA[j];
                    *reinterpret_cast<float2*>(a.y + g0 + (size_t)j * C) = v;
                    s1.x += v.x; s1.y += v.y;
                    s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y);
                }
            }
            // ---- rotate: (A, B, C) <- (B, C, next)
#pragma unroll
            for (int j = 0; j < S2_PW; ++j) { pA[j] = pB[j]; pB[j] = pC[j]; pC[j] = make_float2(0.f, 0.f); }
            if (MODE == S2_WGRAD) {
                d_apply(ti + 2, pC);                               // dN holds raw out(ti+2)
                d_fetch(ti + 3);
            }
            __syncthreads();                                       // everyone is done with the resident frame
            if (ti < t1) {
                s2_cp_async_wait_all();
                plane_land(ti + 1);
                __syncthreads();
                if (ti + 2 <= t1) plane_issue(ti + 2);
            }
        }
        s2_cp_async_wait_all();
        __syncthreads();

        // ---- CTA reductions over the patches, then global atomics (the plane buffers are free)
        float* red = cur;
        if (MODE == S2_WGRAD) {
#pragma unroll
            for (int i = 0; i < 27; ++i) *reinterpret_cast<float2*>(red + ((size_t)patch * 27 + i) * S2_CS + lane * 2) = wreg[i];
            __syncthreads();
            for (int i = tid; i < 27 * S2_CS; i += NT) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < NPATCH; ++q) s += red[(size_t)q * 27 * S2_CS + i];
                const int tap = i / S2_CS, c = i - tap * S2_CS;
                atomicAdd(a.y + (size_t)(cs0 + c) * 27 + tap, s);
            }
        } else if (a.stats_mode != CF_STATS_NONE) {
            *reinterpret_cast<float2*>(red + patch * 2 * S2_CS + lane * 2) = s1;
            *reinterpret_cast<float2*>(red + (patch * 2 + 1) * S2_CS + lane * 2) = s2;
            __syncthreads();
            for (int i = tid; i < 2 * S2_CS; i += NT) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < NPATCH; ++q) s += red[q * 2 * S2_CS + i];
                const int which = i / S2_CS, c = i - which * S2_CS;
                atomicAdd(a.stats + ((size_t)b * C + cs0 + c) * 2 + which, (double)s);
            }
        }
    }   // pieces
}

// ---------------------------------------------------------------------------------------
// Data gradient of the stride-2 convolution.  The small tensor is the OUTPUT gradient here, so this one is
// output-centric like x3d_dw3.cu: a ring of three d' planes at output resolution (5 x 15 positions: the 4 x 14 tile
// plus one row / column towards +h / +w, the only neighbours a transposed stride-2 window reaches) and each thread
// produces the 2 x 14 INPUT positions under its 1 x 7 output patch.  Which taps reach an input position depends only
// on its (row, column) parity: 1, 2, 2 or 4 of the 9 spatial taps (27 / 4 per position on average).
//   dz[ti, 2ho+a, 2wo+b, c] = [a1*y1+b1 > 0] * sum_dt sum_{(dh,ho') in S_a} sum_{(dw,wo') in S_b} d'[ti+1-dt, ho', wo'] w[dt,dh,dw]
//   S_0 = {(1, o)},  S_1 = {(0, o+1), (2, o)}
// ---------------------------------------------------------------------------------------
// FUSED: the same pass also produces the weight gradient.  Every (output position p, tap) pair of
//   dw[c,tap] = sum_p d'[p] * relu(a1*y1+b1)[2p + tap - 1]
// is visited exactly once by the gather form above -- as the term d'[p] * w[tap] of the input position q = 2p + tap - 1 -- so
// the weight gradient is 27 more accumulators fed by the d' values already in registers times the activation at q (the
// `aux` row the ReLU mask needs anyway; aux is re-read from L1/L2 for the epilogue, its registers hold the activations).
// The weights move to shared memory (their registers hold the accumulators).
template <int NPW, bool FUSED>
__global__ void __launch_bounds__(S2_LANES * S2_OTH * NPW) dw3s2_dgrad_kernel(const cf_dw_args a, const S2Params p) {
    cf_pdl_enter();
    constexpr int OTW = S2_PW * NPW, RW = OTW + 1, RH = S2_OTH + 1;
    constexpr int NPATCH = S2_OTH * NPW, NT = S2_LANES * NPATCH;
    constexpr int NPOS = RH * RW;
    constexpr int PLANE = NPOS * S2_CS;
    constexpr int KPT = (NPOS + NPATCH - 1) / NPATCH;
    extern __shared__ __align__(16) float sm[];
    float* ring = sm;                                              // [3][PLANE]
    float* stg0 = ring + 3 * PLANE;
    float* stg1 = stg0 + PLANE;
    float* tabs = stg1 + PLANE;                                    // [5][54]
    float* ws = tabs + 5 * S2_CS;                                  // [27][54]

    const int tid = threadIdx.x;
    const int patch = tid / S2_LANES, lane = tid - patch * S2_LANES;
    const int pr = patch / NPW, pc = patch - pr * NPW;
    const int C = p.C, T = p.T, Hi = p.Hi, Wi = p.Wi, Ho = p.Ho, Wo = p.Wo;
    const bool two_src = a.pro_mode == CF_PRO_AFFINE2;
    const bool drelu = a.epi_mode == CF_EPI_DRELU;
    const bool need_aux = drelu || a.stats_mode == CF_STATS_SUM_AUX;

    const long long step0 = (long long)blockIdx.x * p.steps_per_cta;
    const long long step1 = min(step0 + p.steps_per_cta, p.total_steps);
    for (long long step = step0; step < step1;) {
        const int col = (int)(step / T);
        const int t0 = (int)(step - (long long)col * T);
        const int t1 = (int)min((long long)T, t0 + (step1 - step));
        step += t1 - t0;
        int bx = col;
        const int slab = bx % p.slabs; bx /= p.slabs;
        const int tw_i = bx % p.wtiles; bx /= p.wtiles;
        const int th_i = bx % p.htiles;
        const int b = bx / p.htiles;
        const int ho0 = th_i * S2_OTH, wo0 = tw_i * OTW;
        const int cs0 = slab * S2_CS, c0 = cs0 + lane * 2;
        __syncthreads();

        for (int i = tid; i < S2_CS; i += NT) {
            const size_t tc = (size_t)b * C + cs0 + i;
            float ra = 1.f, rb = 0.f, rc = 0.f, ea = 1.f, eb = 0.f;
            if (a.pro_mode != CF_PRO_NONE) {
                ra = a.pro_a[tc];
                rb = a.pro_b ? a.pro_b[tc] : 0.f;
                rc = a.pro_c ? a.pro_c[tc] : 0.f;
            }
            if (drelu) { ea = a.epi_a[tc]; eb = a.epi_b[tc]; }
            tabs[i] = ra; tabs[S2_CS + i] = rb; tabs[2 * S2_CS + i] = rc; tabs[3 * S2_CS + i] = ea; tabs[4 * S2_CS + i] = eb;
        }
        for (int i = tid; i < 27 * S2_CS; i += NT) {
            const int tap = i / S2_CS, c = i - tap * S2_CS;
            ws[i] = a.w[(size_t)(cs0 + c) * 27 + tap];
        }
        __syncthreads();
        const float2 ra = *reinterpret_cast<const float2*>(tabs + lane * 2);
        const float2 rb = *reinterpret_cast<const float2*>(tabs + S2_CS + lane * 2);
        const float2 rc = *reinterpret_cast<const float2*>(tabs + 2 * S2_CS + lane * 2);
        const float2 ea = *reinterpret_cast<const float2*>(tabs + 3 * S2_CS + lane * 2);
        const float2 eb = *reinterpret_cast<const float2*>(tabs + 4 * S2_CS + lane * 2);

        uint32_t vmask = 0;
        int goff[KPT];
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            const int pos = patch + k * NPATCH;
            const int hh = pos / RW, ww = pos - hh * RW;
            const int h = ho0 + hh, w = wo0 + ww;
            const bool v = pos < NPOS && h < Ho && w < Wo;
            goff[k] = v ? (h * Wo + w) * C + c0 : 0;
            vmask |= v ? (1u << k) : 0u;
        }
        const int soff0 = patch * S2_CS + lane * 2;
        const size_t oframe = (size_t)Ho * Wo * C;
        auto plane_issue = [&](int t) {
            if (t < 0 || t >= T) return;
            const float* f0 = a.x + ((size_t)b * T + t) * oframe;
            const float* f1 = two_src ? a.x2 + ((size_t)b * T + t) * oframe : nullptr;
#pragma unroll
            for (int k = 0; k < KPT; ++k)
                if (vmask & (1u << k)) {
                    s2_cp_async8(stg0 + soff0 + k * (NPATCH * S2_CS), f0 + goff[k]);
                    if (f1) s2_cp_async8(stg1 + soff0 + k * (NPATCH * S2_CS), f1 + goff[k]);
                }
        };
        auto plane_land = [&](int t, float* dst) {
            const uint32_t m = (t >= 0 && t < T) ? vmask : 0u;
#pragma unroll
            for (int k = 0; k < KPT; ++k) {
                if (patch + k * NPATCH >= NPOS) break;
                float2 v = make_float2(0.f, 0.f);
                if (m & (1u << k)) {
                    const float2 x = *reinterpret_cast<const float2*>(stg0 + soff0 + k * (NPATCH * S2_CS));
                    if (two_src) {
                        const float2 x2 = *reinterpret_cast<const float2*>(stg1 + soff0 + k * (NPATCH * S2_CS));
                        v.x = fmaf(ra.x, x.x, fmaf(rb.x, x2.x, rc.x));
                        v.y = fmaf(ra.y, x.y, fmaf(rb.y, x2.y, rc.y));
                    } else if (a.pro_mode != CF_PRO_NONE) {
                        v.x = fmaf(ra.x, x.x, rb.x);
                        v.y = fmaf(ra.y, x.y, rb.y);
                    } else {
                        v = x;
                    }
                }
                *reinterpret_cast<float2*>(dst + soff0 + k * (NPATCH * S2_CS)) = v;
            }
        };
        auto slot = [&](int t) { return ring + ((t - t0 + 1) % 3) * PLANE; };

        float2 wreg[27];                                           // the weights; FUSED: the weight-gradient accumulators
#pragma unroll
        for (int i = 0; i < 27; ++i) wreg[i] = FUSED ? make_float2(0.f, 0.f) : *reinterpret_cast<const float2*>(ws + i * S2_CS + lane * 2);
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
        const int hi_base = 2 * (ho0 + pr), wi_base = 2 * (wo0 + pc * S2_PW);
        const float* pthr = ring + (pr * RW + pc * S2_PW) * S2_CS + lane * 2;

        plane_issue(t0 - 1);
        s2_cp_async_wait_all();
        plane_land(t0 - 1, slot(t0 - 1));
        plane_issue(t0);
        s2_cp_async_wait_all();
        plane_land(t0, slot(t0));
        plane_issue(t0 + 1);

        for (int ti = t0; ti < t1; ++ti) {
            s2_cp_async_wait_all();
            plane_land(ti + 1, slot(ti + 1));
            __syncthreads();
            if (ti + 2 <= t1) plane_issue(ti + 2);

            const int sl0 = (ti - t0) % 3;                         // slot of plane ti-1
            float2 aux[2][2 * S2_PW];
            if (need_aux) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int j = 0; j < 2 * S2_PW; ++j) {
                        const bool v = hi_base + r < Hi && wi_base + j < Wi;
                        aux[r][j] = v ? __ldg(reinterpret_cast<const float2*>(
                                            a.aux + ((((size_t)b * T + ti) * Hi + hi_base + r) * Wi + wi_base + j) * C + c0))
                                      : make_float2(0.f, 0.f);
                    }
            }
            float2 acc[2][2 * S2_PW];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < 2 * S2_PW; ++j) {
                    acc[r][j] = make_float2(0.f, 0.f);
                    if (FUSED) {                                   // aux -> activation relu(ea*aux + eb) (0 outside the image)
                        const bool v = hi_base + r < Hi && wi_base + j < Wi;
                        aux[r][j] = make_float2(v ? fmaxf(fmaf(ea.x, aux[r][j].x, eb.x), 0.f) : 0.f,
                                                v ? fmaxf(fmaf(ea.y, aux[r][j].y, eb.y), 0.f) : 0.f);
                    }
                }
#pragma unroll
            for (int dt = 0; dt < 3; ++dt) {
                // tap dt reads d' frame ti + 1 - dt: dt = 0 -> ti+1 (slot sl0+2), 1 -> ti (sl0+1), 2 -> ti-1 (sl0)
                const float* pl = pthr + ((sl0 + 2 - dt) % 3) * PLANE;
                float2 in0[S2_PW + 1], in1[S2_PW + 1];
#pragma unroll
                for (int j = 0; j < S2_PW + 1; ++j) {
                    in0[j] = *reinterpret_cast<const float2*>(pl + j * S2_CS);
                    in1[j] = *reinterpret_cast<const float2*>(pl + (RW + j) * S2_CS);
                }
                float2 wl[9];                                      // this dt's 9 spatial taps, w[dh * 3 + dw]
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    wl[i] = FUSED ? *reinterpret_cast<const float2*>(ws + (dt * 9 + i) * S2_CS + lane * 2) : wreg[dt * 9 + i];
                float2* g = wreg + dt * 9;                         // FUSED: weight-gradient accumulators of these taps
                // data gradient: ACC += X * w[tap];  weight gradient (FUSED): g[tap] += X * activation at ACC's position
#define S2_FMA(R, J, X, TAP) do { ffma2(acc[R][J], X, wl[TAP]); if (FUSED) ffma2(g[TAP], X, aux[R][J]); } while (0)
#pragma unroll
                for (int j = 0; j < S2_PW; ++j) {
                    S2_FMA(0, 2 * j, in0[j], 1 * 3 + 1);                                           // even row, even col
                    S2_FMA(0, 2 * j + 1, in0[j + 1], 1 * 3 + 0);                                   // even row, odd col
                    S2_FMA(0, 2 * j + 1, in0[j], 1 * 3 + 2);
                    S2_FMA(1, 2 * j, in1[j], 0 * 3 + 1);                                           // odd row, even col
                    S2_FMA(1, 2 * j, in0[j], 2 * 3 + 1);
                    S2_FMA(1, 2 * j + 1, in1[j + 1], 0 * 3 + 0);                                   // odd row, odd col
                    S2_FMA(1, 2 * j + 1, in1[j], 0 * 3 + 2);
                    S2_FMA(1, 2 * j + 1, in0[j + 1], 2 * 3 + 0);
                    S2_FMA(1, 2 * j + 1, in0[j], 2 * 3 + 2);
                }
#undef S2_FMA
            }
            if (FUSED && need_aux) {                               // the raw pre-activation again, for the mask and the statistics
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int j = 0; j < 2 * S2_PW; ++j) {
                        const bool v = hi_base + r < Hi && wi_base + j < Wi;
                        aux[r][j] = v ? __ldg(reinterpret_cast<const float2*>(
                                            a.aux + ((((size_t)b * T + ti) * Hi + hi_base + r) * Wi + wi_base + j) * C + c0))
                                      : make_float2(0.f, 0.f);
                    }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (hi_base + r >= Hi) continue;
#pragma unroll
                for (int j = 0; j < 2 * S2_PW; ++j) {
                    if (wi_base + j >= Wi) continue;
                    float2 v = acc[r][j];
                    if (drelu) {
                        v.x = fmaf(ea.x, aux[r][j].x, eb.x) > 0.f ? v.x : 0.f;
                        v.y = fmaf(ea.y, aux[r][j].y, eb.y) > 0.f ? v.y : 0.f;
                    }
                    *reinterpret_cast<float2*>(a.y + ((((size_t)b * T + ti) * Hi + hi_base + r) * Wi + wi_base + j) * C + c0) = v;
                    s1.x += v.x; s1.y += v.y;
                    if (need_aux) { s2.x = fmaf(v.x, aux[r][j].x, s2.x); s2.y = fmaf(v.y, aux[r][j].y, s2.y); }
                }
            }
            __syncthreads();
        }
        s2_cp_async_wait_all();

        if (FUSED) {                                               // CTA reduction of the weight-gradient partials, then atomics
            float* red = ring;
#pragma unroll
            for (int i = 0; i < 27; ++i) *reinterpret_cast<float2*>(red + ((size_t)patch * 27 + i) * S2_CS + lane * 2) = wreg[i];
            __syncthreads();
            for (int i = tid; i < 27 * S2_CS; i += NT) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < NPATCH; ++q) s += red[(size_t)q * 27 * S2_CS + i];
                const int tap = i / S2_CS, c = i - tap * S2_CS;
                atomicAdd(a.dw_out + (size_t)(cs0 + c) * 27 + tap, s);
            }
            __syncthreads();
        }
        if (a.stats_mode != CF_STATS_NONE) {
            float* red = ring;
            *reinterpret_cast<float2*>(red + patch * 2 * S2_CS + lane * 2) = s1;
            *reinterpret_cast<float2*>(red + (patch * 2 + 1) * S2_CS + lane * 2) = s2;
            __syncthreads();
            for (int i = tid; i < 2 * S2_CS; i += NT) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < NPATCH; ++q) s += red[q * 2 * S2_CS + i];
                const int which = i / S2_CS, c = i - which * S2_CS;
                atomicAdd(a.stats + ((size_t)b * C + cs0 + c) * 2 + which, (double)s);
            }
        }
    }   // pieces
}

template <int NPW, bool FUSED>
static int s2_launch_dgrad(const cf_dw_args* a, const S2Params& p, cudaStream_t stream) {
    constexpr int OTW = S2_PW * NPW;
    constexpr int PLANE = (S2_OTH + 1) * (OTW + 1) * S2_CS;
    const size_t smem = (size_t)(5 * PLANE + 5 * S2_CS + 27 * S2_CS) * sizeof(float);
    static CfOncePerDevice done;
    if (done.need()) {
        cudaError_t e = cudaFuncSetAttribute(dw3s2_dgrad_kernel<NPW, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) { cf_set_error("dw3s2: cannot opt in to shared memory: %s", cudaGetErrorString(e)); return CF_ERR_CUDA; }
        done.mark();
    }
    dim3 grid((unsigned)((p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta));
    cf_launch(dw3s2_dgrad_kernel<NPW, FUSED>, grid, S2_LANES * S2_OTH * NPW, smem, stream, *a, p);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

// ---------------------------------------------------------------------------------------
template <int MODE, int NPW>
static int s2_launch(const cf_dw_args* a, const S2Params& p, cudaStream_t stream) {
    constexpr int OTW = S2_PW * NPW, IW = 2 * OTW + 1;
    constexpr int PLANE = S2_IH * IW * S2_CS;
    const size_t smem = (size_t)(2 * PLANE + 5 * S2_CS + 27 * S2_CS) * sizeof(float);
    static CfOncePerDevice done;
    if (done.need()) {
        cudaError_t e = cudaFuncSetAttribute(dw3s2_kernel<MODE, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) { cf_set_error("dw3s2: cannot opt in to shared memory: %s", cudaGetErrorString(e)); return CF_ERR_CUDA; }
        done.mark();
    }
    dim3 grid((unsigned)((p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta));
    cf_launch(dw3s2_kernel<MODE, NPW>, grid, S2_LANES * S2_OTH * NPW, smem, stream, *a, p);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

// mode: 0 forward, 1 data gradient, 2 weight gradient, 3 data + weight gradient in one pass (a->dw_out).
// Returns CF_OK when launched, -1 when not eligible.
int cf_dw3s2_try(int mode, const cf_dw_args* a, cudaStream_t stream) {
    if (cf_env("CFNET_DW3_OFF", 0)) return -1;
    const cf_geom& g = a->g;
    if (!(g.kt == 3 && g.kh == 3 && g.kw == 3 && g.pt == 1 && g.ph == 1 && g.pw == 1 && g.st == 1 && g.sh == 2 && g.sw == 2)) return -1;
    if (g.T != g.Ti || g.H != (g.Hi - 1) / 2 + 1 || g.W != (g.Wi - 1) / 2 + 1) return -1;
    if (a->C % S2_CS != 0 || a->C < S2_CS) return -1;
    if (!(g.W == 7 || g.W % 14 == 0)) return -1;
    uintptr_t al = (uintptr_t)a->x | (uintptr_t)a->y | (uintptr_t)(a->x2 ? a->x2 : a->x) | (uintptr_t)(a->aux ? a->aux : a->x);
    if (al & 7) return -1;
    if (mode == S2_FWD && !(a->pro_mode == CF_PRO_NONE || a->pro_mode == CF_PRO_AFFINE || a->pro_mode == CF_PRO_AFFINE_RELU)) return -1;
    if (mode == S2_FWD && a->stats_mode == CF_STATS_SUM_AUX) return -1;
    if (mode != S2_FWD && !(a->pro_mode == CF_PRO_NONE || a->pro_mode == CF_PRO_AFFINE || a->pro_mode == CF_PRO_AFFINE2)) return -1;
    if ((mode == 1 || mode == 3) && a->stats_mode == CF_STATS_SUM_SQ) return -1;
    if (mode == 3 && !(a->dw_out && a->aux && a->epi_mode == CF_EPI_DRELU && a->epi_a && a->epi_b)) return -1;
    S2Params p;
    p.B = a->B; p.C = a->C; p.T = g.T; p.Hi = g.Hi; p.Wi = g.Wi; p.Ho = g.H; p.Wo = g.W;
    const int npw = g.W == 7 ? 1 : 2;
    p.htiles = (g.H + S2_OTH - 1) / S2_OTH;
    p.wtiles = g.W / (S2_PW * npw);
    p.slabs = a->C / S2_CS;
    const long long cols = (long long)p.B * p.htiles * p.wtiles * p.slabs;
    p.total_steps = cols * g.T;
    int nsm = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
    if (npw == 1) nsm *= 2;                                      // 65 KB of shared memory and 108 x 254 registers per CTA at the narrow tile
    long long spc = (p.total_steps + nsm - 1) / nsm;
    if (spc < 4) spc = 4;
    p.steps_per_cta = (int)spc;
    if (mode == 1) return npw == 1 ? s2_launch_dgrad<1, false>(a, p, stream) : s2_launch_dgrad<2, false>(a, p, stream);
#ifdef CFNET_AB      // the one-pass stride-2 backward lost its same-box A/B (profiles/r02_ab_same_box.md section 5): experiment build only
    if (mode == 3) return npw == 1 ? s2_launch_dgrad<1, true>(a, p, stream) : s2_launch_dgrad<2, true>(a, p, stream);
#else
    if (mode == 3) return -1;
#endif
    if (npw == 1) return mode == S2_FWD ? s2_launch<S2_FWD, 1>(a, p, stream) : s2_launch<S2_WGRAD, 1>(a, p, stream);
    return mode == S2_FWD ? s2_launch<S2_FWD, 2>(a, p, stream) : s2_launch<S2_WGRAD, 2>(a, p, stream);
}
