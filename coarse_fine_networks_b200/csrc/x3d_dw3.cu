// Depthwise 3x3x3 / stride 1 / pad 1 convolution (every Bottleneck conv2 but the first of a stage,
// x3d_fine.py:89-97,153): forward, data gradient and weight gradient as ONE plane-marching shared-memory kernel family.
//
// The direct-from-global kernels of x3d_dw.cu re-read every input element up to 9 times from L2 (a CTA covered one
// W-row, so neighbouring rows / frames were other CTAs' loads) and ran at 10-20 % of the HBM roofline.  Here a CTA owns
// an (8 x 14) or (8 x 7) spatial tile of a 54-channel slab of one sample and MARCHES ALONG T: the haloed input plane
// (10 x 16 positions x 54 channels, BatchNorm+ReLU resp. BatchNorm-backward prologue applied once per element) of
// frame t+1 arrives with cp.async while frame t is computed out of a ring of three planes in shared memory, so each
// input element is read from global memory once per tile (+ the spatial halo) and 27 times from shared memory /
// registers.  Thread = (2 x 7 output patch, 2 channels): 28 accumulators, the 27 x 2 weights of its channels in
// registers, each shared-memory row of 9 vectors feeds up to 6 x 7 FMAs.
//
//   forward : y[t,h,w,c]  = sum_taps relu(a*x+b)[t+dt-1,h+dh-1,w+dw-1,c] * w[c,dt,dh,dw]        (+ sum y, sum y^2)
//   dgrad   : dz[t,h,w,c] = [a1*y1+b1 > 0] * sum_taps d'[t-dt+1,h-dh+1,w-dw+1,c] * w[c,dt,dh,dw]  (+ sum dz, sum dz*y1)
//             = the same kernel with flipped weights, d' = P*dU + Q*y2 + R (BatchNorm backward as an affine map)
//   wgrad   : dw[c,dt,dh,dw] += sum_pos d'[t,h,w,c] * relu(a1*y1+b1)[t+dt-1,h+dh-1,w+dw-1,c]
//             ring = activated y1 planes, d' read at the patch positions; 27 x 2 accumulators per thread
//   fused   : the data-gradient pass also produces the weight gradient: substituting q = pos + tap - 1,
//             dw[c,tap] = sum_q relu(a1*y1+b1)[q] * d'[q - tap + 1] -- the SAME d' neighbourhood values the data gradient
//             multiplies with the flipped weights, times the activation at the output position q (the `aux` row the
//             ReLU mask needs anyway).  One pass over (dU, y2, y1) instead of two: 27 more FMAs per output from registers,
//             weights read from shared memory (their registers hold the 27 x 2 weight-gradient accumulators).
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include <stdlib.h>

#define D3_LANES 27
#define D3_CS 54                      /* channels per slab = 27 lanes x 2 */
#define D3_PW 7
#define D3_TH 8                       /* tile rows = 4 patch rows of 2 */
#define D3_HH (D3_TH + 2)

enum { D3_FWD = 0, D3_DGRAD = 1, D3_WGRAD = 2, D3_FUSED = 3 };   // FUSED: data gradient + weight gradient in one pass

struct D3Params {
    int B, C, T, H, W;
    int htiles, wtiles, slabs;
    long long total_steps;            // columns * T, a column = (sample, h tile, w tile, channel slab)
    int steps_per_cta;
};

__device__ __forceinline__ void d3_cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void d3_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// MODE: D3_FWD / D3_DGRAD / D3_WGRAD;  NPW: patches along W (tile width 7 * NPW)
template <int MODE, int NPW>
__global__ void __launch_bounds__(D3_LANES * 4 * NPW) dw3_kernel(const cf_dw_args a, const D3Params p) {
    cf_pdl_enter();
    constexpr int TW = D3_PW * NPW, HW = TW + 2;
    constexpr int NPATCH = 4 * NPW, NT = D3_LANES * NPATCH;
    constexpr int NPOS = D3_HH * HW;                               // haloed positions per plane
    constexpr int PLANE = NPOS * D3_CS;                            // floats per plane
    constexpr int KPT = (NPOS + NPATCH - 1) / NPATCH;              // positions loaded per thread per plane
    extern __shared__ __align__(16) float sm[];
    float* ring = sm;                                              // [3][PLANE]
    float* stg0 = ring + 3 * PLANE;                                // raw plane of the first source tensor
    float* stg1 = stg0 + PLANE;                                    // raw plane of the second one (AFFINE2 only)
    constexpr bool DG = MODE == D3_DGRAD || MODE == D3_FUSED;                  // data-gradient form of the ring
    const bool two_src = DG && a.pro_mode == CF_PRO_AFFINE2;                   // must match d3_launch's staging count
    float* tabs = stg0 + (two_src ? 2 : 1) * PLANE;                // [5][54]: ring prologue a,b,c ; epilogue a,b
    float* ws = tabs + 5 * D3_CS;                                  // [27][54]

    const int tid = threadIdx.x;
    const int patch = tid / D3_LANES, lane = tid - patch * D3_LANES;
    const int pr = patch / NPW, pc = patch - pr * NPW;
    const int oh0 = pr * 2, ow0 = pc * D3_PW;                      // patch origin inside the tile
    const int C = p.C, T = p.T, H = p.H, W = p.W;
    // This CTA's share of the flattened (column, t) sequence: contiguous, equal for all CTAs (no wave quantisation: the
    // grid is one CTA per SM); a share may span several columns, each piece restarts the ring (2 halo planes).
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
    const int h0 = th_i * D3_TH, w0 = tw_i * TW;
    const int cs0 = slab * D3_CS, c0 = cs0 + lane * 2;
    __syncthreads();                                               // the previous piece's reductions are done with the ring

    // ---- per-CTA constants: prologue / epilogue tables of this sample and slab, weights of the slab
    //   ring tensor:  FWD: a.x with (pro_a, pro_b);  DGRAD: a.x (+ a.x2) with (pro_a, pro_b, pro_c);
    //                 WGRAD: a.aux with (epi_a, epi_b) as BatchNorm+ReLU
    for (int i = tid; i < D3_CS; i += NT) {
        const size_t tc = (size_t)b * C + cs0 + i;
        float ra = 1.f, rb = 0.f, rc = 0.f, ea = 1.f, eb = 0.f;
        if (MODE == D3_WGRAD) {
            if (a.epi_a) { ra = a.epi_a[tc]; rb = a.epi_b[tc]; }
            if (a.pro_mode != CF_PRO_NONE) {                       // d' tables live in the "epilogue" slots (+ rc)
                ea = a.pro_a[tc];
                eb = a.pro_b ? a.pro_b[tc] : 0.f;
                rc = a.pro_c ? a.pro_c[tc] : 0.f;
            }
        } else {
            if (a.pro_mode != CF_PRO_NONE) {
                ra = a.pro_a[tc];
                rb = a.pro_b ? a.pro_b[tc] : 0.f;
                rc = a.pro_c ? a.pro_c[tc] : 0.f;
            }
            if (DG && a.epi_mode == CF_EPI_DRELU) { ea = a.epi_a[tc]; eb = a.epi_b[tc]; }
        }
        tabs[i] = ra; tabs[D3_CS + i] = rb; tabs[2 * D3_CS + i] = rc; tabs[3 * D3_CS + i] = ea; tabs[4 * D3_CS + i] = eb;
    }
    if (MODE != D3_WGRAD) {
        for (int i = tid; i < 27 * D3_CS; i += NT) {
            const int tap = i / D3_CS, c = i - tap * D3_CS;
            ws[i] = a.w[(size_t)(cs0 + c) * 27 + (DG ? 26 - tap : tap)];     // dgrad = conv with the flipped stencil
        }
    }
    __syncthreads();
    const float2 ra = *reinterpret_cast<const float2*>(tabs + lane * 2);
    const float2 rb = *reinterpret_cast<const float2*>(tabs + D3_CS + lane * 2);
    const float2 rc = *reinterpret_cast<const float2*>(tabs + 2 * D3_CS + lane * 2);
    const float2 ea = *reinterpret_cast<const float2*>(tabs + 3 * D3_CS + lane * 2);
    const float2 eb = *reinterpret_cast<const float2*>(tabs + 4 * D3_CS + lane * 2);

    const float* src0 = MODE == D3_WGRAD ? a.aux : a.x;           // tensor behind the ring
    const float* src1 = two_src ? a.x2 : nullptr;
    const int ring_mode = MODE == D3_WGRAD ? (a.epi_a ? CF_PRO_AFFINE_RELU : CF_PRO_NONE) : a.pro_mode;

    // ---- plane movement: this thread owns positions patch + k * NPATCH of every plane, at its own channel pair.  Which
    // of them lie inside the image, and their element offsets inside a frame, do not depend on t: computed once per piece.
    uint32_t vmask = 0;
    int goff[KPT];
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        const int pos = patch + k * NPATCH;
        const int hh = pos / HW, ww = pos - hh * HW;
        const int h = h0 - 1 + hh, w = w0 - 1 + ww;
        const bool v = pos < NPOS && (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W;
        goff[k] = v ? (h * W + w) * C + c0 : 0;
        vmask |= v ? (1u << k) : 0u;
    }
    const int soff0 = patch * D3_CS + lane * 2;                   // staging / ring offset of position k: soff0 + k * NPATCH * 54
    const size_t frame = (size_t)H * W * C;
    auto plane_issue = [&](int t) {                               // global -> staging (cp.async, 8 bytes per position)
        if (t < 0 || t >= T) return;
        const float* f0 = src0 + ((size_t)b * T + t) * frame;
        const float* f1 = src1 ? src1 + ((size_t)b * T + t) * frame : nullptr;
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            if (vmask & (1u << k)) {
                d3_cp_async8(stg0 + soff0 + k * (NPATCH * D3_CS), f0 + goff[k]);
                if (DG && f1) d3_cp_async8(stg1 + soff0 + k * (NPATCH * D3_CS), f1 + goff[k]);
            }
        }
    };
    auto plane_land = [&](int t, float* dst) {                    // staging -> ring slot with the prologue; zero outside
        const uint32_t m = (t >= 0 && t < T) ? vmask : 0u;
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            if (patch + k * NPATCH >= NPOS) break;
            float2 v = make_float2(0.f, 0.f);
            if (m & (1u << k)) {
                const float2 x = *reinterpret_cast<const float2*>(stg0 + soff0 + k * (NPATCH * D3_CS));
                if (ring_mode == CF_PRO_AFFINE_RELU) {
                    v.x = fmaxf(fmaf(ra.x, x.x, rb.x), 0.f);
                    v.y = fmaxf(fmaf(ra.y, x.y, rb.y), 0.f);
                } else if (ring_mode == CF_PRO_AFFINE2) {
                    const float2 x2 = *reinterpret_cast<const float2*>(stg1 + soff0 + k * (NPATCH * D3_CS));
                    v.x = fmaf(ra.x, x.x, fmaf(rb.x, x2.x, rc.x));
                    v.y = fmaf(ra.y, x.y, fmaf(rb.y, x2.y, rc.y));
                } else if (ring_mode == CF_PRO_AFFINE) {
                    v.x = fmaf(ra.x, x.x, rb.x);
                    v.y = fmaf(ra.y, x.y, rb.y);
                } else {
                    v = x;
                }
            }
            *reinterpret_cast<float2*>(dst + soff0 + k * (NPATCH * D3_CS)) = v;
        }
    };
    auto slot = [&](int t) { return ring + ((t - t0 + 1) % 3) * PLANE; };     // plane t0-1 -> slot 0

    // ---- per-thread state
    float2 wreg[27];                                              // FWD/DGRAD: weights; WGRAD/FUSED: the 27 accumulators
#pragma unroll
    for (int i = 0; i < 27; ++i)
        wreg[i] = (MODE == D3_WGRAD || MODE == D3_FUSED) ? make_float2(0.f, 0.f) : *reinterpret_cast<const float2*>(ws + i * D3_CS + lane * 2);
    float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
    // WGRAD: raw output-gradient values (and the second BatchNorm-backward operand) of the NEXT frame's patch, loaded one
    // step ahead so that their latency hides behind the current frame's FMAs
    float2 dn[MODE == D3_WGRAD ? 2 : 1][MODE == D3_WGRAD ? D3_PW : 1], dn2[MODE == D3_WGRAD ? 2 : 1][MODE == D3_WGRAD ? D3_PW : 1];
    auto d_prefetch = [&](int t) {
        if (MODE != D3_WGRAD) return;
        const size_t ob = (((size_t)b * T + t) * H + h0 + oh0) * W + w0 + ow0;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < D3_PW; ++j) {
                const bool v = t < T && h0 + oh0 + r < H;
                const size_t g = (ob + (size_t)r * W + j) * C + c0;
                dn[MODE == D3_WGRAD ? r : 0][MODE == D3_WGRAD ? j : 0] = v ? __ldg(reinterpret_cast<const float2*>(a.x + g)) : make_float2(0.f, 0.f);
                if (a.pro_mode == CF_PRO_AFFINE2)
                    dn2[MODE == D3_WGRAD ? r : 0][MODE == D3_WGRAD ? j : 0] = v ? __ldg(reinterpret_cast<const float2*>(a.x2 + g)) : make_float2(0.f, 0.f);
            }
    };
    const float* pbase_thr = ring + (oh0 * HW + ow0) * D3_CS + lane * 2;     // patch origin (haloed coordinates) in slot 0

    // ---- prologue: planes t0-1 and t0 synchronously, t0+1 in flight
    plane_issue(t0 - 1);
    d3_cp_async_wait_all();
    plane_land(t0 - 1, slot(t0 - 1));
    plane_issue(t0);
    d3_cp_async_wait_all();
    plane_land(t0, slot(t0));
    plane_issue(t0 + 1);
    d_prefetch(t0);

    for (int t = t0; t < t1; ++t) {
        d3_cp_async_wait_all();
        plane_land(t + 1, slot(t + 1));
        __syncthreads();
        if (t + 2 <= t1) plane_issue(t + 2);                       // arrives while frame t is computed

        const int sl0 = (t - t0) % 3;                              // slot of plane t-1; t -> +1, t+1 -> +2 (mod 3)
        const size_t orow_base = (((size_t)b * T + t) * H + h0 + oh0) * W + w0 + ow0;     // patch origin in the image
        if (MODE != D3_WGRAD) {
            // aux (pre-activation at the output positions) for the dgrad mask / statistics: issue early
            float2 aux[2][D3_PW];
            const bool need_aux = DG && (a.epi_mode == CF_EPI_DRELU || a.stats_mode == CF_STATS_SUM_AUX);
            if (need_aux) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int j = 0; j < D3_PW; ++j)
                        aux[r][j] = (h0 + oh0 + r < H) ? __ldg(reinterpret_cast<const float2*>(a.aux + (orow_base + (size_t)r * W + j) * C + c0))
                                                       : make_float2(0.f, 0.f);
            }
            float2 acc[2][D3_PW];
            float2 actv[MODE == D3_FUSED ? 2 : 1][MODE == D3_FUSED ? D3_PW : 1];       // relu(a1*y1+b1) at the output positions
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < D3_PW; ++j) {
                    acc[r][j] = make_float2(0.f, 0.f);
                    if (MODE == D3_FUSED) {                        // rows outside the image hold aux = 0 -> masked explicitly
                        const bool rv = h0 + oh0 + r < H;
                        actv[MODE == D3_FUSED ? r : 0][MODE == D3_FUSED ? j : 0] =
                            make_float2(rv ? fmaxf(fmaf(ea.x, aux[r][j].x, eb.x), 0.f) : 0.f, rv ? fmaxf(fmaf(ea.y, aux[r][j].y, eb.y), 0.f) : 0.f);
                    }
                }
#pragma unroll
            for (int dt = 0; dt < 3; ++dt) {
                const float* pl = pbase_thr + ((sl0 + dt) % 3) * PLANE;
#pragma unroll
                for (int r = 0; r < 4; ++r) {                      // input rows oh0 + r (haloed) feed output rows r - dh
                    float2 in[D3_PW + 2];
#pragma unroll
                    for (int j = 0; j < D3_PW + 2; ++j) in[j] = *reinterpret_cast<const float2*>(pl + (r * HW + j) * D3_CS);
#pragma unroll
                    for (int dh = 0; dh < 3; ++dh) {
                        const int orow = r - dh;
                        if (orow < 0 || orow > 1) continue;
#pragma unroll
                        for (int dw = 0; dw < 3; ++dw) {
                            if (MODE == D3_FUSED) {
                                const int ti = (dt * 3 + dh) * 3 + dw;
                                const float2 wv = *reinterpret_cast<const float2*>(ws + ti * D3_CS + lane * 2);
                                float2 g = wreg[ti];
#pragma unroll
                                for (int j = 0; j < D3_PW; ++j) {
                                    acc[orow][j].x = fmaf(in[j + dw].x, wv.x, acc[orow][j].x);
                                    acc[orow][j].y = fmaf(in[j + dw].y, wv.y, acc[orow][j].y);
                                    g.x = fmaf(in[j + dw].x, actv[orow][j].x, g.x);
                                    g.y = fmaf(in[j + dw].y, actv[orow][j].y, g.y);
                                }
                                wreg[ti] = g;
                            } else {
                                const float2 wv = wreg[(dt * 3 + dh) * 3 + dw];
#pragma unroll
                                for (int j = 0; j < D3_PW; ++j) {
                                    acc[orow][j].x = fmaf(in[j + dw].x, wv.x, acc[orow][j].x);
                                    acc[orow][j].y = fmaf(in[j + dw].y, wv.y, acc[orow][j].y);
                                }
                            }
                        }
                    }
                }
            }
            // epilogue: mask, store, statistics
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (h0 + oh0 + r >= H) continue;
#pragma unroll
                for (int j = 0; j < D3_PW; ++j) {
                    float2 v = acc[r][j];
                    if (DG && a.epi_mode == CF_EPI_DRELU) {
                        v.x = fmaf(ea.x, aux[r][j].x, eb.x) > 0.f ? v.x : 0.f;
                        v.y = fmaf(ea.y, aux[r][j].y, eb.y) > 0.f ? v.y : 0.f;
                    }
                    *reinterpret_cast<float2*>(a.y + (orow_base + (size_t)r * W + j) * C + c0) = v;
                    s1.x += v.x; s1.y += v.y;
                    if (DG) { s2.x = fmaf(v.x, aux[r][j].x, s2.x); s2.y = fmaf(v.y, aux[r][j].y, s2.y); }
                    else { s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); }
                }
            }
        } else {
            // weight gradient: d' at the patch positions (BatchNorm-backward map of the output gradient)
            float2 dcur[2][D3_PW], dcur2[2][D3_PW];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < D3_PW; ++j) { dcur[r][j] = dn[MODE == D3_WGRAD ? r : 0][MODE == D3_WGRAD ? j : 0]; dcur2[r][j] = dn2[MODE == D3_WGRAD ? r : 0][MODE == D3_WGRAD ? j : 0]; }
            if (t + 1 < t1) d_prefetch(t + 1);
            const float* pl0 = pbase_thr + (sl0 % 3) * PLANE;
            const float* pl1 = pbase_thr + ((sl0 + 1) % 3) * PLANE;
            const float* pl2 = pbase_thr + ((sl0 + 2) % 3) * PLANE;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (h0 + oh0 + r >= H) continue;
                float2 d[D3_PW];
#pragma unroll
                for (int j = 0; j < D3_PW; ++j) {
                    float2 v = dcur[r][j];
                    if (a.pro_mode == CF_PRO_AFFINE2) {
                        const float2 v2 = dcur2[r][j];
                        v.x = fmaf(ea.x, v.x, fmaf(eb.x, v2.x, rc.x));
                        v.y = fmaf(ea.y, v.y, fmaf(eb.y, v2.y, rc.y));
                    } else if (a.pro_mode != CF_PRO_NONE) {
                        v.x = fmaf(ea.x, v.x, eb.x);
                        v.y = fmaf(ea.y, v.y, eb.y);
                    }
                    d[j] = v;
                }
#pragma unroll
                for (int dt = 0; dt < 3; ++dt) {
                    const float* pl = dt == 0 ? pl0 : (dt == 1 ? pl1 : pl2);
#pragma unroll
                    for (int dh = 0; dh < 3; ++dh) {
                        float2 in[D3_PW + 2];
#pragma unroll
                        for (int j = 0; j < D3_PW + 2; ++j) in[j] = *reinterpret_cast<const float2*>(pl + ((r + dh) * HW + j) * D3_CS);
#pragma unroll
                        for (int dw = 0; dw < 3; ++dw) {
                            float2 s = wreg[(dt * 3 + dh) * 3 + dw];
#pragma unroll
                            for (int j = 0; j < D3_PW; ++j) {
                                s.x = fmaf(d[j].x, in[j + dw].x, s.x);
                                s.y = fmaf(d[j].y, in[j + dw].y, s.y);
                            }
                            wreg[(dt * 3 + dh) * 3 + dw] = s;
                        }
                    }
                }
            }
        }
        __syncthreads();                                           // everyone is done with plane t-1's slot
    }
    d3_cp_async_wait_all();

    // ---- CTA reductions over the patches (same channels), then global atomics.  The ring is free now.
    float* red = ring;
    if (MODE == D3_WGRAD || MODE == D3_FUSED) {
        float* dwp = MODE == D3_FUSED ? a.dw_out : a.y;
#pragma unroll
        for (int i = 0; i < 27; ++i) *reinterpret_cast<float2*>(red + ((size_t)patch * 27 + i) * D3_CS + lane * 2) = wreg[i];
        __syncthreads();
        for (int i = tid; i < 27 * D3_CS; i += NT) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < NPATCH; ++q) s += red[(size_t)q * 27 * D3_CS + i];
            const int tap = i / D3_CS, c = i - tap * D3_CS;
            atomicAdd(dwp + (size_t)(cs0 + c) * 27 + (MODE == D3_FUSED ? 26 - tap : tap), s);   // fused accumulators are indexed by the flipped tap
        }
        if (MODE == D3_FUSED) __syncthreads();
    }
    if (MODE != D3_WGRAD && a.stats_mode != CF_STATS_NONE) {
        *reinterpret_cast<float2*>(red + patch * 2 * D3_CS + lane * 2) = s1;
        *reinterpret_cast<float2*>(red + (patch * 2 + 1) * D3_CS + lane * 2) = s2;
        __syncthreads();
        for (int i = tid; i < 2 * D3_CS; i += NT) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < NPATCH; ++q) s += red[q * 2 * D3_CS + i];
            const int which = i / D3_CS, c = i - which * D3_CS;
            atomicAdd(a.stats + ((size_t)b * C + cs0 + c) * 2 + which, (double)s);
        }
    }
    }   // pieces
}

// ---------------------------------------------------------------------------------------
template <int MODE, int NPW>
static int d3_launch(const cf_dw_args* a, const D3Params& p, cudaStream_t stream) {
    constexpr int TW = D3_PW * NPW, HW = TW + 2;
    constexpr int PLANE = D3_HH * HW * D3_CS;
    const int nstg = ((MODE == D3_DGRAD || MODE == D3_FUSED) && a->pro_mode == CF_PRO_AFFINE2) ? 2 : 1;
    const size_t smem = (size_t)((3 + nstg) * PLANE + 5 * D3_CS + 27 * D3_CS) * sizeof(float);
    static CfOncePerDevice done;
    if (done.need()) {
        cudaError_t e = cudaFuncSetAttribute(dw3_kernel<MODE, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { cf_set_error("dw3: cannot opt in to shared memory: %s", cudaGetErrorString(e)); return CF_ERR_CUDA; }
        done.mark();
    }
    dim3 grid((unsigned)((p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta));
    cf_launch(dw3_kernel<MODE, NPW>, grid, D3_LANES * 4 * NPW, smem, stream, *a, p);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

// returns CF_OK when launched, -1 when not eligible (the caller runs the general kernels of x3d_dw.cu)
int cf_dw3_try(int mode, const cf_dw_args* a, cudaStream_t stream) {
    if (cf_env("CFNET_DW3_OFF", 0)) return -1;
    const cf_geom& g = a->g;
    if (!(g.kt == 3 && g.kh == 3 && g.kw == 3 && g.pt == 1 && g.ph == 1 && g.pw == 1 && g.st == 1 && g.sh == 1 && g.sw == 1)) return -1;
    if (g.T != g.Ti || g.H != g.Hi || g.W != g.Wi) return -1;
    if (a->C % D3_CS != 0 || a->C < D3_CS) return -1;
    if (!(g.W == 7 || g.W % 14 == 0)) return -1;
    uintptr_t al = (uintptr_t)a->x | (uintptr_t)a->y | (uintptr_t)(a->x2 ? a->x2 : a->x) | (uintptr_t)(a->aux ? a->aux : a->x);
    if (al & 7) return -1;
    if (mode == D3_FWD && !(a->pro_mode == CF_PRO_NONE || a->pro_mode == CF_PRO_AFFINE || a->pro_mode == CF_PRO_AFFINE_RELU)) return -1;
    if (mode == D3_FWD && a->stats_mode == CF_STATS_SUM_AUX) return -1;
    if (mode != D3_FWD && !(a->pro_mode == CF_PRO_NONE || a->pro_mode == CF_PRO_AFFINE || a->pro_mode == CF_PRO_AFFINE2)) return -1;
    if ((mode == D3_DGRAD || mode == D3_FUSED) && a->stats_mode == CF_STATS_SUM_SQ) return -1;
    if (mode == D3_FUSED && !(a->dw_out && a->aux && a->epi_mode == CF_EPI_DRELU && a->epi_a && a->epi_b)) return -1;
    D3Params p;
    p.B = a->B; p.C = a->C; p.T = g.T; p.H = g.H; p.W = g.W;
    const int force_npw = cf_env("CFNET_DW3_NPW", 0);                // tile width experiment (-DCFNET_AB build only)
    const int npw = (g.W == 7 || force_npw == 1) ? 1 : 2;
    p.htiles = (g.H + D3_TH - 1) / D3_TH;
    p.wtiles = g.W / (D3_PW * npw);
    p.slabs = a->C / D3_CS;
    const long long cols = (long long)p.B * p.htiles * p.wtiles * p.slabs;
    p.total_steps = cols * g.T;
    int nsm = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
    const int cta_per_sm = npw == 1 ? 2 : 1;                       // 78-97 KB of shared memory per CTA at the narrow tile
    nsm *= cta_per_sm;
    long long spc = (p.total_steps + nsm - 1) / nsm;
    if (spc < 4) spc = 4;                                        // two halo planes per piece: keep their share bounded
    p.steps_per_cta = (int)spc;
    if (npw == 1) {
        if (mode == D3_FWD) return d3_launch<D3_FWD, 1>(a, p, stream);
        if (mode == D3_DGRAD) return d3_launch<D3_DGRAD, 1>(a, p, stream);
        if (mode == D3_FUSED) return d3_launch<D3_FUSED, 1>(a, p, stream);
        return d3_launch<D3_WGRAD, 1>(a, p, stream);
    }
    if (mode == D3_FWD) return d3_launch<D3_FWD, 2>(a, p, stream);
    if (mode == D3_DGRAD) return d3_launch<D3_DGRAD, 2>(a, p, stream);
    if (mode == D3_FUSED) return d3_launch<D3_FUSED, 2>(a, p, stream);
    return d3_launch<D3_WGRAD, 2>(a, p, stream);
}
