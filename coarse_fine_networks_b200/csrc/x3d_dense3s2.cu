// The dense 3x3x3, stride-2, pad-1, 24 -> 24 convs of the Grid Pool confidence branch (pool_1.conv1 / conv2,
// x3d_coarse.py:362-365,379-380): forward and data gradient as direct kernels with warp-uniform weight broadcasts.
//
// ---- data gradient, GATHER form ----
// The generic path (pw_conv_kernel with scatter_out) computes a [rows, C*27] GEMM and scatters every element with
// atomicAdd: 648 atomics per output row, 65 M for pool_1.conv1 at the bench shape (2.0 ms, the slowest single launch
// of the coarse stream's backward).  Here every INPUT position collects its contributions instead:
//     dx[b,ti,hi,wi,ci] = sum_{taps (kt,kh,kw) with (ti+1-kt, hi+1-kh, wi+1-kw) even and in range}
//                          sum_co pro(dz[b, (ti+1-kt)/2, (hi+1-kh)/2, (wi+1-kw)/2, co]) * W[co][ci*27 + tap]
// With stride 2 the parity of a coordinate fixes its taps (odd: k in {0,2}; even: k = 1), so a CTA = (sample, input
// frame, parity class of (h,w)) has ONE tap list (1..8 taps) for all its positions: their weights (<= 8 x C x C) are
// staged in shared memory and read as warp-uniform float4 broadcasts; a thread owns two positions (48 accumulators) so
// every weight fetch feeds 8 FMAs.  No atomics, no memset: each input position is written exactly once.
// pro = identity or the BatchNorm-backward map P*dz + Q*y + R (CF_PRO_AFFINE2).
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

#define DG_THREADS 128

struct DgArgs {
    const float* dz;      // dense [B, To*Ho*Wo, C]
    const float* y;       // second input of AFFINE2 or NULL
    const float* w;       // [C][C*27]
    const float* P;       // [B,C] tables or NULL
    const float* Q;
    const float* R;
    float* dx;            // [B] x sample_stride; position stride C
    int To, Ho, Wo, Ti, Hi, Wi;
    long long sample_stride;
    int affine2;
};

template <int C>
__global__ void __launch_bounds__(DG_THREADS) dense3_s2_dgrad_kernel(const DgArgs a) {
    cf_pdl_enter();
    constexpr int C4 = C / 4;
    __shared__ __align__(16) float w_s[8][C][C];          // [tap slot][co][ci]
    __shared__ float tab[3][C];
    const int tid = threadIdx.x;
    const int cls = blockIdx.x & 3, ph = cls >> 1, pw = cls & 1;
    const int ti = (blockIdx.x >> 2) % a.Ti;
    const int b = (blockIdx.x >> 2) / a.Ti;

    // tap lists per dimension: (kernel index, output index offset relative to i = coord / 2)
    int kts[2], ots[2], nt = 0;
    for (int kt = 0; kt < 3; ++kt) {
        const int num = ti + 1 - kt;
        if (num >= 0 && !(num & 1) && (num >> 1) < a.To) { kts[nt] = kt; ots[nt] = num >> 1; ++nt; }
    }
    const int nh = ph ? 2 : 1, nw = pw ? 2 : 1;           // odd coordinate: k = 0 (o = i + 1) and k = 2 (o = i); even: k = 1 (o = i)
    const int ntaps = nt * nh * nw;

    for (int e = tid; e < ntaps * C * C; e += DG_THREADS) {
        const int s = e / (C * C), rem = e - s * C * C, co = rem / C, ci = rem - co * C;
        const int it = s / (nh * nw), ih = (s / nw) % nh, iw = s % nw;
        const int kh = ph ? 2 * ih : 1, kw = pw ? 2 * iw : 1;
        w_s[s][co][ci] = __ldg(a.w + (size_t)co * (C * 27) + ci * 27 + kts[it] * 9 + kh * 3 + kw);
    }
    if (tid < C) {
        tab[0][tid] = a.affine2 ? a.P[(size_t)b * C + tid] : 1.f;
        tab[1][tid] = a.affine2 ? a.Q[(size_t)b * C + tid] : 0.f;
        tab[2][tid] = (a.affine2 && a.R) ? a.R[(size_t)b * C + tid] : 0.f;
    }
    __syncthreads();

    const int ch = (a.Hi - ph + 1) >> 1, cw = (a.Wi - pw + 1) >> 1;     // positions of this parity class
    const int npos = ch * cw;
    const size_t R = (size_t)a.To * a.Ho * a.Wo;
    const float* dzb = a.dz + (size_t)b * R * C;
    const float* yb = a.y ? a.y + (size_t)b * R * C : nullptr;
    float* dxb = a.dx + (size_t)b * a.sample_stride + (size_t)ti * a.Hi * a.Wi * C;

    for (int base = 0; base < npos; base += 2 * DG_THREADS) {
        int pi[2], pj[2];
        bool pv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int p = base + tid + j * DG_THREADS;
            pv[j] = p < npos;
            pi[j] = pv[j] ? p / cw : 0;
            pj[j] = pv[j] ? p - pi[j] * cw : 0;
        }
        float acc[2][C];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[j][c] = 0.f;

        for (int s = 0; s < ntaps; ++s) {
            const int it = s / (nh * nw), ih = (s / nw) % nh, iw = s % nw;
            const int dh = (ph && ih == 0) ? 1 : 0, dw = (pw && iw == 0) ? 1 : 0;
            float v[2][C];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int oh = pi[j] + dh, ow = pj[j] + dw;
                const bool ok = pv[j] && oh < a.Ho && ow < a.Wo;
                const size_t off = (((size_t)ots[it] * a.Ho + oh) * a.Wo + ow) * C;
#pragma unroll
                for (int q = 0; q < C4; ++q) {
                    float4 d = ok ? __ldg(reinterpret_cast<const float4*>(dzb + off) + q) : f4_zero();
                    if (a.affine2) {
                        float4 yy = ok ? __ldg(reinterpret_cast<const float4*>(yb + off) + q) : f4_zero();
                        d.x = fmaf(tab[0][4 * q], d.x, fmaf(tab[1][4 * q], yy.x, tab[2][4 * q]));
                        d.y = fmaf(tab[0][4 * q + 1], d.y, fmaf(tab[1][4 * q + 1], yy.y, tab[2][4 * q + 1]));
                        d.z = fmaf(tab[0][4 * q + 2], d.z, fmaf(tab[1][4 * q + 2], yy.z, tab[2][4 * q + 2]));
                        d.w = fmaf(tab[0][4 * q + 3], d.w, fmaf(tab[1][4 * q + 3], yy.w, tab[2][4 * q + 3]));
                        if (!ok) d = f4_zero();             // positions outside the volume contribute nothing (not R)
                    }
                    v[j][4 * q] = d.x; v[j][4 * q + 1] = d.y; v[j][4 * q + 2] = d.z; v[j][4 * q + 3] = d.w;
                }
            }
#pragma unroll
            for (int co = 0; co < C; ++co) {
#pragma unroll
                for (int q = 0; q < C4; ++q) {
                    const float4 w4 = *reinterpret_cast<const float4*>(&w_s[s][co][4 * q]);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        acc[j][4 * q] = fmaf(v[j][co], w4.x, acc[j][4 * q]);
                        acc[j][4 * q + 1] = fmaf(v[j][co], w4.y, acc[j][4 * q + 1]);
                        acc[j][4 * q + 2] = fmaf(v[j][co], w4.z, acc[j][4 * q + 2]);
                        acc[j][4 * q + 3] = fmaf(v[j][co], w4.w, acc[j][4 * q + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!pv[j]) continue;
            const int hi = 2 * pi[j] + ph, wi = 2 * pj[j] + pw;
            float4* o = reinterpret_cast<float4*>(dxb + ((size_t)hi * a.Wi + wi) * C);
#pragma unroll
            for (int q = 0; q < C4; ++q) o[q] = make_float4(acc[j][4 * q], acc[j][4 * q + 1], acc[j][4 * q + 2], acc[j][4 * q + 3]);
        }
    }
}

// -1: not this kernel's problem (the caller falls through to the generic scatter path)
int cf_dense_s2_dgrad_try(const cf_pw_args* a, cudaStream_t stream) {
    if (cf_env("CFNET_DENSE_DGRAD_OFF", 0)) return -1;
    const cf_geom& g = a->g;
    if (!a->scatter_out || a->accumulate || a->bias || a->stats_mode != CF_STATS_NONE || a->epi_mode != CF_EPI_NONE) return -1;
    if (!(g.kt == 3 && g.kh == 3 && g.kw == 3 && g.st == 2 && g.sh == 2 && g.sw == 2 && g.pt == 1 && g.ph == 1 && g.pw == 1)) return -1;
    if (a->K != 24 || a->N != 24 * 27 || g.ch_stride != 1 || g.pos_stride != 24 || a->w_sn != 1 || a->w_sk != 24 * 27) return -1;
    if (a->pro_mode != CF_PRO_NONE && a->pro_mode != CF_PRO_AFFINE2) return -1;
    if (a->pro_mode == CF_PRO_AFFINE2 && !(a->x2 && a->pro_a && a->pro_b)) return -1;
    if ((g.sample_stride & 3) || ((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15) || (a->x2 && ((uintptr_t)a->x2 & 15))) return -1;
    // every input position must have an output row on the dense side: T = floor((Ti-1)/2)+1 etc.
    if (g.T != (g.Ti - 1) / 2 + 1 || g.H != (g.Hi - 1) / 2 + 1 || g.W != (g.Wi - 1) / 2 + 1) return -1;
    const long long ctas = 4LL * a->B * g.Ti;
    if (ctas > 0x7fffffffLL) return -1;
    DgArgs d;
    d.dz = a->x; d.y = a->x2; d.w = a->w; d.P = a->pro_a; d.Q = a->pro_b; d.R = a->pro_c; d.dx = a->y;
    d.To = g.T; d.Ho = g.H; d.Wo = g.W; d.Ti = g.Ti; d.Hi = g.Hi; d.Wi = g.Wi;
    d.sample_stride = g.sample_stride; d.affine2 = a->pro_mode == CF_PRO_AFFINE2;
    cf_launch(dense3_s2_dgrad_kernel<24>, (unsigned)ctas, DG_THREADS, 0, stream, d);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

// ---- forward ----
// y[b,o,co] = bias[co] + sum_{tap,ci} pro(x[b, 2*o - 1 + tap, ci]) * W[co][ci*27 + tap]   (+ BatchNorm statistics of y)
// The generic path is a tap-gathered CUDA-core GEMM with K = 648 and N = 24 (0.6 ms for pool_1.conv1 at the bench shape).
// Here a thread owns two output positions and all 24 output channels; the 27 x 24 x 24 weights sit in shared memory as
// [tap][ci][co] (tap stride padded to 580 floats so the transposing fill is at most 4-way conflicted) and are read as
// warp-uniform float4 broadcasts, each feeding 8 FMAs.  pro = identity, a*x+b or relu(a*x+b) (bn1 + ReLU in front of
// conv2), applied to in-range inputs only; statistics = per-(sample, channel) sum and sum of squares in fp64 atomics,
// one pair per CTA and channel.
#define FW_THREADS 128
#define FW_TAPSTRIDE 580

struct FwArgs {
    const float* x;       // [B] x sample_stride, position stride C
    const float* w;       // [C][C*27]
    const float* bias;    // [C] or NULL
    const float* pa;      // [B,C] prologue tables or NULL
    const float* pb;
    float* y;             // dense [B, To*Ho*Wo, C]
    double* stats;        // [B,C,2] or NULL
    int To, Ho, Wo, Ti, Hi, Wi;
    long long sample_stride;
    int pro;              // CF_PRO_NONE / AFFINE / AFFINE_RELU
    int tiles_per_sample;
};

template <int C>
__global__ void __launch_bounds__(FW_THREADS) dense3_s2_fwd_kernel(const FwArgs a) {
    cf_pdl_enter();
    constexpr int C4 = C / 4;
    extern __shared__ __align__(16) float fw_smem[];
    float* w_s = fw_smem;                                 // [27][FW_TAPSTRIDE]: [tap][ci][co]
    float* tab = fw_smem + 27 * FW_TAPSTRIDE;             // pa[C], pb[C], bias[C]
    float* red = tab + 3 * C;                             // [FW_THREADS/32][2*C]
    const int tid = threadIdx.x;
    const int b = blockIdx.x / a.tiles_per_sample;
    const int row0 = (blockIdx.x - b * a.tiles_per_sample) * (2 * FW_THREADS);
    const int R = a.To * a.Ho * a.Wo;

    for (int e = tid; e < C * C * 27; e += FW_THREADS) {  // coalesced read of W[co][ci*27+tap], transposing store
        const int co = e / (C * 27), k = e - co * (C * 27), ci = k / 27, tap = k - ci * 27;
        w_s[tap * FW_TAPSTRIDE + ci * C + co] = __ldg(a.w + e);
    }
    if (tid < C) {
        tab[tid] = a.pro != CF_PRO_NONE ? a.pa[(size_t)b * C + tid] : 1.f;
        tab[C + tid] = (a.pro != CF_PRO_NONE && a.pb) ? a.pb[(size_t)b * C + tid] : 0.f;
        tab[2 * C + tid] = a.bias ? a.bias[tid] : 0.f;
    }
    __syncthreads();

    int ot[2], oh[2], ow[2];
    bool pv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int r = row0 + tid + j * FW_THREADS;
        pv[j] = r < R;
        const int rr = pv[j] ? r : 0;
        ow[j] = rr % a.Wo;
        const int q = rr / a.Wo;
        oh[j] = q % a.Ho;
        ot[j] = q / a.Ho;
    }
    float acc[2][C];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[j][c] = tab[2 * C + c];

    const float* xb = a.x + (size_t)b * a.sample_stride;
    for (int tap = 0; tap < 27; ++tap) {
        const int kt = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        float v[2][C];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int ti = 2 * ot[j] - 1 + kt, hi = 2 * oh[j] - 1 + kh, wi = 2 * ow[j] - 1 + kw;
            const bool ok = pv[j] && (unsigned)ti < (unsigned)a.Ti && (unsigned)hi < (unsigned)a.Hi && (unsigned)wi < (unsigned)a.Wi;
            const float4* src = reinterpret_cast<const float4*>(xb + (ok ? (((size_t)ti * a.Hi + hi) * a.Wi + wi) * C : 0));
#pragma unroll
            for (int q = 0; q < C4; ++q) {
                float4 d = f4_zero();
                if (ok) {
                    d = __ldg(src + q);
                    if (a.pro != CF_PRO_NONE) {
                        d.x = fmaf(tab[4 * q], d.x, tab[C + 4 * q]);
                        d.y = fmaf(tab[4 * q + 1], d.y, tab[C + 4 * q + 1]);
                        d.z = fmaf(tab[4 * q + 2], d.z, tab[C + 4 * q + 2]);
                        d.w = fmaf(tab[4 * q + 3], d.w, tab[C + 4 * q + 3]);
                        if (a.pro == CF_PRO_AFFINE_RELU) { d.x = fmaxf(d.x, 0.f); d.y = fmaxf(d.y, 0.f); d.z = fmaxf(d.z, 0.f); d.w = fmaxf(d.w, 0.f); }
                    }
                }
                v[j][4 * q] = d.x; v[j][4 * q + 1] = d.y; v[j][4 * q + 2] = d.z; v[j][4 * q + 3] = d.w;
            }
        }
        const float* wt = w_s + tap * FW_TAPSTRIDE;
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
#pragma unroll
            for (int q = 0; q < C4; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(wt + ci * C + 4 * q);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    acc[j][4 * q] = fmaf(v[j][ci], w4.x, acc[j][4 * q]);
                    acc[j][4 * q + 1] = fmaf(v[j][ci], w4.y, acc[j][4 * q + 1]);
                    acc[j][4 * q + 2] = fmaf(v[j][ci], w4.z, acc[j][4 * q + 2]);
                    acc[j][4 * q + 3] = fmaf(v[j][ci], w4.w, acc[j][4 * q + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (!pv[j]) continue;
        float4* o = reinterpret_cast<float4*>(a.y + ((size_t)b * R + row0 + tid + j * FW_THREADS) * C);
#pragma unroll
        for (int q = 0; q < C4; ++q) o[q] = make_float4(acc[j][4 * q], acc[j][4 * q + 1], acc[j][4 * q + 2], acc[j][4 * q + 3]);
    }
    if (a.stats) {
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float y0 = pv[0] ? acc[0][c] : 0.f, y1 = pv[1] ? acc[1][c] : 0.f;
            const float s1 = warp_sum(y0 + y1), s2 = warp_sum(fmaf(y0, y0, y1 * y1));
            if (lane == 0) { red[warp * 2 * C + c] = s1; red[warp * 2 * C + C + c] = s2; }
        }
        __syncthreads();
        if (tid < 2 * C) {
            double s = 0.0;
#pragma unroll
            for (int wq = 0; wq < FW_THREADS / 32; ++wq) s += (double)red[wq * 2 * C + tid];
            const int c = tid < C ? tid : tid - C;
            atomicAdd(a.stats + ((size_t)b * C + c) * 2 + (tid < C ? 0 : 1), s);
        }
    }
}

// -1: not this kernel's problem (the caller falls through to the generic gathered GEMM)
int cf_dense_s2_fwd_try(const cf_pw_args* a, cudaStream_t stream) {
    if (cf_env("CFNET_DENSE_FWD_OFF", 0)) return -1;
    const cf_geom& g = a->g;
    if (!a->gather_in || a->accumulate || a->epi_mode != CF_EPI_NONE || a->aux) return -1;
    if (a->stats_mode != CF_STATS_NONE && a->stats_mode != CF_STATS_SUM_SQ) return -1;
    if (!(g.kt == 3 && g.kh == 3 && g.kw == 3 && g.st == 2 && g.sh == 2 && g.sw == 2 && g.pt == 1 && g.ph == 1 && g.pw == 1)) return -1;
    if (a->N != 24 || a->K != 24 * 27 || g.ch_stride != 1 || g.pos_stride != 24 || a->w_sn != 24 * 27 || a->w_sk != 1) return -1;
    if (a->pro_mode != CF_PRO_NONE && a->pro_mode != CF_PRO_AFFINE && a->pro_mode != CF_PRO_AFFINE_RELU) return -1;
    if (a->pro_mode != CF_PRO_NONE && !a->pro_a) return -1;
    if ((g.sample_stride & 3) || ((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15)) return -1;
    const long long R = (long long)g.T * g.H * g.W;
    const long long tiles = (R + 2 * FW_THREADS - 1) / (2 * FW_THREADS);
    if (tiles * a->B > 0x7fffffffLL) return -1;
    FwArgs f;
    f.x = a->x; f.w = a->w; f.bias = a->bias; f.pa = a->pro_a; f.pb = a->pro_b; f.y = a->y;
    f.stats = a->stats_mode == CF_STATS_SUM_SQ ? a->stats : nullptr;
    f.To = g.T; f.Ho = g.H; f.Wo = g.W; f.Ti = g.Ti; f.Hi = g.Hi; f.Wi = g.Wi;
    f.sample_stride = g.sample_stride; f.pro = a->pro_mode; f.tiles_per_sample = (int)tiles;
    const size_t smem = (27 * FW_TAPSTRIDE + 3 * 24 + (FW_THREADS / 32) * 2 * 24) * sizeof(float);
    static CfOncePerDevice attr_done;
    if (attr_done.need()) {
        cudaError_t e = cudaFuncSetAttribute(dense3_s2_fwd_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { cf_set_error("cf_dense_s2_fwd: smem opt-in failed: %s", cudaGetErrorString(e)); return CF_ERR_CUDA; }
        attr_done.mark();
    }
    cf_launch(dense3_s2_fwd_kernel<24>, (unsigned)(tiles * a->B), FW_THREADS, smem, stream, f);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

// ---- weight gradient ----
// dw[co][ci*27 + tap] += sum_{b,o} pro_dy(dz[b,o,co]) * pro_x(x[b, 2*o - 1 + tap, ci]);   dbias[co] += sum pro_dy(...)
// The generic path (pw_wgrad_kernel with the tap gather) tiles the 24 x 648 result 64 x 64 and re-gathers the input for
// every tile (1.1 ms for pool_1.conv1 at the bench shape).  Here a persistent CTA stages 32 output rows at a time --
// the prologue'd output gradient [32][24] and the gathered, prologue'd input [32][27][24] (zeros outside the volume) -- in
// shared memory, and 252 threads each own a 4 taps x 4 co x 4 ci block of the result in registers (64 accumulators,
// one float4 of dz and four float4 of x per row: 64 FMAs per 5 shared-memory reads).  One atomicAdd per result element
// and CTA at the end.
#define WG3_THREADS 256
#define WG3_ROWS 32

struct Wg3Args {
    const float* dz;      // dense [B,R,C]
    const float* y;       // second input of AFFINE2 or NULL
    const float* P;       // [B,C] or NULL
    const float* Q;
    const float* Rc;
    const float* x;       // gathered side
    const float* xa;      // [B,C] or NULL
    const float* xb;
    float* dw;            // [C][C*27]
    float* dbias;         // [C] or NULL
    int B, To, Ho, Wo, Ti, Hi, Wi;
    long long sample_stride;
    int dy_affine2, x_mode;
    int chunks_per_sample;
};

template <int C>
__global__ void __launch_bounds__(WG3_THREADS, 2) dense3_s2_wgrad_kernel(const Wg3Args a) {
    cf_pdl_enter();
    constexpr int C4 = C / 4;                                 // 6
    extern __shared__ __align__(16) float wg_smem[];
    float* dzs = wg_smem;                                     // [WG3_ROWS][C]
    float* xs = dzs + WG3_ROWS * C;                           // [WG3_ROWS][27][C]
    const int tid = threadIdx.x;
    const int R = a.To * a.Ho * a.Wo;
    // compute role: (tap group, co group, ci group)
    const int cig = tid % C4, cog = (tid / C4) % C4, tg = tid / (C4 * C4);
    const bool worker = tg < 7;
    const int ntap = worker ? min(4, 27 - tg * 4) : 0;
    float acc[4][4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
    float bsum = 0.f;                                         // threads 0..C-1: bias gradient of channel tid

    const int total = a.B * a.chunks_per_sample;
    for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        const int b = chunk / a.chunks_per_sample;
        const int row0 = (chunk - b * a.chunks_per_sample) * WG3_ROWS;
        // ---- stage dz (prologue applied; rows beyond R are zero) ----
        if (tid < WG3_ROWS * C4) {
            const int r = tid / C4, q = tid - r * C4;
            float4 d = f4_zero();
            if (row0 + r < R) {
                const size_t off = ((size_t)b * R + row0 + r) * C + 4 * q;
                d = __ldg(reinterpret_cast<const float4*>(a.dz + off));
                if (a.dy_affine2) {
                    const float4 yy = __ldg(reinterpret_cast<const float4*>(a.y + off));
                    const float* P = a.P + (size_t)b * C + 4 * q;
                    const float* Q = a.Q + (size_t)b * C + 4 * q;
                    const float4 rc = a.Rc ? *reinterpret_cast<const float4*>(a.Rc + (size_t)b * C + 4 * q) : f4_zero();
                    d.x = fmaf(P[0], d.x, fmaf(Q[0], yy.x, rc.x));
                    d.y = fmaf(P[1], d.y, fmaf(Q[1], yy.y, rc.y));
                    d.z = fmaf(P[2], d.z, fmaf(Q[2], yy.z, rc.z));
                    d.w = fmaf(P[3], d.w, fmaf(Q[3], yy.w, rc.w));
                }
            }
            *reinterpret_cast<float4*>(dzs + r * C + 4 * q) = d;
        }
        // ---- stage the gathered input: item = (row, tap, channel quad) ----
        const float* xbase = a.x + (size_t)b * a.sample_stride;
        for (int e = tid; e < WG3_ROWS * 27 * C4; e += WG3_THREADS) {
            const int q = e % C4, rt = e / C4, tap = rt % 27, r = rt / 27;
            float4 v = f4_zero();
            const int row = row0 + r;
            if (row < R) {
                const int ow = row % a.Wo, qq = row / a.Wo, oh = qq % a.Ho, ot = qq / a.Ho;
                const int ti = 2 * ot - 1 + tap / 9, hi = 2 * oh - 1 + (tap / 3) % 3, wi = 2 * ow - 1 + tap % 3;
                if ((unsigned)ti < (unsigned)a.Ti && (unsigned)hi < (unsigned)a.Hi && (unsigned)wi < (unsigned)a.Wi) {
                    v = __ldg(reinterpret_cast<const float4*>(xbase + (((size_t)ti * a.Hi + hi) * a.Wi + wi) * C + 4 * q));
                    if (a.x_mode != CF_PRO_NONE) {
                        const float* xa = a.xa + (size_t)b * C + 4 * q;
                        const float* xb = a.xb + (size_t)b * C + 4 * q;
                        v.x = fmaf(xa[0], v.x, xb[0]); v.y = fmaf(xa[1], v.y, xb[1]);
                        v.z = fmaf(xa[2], v.z, xb[2]); v.w = fmaf(xa[3], v.w, xb[3]);
                        if (a.x_mode == CF_PRO_AFFINE_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    }
                }
            }
            *reinterpret_cast<float4*>(xs + (size_t)rt * C + 4 * q) = v;
        }
        __syncthreads();
        // ---- accumulate ----
        if (worker) {
#pragma unroll 2
            for (int r = 0; r < WG3_ROWS; ++r) {
                const float4 d = *reinterpret_cast<const float4*>(dzs + r * C + 4 * cog);
                const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (t < ntap) {
                        const float4 xv = *reinterpret_cast<const float4*>(xs + ((size_t)r * 27 + tg * 4 + t) * C + 4 * cig);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            acc[t][i][0] = fmaf(dd[i], xv.x, acc[t][i][0]);
                            acc[t][i][1] = fmaf(dd[i], xv.y, acc[t][i][1]);
                            acc[t][i][2] = fmaf(dd[i], xv.z, acc[t][i][2]);
                            acc[t][i][3] = fmaf(dd[i], xv.w, acc[t][i][3]);
                        }
                    }
                }
            }
        }
        if (a.dbias && tid < C) {
#pragma unroll 8
            for (int r = 0; r < WG3_ROWS; ++r) bsum += dzs[r * C + tid];
        }
        __syncthreads();
    }
    if (worker) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (t < ntap)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        atomicAdd(a.dw + (size_t)(4 * cog + i) * (C * 27) + (4 * cig + j) * 27 + tg * 4 + t, acc[t][i][j]);
    }
    if (a.dbias && tid < C) atomicAdd(a.dbias + tid, bsum);
}

// -1: not this kernel's problem (the caller falls through to the generic gathered weight gradient)
int cf_dense_s2_wgrad_try(const cf_pw_wgrad_args* a, cudaStream_t stream) {
    if (cf_env("CFNET_DENSE_WGRAD_OFF", 0)) return -1;
    const cf_geom& g = a->g;
    if (!a->gather_in) return -1;
    if (!(g.kt == 3 && g.kh == 3 && g.kw == 3 && g.st == 2 && g.sh == 2 && g.sw == 2 && g.pt == 1 && g.ph == 1 && g.pw == 1)) return -1;
    if (a->N != 24 || a->K != 24 * 27 || g.ch_stride != 1 || g.pos_stride != 24) return -1;
    if (a->dy_mode != CF_PRO_NONE && a->dy_mode != CF_PRO_AFFINE2) return -1;
    if (a->dy_mode == CF_PRO_AFFINE2 && !(a->dy2 && a->dy_a && a->dy_b)) return -1;
    if (a->x_mode != CF_PRO_NONE && a->x_mode != CF_PRO_AFFINE && a->x_mode != CF_PRO_AFFINE_RELU) return -1;
    if (a->x_mode != CF_PRO_NONE && !(a->x_a && a->x_b)) return -1;
    if ((g.sample_stride & 3) || ((uintptr_t)a->x & 15) || ((uintptr_t)a->dy & 15) || (a->dy2 && ((uintptr_t)a->dy2 & 15))) return -1;
    if (a->dy_c && ((uintptr_t)a->dy_c & 15)) return -1;
    const long long R = (long long)g.T * g.H * g.W;
    const long long cps = (R + WG3_ROWS - 1) / WG3_ROWS;
    if (cps * a->B > 0x7fffffffLL) return -1;
    Wg3Args w;
    w.dz = a->dy; w.y = a->dy2; w.P = a->dy_a; w.Q = a->dy_b; w.Rc = a->dy_c; w.x = a->x; w.xa = a->x_a; w.xb = a->x_b;
    w.dw = a->dw; w.dbias = a->dbias; w.B = a->B;
    w.To = g.T; w.Ho = g.H; w.Wo = g.W; w.Ti = g.Ti; w.Hi = g.Hi; w.Wi = g.Wi;
    w.sample_stride = g.sample_stride; w.dy_affine2 = a->dy_mode == CF_PRO_AFFINE2; w.x_mode = a->x_mode;
    w.chunks_per_sample = (int)cps;
    const size_t smem = (size_t)(WG3_ROWS * 24 + WG3_ROWS * 27 * 24) * sizeof(float);
    static CfOncePerDevice attr_done;
    if (attr_done.need()) {
        cudaError_t e = cudaFuncSetAttribute(dense3_s2_wgrad_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { cf_set_error("cf_dense_s2_wgrad: smem opt-in failed: %s", cudaGetErrorString(e)); return CF_ERR_CUDA; }
        attr_done.mark();
    }
    long long ctas = cps * a->B;
    if (ctas > 148 * 2) ctas = 148 * 2;
    cf_launch(dense3_s2_wgrad_kernel<24>, (unsigned)ctas, WG3_THREADS, smem, stream, w);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
