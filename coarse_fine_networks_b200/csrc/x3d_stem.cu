// Stem spatial convolution conv1_s (x3d_fine.py:210-215: Conv3d(3, 24, (1,3,3), stride (1,2,2), pad (0,1,1), no bias)
// reading the network's NCTHW clip (possibly a temporal window of a longer clip, through its strides) and writing
// channels-last rows: forward and weight gradient.
//
// In the general tap-gather GEMM (pw_conv_kernel / pw_wgrad_kernel<GATHER>, x3d_pw.cu) these two launches took 6.4 ms
// and 11.2 ms of the 156 ms step for the fine stream alone (K = 27 gathered scalars per row through the generic
// geometry code, 64-wide tiles for a 24 x 27 problem).  Here:
//   forward : one thread = one output position, 27 gathered inputs (coalesced along W across the warp) against the
//             [27][24] weights broadcast from shared memory, 24 accumulators, six 16-byte stores (96 contiguous bytes).
//   wgrad   : persistent CTAs stage tiles of 64 positions (dy rows [64][24], gathered x [64][28]) in shared memory;
//             thread = (4 channels x 4 taps) register tile of the 24 x 28 outer product, six thread groups split
//             the positions; 16 FMA per two 16-byte shared loads; one 24 x 27 atomic flush per CTA.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include <stdlib.h>

#define ST_CI 3
#define ST_CO 24
#define ST_K 27
#define ST_KP 28
#define ST_TILE 64

struct StemParams {
    int B, T, Hi, Wi, Ho, Wo;
    long long R;                     // rows per sample = T*Ho*Wo
    long long ch_stride, sample_stride, t_stride;   // element strides of the NCTHW input (t_stride = Hi*Wi)
};

__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                       const StemParams p) {
    cf_pdl_enter();
    __shared__ __align__(16) float ws[ST_K * ST_CO];             // [k][n]
    for (int i = threadIdx.x; i < ST_K * ST_CO; i += 256) {
        const int k = i / ST_CO, n = i - k * ST_CO;
        ws[i] = w[n * ST_K + k];
    }
    __syncthreads();
    const long long pos = (long long)blockIdx.x * 256 + threadIdx.x;
    if (pos >= p.R) return;
    const int b = blockIdx.y;
    const int wo = (int)(pos % p.Wo);
    const long long q = pos / p.Wo;
    const int ho = (int)(q % p.Ho);
    const int t = (int)(q / p.Ho);
    const float* xb = x + (long long)b * p.sample_stride + (long long)t * p.t_stride;
    float xin[ST_K];
#pragma unroll
    for (int c = 0; c < ST_CI; ++c)
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
            const int hi = 2 * ho - 1 + dh;
            const bool hv = (unsigned)hi < (unsigned)p.Hi;
            const float* row = xb + (long long)c * p.ch_stride + (long long)hi * p.Wi;
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                const int wi = 2 * wo - 1 + dw;
                xin[(c * 3 + dh) * 3 + dw] = (hv && (unsigned)wi < (unsigned)p.Wi) ? __ldg(row + wi) : 0.f;
            }
        }
    float4 acc[ST_CO / 4];
#pragma unroll
    for (int j = 0; j < ST_CO / 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < ST_K; ++k) {
        const float xv = xin[k];
#pragma unroll
        for (int j = 0; j < ST_CO / 4; ++j) {
            const float4 wv = *reinterpret_cast<const float4*>(ws + k * ST_CO + 4 * j);
            acc[j].x = fmaf(xv, wv.x, acc[j].x); acc[j].y = fmaf(xv, wv.y, acc[j].y);
            acc[j].z = fmaf(xv, wv.z, acc[j].z); acc[j].w = fmaf(xv, wv.w, acc[j].w);
        }
    }
    float4* dst = reinterpret_cast<float4*>(y + ((long long)b * p.R + pos) * ST_CO);
#pragma unroll
    for (int j = 0; j < ST_CO / 4; ++j) dst[j] = acc[j];
}

__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                         const StemParams p, long long total_tiles, int tiles_per_sample) {
    cf_pdl_enter();
    __shared__ __align__(16) float dys[ST_TILE * ST_CO];          // [pos][24]
    __shared__ __align__(16) float xgs[ST_TILE * ST_KP];          // [pos][28] (tap 27 = 0)
    __shared__ float red[6 * ST_CO * ST_KP];
    const int tid = threadIdx.x;
    const int n4 = tid % 6, k4 = (tid / 6) % 7, grp = tid / 42;   // grp 6 = the 4 spare threads (load only)
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lp = tid & (ST_TILE - 1);                           // the tile position this thread gathers for
    const int kq = tid >> 6;                                      // taps kq, kq + 4, ..., < 27
    if (tid < ST_TILE) xgs[tid * ST_KP + ST_K] = 0.f;             // padded tap
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_sample);
        const long long pos0 = (tile - (long long)b * tiles_per_sample) * ST_TILE;
        const int npos = (int)min((long long)ST_TILE, p.R - pos0);
        __syncthreads();                                          // previous tile consumed
        // dy rows: contiguous [npos][24]
        {
            const float4* src = reinterpret_cast<const float4*>(dy + ((long long)b * p.R + pos0) * ST_CO);
            for (int i = tid; i < ST_TILE * ST_CO / 4; i += 256)
                reinterpret_cast<float4*>(dys)[i] = (i < npos * (ST_CO / 4)) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // gathered inputs: this thread's position, its share of the 27 taps
        {
            const long long pos = pos0 + lp;
            const bool pv = lp < npos;
            const int wo = (int)(pos % p.Wo);
            const long long q = pos / p.Wo;
            const int ho = (int)(q % p.Ho);
            const int t = (int)(q / p.Ho);
            const float* xb = x + (long long)b * p.sample_stride + (long long)t * p.t_stride;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int k = kq + 4 * j;
                if (k < ST_K) {
                    const int c = k / 9, r = k - c * 9, dh = r / 3, dwi = r - dh * 3;
                    const int hi = 2 * ho - 1 + dh, wi = 2 * wo - 1 + dwi;
                    const bool v = pv && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi;
                    xgs[lp * ST_KP + k] = v ? __ldg(xb + (long long)c * p.ch_stride + (long long)hi * p.Wi + wi) : 0.f;
                }
            }
        }
        __syncthreads();
        if (grp < 6) {
            for (int pp = grp; pp < ST_TILE; pp += 6) {
                const float4 d = *reinterpret_cast<const float4*>(dys + pp * ST_CO + 4 * n4);
                const float4 xv = *reinterpret_cast<const float4*>(xgs + pp * ST_KP + 4 * k4);
                const float dd[4] = {d.x, d.y, d.z, d.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dd[i], xx[j], acc[i][j]);
            }
        }
    }
    __syncthreads();
    if (grp < 6) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(grp * ST_CO + 4 * n4 + i) * ST_KP + 4 * k4 + j] = acc[i][j];
    }
    __syncthreads();
    for (int i = tid; i < ST_CO * ST_KP; i += 256) {
        const int n = i / ST_KP, k = i - n * ST_KP;
        if (k < ST_K) {
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < 6; ++g) s += red[g * ST_CO * ST_KP + i];
            atomicAdd(dw + n * ST_K + k, s);
        }
    }
}

// ---------------------------------------------------------------------------------------
static bool stem_geom_ok(const cf_geom& g, int K, int N) {
    return K == ST_K && N == ST_CO && g.kt == 1 && g.kh == 3 && g.kw == 3 && g.st == 1 && g.sh == 2 && g.sw == 2 && g.pt == 0 && g.ph == 1 &&
           g.pw == 1 && g.pos_stride == 1 && g.T == g.Ti && g.H == (g.Hi - 1) / 2 + 1 && g.W == (g.Wi - 1) / 2 + 1;
}
static bool stem_enabled() {
    return !cf_env("CFNET_STEM_OFF", 0);
}
static StemParams stem_params(int B, const cf_geom& g) {
    StemParams p;
    p.B = B; p.T = g.T; p.Hi = g.Hi; p.Wi = g.Wi; p.Ho = g.H; p.Wo = g.W;
    p.R = (long long)g.T * g.H * g.W;
    p.ch_stride = g.ch_stride; p.sample_stride = g.sample_stride;
    p.t_stride = (long long)g.Hi * g.Wi;
    return p;
}

// returns CF_OK when launched, -1 when the call is not the stem convolution
int cf_stem_fwd_try(const cf_pw_args* a, cudaStream_t stream) {
    if (!stem_enabled() || !a->gather_in || !stem_geom_ok(a->g, a->K, a->N)) return -1;
    if (a->pro_mode != CF_PRO_NONE || a->epi_mode != CF_EPI_NONE || a->stats_mode != CF_STATS_NONE || a->bias || a->accumulate ||
        a->w_sn != ST_K || a->w_sk != 1 || (((uintptr_t)a->y) & 15))
        return -1;
    StemParams p = stem_params(a->B, a->g);
    if (a->B > 65535) return -1;
    dim3 grid((unsigned)cf_cdiv64(p.R, 256), (unsigned)a->B);
    cf_launch(stem_fwd_kernel, grid, 256, 0, stream, a->x, a->w, a->y, p);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_stem_wgrad_try(const cf_pw_wgrad_args* a, cudaStream_t stream) {
    if (!stem_enabled() || !a->gather_in || !stem_geom_ok(a->g, a->K, a->N)) return -1;
    if (a->dy_mode != CF_PRO_NONE || a->x_mode != CF_PRO_NONE || a->dbias || (((uintptr_t)a->dy) & 15)) return -1;
    StemParams p = stem_params(a->B, a->g);
    const int tps = (int)cf_cdiv64(p.R, ST_TILE);
    const long long total = (long long)a->B * tps;
    int nsm = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
    long long grid = (long long)nsm * 4;
    if (grid > total) grid = total;
    cf_launch(stem_wgrad_kernel, (unsigned)grid, 256, 0, stream, a->dy, a->x, a->dw, p, total, tps);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
