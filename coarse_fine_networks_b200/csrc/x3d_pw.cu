// Convolution-as-GEMM kernels of the X3D stacks, fp32 CUDA-core path (sm_100a).
//
//   cf_pw_conv   y[b,r,n] = epi( sum_k pro(x[b,gather(r,k)]) * w[n,k] )     forward + data gradient
//   cf_pw_wgrad  dw[n,k] += sum_{b,r} pro(dy[b,r,n]) * pro(x[b,gather(r,k)])  weight gradient
//
// Reference call sites: conv1x1x1 (x3d_fine.py:100-105) used at :149,166,286,356,370; fc2
// (nn.Linear, :380); conv1_s (:210-215); pool_1.conv1-3 (x3d_coarse.py:362-366).  BatchNorm
// (train-mode batch statistics, x3d_fine.py:51-62), ReLU, SE gate and Swish are folded into the
// prologue (per-(sample,channel) affine tables) and the epilogue (statistics via double atomics).
//
// Tiling: one CTA = 128 rows of ONE sample x BN output channels, K in chunks of 16; 256 threads,
// 8 x (BN/16) register tile per thread.  Row tiles never straddle samples, so the per-sample
// tables live in shared memory and the statistics need one atomic per (CTA, channel).
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

#define PW_BM 128
#define PW_BK 16
#define PW_LDA (PW_BM + 4)

__device__ __forceinline__ float cf_swish(float v) { return v * cf_sigmoid(v); }
__device__ __forceinline__ float cf_dswish(float v) {
    float s = cf_sigmoid(v);
    return s * (1.0f + v * (1.0f - s));
}

__device__ __forceinline__ float apply_pro(int mode, float x, float x2, float a, float b, float c) {
    switch (mode) {
        case CF_PRO_AFFINE: return fmaf(a, x, b);
        case CF_PRO_AFFINE_RELU: return fmaxf(fmaf(a, x, b), 0.f);
        case CF_PRO_AFFINE_SWISH: return cf_swish(fmaf(a, x, b));
        case CF_PRO_AFFINE2: return fmaf(a, x, fmaf(b, x2, c));
        default: return x;
    }
}

__device__ __forceinline__ unsigned pack_pos(const cf_geom& g, int r) {
    int w = r % g.W;
    int q = r / g.W;
    int h = q % g.H;
    int t = q / g.H;
    return ((unsigned)t << 22) | ((unsigned)h << 11) | (unsigned)w;
}

// offset (floats, within the sample) of tap `tap`, channel c seen from dense position `pos`; -1 if outside
__device__ __forceinline__ long long gathered_off(const cf_geom& g, unsigned pos, int tap, int c) {
    int t = pos >> 22, h = (pos >> 11) & 2047, w = pos & 2047;
    int dw = tap % g.kw;
    int q = tap / g.kw;
    int dh = q % g.kh;
    int dt = q / g.kh;
    int ti = t * g.st - g.pt + dt, hi = h * g.sh - g.ph + dh, wi = w * g.sw - g.pw + dw;
    if ((unsigned)ti >= (unsigned)g.Ti || (unsigned)hi >= (unsigned)g.Hi || (unsigned)wi >= (unsigned)g.Wi) return -1;
    return ((long long)(ti * g.Hi + hi) * g.Wi + wi) * g.pos_stride + (long long)c * g.ch_stride;
}

template <int BN_, bool GATHER>
__global__ void __launch_bounds__(256) pw_conv_kernel(const cf_pw_args a, int tiles_per_sample, int R) {
    cf_pdl_enter();
    constexpr int TN = BN_ / 16;
    constexpr int LDB = BN_ + 4;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                       // [PW_BK][PW_LDA]
    float* Bs = As + PW_BK * PW_LDA;        // [PW_BK][LDB]
    float* tab = Bs + PW_BK * LDB;          // 3 * Cin
    const int tid = threadIdx.x;
    const int b = blockIdx.x / tiles_per_sample;
    const int r0 = (blockIdx.x - b * tiles_per_sample) * PW_BM;
    const int n0 = blockIdx.y * BN_;
    const int K = a.K, N = a.N;
    const int taps = a.g.kt * a.g.kh * a.g.kw;
    const int cin = GATHER ? K / taps : K;
    const int pro = a.pro_mode;
    if (pro != CF_PRO_NONE) {
        for (int i = tid; i < cin; i += 256) {
            tab[i] = a.pro_a[(size_t)b * cin + i];
            tab[cin + i] = a.pro_b ? a.pro_b[(size_t)b * cin + i] : 0.f;
            tab[2 * cin + i] = a.pro_c ? a.pro_c[(size_t)b * cin + i] : 0.f;
        }
    }
    const float* xb = a.x + (GATHER ? (long long)b * a.g.sample_stride : (long long)b * R * K);
    const float* x2b = a.x2 ? a.x2 + (long long)b * R * K : nullptr;

    // rows this thread loads into the A tile: (tid>>4) + 16*i
    unsigned posr[8];
    bool rv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int r = r0 + (tid >> 4) + 16 * i;
        rv[i] = r < R;
        posr[i] = GATHER ? pack_pos(a.g, rv[i] ? r : 0) : (unsigned)(rv[i] ? r : 0);
    }
    const int kk = tid & 15;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    __syncthreads();

    for (int k0 = 0; k0 < K; k0 += PW_BK) {
        {   // ---- A tile (prologue applied; zero outside rows / K / volume)
            int k = k0 + kk;
            bool kv = k < K;
            int c = GATHER ? k / taps : k;
            int tap = GATHER ? k - c * taps : 0;
            float pa = 1.f, pb = 0.f, pc = 0.f;
            if (pro != CF_PRO_NONE && kv) { pa = tab[c]; pb = tab[cin + c]; pc = tab[2 * cin + c]; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = 0.f;
                if (rv[i] && kv) {
                    long long off = GATHER ? gathered_off(a.g, posr[i], tap, c) : (long long)posr[i] * K + k;
                    if (off >= 0) {
                        float xv = __ldg(xb + off);
                        float x2v = (pro == CF_PRO_AFFINE2) ? __ldg(x2b + off) : 0.f;
                        v = apply_pro(pro, xv, x2v, pa, pb, pc);
                    }
                }
                As[kk * PW_LDA + (tid >> 4) + 16 * i] = v;
            }
        }
        {   // ---- W tile
#pragma unroll
            for (int i = 0; i < BN_ / 16; ++i) {
                int e = tid + i * 256;
                int kk2, nn;
                if (a.w_sk == 1) { kk2 = e & 15; nn = e >> 4; } else { nn = e % BN_; kk2 = e / BN_; }
                int k = k0 + kk2, n = n0 + nn;
                float v = (k < K && n < N) ? __ldg(a.w + (long long)n * a.w_sn + (long long)k * a.w_sk) : 0.f;
                Bs[kk2 * LDB + nn] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PW_BK; ++q) {
            float av[8], bv[TN];
            float4 a0 = *reinterpret_cast<const float4*>(As + q * PW_LDA + ty * 8);
            float4 a1 = *reinterpret_cast<const float4*>(As + q * PW_LDA + ty * 8 + 4);
            av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
            av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
            if constexpr (TN == 8) {
                float4 b0 = *reinterpret_cast<const float4*>(Bs + q * LDB + tx * 8);
                float4 b1 = *reinterpret_cast<const float4*>(Bs + q * LDB + tx * 8 + 4);
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            } else if constexpr (TN == 4) {
                float4 b0 = *reinterpret_cast<const float4*>(Bs + q * LDB + tx * 4);
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
            } else {
                float2 b0 = *reinterpret_cast<const float2*>(Bs + q * LDB + tx * 2);
                bv[0] = b0.x; bv[1] = b0.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue
    const int epi = a.epi_mode, smode = a.stats_mode;
    float ea[TN], eb[TN], bi[TN], s1[TN], s2[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        int n = n0 + tx * TN + j;
        bool nv = n < N;
        ea[j] = (nv && a.epi_a) ? a.epi_a[(size_t)b * N + n] : 1.f;
        eb[j] = (nv && a.epi_b) ? a.epi_b[(size_t)b * N + n] : 0.f;
        bi[j] = (nv && a.bias) ? a.bias[n] : 0.f;
        s1[j] = 0.f;
        s2[j] = 0.f;
    }
    const int ntaps_out = a.scatter_out ? taps : 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int r = r0 + ty * 8 + i;
        if (r >= R) continue;
        long long dense = ((long long)b * R + r) * N;
        unsigned pos = a.scatter_out ? pack_pos(a.g, r) : 0u;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float v = acc[i][j] + bi[j];
            float auxv = 0.f;
            if ((epi >= CF_EPI_DRELU && epi <= CF_EPI_ADD_AUX) || epi == CF_EPI_AFFINE_ADD_RELU || smode == CF_STATS_SUM_AUX)
                auxv = __ldg(a.aux + dense + n);
            if (epi == CF_EPI_RELU) v = fmaxf(v, 0.f);
            else if (epi == CF_EPI_DRELU) v = (fmaf(ea[j], auxv, eb[j]) > 0.f) ? v : 0.f;
            else if (epi == CF_EPI_DSWISH) v *= cf_dswish(fmaf(ea[j], auxv, eb[j]));
            else if (epi == CF_EPI_ADD_AUX) v += auxv;
            else if (epi == CF_EPI_SIGMOID) v = cf_sigmoid(v);
            else if (epi == CF_EPI_AFFINE) v = fmaf(ea[j], v, eb[j]);
            else if (epi == CF_EPI_AFFINE_ADD_RELU) v = fmaxf(fmaf(ea[j], v, eb[j]) + auxv, 0.f);
            s1[j] += v;
            s2[j] += (smode == CF_STATS_SUM_AUX) ? v * auxv : v * v;
            if (!a.scatter_out) {
                float* yp = a.y + dense + n;
                *yp = a.accumulate ? (*yp + v) : v;
            } else {
                int c = n / ntaps_out;
                long long off = gathered_off(a.g, pos, n - c * ntaps_out, c);
                if (off >= 0) {
                    float* yp = a.y + (long long)b * a.g.sample_stride + off;
                    if (ntaps_out > 1) atomicAdd(yp, v);
                    else *yp = a.accumulate ? (*yp + v) : v;
                }
            }
        }
    }
    if (smode != CF_STATS_NONE) {
        __syncthreads();
        float* red1 = smem;                  // [16][BN_]
        float* red2 = smem + 16 * BN_;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            red1[ty * BN_ + tx * TN + j] = s1[j];
            red2[ty * BN_ + tx * TN + j] = s2[j];
        }
        __syncthreads();
        if (tid < BN_ && n0 + tid < N) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) { t1 += red1[q * BN_ + tid]; t2 += red2[q * BN_ + tid]; }
            double* st = a.stats + ((size_t)b * N + n0 + tid) * 2;
            atomicAdd(st, (double)t1);
            atomicAdd(st + 1, (double)t2);
        }
    }
}

// ---------------------------------------------------------------------------------------
// weight gradient: CTA tile 64 (n) x 64 (k), rows of one sample split across grid.z
// ---------------------------------------------------------------------------------------
#define WG_T 64
#define WG_LD (WG_T + 4)
#define WG_BM 16

template <bool GATHER>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const cf_pw_wgrad_args a, int splits, int rows_per_split, int R) {
    cf_pdl_enter();
    __shared__ __align__(16) float Ds[WG_BM * WG_LD];
    __shared__ __align__(16) float Xs[WG_BM * WG_LD];
    const int tid = threadIdx.x;
    const int b = blockIdx.z / splits;
    const int sp = blockIdx.z - b * splits;
    const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
    const int K = a.K, N = a.N;
    const int taps = a.g.kt * a.g.kh * a.g.kw;
    const int cin = GATHER ? K / taps : K;
    const int rbeg = sp * rows_per_split;
    const int rend = min(R, rbeg + rows_per_split);
    const int col = tid & 63;            // channel (n or k) this thread loads: fixed over the loop
    const int mrow = tid >> 6;           // 0..3, rows mrow + 4*i
    // prologue constants
    const int n_ld = n0 + col, k_ld = k0 + col;
    const bool nv = n_ld < N, kv = k_ld < K;
    float da = 1.f, db = 0.f, dc = 0.f, xa = 1.f, xbb = 0.f;
    if (a.dy_mode != CF_PRO_NONE && nv) {
        da = a.dy_a[(size_t)b * N + n_ld];
        db = a.dy_b ? a.dy_b[(size_t)b * N + n_ld] : 0.f;
        dc = a.dy_c ? a.dy_c[(size_t)b * N + n_ld] : 0.f;
    }
    const int c_ld = GATHER ? k_ld / taps : k_ld;
    const int tap_ld = GATHER ? k_ld - c_ld * taps : 0;
    if (a.x_mode != CF_PRO_NONE && kv) {
        xa = a.x_a[(size_t)b * cin + c_ld];
        xbb = a.x_b ? a.x_b[(size_t)b * cin + c_ld] : 0.f;
    }
    const float* dyb = a.dy + (long long)b * R * N;
    const float* dy2b = a.dy2 ? a.dy2 + (long long)b * R * N : nullptr;
    const float* xb = a.x + (GATHER ? (long long)b * a.g.sample_stride : (long long)b * R * K);
    const int ty = tid >> 4, tx = tid & 15;       // 4 n-rows x 4 k-cols per thread
    float acc[4][4];
    float dsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int m0 = rbeg; m0 < rend; m0 += WG_BM) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int mm = mrow + 4 * i;
            int r = m0 + mm;
            float dv = 0.f, xv = 0.f;
            if (r < rend) {
                if (nv) {
                    float d1 = __ldg(dyb + (long long)r * N + n_ld);
                    float d2 = (a.dy_mode == CF_PRO_AFFINE2) ? __ldg(dy2b + (long long)r * N + n_ld) : 0.f;
                    dv = apply_pro(a.dy_mode, d1, d2, da, db, dc);
                }
                if (kv) {
                    long long off = GATHER ? gathered_off(a.g, pack_pos(a.g, r), tap_ld, c_ld) : (long long)r * K + k_ld;
                    if (off >= 0) xv = apply_pro(a.x_mode, __ldg(xb + off), 0.f, xa, xbb, 0.f);
                }
            }
            Ds[mm * WG_LD + col] = dv;
            Xs[mm * WG_LD + col] = xv;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < WG_BM; ++mm) {
            float4 d = *reinterpret_cast<const float4*>(Ds + mm * WG_LD + ty * 4);
            float4 x = *reinterpret_cast<const float4*>(Xs + mm * WG_LD + tx * 4);
            float dvv[4] = {d.x, d.y, d.z, d.w};
            float xvv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dsum[i] += dvv[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dvv[i], xvv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(a.dw + (size_t)n * K + k, acc[i][j]);
        }
        if (a.dbias && blockIdx.y == 0 && tx == 0) atomicAdd(a.dbias + n, dsum[i]);
    }
}

// ---------------------------------------------------------------------------------------
int cf_pw_conv_tc(const cf_pw_args* a, cudaStream_t stream);      // x3d_pw_tc2.cu
int cf_pw_wgrad_tc(const cf_pw_wgrad_args* a, cudaStream_t stream);   // x3d_pw_wgrad_tc.cu (-1: not eligible)
int cf_dense_s2_dgrad_try(const cf_pw_args* a, cudaStream_t stream);  // x3d_dense3s2.cu (-1: not eligible)
int cf_dense_s2_fwd_try(const cf_pw_args* a, cudaStream_t stream);    // x3d_dense3s2.cu (-1: not eligible)
int cf_dense_s2_wgrad_try(const cf_pw_wgrad_args* a, cudaStream_t stream);   // x3d_dense3s2.cu (-1: not eligible)
int cf_stem_fwd_try(const cf_pw_args* a, cudaStream_t stream);         // x3d_stem.cu (-1: not the stem conv)
int cf_stem_wgrad_try(const cf_pw_wgrad_args* a, cudaStream_t stream);

extern "C" size_t cf_sizeof_pw_args(void) { return sizeof(cf_pw_args); }
size_t cf_sizeof_pw_wgrad_args(void) { return sizeof(cf_pw_wgrad_args); }

// the packed (t,h,w) position (10+11+11 bits) is only used by the gathered / scattered paths
static bool geom_ok(const cf_geom& g, bool packed) {
    if (!(g.T > 0 && g.H > 0 && g.W > 0 && (long long)g.T * g.H * g.W < (1LL << 31))) return false;
    if (!packed) return true;
    return g.T < 1024 && g.H < 2048 && g.W < 2048 && g.kt > 0 && g.kh > 0 && g.kw > 0 && g.st > 0 && g.sh > 0 && g.sw > 0;
}

template <int BN_, bool GATHER>
static int launch_pw(const cf_pw_args* a, int R, cudaStream_t stream) {
    int tps = cf_cdiv(R, PW_BM);
    int taps = a->g.kt * a->g.kh * a->g.kw;
    int cin = GATHER ? a->K / taps : a->K;
    size_t smem = (size_t)(PW_BK * PW_LDA + PW_BK * (BN_ + 4) + 3 * cin) * sizeof(float);
    size_t red = (size_t)2 * 16 * BN_ * sizeof(float);
    if (smem < red) smem = red;
    static CfOncePerDevice attr_done;
    if (attr_done.need()) {
        cudaFuncSetAttribute(pw_conv_kernel<BN_, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_done.mark();
    }
    if (smem > 100 * 1024) { cf_set_error("cf_pw_conv: K too large for the table cache"); return CF_ERR_ARG; }
    dim3 grid((unsigned)(tps * a->B), (unsigned)cf_cdiv(a->N, BN_));
    cf_launch(pw_conv_kernel<BN_, GATHER>, grid, 256, smem, stream, *a, tps, R);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_pw_conv(const cf_pw_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->x && a->w && a->y, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->K > 0 && a->N > 0 && geom_ok(a->g, a->gather_in || a->scatter_out), "bad shape");
    CF_CHECK_ARG(a->pro_mode == CF_PRO_NONE || a->pro_a, "prologue tables missing");
    CF_CHECK_ARG(a->pro_mode != CF_PRO_AFFINE2 || (a->x2 && !a->gather_in), "AFFINE2 needs a dense second input");
    CF_CHECK_ARG(((a->epi_mode < CF_EPI_DRELU || a->epi_mode > CF_EPI_ADD_AUX) && a->epi_mode != CF_EPI_AFFINE_ADD_RELU &&
                  a->stats_mode != CF_STATS_SUM_AUX) || a->aux, "aux tensor missing");
    CF_CHECK_ARG((a->epi_mode != CF_EPI_AFFINE && a->epi_mode != CF_EPI_AFFINE_ADD_RELU) || (a->epi_a && a->epi_b), "epilogue tables missing");
    CF_CHECK_ARG(a->stats_mode == CF_STATS_NONE || a->stats, "stats buffer missing");
    CF_CHECK_ARG(!(a->gather_in && a->scatter_out), "gather_in and scatter_out are exclusive");
    int taps = a->g.kt * a->g.kh * a->g.kw;
    CF_CHECK_ARG(!a->gather_in || a->K % taps == 0, "K must be channels*taps");
    CF_CHECK_ARG(!a->scatter_out || a->N % taps == 0, "N must be channels*taps");
    int R = a->g.T * a->g.H * a->g.W;
    if (a->wpack && !a->gather_in && !a->scatter_out) return cf_pw_conv_tc(a, stream);
    if (a->wpack && taps == 1 && a->g.ch_stride == 1 && a->g.pt == 0 && a->g.ph == 0 && a->g.pw == 0 &&
        (a->gather_in ? (a->g.pos_stride == a->K && a->pro_mode != CF_PRO_AFFINE2)
                      : (a->g.pos_stride == a->N && a->accumulate && a->stats_mode == CF_STATS_NONE && !a->aux))) {
        int rct = cf_pw_conv_tc(a, stream);                  // strided 1x1x1 conv (downsample branch): row gather / scatter in the GEMM
        if (rct >= 0) return rct;
    }
    if (a->scatter_out) {
        int rcd = cf_dense_s2_dgrad_try(a, stream);          // pool_1.conv1/conv2 data gradient: gather form, no atomics
        if (rcd >= 0) return rcd;
    }
    if (a->gather_in) {
        int rcf = cf_dense_s2_fwd_try(a, stream);            // pool_1.conv1/conv2 forward: direct kernel
        if (rcf >= 0) return rcf;
        int rcs = cf_stem_fwd_try(a, stream);               // conv1_s: specialised kernel
        if (rcs >= 0) return rcs;
        if (a->N <= 32) return launch_pw<32, true>(a, R, stream);
        if (a->N <= 64) return launch_pw<64, true>(a, R, stream);
        return launch_pw<128, true>(a, R, stream);
    }
    if (a->N <= 32) return launch_pw<32, false>(a, R, stream);
    if (a->N <= 64) return launch_pw<64, false>(a, R, stream);
    return launch_pw<128, false>(a, R, stream);
}

// dbias[n] += sum_rows dy[row,n] for a dense [rows,N] gradient: the bias half of a weight gradient whose GEMM half runs
// on the tensor-core kernel.  CTA = 32 columns x a chunk of rows, 8 row lanes, 128-byte row segments per warp.
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ dbias, long long rows,
                                                        int N, int chunk) {
    cf_pdl_enter();
    __shared__ float part[8][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lx;
    const long long r0 = (long long)blockIdx.y * chunk;
    const long long r1 = r0 + chunk < rows ? r0 + chunk : rows;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (col < N) {
        const float* p = dy + col;
        long long r = r0 + ly;
        for (; r + 24 < r1; r += 32) {
            a0 += __ldg(p + r * N);
            a1 += __ldg(p + (r + 8) * N);
            a2 += __ldg(p + (r + 16) * N);
            a3 += __ldg(p + (r + 24) * N);
        }
        for (; r < r1; r += 8) a0 += __ldg(p + r * N);
    }
    part[ly][lx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ly == 0 && col < N) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += part[q][lx];
        atomicAdd(dbias + col, s);
    }
}

extern "C" int cf_pw_wgrad(const cf_pw_wgrad_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->dy && a->x && a->dw, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->K > 0 && a->N > 0 && geom_ok(a->g, a->gather_in != 0), "bad shape");
    CF_CHECK_ARG(a->dy_mode == CF_PRO_NONE || a->dy_a, "dy tables missing");
    CF_CHECK_ARG(a->dy_mode != CF_PRO_AFFINE2 || a->dy2, "AFFINE2 needs dy2");
    CF_CHECK_ARG(a->x_mode == CF_PRO_NONE || a->x_a, "x tables missing");
    int taps = a->g.kt * a->g.kh * a->g.kw;
    CF_CHECK_ARG(!a->gather_in || a->K % taps == 0, "K must be channels*taps");
    {
        int rc = cf_pw_wgrad_tc(a, stream);                 // dense problems: tensor cores
        if (rc >= 0) return rc;
        if (a->dbias && a->dy_mode == CF_PRO_NONE) {        // GEMM half on the tensor cores, bias half as a column sum
            cf_pw_wgrad_args nb = *a;
            nb.dbias = nullptr;
            rc = cf_pw_wgrad_tc(&nb, stream);
            if (rc > 0) return rc;
            if (rc == 0) {
                const long long rows = (long long)a->B * a->g.T * a->g.H * a->g.W;
                const int ct = cf_cdiv(a->N, 32);
                long long want = cf_cdiv64(148 * 4, ct);
                int chunk = (int)cf_cdiv64(rows, want < 1 ? 1 : want);
                chunk = chunk < 256 ? 256 : ((chunk + 31) / 32) * 32;
                dim3 grid((unsigned)ct, (unsigned)cf_cdiv64(rows, chunk));
                cf_launch(bias_grad_kernel, grid, 256, 0, stream, a->dy, a->dbias, rows, a->N, chunk);
                CF_COUNT_LAUNCH(1);
                CF_CHECK_LAUNCH();
                return CF_OK;
            }
        }
        rc = cf_stem_wgrad_try(a, stream);                  // conv1_s: specialised kernel
        if (rc >= 0) return rc;
        rc = cf_dense_s2_wgrad_try(a, stream);              // pool_1.conv1/conv2: staged rows, register-resident result
        if (rc >= 0) return rc;
    }
    int R = a->g.T * a->g.H * a->g.W;
    int nt = cf_cdiv(a->N, WG_T), kt = cf_cdiv(a->K, WG_T);
    int want = cf_cdiv(148 * 3, nt * kt * a->B);
    int maxs = cf_cdiv(R, 4 * WG_BM);
    int splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    if (splits < 1) splits = 1;
    int rps = cf_cdiv(R, splits);
    rps = cf_cdiv(rps, WG_BM) * WG_BM;
    splits = cf_cdiv(R, rps);
    CF_CHECK_ARG((long long)a->B * splits <= 65535, "grid.z overflow");
    dim3 grid((unsigned)nt, (unsigned)kt, (unsigned)(a->B * splits));
    if (a->gather_in) cf_launch(pw_wgrad_kernel<true>, grid, 256, 0, stream, *a, splits, rps, R);
    else cf_launch(pw_wgrad_kernel<false>, grid, 256, 0, stream, *a, splits, rps, R);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

