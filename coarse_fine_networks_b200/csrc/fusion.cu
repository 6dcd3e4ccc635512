// Multi-stage Fusion kernels (sm_100a): Gaussian temporal alignment, self-attention-filtered
// aligned aggregation (RewightLayer), scale/shift modulation (FiLM), nearest replication.
//
// Reference semantics (paths relative to the reference repo):
//   Gaussian.forward ............. x3d_coarse.py:256-286   (cf_gaussian_*)
//   RewightLayer aggregation ..... x3d_coarse.py:221-225   (cf_rewight_agg_*)
//   x*m + c ...................... x3d_coarse.py:664-679,721 (cf_film_*)
//   adaptive_max_pool2d as a nearest up-sampler .. :214,315,322 (cf_nearest_up*)
//
// The reference builds a 6-D [B,C,Tf,Tl,h,w] broadcast product at the *up-sampled* resolution
// (655 MB per clip for rw2 at Tf=128).  Here the contraction over Tf runs at the 7x7 base
// resolution of the fine features, per (sample, pixel), straight from the channels-last
// feature rows: x is read once, nothing 6-D is ever materialised.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

// ---------------------------------------------------------------------------------------
// K13 Gaussian: one warp per (b,k)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float gauss_f(float t, float mu, float den) {
    float d = __fsub_rn(t, mu);
    return expf(-__fdiv_rn(__fmul_rn(d, d), den));
}

__global__ void gaussian_fwd_kernel(const float* __restrict__ cdf, const float* __restrict__ start,
                                    const float* __restrict__ mask, float* __restrict__ gx, int B, int Tf, int Tl,
                                    float tx, float ratio) {
    int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= B * Tl) return;
    int b = wid / Tl, k = wid - b * Tl;
    float s = 0.f;
    for (int t = lane; t < Tf; t += 32) s += mask[(size_t)b * Tf + t];
    s = warp_sum(s);
    float sd = 0.125f * s;
    float den = __fadd_rn(__fmul_rn(2.0f, __fmul_rn(sd, sd)), 1e-16f);
    float mu = __fdiv_rn(__fadd_rn(__fmul_rn(cdf[(size_t)b * Tl + k], tx), start[b]), ratio);
    float m = 0.f;
    for (int t = lane; t < Tf; t += 32) m = fmaxf(m, gauss_f((float)t, mu, den));
    m = warp_max(m);
    float dn = __fadd_rn(m, 1e-16f);
    for (int t = lane; t < Tf; t += 32) gx[((size_t)b * Tf + t) * Tl + k] = __fdiv_rn(gauss_f((float)t, mu, den), dn);
}

// out_t = f_t / (M + e), M = f_{t*}:  dL/dmu = sum_t dout_t f'_t/(M+e) - (sum_t dout_t f_t) f'_{t*}/(M+e)^2,
// f'_t = f_t * 2 (t - mu) / den;  dcdf = dL/dmu * tx / ratio
__global__ void gaussian_bwd_kernel(const float* __restrict__ cdf, const float* __restrict__ start,
                                    const float* __restrict__ mask, const float* __restrict__ dgx,
                                    float* __restrict__ dcdf, int B, int Tf, int Tl, float tx, float ratio) {
    int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= B * Tl) return;
    int b = wid / Tl, k = wid - b * Tl;
    float s = 0.f;
    for (int t = lane; t < Tf; t += 32) s += mask[(size_t)b * Tf + t];
    s = warp_sum(s);
    float sd = 0.125f * s;
    float den = __fadd_rn(__fmul_rn(2.0f, __fmul_rn(sd, sd)), 1e-16f);
    float mu = __fdiv_rn(__fadd_rn(__fmul_rn(cdf[(size_t)b * Tl + k], tx), start[b]), ratio);
    float m = -1.f;
    int targ = 0;
    for (int t = lane; t < Tf; t += 32) {
        float f = gauss_f((float)t, mu, den);
        if (f > m) { m = f; targ = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {        // arg-max, first index wins ties (torch.max)
        float om = __shfl_xor_sync(0xffffffffu, m, o);
        int ot = __shfl_xor_sync(0xffffffffu, targ, o);
        if (om > m || (om == m && ot < targ)) { m = om; targ = ot; }
    }
    float dn = m + 1e-16f;
    float s1 = 0.f, s2 = 0.f;
    for (int t = lane; t < Tf; t += 32) {
        float f = gauss_f((float)t, mu, den);
        float go = dgx[((size_t)b * Tf + t) * Tl + k];
        s1 += go * f * 2.0f * ((float)t - mu) / den;
        s2 += go * f;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        float fpm = m * 2.0f * ((float)targ - mu) / den;
        float dmu = s1 / dn - s2 * fpm / (dn * dn);
        dcdf[(size_t)b * Tl + k] += dmu * tx / ratio;
    }
}

// ---------------------------------------------------------------------------------------
// K14 aligned aggregation, forward.  grid (B*P, ceil(C/32)), block (32 channels, 8 t-slices)
// smem: wgt[Tf][Tl] | den[Tl] | red[8][KC][32]
// ---------------------------------------------------------------------------------------
#define RW_KC 17

__global__ void __launch_bounds__(256) rewight_agg_fwd_kernel(const cf_rewight_args a) {
    extern __shared__ __align__(16) float sm[];
    const int Tf = a.Tf, Tl = a.Tl, P = a.P, C = a.C;
    float* wgt = sm;                         // [Tf][Tl]
    float* dens = wgt + (size_t)Tf * Tl;     // [Tl]
    float* red = dens + ((Tl + 3) & ~3);     // [8][RW_KC][32]
    const int b = blockIdx.x / P, p = blockIdx.x - b * P;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const float* att = a.att + (size_t)b * Tf * P + p;
    const float* gx = a.gx + (size_t)b * Tf * Tl;
    const float* mask = a.mask + (size_t)b * Tf;
    for (int i = tid; i < Tf * Tl; i += 256) {
        int t = i / Tl;
        wgt[i] = __ldg(att + (size_t)t * P) * __ldg(gx + i) * __ldg(mask + t);     // A * mask
    }
    __syncthreads();
    for (int k = ty; k < Tl; k += 8) {
        float s = 0.f;
        for (int t = tx; t < Tf; t += 32) s += wgt[t * Tl + k];
        s = warp_sum(s);
        if (tx == 0) {
            float d = s + 1e-6f;
            dens[k] = d;
            if (blockIdx.y == 0) a.den[((size_t)b * Tl + k) * P + p] = d;
        }
    }
    __syncthreads();
    for (int i = tid; i < Tf * Tl; i += 256) wgt[i] = wgt[i] / dens[i % Tl];
    __syncthreads();
    const int c = blockIdx.y * 32 + tx;
    const bool cv = c < C;
    const float* xs = a.x + ((size_t)b * Tf * P + p) * C + (cv ? c : 0);
    for (int k0 = 0; k0 < Tl; k0 += RW_KC) {
        const int kn = min(RW_KC, Tl - k0);
        float acc[RW_KC];
#pragma unroll
        for (int j = 0; j < RW_KC; ++j) acc[j] = 0.f;
        if (cv) {
#pragma unroll 4
            for (int t = ty; t < Tf; t += 8) {
                float xv = __ldg(xs + (size_t)t * P * C);
                const float* wr = wgt + t * Tl + k0;
#pragma unroll
                for (int j = 0; j < RW_KC; ++j)
                    if (j < kn) acc[j] = fmaf(wr[j], xv, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < RW_KC; ++j) red[(ty * RW_KC + j) * 32 + tx] = acc[j];
        __syncthreads();
        for (int j = ty; j < kn; j += 8) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) s += red[(q * RW_KC + j) * 32 + tx];
            if (cv) a.agg[(((size_t)b * Tl + k0 + j) * P + p) * C + c] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K14 backward.  grid (B*P), block 256 (8 warps; warp = t-slice, lanes = channels / k)
//   s[t,k] = sum_c dagg[k,c] x[t,c];  r[k] = sum_c dagg[k,c] agg[k,c]
//   dA[t,k] = mask_t (s[t,k] - r[k]) / den[k]
//   datt[t] = sum_k dA gx[t,k];  dgx[t,k] += dA att[t];  dx[t,c] = sum_k wgt[t,k] dagg[k,c]
// smem: dagg_s[Tl][C] | r[Tl] | den[Tl] | wrow[8][Tl]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rewight_agg_bwd_kernel(const cf_rewight_bwd_args a) {
    extern __shared__ __align__(16) float sm[];
    const int Tf = a.Tf, Tl = a.Tl, P = a.P, C = a.C;
    const int Tl4 = (Tl + 3) & ~3;
    float* dg = sm;                          // [Tl][C]
    float* rs = dg + (size_t)Tl * C;         // [Tl]
    float* dens = rs + Tl4;                  // [Tl]
    float* wrow = dens + Tl4;                // [8][Tl4]
    const int b = blockIdx.x / P, p = blockIdx.x - b * P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < Tl * C; i += 256) {
        int k = i / C, c = i - k * C;
        dg[i] = __ldg(a.dagg + (((size_t)b * Tl + k) * P + p) * C + c);
    }
    __syncthreads();
    for (int k = warp; k < Tl; k += 8) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += dg[k * C + c] * __ldg(a.agg + (((size_t)b * Tl + k) * P + p) * C + c);
        s = warp_sum(s);
        if (lane == 0) { rs[k] = s; dens[k] = a.den[((size_t)b * Tl + k) * P + p]; }
    }
    __syncthreads();
    float* wr = wrow + warp * Tl4;
    for (int t = warp; t < Tf; t += 8) {
        const float* xs = a.x + (((size_t)b * Tf + t) * P + p) * C;
        const float at = __ldg(a.att + ((size_t)b * Tf + t) * P + p);
        const float mk = __ldg(a.mask + (size_t)b * Tf + t);
        const float* gxr = a.gx + ((size_t)b * Tf + t) * Tl;
        float datt = 0.f;
        for (int k0 = 0; k0 < Tl; k0 += RW_KC) {
            const int kn = min(RW_KC, Tl - k0);
            float s[RW_KC];
#pragma unroll
            for (int j = 0; j < RW_KC; ++j) s[j] = 0.f;
            for (int c = lane; c < C; c += 32) {
                float xv = __ldg(xs + c);
#pragma unroll
                for (int j = 0; j < RW_KC; ++j)
                    if (j < kn) s[j] = fmaf(dg[(k0 + j) * C + c], xv, s[j]);
            }
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < RW_KC; ++j) {
                float v = warp_sum(s[j]);
                if (lane == j) mine = v;
            }
            if (lane < kn) {
                int k = k0 + lane;
                float g = __ldg(gxr + k);
                float dA = mk * (mine - rs[k]) / dens[k];
                atomicAdd(a.dgx + ((size_t)b * Tf + t) * Tl + k, dA * at);
                datt += dA * g;
                wr[k] = at * g * mk / dens[k];
            }
        }
        datt = warp_sum(datt);
        if (lane == 0) a.datt[((size_t)b * Tf + t) * P + p] = datt;
        if (a.dx) {
            __syncwarp();
            float* dxs = a.dx + (((size_t)b * Tf + t) * P + p) * C;
            for (int c = lane; c < C; c += 32) {
                float acc = 0.f;
                for (int k = 0; k < Tl; ++k) acc = fmaf(wr[k], dg[k * C + c], acc);
                dxs[c] = acc;
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------
// FiLM: out = x * scale[base] + shift[base]
// ---------------------------------------------------------------------------------------
template <int V> struct FV { float v[V]; };
template <int V> __device__ __forceinline__ FV<V> fld(const float* p);
template <> __device__ __forceinline__ FV<4> fld<4>(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    FV<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <> __device__ __forceinline__ FV<1> fld<1>(const float* p) { FV<1> r; r.v[0] = __ldg(p); return r; }
template <int V> __device__ __forceinline__ void fst(float* p, const FV<V>& x);
template <> __device__ __forceinline__ void fst<4>(float* p, const FV<4>& x) {
    *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
template <> __device__ __forceinline__ void fst<1>(float* p, const FV<1>& x) { *p = x.v[0]; }

template <int V>
__global__ void __launch_bounds__(256) film_fwd_kernel(const cf_film_args a, long long total) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int CV = a.C / V;
    int cv = (int)(i % CV);
    long long q = i / CV;
    int x = (int)(q % a.W); q /= a.W;
    int y = (int)(q % a.H); q /= a.H;            // q = b*T + t
    const int rh = a.H / a.Hb, rw = a.W / a.Wb;
    long long base = ((q * a.Hb + y / rh) * a.Wb + x / rw) * a.C + (long long)cv * V;
    FV<V> xv = fld<V>(a.x + i * V), sc = fld<V>(a.scale + base), sh = fld<V>(a.shift + base), o;
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = fmaf(xv.v[j], sc.v[j], sh.v[j]);
    fst<V>(a.out + i * V, o);
}

// one thread per base element x V channels: walks its (rh x rw) block
template <int V>
__global__ void __launch_bounds__(256) film_bwd_kernel(const cf_film_bwd_args a, long long total) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int CV = a.C / V;
    int cv = (int)(i % CV);
    long long q = i / CV;
    int xb = (int)(q % a.Wb); q /= a.Wb;
    int yb = (int)(q % a.Hb); q /= a.Hb;         // q = b*T + t
    const int rh = a.H / a.Hb, rw = a.W / a.Wb;
    FV<V> sc = fld<V>(a.scale + i * V), dsc, dsh;
#pragma unroll
    for (int j = 0; j < V; ++j) { dsc.v[j] = 0.f; dsh.v[j] = 0.f; }
    for (int dy = 0; dy < rh; ++dy)
        for (int dx = 0; dx < rw; ++dx) {
            long long off = ((q * a.H + yb * rh + dy) * a.W + xb * rw + dx) * a.C + (long long)cv * V;
            FV<V> g = fld<V>(a.dout + off), xv = fld<V>(a.x + off), o;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                dsc.v[j] = fmaf(g.v[j], xv.v[j], dsc.v[j]);
                dsh.v[j] += g.v[j];
                o.v[j] = g.v[j] * sc.v[j];
            }
            if (a.dx) fst<V>(a.dx + off, o);
        }
    fst<V>(a.dscale + i * V, dsc);
    fst<V>(a.dshift + i * V, dsh);
}

template <int V>
__global__ void __launch_bounds__(256) nearest_up_kernel(const float* __restrict__ x, float* __restrict__ out, int Hb,
                                                          int Wb, int H, int W, int C, long long total) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int CV = C / V;
    int cv = (int)(i % CV);
    long long q = i / CV;
    int xx = (int)(q % W); q /= W;
    int yy = (int)(q % H); q /= H;
    long long base = ((q * Hb + yy / (H / Hb)) * Wb + xx / (W / Wb)) * C + (long long)cv * V;
    fst<V>(out + i * V, fld<V>(x + base));
}

template <int V>
__global__ void __launch_bounds__(256) nearest_up_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int Hb,
                                                              int Wb, int H, int W, int C, long long total) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int CV = C / V;
    int cv = (int)(i % CV);
    long long q = i / CV;
    int xb = (int)(q % Wb); q /= Wb;
    int yb = (int)(q % Hb); q /= Hb;
    const int rh = H / Hb, rw = W / Wb;
    FV<V> acc;
#pragma unroll
    for (int j = 0; j < V; ++j) acc.v[j] = 0.f;
    for (int dy = 0; dy < rh; ++dy)
        for (int dxx = 0; dxx < rw; ++dxx) {
            FV<V> g = fld<V>(dout + ((q * H + yb * rh + dy) * W + xb * rw + dxx) * C + (long long)cv * V);
#pragma unroll
            for (int j = 0; j < V; ++j) acc.v[j] += g.v[j];
        }
    fst<V>(dx + i * V, acc);
}

// block max pooling over (H,W) (F.adaptive_max_pool2d used as a down-sampler at
// x3d_coarse.py:315,322 when the map is larger than the target): first maximum wins, as in ATen.
__global__ void __launch_bounds__(256) block_maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                 int* __restrict__ idx, int H, int W, int C, int rh, int rw,
                                                                 long long total) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int Ho = H / rh, Wo = W / rw;
    int c = (int)(i % C);
    long long q = i / C;
    int xo = (int)(q % Wo); q /= Wo;
    int yo = (int)(q % Ho); q /= Ho;
    float best = -INFINITY;
    int bi = 0;
    for (int dy = 0; dy < rh; ++dy)
        for (int dx = 0; dx < rw; ++dx) {
            float v = __ldg(x + ((q * H + yo * rh + dy) * W + xo * rw + dx) * C + c);
            if (v > best || (dy == 0 && dx == 0)) { best = v; bi = dy * rw + dx; }
        }
    out[i] = best;
    idx[i] = bi;
}

__global__ void __launch_bounds__(256) block_maxpool_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ idx,
                                                                 float* __restrict__ dx, int H, int W, int C, int rh, int rw,
                                                                 long long total) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int Ho = H / rh, Wo = W / rw;
    int c = (int)(i % C);
    long long q = i / C;
    int xo = (int)(q % Wo); q /= Wo;
    int yo = (int)(q % Ho); q /= Ho;
    float g = dout[i];
    int bi = idx[i];
    for (int dy = 0; dy < rh; ++dy)
        for (int dxx = 0; dxx < rw; ++dxx)
            dx[((q * H + yo * rh + dy) * W + xo * rw + dxx) * C + c] = (dy * rw + dxx == bi) ? g : 0.f;
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out,
                                   long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float s = y[i]; out[i] = dy[i] * s * (1.0f - s); }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

extern "C" {

size_t cf_sizeof_rewight_args(void) { return sizeof(cf_rewight_args); }
size_t cf_sizeof_rewight_bwd_args(void) { return sizeof(cf_rewight_bwd_args); }
size_t cf_sizeof_film_args(void) { return sizeof(cf_film_args); }
size_t cf_sizeof_film_bwd_args(void) { return sizeof(cf_film_bwd_args); }

int cf_sigmoid_bwd(const float* dy, const float* y, float* out, int64_t n, cudaStream_t stream) {
    CF_CHECK_ARG(dy && y && out && n > 0, "bad argument");
    sigmoid_bwd_kernel<<<(unsigned)cf_cdiv64(n, 256), 256, 0, stream>>>(dy, y, out, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_gaussian_fwd(const float* cdf, const float* start, const float* mask, float* gx, int B, int Tf, int Tl, float tx,
                    float ratio, cudaStream_t stream) {
    CF_CHECK_ARG(cdf && start && mask && gx && B > 0 && Tf > 0 && Tl > 0 && ratio != 0.f, "bad argument");
    gaussian_fwd_kernel<<<cf_cdiv((long long)B * Tl * 32, 128), 128, 0, stream>>>(cdf, start, mask, gx, B, Tf, Tl, tx, ratio);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_gaussian_bwd(const float* cdf, const float* start, const float* mask, const float* dgx, float* dcdf_accum, int B,
                    int Tf, int Tl, float tx, float ratio, cudaStream_t stream) {
    CF_CHECK_ARG(cdf && start && mask && dgx && dcdf_accum && B > 0 && Tf > 0 && Tl > 0 && ratio != 0.f, "bad argument");
    gaussian_bwd_kernel<<<cf_cdiv((long long)B * Tl * 32, 128), 128, 0, stream>>>(cdf, start, mask, dgx, dcdf_accum, B, Tf,
                                                                                Tl, tx, ratio);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_rewight_agg_fwd(const cf_rewight_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->x && a->att && a->gx && a->mask && a->agg && a->den, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->C > 0 && a->Tf > 0 && a->Tl > 0 && a->P > 0, "bad shape");
    size_t smem = ((size_t)a->Tf * a->Tl + ((a->Tl + 3) & ~3) + 8 * RW_KC * 32) * sizeof(float);
    CF_CHECK_ARG(smem <= 200 * 1024, "Tf*Tl too large for shared memory");
    static CfOncePerDevice done;
    if (done.need()) { cudaFuncSetAttribute(rewight_agg_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); done.mark(); }
    dim3 grid((unsigned)(a->B * a->P), (unsigned)cf_cdiv(a->C, 32));
    rewight_agg_fwd_kernel<<<grid, dim3(32, 8), smem, stream>>>(*a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_rewight_agg_bwd(const cf_rewight_bwd_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->x && a->att && a->gx && a->mask && a->agg && a->den && a->dagg && a->datt && a->dgx, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->C > 0 && a->Tf > 0 && a->Tl > 0 && a->P > 0, "bad shape");
    int Tl4 = (a->Tl + 3) & ~3;
    size_t smem = ((size_t)a->Tl * a->C + 2 * Tl4 + 8 * Tl4) * sizeof(float);
    CF_CHECK_ARG(smem <= 200 * 1024, "Tl*C too large for shared memory");
    static CfOncePerDevice done;
    if (done.need()) { cudaFuncSetAttribute(rewight_agg_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); done.mark(); }
    rewight_agg_bwd_kernel<<<(unsigned)(a->B * a->P), 256, smem, stream>>>(*a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

static bool film_dims_ok(int B, int C, int T, int H, int W, int Hb, int Wb) {
    return B > 0 && C > 0 && T > 0 && H > 0 && W > 0 && Hb > 0 && Wb > 0 && H % Hb == 0 && W % Wb == 0;
}

int cf_film_fwd(const cf_film_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->x && a->scale && a->shift && a->out, "null pointer");
    CF_CHECK_ARG(film_dims_ok(a->B, a->C, a->T, a->H, a->W, a->Hb, a->Wb), "bad shape (H,W must be multiples of Hb,Wb)");
    long long n = (long long)a->B * a->T * a->H * a->W * a->C;
    if ((a->C & 3) == 0 && al16(a->x) && al16(a->scale) && al16(a->shift) && al16(a->out))
        film_fwd_kernel<4><<<(unsigned)cf_cdiv64(n / 4, 256), 256, 0, stream>>>(*a, n / 4);
    else
        film_fwd_kernel<1><<<(unsigned)cf_cdiv64(n, 256), 256, 0, stream>>>(*a, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_film_bwd(const cf_film_bwd_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->dout && a->x && a->scale && a->dscale && a->dshift, "null pointer");
    CF_CHECK_ARG(film_dims_ok(a->B, a->C, a->T, a->H, a->W, a->Hb, a->Wb), "bad shape (H,W must be multiples of Hb,Wb)");
    long long n = (long long)a->B * a->T * a->Hb * a->Wb * a->C;
    if ((a->C & 3) == 0 && al16(a->x) && al16(a->scale) && al16(a->dout) && al16(a->dscale) && al16(a->dshift) &&
        (!a->dx || al16(a->dx)))
        film_bwd_kernel<4><<<(unsigned)cf_cdiv64(n / 4, 256), 256, 0, stream>>>(*a, n / 4);
    else
        film_bwd_kernel<1><<<(unsigned)cf_cdiv64(n, 256), 256, 0, stream>>>(*a, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_nearest_up(const float* x, float* out, int B, int T, int Hb, int Wb, int H, int W, int C, cudaStream_t stream) {
    CF_CHECK_ARG(x && out, "null pointer");
    CF_CHECK_ARG(film_dims_ok(B, C, T, H, W, Hb, Wb), "bad shape (H,W must be multiples of Hb,Wb)");
    long long n = (long long)B * T * H * W * C;
    if ((C & 3) == 0 && al16(x) && al16(out))
        nearest_up_kernel<4><<<(unsigned)cf_cdiv64(n / 4, 256), 256, 0, stream>>>(x, out, Hb, Wb, H, W, C, n / 4);
    else
        nearest_up_kernel<1><<<(unsigned)cf_cdiv64(n, 256), 256, 0, stream>>>(x, out, Hb, Wb, H, W, C, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_nearest_up_bwd(const float* dout, float* dx, int B, int T, int Hb, int Wb, int H, int W, int C, cudaStream_t stream) {
    CF_CHECK_ARG(dout && dx, "null pointer");
    CF_CHECK_ARG(film_dims_ok(B, C, T, H, W, Hb, Wb), "bad shape (H,W must be multiples of Hb,Wb)");
    long long n = (long long)B * T * Hb * Wb * C;
    if ((C & 3) == 0 && al16(dout) && al16(dx))
        nearest_up_bwd_kernel<4><<<(unsigned)cf_cdiv64(n / 4, 256), 256, 0, stream>>>(dout, dx, Hb, Wb, H, W, C, n / 4);
    else
        nearest_up_bwd_kernel<1><<<(unsigned)cf_cdiv64(n, 256), 256, 0, stream>>>(dout, dx, Hb, Wb, H, W, C, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_block_maxpool_fwd(const float* x, float* out, int32_t* idx, int B, int T, int H, int W, int C, int rh, int rw,
                         cudaStream_t stream) {
    CF_CHECK_ARG(x && out && idx, "null pointer");
    CF_CHECK_ARG(B > 0 && T > 0 && C > 0 && rh > 0 && rw > 0 && H % rh == 0 && W % rw == 0, "bad shape");
    long long n = (long long)B * T * (H / rh) * (W / rw) * C;
    block_maxpool_fwd_kernel<<<(unsigned)cf_cdiv64(n, 256), 256, 0, stream>>>(x, out, idx, H, W, C, rh, rw, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_block_maxpool_bwd(const float* dout, const int32_t* idx, float* dx, int B, int T, int H, int W, int C, int rh, int rw,
                         cudaStream_t stream) {
    CF_CHECK_ARG(dout && dx && idx, "null pointer");
    CF_CHECK_ARG(B > 0 && T > 0 && C > 0 && rh > 0 && rw > 0 && H % rh == 0 && W % rw == 0, "bad shape");
    long long n = (long long)B * T * (H / rh) * (W / rw) * C;
    block_maxpool_bwd_kernel<<<(unsigned)cf_cdiv64(n, 256), 256, 0, stream>>>(dout, idx, dx, H, W, C, rh, rw, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

}  // extern "C"
