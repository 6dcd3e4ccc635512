// Depthwise 3-D convolution kernels (channels-last fp32, CUDA cores; 27 MAC/output is far
// below any tensor-core ridge, the op is bandwidth/L1 bound).
//
// Reference call sites: conv3x3x3(groups=C, stride (1,s,s)) x3d_fine.py:89-97 used at :153;
// conv1_t (5,1,1) depthwise x3d_fine.py:216-222.  The preceding BatchNorm+ReLU is applied in
// the prologue (zero padding is applied AFTER it, as in the reference where the padded tensor
// is the activated one); the statistics of the following BatchNorm (and the SE average pool)
// are accumulated in the epilogue.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

template <int V> struct Vec { float v[V]; };

template <int V> __device__ __forceinline__ Vec<V> vload(const float* p);
template <> __device__ __forceinline__ Vec<4> vload<4>(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <> __device__ __forceinline__ Vec<2> vload<2>(const float* p) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    Vec<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <> __device__ __forceinline__ Vec<1> vload<1>(const float* p) { Vec<1> r; r.v[0] = __ldg(p); return r; }

template <int V> __device__ __forceinline__ Vec<V> vload_s(const float* p) {   // shared / generic
    Vec<V> r;
#pragma unroll
    for (int i = 0; i < V; ++i) r.v[i] = p[i];
    return r;
}
template <int V> __device__ __forceinline__ void vstore(float* p, const Vec<V>& x);
template <> __device__ __forceinline__ void vstore<4>(float* p, const Vec<4>& x) {
    *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
template <> __device__ __forceinline__ void vstore<2>(float* p, const Vec<2>& x) {
    *reinterpret_cast<float2*>(p) = make_float2(x.v[0], x.v[1]);
}
template <> __device__ __forceinline__ void vstore<1>(float* p, const Vec<1>& x) { *p = x.v[0]; }

__device__ __forceinline__ float dw_pro(int mode, float x, float x2, float a, float b, float c) {
    switch (mode) {
        case CF_PRO_AFFINE: return fmaf(a, x, b);
        case CF_PRO_AFFINE_RELU: return fmaxf(fmaf(a, x, b), 0.f);
        case CF_PRO_AFFINE2: return fmaf(a, x, fmaf(b, x2, c));
        default: return x;
    }
}

// ---------------------------------------------------------------------------------------
// forward: each thread produces TW consecutive outputs along W for V channels
// ---------------------------------------------------------------------------------------
template <int V, int KW, int SW, int TW>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const cf_dw_args a) {
    cf_pdl_enter();
    extern __shared__ __align__(16) float sm[];
    const cf_geom& g = a.g;
    const int C = a.C, taps = g.kt * g.kh * KW;
    float* ws = sm;                 // [taps][C]
    float* sst = sm + taps * C;     // [2][C]
    const int tid = threadIdx.x, b = blockIdx.y;
    for (int i = tid; i < taps * C; i += 256) {
        int tap = i / C, c = i - tap * C;
        ws[i] = a.w[(size_t)c * taps + tap];
    }
    const bool do_stats = a.stats_mode != CF_STATS_NONE;
    if (do_stats) for (int i = tid; i < 2 * C; i += 256) sst[i] = 0.f;
    __syncthreads();
    const int CV = C / V, WG = (g.W + TW - 1) / TW;
    long long task = (long long)blockIdx.x * 256 + tid;
    long long ntask = (long long)g.T * g.H * WG * CV;
    if (task < ntask) {
        int cv = (int)(task % CV);
        long long q = task / CV;
        int wg = (int)(q % WG); q /= WG;
        int h = (int)(q % g.H);
        int t = (int)(q / g.H);
        const int c0 = cv * V;
        Vec<V> pa, pb;
#pragma unroll
        for (int i = 0; i < V; ++i) { pa.v[i] = 1.f; pb.v[i] = 0.f; }
        if (a.pro_mode != CF_PRO_NONE) {
            pa = vload_s<V>(a.pro_a + (size_t)b * C + c0);
            pb = vload_s<V>(a.pro_b + (size_t)b * C + c0);
        }
        Vec<V> acc[TW];
#pragma unroll
        for (int u = 0; u < TW; ++u)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[u].v[i] = 0.f;
        constexpr int SPAN = (TW - 1) * SW + KW;
        const int wi0 = wg * TW * SW - g.pw;
        for (int dt = 0; dt < g.kt; ++dt) {
            int ti = t * g.st - g.pt + dt;
            if ((unsigned)ti >= (unsigned)g.Ti) continue;
            for (int dh = 0; dh < g.kh; ++dh) {
                int hi = h * g.sh - g.ph + dh;
                if ((unsigned)hi >= (unsigned)g.Hi) continue;
                const float* xrow = a.x + ((((long long)b * g.Ti + ti) * g.Hi + hi) * g.Wi) * C + c0;
                Vec<V> in[SPAN];
#pragma unroll
                for (int j = 0; j < SPAN; ++j) {
                    int wi = wi0 + j;
                    if ((unsigned)wi < (unsigned)g.Wi) {
                        Vec<V> xv = vload<V>(xrow + (long long)wi * C);
#pragma unroll
                        for (int i = 0; i < V; ++i) in[j].v[i] = dw_pro(a.pro_mode, xv.v[i], 0.f, pa.v[i], pb.v[i], 0.f);
                    } else {
#pragma unroll
                        for (int i = 0; i < V; ++i) in[j].v[i] = 0.f;
                    }
                }
                const float* wrow = ws + (size_t)((dt * g.kh + dh) * KW) * C + c0;
#pragma unroll
                for (int dw = 0; dw < KW; ++dw) {
                    Vec<V> wv = vload_s<V>(wrow + dw * C);
#pragma unroll
                    for (int u = 0; u < TW; ++u)
#pragma unroll
                        for (int i = 0; i < V; ++i) acc[u].v[i] = fmaf(in[u * SW + dw].v[i], wv.v[i], acc[u].v[i]);
                }
            }
        }
        float* yrow = a.y + ((((long long)b * g.T + t) * g.H + h) * g.W) * C + c0;
        Vec<V> s1, s2;
#pragma unroll
        for (int i = 0; i < V; ++i) { s1.v[i] = 0.f; s2.v[i] = 0.f; }
#pragma unroll
        for (int u = 0; u < TW; ++u) {
            int w = wg * TW + u;
            if (w < g.W) {
                vstore<V>(yrow + (long long)w * C, acc[u]);
#pragma unroll
                for (int i = 0; i < V; ++i) { s1.v[i] += acc[u].v[i]; s2.v[i] += acc[u].v[i] * acc[u].v[i]; }
            }
        }
        if (do_stats) {
#pragma unroll
            for (int i = 0; i < V; ++i) { atomicAdd(sst + c0 + i, s1.v[i]); atomicAdd(sst + C + c0 + i, s2.v[i]); }
        }
    }
    if (do_stats) {
        __syncthreads();
        for (int i = tid; i < C; i += 256) {
            double* st = a.stats + ((size_t)b * C + i) * 2;
            atomicAdd(st, (double)sst[i]);
            atomicAdd(st + 1, (double)sst[C + i]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// data gradient (transposed depthwise conv, gather form): one thread = one input position x V ch
//   dx[ti,hi,wi,c] = sum_taps pro(dy[t,h,w,c]) * w[c,tap],  t = (ti+pt-dt)/st (exact), ...
// epilogue: CF_EPI_DRELU with aux = pre-activation at the same position; stats (sum, sum*aux)
// ---------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) dw_dgrad_kernel(const cf_dw_args a) {
    cf_pdl_enter();
    extern __shared__ __align__(16) float sm[];
    const cf_geom& g = a.g;
    const int C = a.C, taps = g.kt * g.kh * g.kw;
    float* ws = sm;
    float* sst = sm + taps * C;
    const int tid = threadIdx.x, b = blockIdx.y;
    for (int i = tid; i < taps * C; i += 256) {
        int tap = i / C, c = i - tap * C;
        ws[i] = a.w[(size_t)c * taps + tap];
    }
    const bool do_stats = a.stats_mode != CF_STATS_NONE;
    if (do_stats) for (int i = tid; i < 2 * C; i += 256) sst[i] = 0.f;
    __syncthreads();
    const int CV = C / V;
    long long task = (long long)blockIdx.x * 256 + tid;
    long long ntask = (long long)g.Ti * g.Hi * g.Wi * CV;
    if (task < ntask) {
        int cv = (int)(task % CV);
        long long q = task / CV;
        int wi = (int)(q % g.Wi); q /= g.Wi;
        int hi = (int)(q % g.Hi);
        int ti = (int)(q / g.Hi);
        const int c0 = cv * V;
        Vec<V> pa, pb, pc;
#pragma unroll
        for (int i = 0; i < V; ++i) { pa.v[i] = 1.f; pb.v[i] = 0.f; pc.v[i] = 0.f; }
        if (a.pro_mode != CF_PRO_NONE) {
            pa = vload_s<V>(a.pro_a + (size_t)b * C + c0);
            if (a.pro_b) pb = vload_s<V>(a.pro_b + (size_t)b * C + c0);
            if (a.pro_c) pc = vload_s<V>(a.pro_c + (size_t)b * C + c0);
        }
        Vec<V> acc;
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
        for (int dt = 0; dt < g.kt; ++dt) {
            int tn = ti + g.pt - dt;
            if (tn < 0 || tn % g.st) continue;
            int t = tn / g.st;
            if (t >= g.T) continue;
            for (int dh = 0; dh < g.kh; ++dh) {
                int hn = hi + g.ph - dh;
                if (hn < 0 || hn % g.sh) continue;
                int h = hn / g.sh;
                if (h >= g.H) continue;
                for (int dw = 0; dw < g.kw; ++dw) {
                    int wn = wi + g.pw - dw;
                    if (wn < 0 || wn % g.sw) continue;
                    int w = wn / g.sw;
                    if (w >= g.W) continue;
                    long long off = ((((long long)b * g.T + t) * g.H + h) * g.W + w) * C + c0;
                    Vec<V> dv = vload<V>(a.x + off);
                    Vec<V> d2;
                    if (a.pro_mode == CF_PRO_AFFINE2) d2 = vload<V>(a.x2 + off);
                    Vec<V> wv = vload_s<V>(ws + (size_t)((dt * g.kh + dh) * g.kw + dw) * C + c0);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        float d = dw_pro(a.pro_mode, dv.v[i], a.pro_mode == CF_PRO_AFFINE2 ? d2.v[i] : 0.f, pa.v[i], pb.v[i], pc.v[i]);
                        acc.v[i] = fmaf(d, wv.v[i], acc.v[i]);
                    }
                }
            }
        }
        long long ooff = ((((long long)b * g.Ti + ti) * g.Hi + hi) * g.Wi + wi) * C + c0;
        Vec<V> auxv;
#pragma unroll
        for (int i = 0; i < V; ++i) auxv.v[i] = 0.f;
        if (a.epi_mode == CF_EPI_DRELU || a.stats_mode == CF_STATS_SUM_AUX) auxv = vload<V>(a.aux + ooff);
        if (a.epi_mode == CF_EPI_DRELU) {
            Vec<V> ea = vload_s<V>(a.epi_a + (size_t)b * C + c0), eb = vload_s<V>(a.epi_b + (size_t)b * C + c0);
#pragma unroll
            for (int i = 0; i < V; ++i) acc.v[i] = (fmaf(ea.v[i], auxv.v[i], eb.v[i]) > 0.f) ? acc.v[i] : 0.f;
        }
        vstore<V>(a.y + ooff, acc);
        if (do_stats) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                atomicAdd(sst + c0 + i, acc.v[i]);
                atomicAdd(sst + C + c0 + i, a.stats_mode == CF_STATS_SUM_AUX ? acc.v[i] * auxv.v[i] : acc.v[i] * acc.v[i]);
            }
        }
    }
    if (do_stats) {
        __syncthreads();
        for (int i = tid; i < C; i += 256) {
            double* st = a.stats + ((size_t)b * C + i) * 2;
            atomicAdd(st, (double)sst[i]);
            atomicAdd(st + 1, (double)sst[C + i]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// weight gradient: dw[c,tap] += sum_pos pro2(dy[pos,c], y2[pos,c]) * pro(x[pos_in(tap),c])
// blockDim 256 = PY position lanes x CV channel vectors; grid = (chunks, B)
// ---------------------------------------------------------------------------------------
template <int V, int TAPS>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const cf_dw_args a, int chunk) {
    cf_pdl_enter();
    extern __shared__ __align__(16) float sm[];     // [TAPS][C]
    const cf_geom& g = a.g;
    const int C = a.C;
    const int tid = threadIdx.x, b = blockIdx.y;
    for (int i = tid; i < TAPS * C; i += 256) sm[i] = 0.f;
    __syncthreads();
    const int CV = C / V, CVb = CV < 256 ? CV : 256, PY = 256 / CVb;          // channel vectors in slabs of 256 (X3D-XL: 630 channels)
    const int cvl = tid % CVb, lane = tid / CVb;
    const long long R = (long long)g.T * g.H * g.W;
    const long long p0 = (long long)blockIdx.x * chunk;
    const long long p1 = p0 + chunk < R ? p0 + chunk : R;
    if (lane < PY) for (int cv = cvl; cv < CV; cv += CVb) {
        const int c0 = cv * V;
        Vec<V> da, db, dc, xa, xb;
#pragma unroll
        for (int i = 0; i < V; ++i) { da.v[i] = 1.f; db.v[i] = 0.f; dc.v[i] = 0.f; xa.v[i] = 1.f; xb.v[i] = 0.f; }
        if (a.pro_mode != CF_PRO_NONE) {          // tables of the dy side
            da = vload_s<V>(a.pro_a + (size_t)b * C + c0);
            if (a.pro_b) db = vload_s<V>(a.pro_b + (size_t)b * C + c0);
            if (a.pro_c) dc = vload_s<V>(a.pro_c + (size_t)b * C + c0);
        }
        if (a.epi_a) {                            // tables of the activation side (bn+relu)
            xa = vload_s<V>(a.epi_a + (size_t)b * C + c0);
            xb = vload_s<V>(a.epi_b + (size_t)b * C + c0);
        }
        Vec<V> acc[TAPS];
#pragma unroll
        for (int tp = 0; tp < TAPS; ++tp)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[tp].v[i] = 0.f;
        for (long long p = p0 + lane; p < p1; p += PY) {
            int w = (int)(p % g.W);
            long long q = p / g.W;
            int h = (int)(q % g.H);
            int t = (int)(q / g.H);
            long long off = ((long long)b * R + p) * C + c0;
            Vec<V> dv = vload<V>(a.x + off);
            if (a.pro_mode == CF_PRO_AFFINE2) {
                Vec<V> d2 = vload<V>(a.x2 + off);
#pragma unroll
                for (int i = 0; i < V; ++i) dv.v[i] = fmaf(da.v[i], dv.v[i], fmaf(db.v[i], d2.v[i], dc.v[i]));
            }
#pragma unroll
            for (int tp = 0; tp < TAPS; ++tp) {
                int dw = tp % g.kw, dh = (tp / g.kw) % g.kh, dt = tp / (g.kw * g.kh);
                int ti = t * g.st - g.pt + dt, hi = h * g.sh - g.ph + dh, wi = w * g.sw - g.pw + dw;
                if ((unsigned)ti < (unsigned)g.Ti && (unsigned)hi < (unsigned)g.Hi && (unsigned)wi < (unsigned)g.Wi) {
                    Vec<V> xv = vload<V>(a.aux + ((((long long)b * g.Ti + ti) * g.Hi + hi) * g.Wi + wi) * C + c0);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        float act = a.epi_a ? fmaxf(fmaf(xa.v[i], xv.v[i], xb.v[i]), 0.f) : xv.v[i];
                        acc[tp].v[i] = fmaf(dv.v[i], act, acc[tp].v[i]);
                    }
                }
            }
        }
#pragma unroll
        for (int tp = 0; tp < TAPS; ++tp)
#pragma unroll
            for (int i = 0; i < V; ++i) atomicAdd(sm + tp * C + c0 + i, acc[tp].v[i]);
    }
    __syncthreads();
    for (int i = tid; i < TAPS * C; i += 256) {
        int tp = i / C, c = i - tp * C;
        atomicAdd(a.y + (size_t)c * TAPS + tp, sm[i]);
    }
}


// ---------------------------------------------------------------------------------------
// 3x3x3 / pad 1 / temporal stride 1 specialisations (every Bottleneck conv2, x3d_fine.py:89-97): the window
// geometry is compile time (no runtime div / mod per tap) and each thread produces TW = 4 consecutive
// positions along W so that one row of loaded inputs serves all of them.
// ---------------------------------------------------------------------------------------
#define DW3_TW 4

// data gradient: dx[ti,hi,wi] = sum_taps pro(d[t,h,w]) * w[tap], t = ti+1-dt, h = (hi+1-dh)/ST, w = (wi+1-dw)/ST
template <int V, int ST>
__global__ void __launch_bounds__(256) dw_dgrad3_kernel(const cf_dw_args a) {
    cf_pdl_enter();
    extern __shared__ __align__(16) float sm[];
    const cf_geom& g = a.g;
    const int C = a.C;
    float* ws = sm;                 // [27][C]
    float* sst = sm + 27 * C;       // [2][C]
    const int tid = threadIdx.x, b = blockIdx.y;
    for (int i = tid; i < 27 * C; i += 256) {
        int tap = i / C, c = i - tap * C;
        ws[i] = a.w[(size_t)c * 27 + tap];
    }
    const bool do_stats = a.stats_mode != CF_STATS_NONE;
    if (do_stats) for (int i = tid; i < 2 * C; i += 256) sst[i] = 0.f;
    __syncthreads();
    constexpr int TW = DW3_TW;
    constexpr int SPAN = (ST == 1) ? TW + 2 : TW / 2 + 1;
    const int CV = C / V, WG = (g.Wi + TW - 1) / TW;
    const long long task = (long long)blockIdx.x * 256 + tid;
    const long long ntask = (long long)g.Ti * g.Hi * WG * CV;
    if (task < ntask) {
        const int cv = (int)(task % CV);
        long long q = task / CV;
        const int wg = (int)(q % WG); q /= WG;
        const int hi = (int)(q % g.Hi);
        const int ti = (int)(q / g.Hi);
        const int c0 = cv * V;
        const int wi0 = wg * TW;
        const int wbase = (ST == 1) ? wi0 - 1 : wi0 / 2;
        const bool aff2 = a.pro_mode == CF_PRO_AFFINE2;
        Vec<V> pa, pb, pc;
#pragma unroll
        for (int i = 0; i < V; ++i) { pa.v[i] = 1.f; pb.v[i] = 0.f; pc.v[i] = 0.f; }
        if (a.pro_mode != CF_PRO_NONE) {
            pa = vload_s<V>(a.pro_a + (size_t)b * C + c0);
            if (a.pro_b) pb = vload_s<V>(a.pro_b + (size_t)b * C + c0);
            if (a.pro_c) pc = vload_s<V>(a.pro_c + (size_t)b * C + c0);
        }
        Vec<V> acc[TW];
#pragma unroll
        for (int u = 0; u < TW; ++u)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[u].v[i] = 0.f;
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
            const int t = ti + 1 - dt;
            if ((unsigned)t >= (unsigned)g.T) continue;
#pragma unroll
            for (int dh = 0; dh < 3; ++dh) {
                const int hn = hi + 1 - dh;
                if (hn < 0 || (ST == 2 && (hn & 1))) continue;
                const int h = hn / ST;
                if (h >= g.H) continue;
                const long long rowoff = ((((long long)b * g.T + t) * g.H + h) * g.W) * C + c0;
                Vec<V> in[SPAN];
#pragma unroll
                for (int j = 0; j < SPAN; ++j) {
                    const int w = wbase + j;
                    if ((unsigned)w < (unsigned)g.W) {
                        Vec<V> dv = vload<V>(a.x + rowoff + (long long)w * C);
                        if (aff2) {
                            Vec<V> d2 = vload<V>(a.x2 + rowoff + (long long)w * C);
#pragma unroll
                            for (int i = 0; i < V; ++i) in[j].v[i] = fmaf(pa.v[i], dv.v[i], fmaf(pb.v[i], d2.v[i], pc.v[i]));
                        } else {
#pragma unroll
                            for (int i = 0; i < V; ++i) in[j].v[i] = dw_pro(a.pro_mode, dv.v[i], 0.f, pa.v[i], pb.v[i], pc.v[i]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < V; ++i) in[j].v[i] = 0.f;
                    }
                }
#pragma unroll
                for (int dw = 0; dw < 3; ++dw) {
                    const Vec<V> wv = vload_s<V>(ws + (size_t)((dt * 3 + dh) * 3 + dw) * C + c0);
#pragma unroll
                    for (int u = 0; u < TW; ++u) {
                        constexpr int dummy = 0;
                        (void)dummy;
                        const int wn = u + 1 - dw;                 // relative to wi0 (a multiple of 4)
                        if (ST == 2 && (wn & 1)) continue;
                        const int j = (ST == 1) ? wn + 1 : wn / 2;
                        if (j < 0 || j >= SPAN) continue;
#pragma unroll
                        for (int i = 0; i < V; ++i) acc[u].v[i] = fmaf(in[j].v[i], wv.v[i], acc[u].v[i]);
                    }
                }
            }
        }
        Vec<V> ea, eb, s1, s2;
#pragma unroll
        for (int i = 0; i < V; ++i) { ea.v[i] = 1.f; eb.v[i] = 0.f; s1.v[i] = 0.f; s2.v[i] = 0.f; }
        if (a.epi_mode == CF_EPI_DRELU) { ea = vload_s<V>(a.epi_a + (size_t)b * C + c0); eb = vload_s<V>(a.epi_b + (size_t)b * C + c0); }
        const long long orow = ((((long long)b * g.Ti + ti) * g.Hi + hi) * g.Wi) * C + c0;
#pragma unroll
        for (int u = 0; u < TW; ++u) {
            const int wi = wi0 + u;
            if (wi >= g.Wi) continue;
            const long long ooff = orow + (long long)wi * C;
            Vec<V> auxv;
#pragma unroll
            for (int i = 0; i < V; ++i) auxv.v[i] = 0.f;
            if (a.epi_mode == CF_EPI_DRELU || a.stats_mode == CF_STATS_SUM_AUX) auxv = vload<V>(a.aux + ooff);
            if (a.epi_mode == CF_EPI_DRELU) {
#pragma unroll
                for (int i = 0; i < V; ++i) acc[u].v[i] = (fmaf(ea.v[i], auxv.v[i], eb.v[i]) > 0.f) ? acc[u].v[i] : 0.f;
            }
            vstore<V>(a.y + ooff, acc[u]);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                s1.v[i] += acc[u].v[i];
                s2.v[i] += a.stats_mode == CF_STATS_SUM_AUX ? acc[u].v[i] * auxv.v[i] : acc[u].v[i] * acc[u].v[i];
            }
        }
        if (do_stats) {
#pragma unroll
            for (int i = 0; i < V; ++i) { atomicAdd(sst + c0 + i, s1.v[i]); atomicAdd(sst + C + c0 + i, s2.v[i]); }
        }
    }
    if (do_stats) {
        __syncthreads();
        for (int i = tid; i < C; i += 256) {
            double* st = a.stats + ((size_t)b * C + i) * 2;
            atomicAdd(st, (double)sst[i]);
            atomicAdd(st + 1, (double)sst[C + i]);
        }
    }
}

// weight gradient: dw[c,tap] += sum_pos pro2(d[pos,c], y2[pos,c]) * act(x[pos_in(tap),c]); each loop iteration takes
// TWG = 2 neighbouring output positions along W so the 3 input rows are loaded once for both
template <int V, int ST>
__global__ void __launch_bounds__(256) dw_wgrad3_kernel(const cf_dw_args a, int chunk) {
    cf_pdl_enter();
    extern __shared__ __align__(16) float sm[];     // [27][C]
    const cf_geom& g = a.g;
    const int C = a.C;
    const int tid = threadIdx.x, b = blockIdx.y;
    for (int i = tid; i < 27 * C; i += 256) sm[i] = 0.f;
    __syncthreads();
    constexpr int TWG = 2;
    constexpr int SPAN = (TWG - 1) * ST + 3;
    const int CV = C / V, CVb = CV < 256 ? CV : 256, PY = 256 / CVb;          // channel vectors in slabs of 256 (X3D-XL: 630 channels)
    const int cvl = tid % CVb, lane = tid / CVb;
    const int WG = (g.W + TWG - 1) / TWG;
    const long long NG = (long long)g.T * g.H * WG;          // position groups per sample
    const long long p0 = (long long)blockIdx.x * chunk;
    const long long p1 = p0 + chunk < NG ? p0 + chunk : NG;
    if (lane < PY) for (int cv = cvl; cv < CV; cv += CVb) {
        const int c0 = cv * V;
        const bool aff2 = a.pro_mode == CF_PRO_AFFINE2;
        const bool act = a.epi_a != nullptr;
        Vec<V> da, db, dc, xa, xb;
#pragma unroll
        for (int i = 0; i < V; ++i) { da.v[i] = 1.f; db.v[i] = 0.f; dc.v[i] = 0.f; xa.v[i] = 1.f; xb.v[i] = 0.f; }
        if (a.pro_mode != CF_PRO_NONE) {
            da = vload_s<V>(a.pro_a + (size_t)b * C + c0);
            if (a.pro_b) db = vload_s<V>(a.pro_b + (size_t)b * C + c0);
            if (a.pro_c) dc = vload_s<V>(a.pro_c + (size_t)b * C + c0);
        }
        if (act) { xa = vload_s<V>(a.epi_a + (size_t)b * C + c0); xb = vload_s<V>(a.epi_b + (size_t)b * C + c0); }
        Vec<V> acc[27];
#pragma unroll
        for (int tp = 0; tp < 27; ++tp)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[tp].v[i] = 0.f;
        for (long long p = p0 + lane; p < p1; p += PY) {
            const int wgi = (int)(p % WG);
            const long long q = p / WG;
            const int h = (int)(q % g.H);
            const int t = (int)(q / g.H);
            const int w0 = wgi * TWG;
            Vec<V> d[TWG];
            const long long drow = ((((long long)b * g.T + t) * g.H + h) * g.W) * C + c0;
#pragma unroll
            for (int u = 0; u < TWG; ++u) {
                if (w0 + u < g.W) {
                    d[u] = vload<V>(a.x + drow + (long long)(w0 + u) * C);
                    if (aff2) {
                        Vec<V> d2 = vload<V>(a.x2 + drow + (long long)(w0 + u) * C);
#pragma unroll
                        for (int i = 0; i < V; ++i) d[u].v[i] = fmaf(da.v[i], d[u].v[i], fmaf(db.v[i], d2.v[i], dc.v[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < V; ++i) d[u].v[i] = 0.f;
                }
            }
            const int wib = w0 * ST - 1;
#pragma unroll
            for (int dt = 0; dt < 3; ++dt) {
                const int ti = t - 1 + dt;
                if ((unsigned)ti >= (unsigned)g.Ti) continue;
#pragma unroll
                for (int dh = 0; dh < 3; ++dh) {
                    const int hi = h * ST - 1 + dh;
                    if ((unsigned)hi >= (unsigned)g.Hi) continue;
                    const long long xrow = ((((long long)b * g.Ti + ti) * g.Hi + hi) * g.Wi) * C + c0;
                    Vec<V> xin[SPAN];
#pragma unroll
                    for (int j = 0; j < SPAN; ++j) {
                        const int wi = wib + j;
                        if ((unsigned)wi < (unsigned)g.Wi) {
                            Vec<V> xv = vload<V>(a.aux + xrow + (long long)wi * C);
#pragma unroll
                            for (int i = 0; i < V; ++i) xin[j].v[i] = act ? fmaxf(fmaf(xa.v[i], xv.v[i], xb.v[i]), 0.f) : xv.v[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < V; ++i) xin[j].v[i] = 0.f;
                        }
                    }
#pragma unroll
                    for (int dw = 0; dw < 3; ++dw)
#pragma unroll
                        for (int u = 0; u < TWG; ++u)
#pragma unroll
                            for (int i = 0; i < V; ++i)
                                acc[(dt * 3 + dh) * 3 + dw].v[i] = fmaf(d[u].v[i], xin[u * ST + dw].v[i], acc[(dt * 3 + dh) * 3 + dw].v[i]);
                }
            }
        }
#pragma unroll
        for (int tp = 0; tp < 27; ++tp)
#pragma unroll
            for (int i = 0; i < V; ++i) atomicAdd(sm + tp * C + c0 + i, acc[tp].v[i]);
    }
    __syncthreads();
    for (int i = tid; i < 27 * C; i += 256) {
        int tp = i / C, c = i - tp * C;
        atomicAdd(a.y + (size_t)c * 27 + tp, sm[i]);
    }
}

// ---------------------------------------------------------------------------------------
extern "C" size_t cf_sizeof_dw_args(void) { return sizeof(cf_dw_args); }

static int pick_vec(int C, const void* p0, const void* p1) {
    bool al16 = ((((uintptr_t)p0) | ((uintptr_t)p1)) & 15) == 0;
    bool al8 = ((((uintptr_t)p0) | ((uintptr_t)p1)) & 7) == 0;
    if ((C & 3) == 0 && al16) return 4;
    if ((C & 1) == 0 && al8) return 2;
    return 1;
}

template <typename F>
static void set_smem(F f) { cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); }

template <int V>
static int launch_dw_fwd(const cf_dw_args* a, cudaStream_t stream) {
    const cf_geom& g = a->g;
    int taps = g.kt * g.kh * g.kw;
    size_t smem = (size_t)(taps + 2) * a->C * sizeof(float);
    constexpr int TW = 4;
    long long ntask = (long long)g.T * g.H * ((g.W + TW - 1) / TW) * (a->C / V);
    dim3 grid((unsigned)cf_cdiv64(ntask, 256), (unsigned)a->B);
#define CF_DW_LAUNCH(KW_, SW_)                                                          \
    do {                                                                                \
        static CfOncePerDevice done;                                                       \
        if (done.need()) { set_smem(dw_fwd_kernel<V, KW_, SW_, TW>); done.mark(); }           \
        cf_launch(dw_fwd_kernel<V, KW_, SW_, TW>, grid, 256, smem, stream, *a);                \
    } while (0)
    if (g.kw == 3 && g.sw == 1) CF_DW_LAUNCH(3, 1);
    else if (g.kw == 3 && g.sw == 2) CF_DW_LAUNCH(3, 2);
    else if (g.kw == 1 && g.sw == 1) CF_DW_LAUNCH(1, 1);
    else { cf_set_error("cf_dw_conv: unsupported (kw,sw)=(%d,%d)", g.kw, g.sw); return CF_ERR_ARG; }
#undef CF_DW_LAUNCH
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

template <int V>
static int launch_dw_dgrad(const cf_dw_args* a, cudaStream_t stream) {
    const cf_geom& g = a->g;
    int taps = g.kt * g.kh * g.kw;
    size_t smem = (size_t)(taps + 2) * a->C * sizeof(float);
    long long ntask = (long long)g.Ti * g.Hi * g.Wi * (a->C / V);
    dim3 grid((unsigned)cf_cdiv64(ntask, 256), (unsigned)a->B);
    static CfOncePerDevice done;
    if (done.need()) { set_smem(dw_dgrad_kernel<V>); done.mark(); }
    cf_launch(dw_dgrad_kernel<V>, grid, 256, smem, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

template <int V, int TAPS>
static int launch_dw_wgrad(const cf_dw_args* a, cudaStream_t stream) {
    const cf_geom& g = a->g;
    size_t smem = (size_t)TAPS * a->C * sizeof(float);
    long long R = (long long)g.T * g.H * g.W;
    int PY = 256 / ((a->C / V) > 256 ? 256 : (a->C / V));
    long long want_ctas = cf_cdiv64(148 * 4, a->B);
    long long chunk = cf_cdiv64(R, want_ctas);
    if (chunk < 4LL * PY) chunk = 4LL * PY;
    dim3 grid((unsigned)cf_cdiv64(R, chunk), (unsigned)a->B);
    static CfOncePerDevice done;
    if (done.need()) { set_smem(dw_wgrad_kernel<V, TAPS>); done.mark(); }
    cf_launch(dw_wgrad_kernel<V, TAPS>, grid, 256, smem, stream, *a, (int)chunk);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

static bool is_333(const cf_geom& g) {
    return g.kt == 3 && g.kh == 3 && g.kw == 3 && g.pt == 1 && g.ph == 1 && g.pw == 1 && g.st == 1 && g.sh == g.sw &&
           (g.sh == 1 || g.sh == 2);
}

template <int V, int ST>
static int launch_dw_dgrad3(const cf_dw_args* a, cudaStream_t stream) {
    const cf_geom& g = a->g;
    size_t smem = (size_t)(27 + 2) * a->C * sizeof(float);
    long long ntask = (long long)g.Ti * g.Hi * ((g.Wi + DW3_TW - 1) / DW3_TW) * (a->C / V);
    dim3 grid((unsigned)cf_cdiv64(ntask, 256), (unsigned)a->B);
    static CfOncePerDevice done;
    if (done.need()) { set_smem(dw_dgrad3_kernel<V, ST>); done.mark(); }
    cf_launch(dw_dgrad3_kernel<V, ST>, grid, 256, smem, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

template <int V, int ST>
static int launch_dw_wgrad3(const cf_dw_args* a, cudaStream_t stream) {
    const cf_geom& g = a->g;
    size_t smem = (size_t)27 * a->C * sizeof(float);
    long long NG = (long long)g.T * g.H * ((g.W + 1) / 2);
    int PY = 256 / ((a->C / V) > 256 ? 256 : (a->C / V));
    long long want_ctas = cf_cdiv64(148 * 4, a->B);
    long long chunk = cf_cdiv64(NG, want_ctas);
    if (chunk < 4LL * PY) chunk = 4LL * PY;
    dim3 grid((unsigned)cf_cdiv64(NG, chunk), (unsigned)a->B);
    static CfOncePerDevice done;
    if (done.need()) { set_smem(dw_wgrad3_kernel<V, ST>); done.mark(); }
    cf_launch(dw_wgrad3_kernel<V, ST>, grid, 256, smem, stream, *a, (int)chunk);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

static int dw_common_checks(const cf_dw_args* a) {
    CF_CHECK_ARG(a && a->x && a->w && a->y, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->C > 0 && a->B <= 65535, "bad shape");
    CF_CHECK_ARG(a->pro_mode == CF_PRO_NONE || a->pro_a, "prologue tables missing");
    CF_CHECK_ARG(a->pro_mode != CF_PRO_AFFINE2 || a->x2, "AFFINE2 needs x2");
    CF_CHECK_ARG(a->stats_mode == CF_STATS_NONE || a->stats, "stats buffer missing");
    CF_CHECK_ARG((size_t)(a->g.kt * a->g.kh * a->g.kw + 2) * a->C * 4 <= 160 * 1024, "C*taps too large for shared memory");
    return CF_OK;
}

int cf_dw3_try(int mode, const cf_dw_args* a, cudaStream_t stream);   // x3d_dw3.cu: plane-marching 3x3x3 stride-1 kernels (-1: not eligible)
int cf_dw3s2_try(int mode, const cf_dw_args* a, cudaStream_t stream); // x3d_dw3s2.cu: spatial stride 2
int cf_dwt5_try(int mode, const cf_dw_args* a, cudaStream_t stream);  // x3d_dwt5.cu: temporal 5x1x1 (stem conv1_t)

extern "C" int cf_dw_conv_fwd(const cf_dw_args* a, cudaStream_t stream) {
    int rc = dw_common_checks(a);
    if (rc) return rc;
    CF_CHECK_ARG(a->pro_mode != CF_PRO_AFFINE2, "forward takes NONE/AFFINE/AFFINE_RELU");
    rc = cf_dw3_try(0, a, stream);
    if (rc >= 0) return rc;
    rc = cf_dw3s2_try(0, a, stream);
    if (rc >= 0) return rc;
    rc = cf_dwt5_try(0, a, stream);
    if (rc >= 0) return rc;
    int v = pick_vec(a->C, a->x, a->y);
    if (v == 4) return launch_dw_fwd<4>(a, stream);
    if (v == 2) return launch_dw_fwd<2>(a, stream);
    return launch_dw_fwd<1>(a, stream);
}

extern "C" int cf_dw_conv_wgrad(const cf_dw_args* a, cudaStream_t stream);
static int dw_dgrad_impl(const cf_dw_args* a, cudaStream_t stream);

extern "C" int cf_dw_conv_dgrad(const cf_dw_args* a, cudaStream_t stream) {
    int rc = dw_common_checks(a);
    if (rc) return rc;
    CF_CHECK_ARG(a->epi_mode == CF_EPI_NONE || (a->epi_mode == CF_EPI_DRELU && a->aux && a->epi_a && a->epi_b), "bad epilogue");
    CF_CHECK_ARG(a->stats_mode != CF_STATS_SUM_AUX || a->aux, "aux missing");
    if (!a->dw_out) return dw_dgrad_impl(a, stream);
    CF_CHECK_ARG(a->aux, "dw_out: aux (the forward input) missing");
    rc = cf_env("CFNET_DW3_NOFUSE", 0) ? -1 : cf_dw3_try(3, a, stream);     // data gradient + weight gradient in one pass
    if (rc >= 0) return rc;
    // The stride-(1,2,2) kernel has the same one-pass form (dw3s2_dgrad_kernel<*, true>), but there it LOSES: 28 outputs
    // per thread leave no registers for the 28 activations next to the mask operands (they are re-read) and the same-box
    // A/B measured 103.2 vs 101.8 ms per step (profiles/r02_ab_same_box.md): experiment build only.
    rc = cf_env("CFNET_DW3S2_FUSE", 0) ? cf_dw3s2_try(3, a, stream) : -1;
    if (rc >= 0) return rc;
    rc = cf_env("CFNET_DWT5_NOFUSE", 0) ? -1 : cf_dwt5_try(3, a, stream);    // stem conv1_t: one march over (dz, yt, y0)
    if (rc >= 0) return rc;
    cf_dw_args d = *a;
    d.dw_out = nullptr;
    rc = dw_dgrad_impl(&d, stream);
    if (rc) return rc;
    cf_dw_args w = *a;                                       // wgrad's contract: y = the weight gradient, act tables in epi_a / epi_b
    w.y = a->dw_out;
    w.dw_out = nullptr;
    w.stats = nullptr;
    w.stats_mode = CF_STATS_NONE;
    w.epi_mode = CF_EPI_NONE;
    return cf_dw_conv_wgrad(&w, stream);
}

static int dw_dgrad_impl(const cf_dw_args* a, cudaStream_t stream) {
    int rc = cf_dw3_try(1, a, stream);
    if (rc >= 0) return rc;
    rc = cf_dw3s2_try(1, a, stream);
    if (rc >= 0) return rc;
    rc = cf_dwt5_try(1, a, stream);
    if (rc >= 0) return rc;
    int v = pick_vec(a->C, a->x, a->y);
    if (a->x2 && (((uintptr_t)a->x2) & 15)) v = v > 2 ? 2 : v;
    if (a->aux && (((uintptr_t)a->aux) & 15)) v = v > 2 ? 2 : v;
    if (is_333(a->g)) {
        const bool s2 = a->g.sh == 2;
        if (v == 4) return s2 ? launch_dw_dgrad3<4, 2>(a, stream) : launch_dw_dgrad3<4, 1>(a, stream);
        if (v == 2) return s2 ? launch_dw_dgrad3<2, 2>(a, stream) : launch_dw_dgrad3<2, 1>(a, stream);
        return s2 ? launch_dw_dgrad3<1, 2>(a, stream) : launch_dw_dgrad3<1, 1>(a, stream);
    }
    if (v == 4) return launch_dw_dgrad<4>(a, stream);
    if (v == 2) return launch_dw_dgrad<2>(a, stream);
    return launch_dw_dgrad<1>(a, stream);
}

extern "C" int cf_dw_conv_wgrad(const cf_dw_args* a, cudaStream_t stream) {
    int rc = dw_common_checks(a);
    if (rc) return rc;
    CF_CHECK_ARG(a->aux, "aux (the forward input) missing");
    CF_CHECK_ARG(a->C <= 1024, "C too large");
    rc = cf_dw3_try(2, a, stream);
    if (rc >= 0) return rc;
    rc = cf_dw3s2_try(2, a, stream);
    if (rc >= 0) return rc;
    rc = cf_dwt5_try(2, a, stream);
    if (rc >= 0) return rc;
    int taps = a->g.kt * a->g.kh * a->g.kw;
    int v = pick_vec(a->C, a->x, a->aux);
    if (is_333(a->g)) {
        const bool s2 = a->g.sh == 2;
        if (v == 4) return s2 ? launch_dw_wgrad3<4, 2>(a, stream) : launch_dw_wgrad3<4, 1>(a, stream);
        if (v == 2) return s2 ? launch_dw_wgrad3<2, 2>(a, stream) : launch_dw_wgrad3<2, 1>(a, stream);
        return s2 ? launch_dw_wgrad3<1, 2>(a, stream) : launch_dw_wgrad3<1, 1>(a, stream);
    }
    if (taps == 27) {
        if (v == 4) return launch_dw_wgrad<4, 27>(a, stream);
        if (v == 2) return launch_dw_wgrad<2, 27>(a, stream);
        return launch_dw_wgrad<1, 27>(a, stream);
    }
    if (taps == 5) {
        if (v == 4) return launch_dw_wgrad<4, 5>(a, stream);
        if (v == 2) return launch_dw_wgrad<2, 5>(a, stream);
        return launch_dw_wgrad<1, 5>(a, stream);
    }
    cf_set_error("cf_dw_conv_wgrad: unsupported tap count %d", taps);
    return CF_ERR_ARG;
}

