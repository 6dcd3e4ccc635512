// BatchNorm / SE / residual / pooling glue kernels of the X3D stacks (channels-last fp32).
//
// Train-mode SubBatchNorm3d (x3d_fine.py:13-62) is split in two: statistics are accumulated by
// the producing conv kernel (per-sample double sums), cf_bn_finalize turns them into
// per-(sample,channel) affine tables, and the consuming kernel applies the table in its prologue.
// The backward pass mirrors this: producers accumulate (sum dz, sum dz*y), cf_bn_bwd_coeffs
// turns them into the coefficients of dy = P*d + Q*y + R, consumers apply them on load.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

template <int V> struct VecF { float v[V]; };
template <int V> __device__ __forceinline__ VecF<V> ldv(const float* p);
template <> __device__ __forceinline__ VecF<4> ldv<4>(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    VecF<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <> __device__ __forceinline__ VecF<2> ldv<2>(const float* p) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    VecF<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <> __device__ __forceinline__ VecF<1> ldv<1>(const float* p) { VecF<1> r; r.v[0] = __ldg(p); return r; }
template <int V> __device__ __forceinline__ void stv(float* p, const VecF<V>& x);
template <> __device__ __forceinline__ void stv<4>(float* p, const VecF<4>& x) {
    *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
template <> __device__ __forceinline__ void stv<2>(float* p, const VecF<2>& x) {
    *reinterpret_cast<float2*>(p) = make_float2(x.v[0], x.v[1]);
}
template <> __device__ __forceinline__ void stv<1>(float* p, const VecF<1>& x) { *p = x.v[0]; }
template <int V> __device__ __forceinline__ VecF<V> ldtab(const float* p) {
    VecF<V> r;
#pragma unroll
    for (int i = 0; i < V; ++i) r.v[i] = __ldg(p + i);
    return r;
}

// ---------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const cf_bn_args a) {
    cf_pdl_enter_early();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    const int C = a.C, S = a.splits;
    float gamma = a.gamma[c], beta = a.beta[c];
    if (!a.training) {
        float m = a.running_mean[c];
        float is = rsqrtf(a.running_var[c] + a.eps);
        a.mean[c] = m;
        a.invstd[c] = is;
        float A = gamma * is, Bc = beta - m * A;
        for (int b = 0; b < a.B; ++b) { a.tab_a[(size_t)b * C + c] = A; a.tab_b[(size_t)b * C + c] = Bc; }
        return;
    }
    for (int g = 0; g < S; ++g) {
        double s1 = 0.0, s2 = 0.0;
        int nb = 0;
        for (int b = g; b < a.B; b += S) {
            s1 += a.stats[((size_t)b * C + c) * 2];
            s2 += a.stats[((size_t)b * C + c) * 2 + 1];
            ++nb;
        }
        double cnt = (double)nb * (double)a.rows_per_sample;
        double mean = s1 / cnt;
        double var = s2 / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        float is = (float)(1.0 / sqrt(var + (double)a.eps));
        float mf = (float)mean;
        a.mean[(size_t)g * C + c] = mf;
        a.invstd[(size_t)g * C + c] = is;
        if (a.running_mean) {
            size_t ri = (size_t)g * C + c;
            double unb = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
            a.running_mean[ri] = (1.f - a.momentum) * a.running_mean[ri] + a.momentum * mf;
            a.running_var[ri] = (1.f - a.momentum) * a.running_var[ri] + a.momentum * (float)unb;
        }
        float A = gamma * is, Bc = beta - mf * A;
        for (int b = g; b < a.B; b += S) { a.tab_a[(size_t)b * C + c] = A; a.tab_b[(size_t)b * C + c] = Bc; }
    }
}

__global__ void bn_bwd_coeffs_kernel(const cf_bn_bwd_args a) {
    cf_pdl_enter_early();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    const int C = a.C, S = a.splits;
    float gamma = a.gamma[c];
    double dgam = 0.0, dbet = 0.0;
    for (int g = 0; g < S; ++g) {
        double sdz = 0.0, sdzy = 0.0;
        int nb = 0;
        for (int b = g; b < a.B; b += S) {
            sdz += a.sums[((size_t)b * C + c) * 2];
            sdzy += a.sums[((size_t)b * C + c) * 2 + 1];
            ++nb;
        }
        double cnt = (double)nb * (double)a.rows_per_sample;
        double mean = a.mean[(size_t)g * C + c], is = a.invstd[(size_t)g * C + c];
        double sdzh = is * (sdzy - mean * sdz);          // sum dz * yhat
        dgam += sdzh;
        dbet += sdz;
        double c1 = gamma * is, c2 = 0.0, c3 = 0.0;
        if (a.training) {
            c2 = -gamma * is * is * (sdzh / cnt);
            c3 = -gamma * is * (sdz / cnt) + gamma * is * is * mean * (sdzh / cnt);
        }
        for (int b = g; b < a.B; b += S) {
            size_t i = (size_t)b * C + c;
            float gate = a.gate ? a.gate[i] : 1.f;
            float cst = a.cst ? a.cst[i] : 0.f;
            a.tab_p[i] = (float)(c1 * gate);
            a.tab_q[i] = (float)c2;
            a.tab_r[i] = (float)(c3 + c1 * cst);
        }
    }
    if (a.dgamma) a.dgamma[c] += (float)dgam;
    if (a.dbeta) a.dbeta[c] += (float)dbet;
}

// ---------------------------------------------------------------------------------------
// SE forward: one CTA per sample
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_fwd_kernel(const cf_se_args a) {
    cf_pdl_enter_early();
    extern __shared__ float sm[];            // pooled[C] | hidden[Wd]
    const int b = blockIdx.x, tid = threadIdx.x, C = a.C, Wd = a.Wd;
    float* pooled = sm;
    float* hid = sm + C;
    for (int c = tid; c < C; c += 256) {
        size_t i = (size_t)b * C + c;
        float m = (float)(a.stats[i * 2] / (double)a.rows_per_sample);
        float p = fmaf(a.tab_a[i], m, a.tab_b[i]);
        pooled[c] = p;
        a.pooled[i] = p;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < Wd; j += 8) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(a.w1[(size_t)j * C + c], pooled[c], s);
        s = warp_sum(s);
        if (lane == 0) {
            float h = fmaxf(s + a.b1[j], 0.f);
            hid[j] = h;
            a.hidden[(size_t)b * Wd + j] = h;
        }
    }
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
        float s = a.b2[c];
        for (int j = 0; j < Wd; ++j) s = fmaf(a.w2[(size_t)c * Wd + j], hid[j], s);
        float gt = cf_sigmoid(s);
        size_t i = (size_t)b * C + c;
        a.gate[i] = gt;
        a.out_a[i] = gt * a.tab_a[i];
        a.out_b[i] = gt * a.tab_b[i];
    }
}

__global__ void __launch_bounds__(256) se_bwd_kernel(const cf_se_bwd_args a) {
    cf_pdl_enter_early();
    extern __shared__ float sm[];            // dlogit[C] | dpre[Wd]
    const int b = blockIdx.x, tid = threadIdx.x, C = a.C, Wd = a.Wd;
    float* dlog = sm;
    float* dpre = sm + C;
    for (int c = tid; c < C; c += 256) {
        size_t i = (size_t)b * C + c;
        double sdu = a.sums[i * 2], sduy = a.sums[i * 2 + 1];
        float dgate = (float)((double)a.tab_a[i] * sduy + (double)a.tab_b[i] * sdu);
        float gt = a.gate[i];
        float dl = dgate * gt * (1.f - gt);
        dlog[c] = dl;
        atomicAdd(a.db2 + c, dl);
        for (int j = 0; j < Wd; ++j) atomicAdd(a.dw2 + (size_t)c * Wd + j, dl * a.hidden[(size_t)b * Wd + j]);
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < Wd; j += 8) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(dlog[c], a.w2[(size_t)c * Wd + j], s);
        s = warp_sum(s);
        float dp = (a.hidden[(size_t)b * Wd + j] > 0.f) ? s : 0.f;
        if (lane == 0) {
            dpre[j] = dp;
            atomicAdd(a.db1 + j, dp);
        }
        for (int c = lane; c < C; c += 32) atomicAdd(a.dw1 + (size_t)j * C + c, dp * a.pooled[(size_t)b * C + c]);
    }
    __syncthreads();
    const double R = (double)a.rows_per_sample;
    for (int c = tid; c < C; c += 256) {
        float dpool = 0.f;
        for (int j = 0; j < Wd; ++j) dpool = fmaf(dpre[j], a.w1[(size_t)j * C + c], dpool);
        size_t i = (size_t)b * C + c;
        float cst = (float)((double)dpool / R);
        a.cst[i] = cst;
        double gt = a.gate[i];
        double sdu = a.sums[i * 2], sduy = a.sums[i * 2 + 1];
        a.sums[i * 2] = gt * sdu + R * (double)cst;
        a.sums[i * 2 + 1] = gt * sduy + (double)cst * a.stats_y[i * 2];
    }
}

// ---------------------------------------------------------------------------------------
// residual join
// ---------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) residual_fwd_kernel(const cf_residual_args a) {
    cf_pdl_enter();
    const int b = blockIdx.y, C = a.C, CV = C / V;
    long long n = a.rows_per_sample * CV;
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    int c0 = (int)(idx % CV) * V;
    long long off = (long long)b * a.rows_per_sample * C + (idx / CV) * C + c0;
    VecF<V> y = ldv<V>(a.y + off);
    VecF<V> ta = ldtab<V>(a.tab_a + (size_t)b * C + c0), tb = ldtab<V>(a.tab_b + (size_t)b * C + c0);
    VecF<V> o;
#pragma unroll
    for (int i = 0; i < V; ++i) o.v[i] = fmaf(ta.v[i], y.v[i], tb.v[i]);
    if (a.res) {
        VecF<V> r = ldv<V>(a.res + off);
        if (a.res_a) {
            VecF<V> ra = ldtab<V>(a.res_a + (size_t)b * C + c0), rb = ldtab<V>(a.res_b + (size_t)b * C + c0);
#pragma unroll
            for (int i = 0; i < V; ++i) o.v[i] += fmaf(ra.v[i], r.v[i], rb.v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) o.v[i] += r.v[i];
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) o.v[i] = fmaxf(o.v[i], 0.f);
    stv<V>(a.out + off, o);
}

// the same join for a stage-final block of the global tower: a thread owns one (t, pooling block, channel vector), writes
// the rh x rw outputs of its block and their average (x3d_fine.py:345-354) -- the features leave in the pass that produces
// the stage output instead of in a second full read of it
template <int V>
__global__ void __launch_bounds__(256) residual_pool_fwd_kernel(const cf_residual_args a) {
    cf_pdl_enter();
    const int b = blockIdx.y, C = a.C, CV = C / V;
    const int Ho = a.H / a.rh, Wo = a.W / a.rw;
    long long n = (long long)a.T * Ho * Wo * CV;
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    int c0 = (int)(idx % CV) * V;
    long long q = idx / CV;
    int wo = (int)(q % Wo); q /= Wo;
    int ho = (int)(q % Ho);
    int t = (int)(q / Ho);
    const VecF<V> ta = ldtab<V>(a.tab_a + (size_t)b * C + c0), tb = ldtab<V>(a.tab_b + (size_t)b * C + c0);
    VecF<V> ra, rb;
    if (a.res && a.res_a) { ra = ldtab<V>(a.res_a + (size_t)b * C + c0); rb = ldtab<V>(a.res_b + (size_t)b * C + c0); }
    VecF<V> acc;
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
    for (int dh = 0; dh < a.rh; ++dh)
        for (int dw = 0; dw < a.rw; ++dw) {
            const long long off = ((((long long)b * a.T + t) * a.H + ho * a.rh + dh) * a.W + wo * a.rw + dw) * C + c0;
            const VecF<V> y = ldv<V>(a.y + off);
            VecF<V> o;
#pragma unroll
            for (int i = 0; i < V; ++i) o.v[i] = fmaf(ta.v[i], y.v[i], tb.v[i]);
            if (a.res) {
                const VecF<V> r = ldv<V>(a.res + off);
#pragma unroll
                for (int i = 0; i < V; ++i) o.v[i] += a.res_a ? fmaf(ra.v[i], r.v[i], rb.v[i]) : r.v[i];
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                o.v[i] = fmaxf(o.v[i], 0.f);
                acc.v[i] += o.v[i];
            }
            stv<V>(a.out + off, o);
        }
    const float inv = 1.0f / (float)(a.rh * a.rw);
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] *= inv;
    stv<V>(a.pooled + ((((long long)b * a.T + t) * Ho + ho) * Wo + wo) * C + c0, acc);
}

// rows of one sample are split in chunks over grid.x; 256 threads = PY row lanes x CV channel vectors
template <int V>
__global__ void __launch_bounds__(256) residual_bwd_kernel(const cf_residual_bwd_args a, int chunk) {
    cf_pdl_enter();
    extern __shared__ float sm[];            // [4][C]
    const int b = blockIdx.y, C = a.C, CV = C / V, tid = threadIdx.x;
    for (int i = tid; i < 4 * C; i += 256) sm[i] = 0.f;
    __syncthreads();
    // 256 threads = PY row lanes x CVb channel vectors; more than 256 channel vectors (X3D-XL: 630 channels) are taken in slabs
    const int CVb = CV < 256 ? CV : 256, PY = 256 / CVb, cvl = tid % CVb, lane = tid / CVb;
    long long r0 = (long long)blockIdx.x * chunk;
    long long r1 = r0 + chunk < a.rows_per_sample ? r0 + chunk : a.rows_per_sample;
    if (lane < PY) for (int cv = cvl; cv < CV; cv += CVb) {
        const int c0 = cv * V;
        VecF<V> s0, s1, s2;
#pragma unroll
        for (int i = 0; i < V; ++i) { s0.v[i] = 0.f; s1.v[i] = 0.f; s2.v[i] = 0.f; }
        for (long long r = r0 + lane; r < r1; r += PY) {
            long long off = ((long long)b * a.rows_per_sample + r) * C + c0;
            VecF<V> d, o = ldv<V>(a.out + off), y = ldv<V>(a.y + off);
            if (a.dout) d = ldv<V>(a.dout + off);
            else {
#pragma unroll
                for (int i = 0; i < V; ++i) d.v[i] = 0.f;
            }
            if (a.dpool) {                       // gradient of the pooled features, spread over their rh x rw block
                const int w = (int)(r % a.W);
                const long long q = r / a.W;
                const int h = (int)(q % a.H), t = (int)(q / a.H);
                const int Ho = a.H / a.rh, Wo = a.W / a.rw;
                const VecF<V> dp = ldv<V>(a.dpool + ((((long long)b * a.T + t) * Ho + h / a.rh) * Wo + w / a.rw) * C + c0);
                const float inv = 1.0f / (float)(a.rh * a.rw);
#pragma unroll
                for (int i = 0; i < V; ++i) d.v[i] = fmaf(dp.v[i], inv, d.v[i]);
            }
            VecF<V> rs;
            if (a.res) rs = ldv<V>(a.res + off);
            VecF<V> dz;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                dz.v[i] = o.v[i] > 0.f ? d.v[i] : 0.f;
                s0.v[i] += dz.v[i];
                s1.v[i] = fmaf(dz.v[i], y.v[i], s1.v[i]);
                if (a.res) s2.v[i] = fmaf(dz.v[i], rs.v[i], s2.v[i]);
            }
            stv<V>(a.dz + off, dz);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            atomicAdd(sm + c0 + i, s0.v[i]);
            atomicAdd(sm + C + c0 + i, s1.v[i]);
            if (a.res) atomicAdd(sm + 2 * C + c0 + i, s2.v[i]);
        }
    }
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
        size_t i = ((size_t)b * C + c) * 2;
        atomicAdd(a.sums_y + i, (double)sm[c]);
        atomicAdd(a.sums_y + i + 1, (double)sm[C + c]);
        if (a.sums_res) {
            atomicAdd(a.sums_res + i, (double)sm[c]);
            atomicAdd(a.sums_res + i + 1, (double)sm[2 * C + c]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// block average pooling over (H,W)
// ---------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const cf_pool_args a) {
    cf_pdl_enter();
    const int b = blockIdx.y, C = a.C, CV = C / V;
    const int Ho = a.H / a.rh, Wo = a.W / a.rw;
    long long n = (long long)a.T * Ho * Wo * CV;
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    int c0 = (int)(idx % CV) * V;
    long long q = idx / CV;
    int wo = (int)(q % Wo); q /= Wo;
    int ho = (int)(q % Ho);
    int t = (int)(q / Ho);
    VecF<V> ta, tb;
    const bool pro = a.tab_a != nullptr;
    if (pro) { ta = ldtab<V>(a.tab_a + (size_t)b * C + c0); tb = ldtab<V>(a.tab_b + (size_t)b * C + c0); }
    VecF<V> acc;
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
    for (int dh = 0; dh < a.rh; ++dh)
        for (int dw = 0; dw < a.rw; ++dw) {
            long long off = ((((long long)b * a.T + t) * a.H + ho * a.rh + dh) * a.W + wo * a.rw + dw) * C + c0;
            VecF<V> x = ldv<V>(a.x + off);
#pragma unroll
            for (int i = 0; i < V; ++i) acc.v[i] += pro ? fmaxf(fmaf(ta.v[i], x.v[i], tb.v[i]), 0.f) : x.v[i];
        }
    float inv = 1.0f / (float)(a.rh * a.rw);
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] *= inv;
    stv<V>(a.y + ((((long long)b * a.T + t) * Ho + ho) * Wo + wo) * C + c0, acc);
}

template <int V>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const cf_pool_bwd_args a, int chunk) {
    cf_pdl_enter();
    extern __shared__ float sm[];            // [2][C]
    const int b = blockIdx.y, C = a.C, CV = C / V, tid = threadIdx.x;
    const bool do_sums = a.sums != nullptr;
    if (do_sums) for (int i = tid; i < 2 * C; i += 256) sm[i] = 0.f;
    __syncthreads();
    const int CVb = CV < 256 ? CV : 256, PY = 256 / CVb, cvl = tid % CVb, lane = tid / CVb;       // channel vectors in slabs of 256
    const int Ho = a.H / a.rh, Wo = a.W / a.rw;
    const long long R = (long long)a.T * a.H * a.W;
    long long r0 = (long long)blockIdx.x * chunk;
    long long r1 = r0 + chunk < R ? r0 + chunk : R;
    const float inv = 1.0f / (float)(a.rh * a.rw);
    if (lane < PY) for (int cv = cvl; cv < CV; cv += CVb) {
        const int c0 = cv * V;
        VecF<V> ta, tb, s0, s1;
        const bool pro = a.tab_a != nullptr;
        if (pro) { ta = ldtab<V>(a.tab_a + (size_t)b * C + c0); tb = ldtab<V>(a.tab_b + (size_t)b * C + c0); }
#pragma unroll
        for (int i = 0; i < V; ++i) { s0.v[i] = 0.f; s1.v[i] = 0.f; }
        for (long long r = r0 + lane; r < r1; r += PY) {
            int w = (int)(r % a.W);
            long long q = r / a.W;
            int h = (int)(q % a.H);
            int t = (int)(q / a.H);
            long long off = ((long long)b * R + r) * C + c0;
            VecF<V> dy = ldv<V>(a.dy + ((((long long)b * a.T + t) * Ho + h / a.rh) * Wo + w / a.rw) * C + c0);
            VecF<V> x;
            if (a.x) x = ldv<V>(a.x + off);
            VecF<V> dz;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float d = dy.v[i] * inv;
                if (pro) d = (fmaf(ta.v[i], x.v[i], tb.v[i]) > 0.f) ? d : 0.f;
                dz.v[i] = d;
                s0.v[i] += d;
                if (a.x) s1.v[i] = fmaf(d, x.v[i], s1.v[i]);
            }
            if (a.accumulate) {
                VecF<V> old = ldv<V>(a.dz + off);
#pragma unroll
                for (int i = 0; i < V; ++i) dz.v[i] += old.v[i];
            }
            stv<V>(a.dz + off, dz);
        }
        if (do_sums) {
#pragma unroll
            for (int i = 0; i < V; ++i) { atomicAdd(sm + c0 + i, s0.v[i]); atomicAdd(sm + C + c0 + i, s1.v[i]); }
        }
    }
    if (do_sums) {
        __syncthreads();
        for (int c = tid; c < C; c += 256) {
            size_t i = ((size_t)b * C + c) * 2;
            atomicAdd(a.sums + i, (double)sm[c]);
            atomicAdd(a.sums + i + 1, (double)sm[C + c]);
        }
    }
}

// ---------------------------------------------------------------------------------------
static int vec_for(int C, const void* p0, const void* p1, const void* p2, const void* p3) {
    uintptr_t m = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)p3;
    if ((C & 3) == 0 && (m & 15) == 0) return 4;
    if ((C & 1) == 0 && (m & 7) == 0) return 2;
    return 1;
}
static int chunk_for(long long R, int B, int PY) {
    long long want = cf_cdiv64(148 * 4, B);
    long long chunk = cf_cdiv64(R, want);
    if (chunk < 4LL * PY) chunk = 4LL * PY;
    return (int)chunk;
}

extern "C" int cf_bn_finalize(const cf_bn_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->gamma && a->beta && a->tab_a && a->tab_b && a->mean && a->invstd, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->C > 0 && a->splits > 0 && a->B % a->splits == 0, "bad shape (B % splits)");
    CF_CHECK_ARG(a->training ? (a->stats != nullptr) : (a->running_mean && a->running_var), "missing statistics");
    if (cf_env("CFNET_BN_SKIP", 0)) return CF_OK;       // timing experiment (experiment build): what the table launches cost
    cf_launch(bn_finalize_kernel, cf_cdiv(a->C, 128), 128, 0, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_bn_bwd_coeffs(const cf_bn_bwd_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->sums && a->gamma && a->mean && a->invstd && a->tab_p && a->tab_q && a->tab_r, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->C > 0 && a->splits > 0 && a->B % a->splits == 0, "bad shape (B % splits)");
    if (cf_env("CFNET_BN_SKIP", 0)) return CF_OK;
    cf_launch(bn_bwd_coeffs_kernel, cf_cdiv(a->C, 128), 128, 0, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_se_fwd(const cf_se_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->stats && a->tab_a && a->tab_b && a->w1 && a->b1 && a->w2 && a->b2, "null pointer");
    CF_CHECK_ARG(a->pooled && a->hidden && a->gate && a->out_a && a->out_b, "null output");
    CF_CHECK_ARG(a->B > 0 && a->C > 0 && a->Wd > 0 && (size_t)(a->C + a->Wd) * 4 <= 48 * 1024, "bad shape");
    cf_launch(se_fwd_kernel, a->B, 256, (size_t)(a->C + a->Wd) * 4, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_se_bwd(const cf_se_bwd_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->sums && a->stats_y && a->tab_a && a->tab_b && a->w1 && a->w2 && a->pooled && a->hidden && a->gate, "null pointer");
    CF_CHECK_ARG(a->dw1 && a->db1 && a->dw2 && a->db2 && a->cst, "null output");
    CF_CHECK_ARG(a->B > 0 && a->C > 0 && a->Wd > 0 && (size_t)(a->C + a->Wd) * 4 <= 48 * 1024, "bad shape");
    cf_launch(se_bwd_kernel, a->B, 256, (size_t)(a->C + a->Wd) * 4, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_residual_fwd(const cf_residual_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->y && a->tab_a && a->tab_b && a->out, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->B <= 65535 && a->C > 0 && a->rows_per_sample > 0, "bad shape");
    int v = vec_for(a->C, a->y, a->out, a->res, nullptr);
    if (a->pooled) {
        CF_CHECK_ARG(a->T > 0 && a->H > 0 && a->W > 0 && a->rh > 0 && a->rw > 0 && a->H % a->rh == 0 && a->W % a->rw == 0 &&
                         (int64_t)a->T * a->H * a->W == a->rows_per_sample,
                     "pooled output: rows_per_sample must be T*H*W with H,W multiples of the pooling block");
        if ((((uintptr_t)a->pooled) & 15) && v == 4) v = 2;
        long long np = (long long)a->T * (a->H / a->rh) * (a->W / a->rw) * (a->C / v);
        dim3 gp((unsigned)cf_cdiv64(np, 256), (unsigned)a->B);
        if (v == 4) cf_launch(residual_pool_fwd_kernel<4>, gp, 256, 0, stream, *a);
        else if (v == 2) cf_launch(residual_pool_fwd_kernel<2>, gp, 256, 0, stream, *a);
        else cf_launch(residual_pool_fwd_kernel<1>, gp, 256, 0, stream, *a);
        CF_COUNT_LAUNCH(1);
        CF_CHECK_LAUNCH();
        return CF_OK;
    }
    long long n = a->rows_per_sample * (a->C / v);
    dim3 grid((unsigned)cf_cdiv64(n, 256), (unsigned)a->B);
    if (v == 4) cf_launch(residual_fwd_kernel<4>, grid, 256, 0, stream, *a);
    else if (v == 2) cf_launch(residual_fwd_kernel<2>, grid, 256, 0, stream, *a);
    else cf_launch(residual_fwd_kernel<1>, grid, 256, 0, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_residual_bwd(const cf_residual_bwd_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && (a->dout || a->dpool) && a->out && a->y && a->dz && a->sums_y, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->B <= 65535 && a->C > 0 && a->rows_per_sample > 0, "bad shape");
    CF_CHECK_ARG(!a->sums_res || a->res, "sums_res without res");
    CF_CHECK_ARG(!a->dpool || (a->T > 0 && a->rh > 0 && a->rw > 0 && a->H % a->rh == 0 && a->W % a->rw == 0 &&
                               (int64_t)a->T * a->H * a->W == a->rows_per_sample),
                 "dpool: rows_per_sample must be T*H*W with H,W multiples of the pooling block");
    int v = vec_for(a->C, a->dout ? a->dout : a->out, a->out, a->y, a->dz);
    if (a->dpool && (((uintptr_t)a->dpool) & 15) && v == 4) v = 2;
    if (a->res && (((uintptr_t)a->res) & 15) && v == 4) v = 2;
    int chunk = chunk_for(a->rows_per_sample, a->B, 256 / (a->C / v > 256 ? 256 : a->C / v));
    dim3 grid((unsigned)cf_cdiv64(a->rows_per_sample, chunk), (unsigned)a->B);
    size_t smem = (size_t)4 * a->C * 4;
    if (v == 4) cf_launch(residual_bwd_kernel<4>, grid, 256, smem, stream, *a, chunk);
    else if (v == 2) cf_launch(residual_bwd_kernel<2>, grid, 256, smem, stream, *a, chunk);
    else cf_launch(residual_bwd_kernel<1>, grid, 256, smem, stream, *a, chunk);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_block_avgpool_fwd(const cf_pool_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->x && a->y, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->B <= 65535 && a->C > 0 && a->rh > 0 && a->rw > 0 && a->H % a->rh == 0 && a->W % a->rw == 0,
                 "bad shape (H,W must be multiples of the pooling block)");
    int v = vec_for(a->C, a->x, a->y, nullptr, nullptr);
    long long n = (long long)a->T * (a->H / a->rh) * (a->W / a->rw) * (a->C / v);
    dim3 grid((unsigned)cf_cdiv64(n, 256), (unsigned)a->B);
    if (v == 4) cf_launch(avgpool_fwd_kernel<4>, grid, 256, 0, stream, *a);
    else if (v == 2) cf_launch(avgpool_fwd_kernel<2>, grid, 256, 0, stream, *a);
    else cf_launch(avgpool_fwd_kernel<1>, grid, 256, 0, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_block_avgpool_bwd(const cf_pool_bwd_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->dy && a->dz, "null pointer");
    CF_CHECK_ARG(a->B > 0 && a->B <= 65535 && a->C > 0 && a->rh > 0 && a->rw > 0 && a->H % a->rh == 0 && a->W % a->rw == 0, "bad shape");
    CF_CHECK_ARG(!a->tab_a || (a->x && a->tab_b), "prologue tables need x");
    CF_CHECK_ARG(!a->sums || a->x, "sums need x");
    int v = vec_for(a->C, a->dy, a->dz, a->x, nullptr);
    long long R = (long long)a->T * a->H * a->W;
    int chunk = chunk_for(R, a->B, 256 / (a->C / v > 256 ? 256 : a->C / v));
    dim3 grid((unsigned)cf_cdiv64(R, chunk), (unsigned)a->B);
    size_t smem = (size_t)2 * a->C * 4;
    if (v == 4) cf_launch(avgpool_bwd_kernel<4>, grid, 256, smem, stream, *a, chunk);
    else if (v == 2) cf_launch(avgpool_bwd_kernel<2>, grid, 256, smem, stream, *a, chunk);
    else cf_launch(avgpool_bwd_kernel<1>, grid, 256, smem, stream, *a, chunk);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out, long long n) {
    cf_pdl_enter();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = y[i] > 0.f ? dy[i] : 0.f;
}

extern "C" int cf_relu_bwd(const float* dy, const float* y, float* out, int64_t n, cudaStream_t stream) {
    CF_CHECK_ARG(dy && y && out && n > 0, "bad argument");
    cf_launch(relu_bwd_kernel, (unsigned)cf_cdiv64(n, 256), 256, 0, stream, dy, y, out, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

// ---------------------------------------------------------------------------------------
// standalone surfaces
// ---------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) channel_stats_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            double* __restrict__ stats, int C, long long rows, int chunk) {
    cf_pdl_enter();
    extern __shared__ float sm[];            // [2][C]
    const int b = blockIdx.y, CV = C / V, tid = threadIdx.x;
    for (int i = tid; i < 2 * C; i += 256) sm[i] = 0.f;
    __syncthreads();
    const int CVb = CV < 256 ? CV : 256, PY = 256 / CVb, cvl = tid % CVb, lane = tid / CVb;
    long long r0 = (long long)blockIdx.x * chunk;
    long long r1 = r0 + chunk < rows ? r0 + chunk : rows;
    if (lane < PY) for (int cv = cvl; cv < CV; cv += CVb) {
        const int c0 = cv * V;
        VecF<V> s0, s1;
#pragma unroll
        for (int i = 0; i < V; ++i) { s0.v[i] = 0.f; s1.v[i] = 0.f; }
        for (long long r = r0 + lane; r < r1; r += PY) {
            long long off = ((long long)b * rows + r) * C + c0;
            VecF<V> xv = ldv<V>(x + off);
            VecF<V> yv = y ? ldv<V>(y + off) : xv;
#pragma unroll
            for (int i = 0; i < V; ++i) { s0.v[i] += xv.v[i]; s1.v[i] = fmaf(xv.v[i], yv.v[i], s1.v[i]); }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) { atomicAdd(sm + c0 + i, s0.v[i]); atomicAdd(sm + C + c0 + i, s1.v[i]); }
    }
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
        size_t i = ((size_t)b * C + c) * 2;
        atomicAdd(stats + i, (double)sm[c]);
        atomicAdd(stats + i + 1, (double)sm[C + c]);
    }
}

__global__ void __launch_bounds__(256) affine_apply_kernel(const cf_affine_args a) {
    cf_pdl_enter();
    const int b = blockIdx.y, C = a.C;
    long long n = a.rows_per_sample * C;
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    int c = (int)(idx % C);
    long long off = (long long)b * n + idx;
    float ta = a.tab_a[(size_t)b * C + c];
    float tb = a.tab_b ? a.tab_b[(size_t)b * C + c] : 0.f;
    float tc = a.tab_c ? a.tab_c[(size_t)b * C + c] : 0.f;
    float x = a.x[off], x2 = a.x2 ? a.x2[off] : 0.f, o;
    switch (a.mode) {
        case CF_PRO_AFFINE: o = fmaf(ta, x, tb); break;
        case CF_PRO_AFFINE_RELU: o = fmaxf(fmaf(ta, x, tb), 0.f); break;
        case CF_PRO_AFFINE_SWISH: { float v = fmaf(ta, x, tb); o = v * cf_sigmoid(v); break; }
        case CF_PRO_AFFINE2: o = fmaf(ta, x, fmaf(tb, x2, tc)); break;
        default: o = x;
    }
    a.out[off] = o;
}

__global__ void swish_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
    cf_pdl_enter();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float v = x[i]; out[i] = v * cf_sigmoid(v); }
}
__global__ void swish_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long n) {
    cf_pdl_enter();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float v = x[i], s = cf_sigmoid(v); dx[i] = dy[i] * (s * (1.f + v * (1.f - s))); }
}

extern "C" int cf_channel_stats(const float* x, const float* y, double* stats, int B, int C, int64_t rows, cudaStream_t stream) {
    CF_CHECK_ARG(x && stats && B > 0 && B <= 65535 && C > 0 && rows > 0, "bad argument");
    int v = vec_for(C, x, y, nullptr, nullptr);
    if (C % v) { cf_set_error("cf_channel_stats: bad channel count"); return CF_ERR_ARG; }
    int chunk = chunk_for(rows, B, 256 / ((C / v) > 256 ? 256 : (C / v)));
    dim3 grid((unsigned)cf_cdiv64(rows, chunk), (unsigned)B);
    size_t smem = (size_t)2 * C * 4;
    if (v == 4) cf_launch(channel_stats_kernel<4>, grid, 256, smem, stream, x, y, stats, C, rows, chunk);
    else if (v == 2) cf_launch(channel_stats_kernel<2>, grid, 256, smem, stream, x, y, stats, C, rows, chunk);
    else cf_launch(channel_stats_kernel<1>, grid, 256, smem, stream, x, y, stats, C, rows, chunk);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_affine_apply(const cf_affine_args* a, cudaStream_t stream) {
    CF_CHECK_ARG(a && a->x && a->out && a->tab_a && a->B > 0 && a->B <= 65535 && a->C > 0 && a->rows_per_sample > 0, "bad argument");
    CF_CHECK_ARG(a->mode != CF_PRO_AFFINE2 || a->x2, "AFFINE2 needs x2");
    dim3 grid((unsigned)cf_cdiv64(a->rows_per_sample * a->C, 256), (unsigned)a->B);
    cf_launch(affine_apply_kernel, grid, 256, 0, stream, *a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_swish_fwd(const float* x, float* out, int64_t n, cudaStream_t stream) {
    CF_CHECK_ARG(x && out && n > 0, "bad argument");
    cf_launch(swish_fwd_kernel, (unsigned)cf_cdiv64(n, 256), 256, 0, stream, x, out, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

extern "C" int cf_swish_bwd(const float* x, const float* dy, float* dx, int64_t n, cudaStream_t stream) {
    CF_CHECK_ARG(x && dy && dx && n > 0, "bad argument");
    cf_launch(swish_bwd_kernel, (unsigned)cf_cdiv64(n, 256), 256, 0, stream, x, dy, dx, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
