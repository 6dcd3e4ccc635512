// Grid Pool / Grid Unpool kernels (sm_100a).
//
// Reference semantics (paths relative to the reference repo):
//   confidence -> CDF ............ x3d_coarse.py:384-392   (cf_gridpool_cdf_*)
//   CDF -> sample coordinate ..... x3d_coarse.py:394,440 + ATen grid_sampler_unnormalize
//                                  (align_corners=True)     (cf_sample_bins)
//   grid_sample along T .......... x3d_coarse.py:396-403, 442-445 (closed form: temporal lerp
//                                  with zero padding)       (cf_temporal_gather_*)
//   inverse CDF (Interp1d) ....... interp1d.py:100-141, x3d_coarse.py:435-438 (cf_inverse_cdf_*)
//   linear / trilinear upsample .. x3d_coarse.py:449,725    (cf_linear_bins + gather)
//
// All big tensors are viewed as [outer, T, inner] with inner contiguous:
//   NCTHW   : outer = B*C, inner = H*W,   outer_per_b = C
//   NTHWC   : outer = B,   inner = H*W*C, outer_per_b = 1
// The gather is a pure HBM-bandwidth op: every thread owns one 128-bit column of `inner`
// and walks the K sample points, so a frame that two consecutive sample points share is
// re-read from L1/L2, not DRAM.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

// ---------------------------------------------------------------------------------------
// K10: confidences g [B,n]  ->  cdf [B,n+1]        one warp per row
// ---------------------------------------------------------------------------------------
__global__ void cdf_fwd_kernel(const float* __restrict__ g, float* __restrict__ cdf, int B, int n) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* gr = g + (size_t)row * n;
    float* cr = cdf + (size_t)row * (n + 1);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += 1.0f - cf_sigmoid(0.5f * gr[j]);
    s = warp_sum(s);
    float den = s + 1e-16f;
    if (lane == 0) cr[0] = 0.f;
    float carry = 0.f;
    for (int base = 0; base < n; base += 32) {
        int j = base + lane;
        float v = (j < n) ? __fdiv_rn(1.0f - cf_sigmoid(0.5f * gr[j]), den) : 0.f;
        // Kogge-Stone inclusive scan across the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        v += carry;
        if (j < n) cr[j + 1] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// dcdf [B,n+1] -> dg [B,n]
__global__ void cdf_bwd_kernel(const float* __restrict__ g, const float* __restrict__ dcdf,
                               float* __restrict__ dg, int B, int n) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* gr = g + (size_t)row * n;
    const float* dc = dcdf + (size_t)row * (n + 1);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += 1.0f - cf_sigmoid(0.5f * gr[j]);
    s = warp_sum(s);
    float den = s + 1e-16f;
    // dq_j = sum_{k=j+1..n} dcdf[k];  dot = sum_j dq_j * p_j
    float dot = 0.f;
    for (int j = lane; j < n; j += 32) {
        float dq = 0.f;
        for (int k = j + 1; k <= n; ++k) dq += dc[k];
        dot += dq * (1.0f - cf_sigmoid(0.5f * gr[j]));
    }
    dot = warp_sum(dot);
    for (int j = lane; j < n; j += 32) {
        float dq = 0.f;
        for (int k = j + 1; k <= n; ++k) dq += dc[k];
        float sg = cf_sigmoid(0.5f * gr[j]);
        float dp = dq / den - dot / (den * den);
        dg[(size_t)row * n + j] = -0.5f * sg * (1.0f - sg) * dp;
    }
}

// ---------------------------------------------------------------------------------------
// cdf -> (i0, w1): z = (((cdf-0.5)*2 + 1)/2)*(T-1) evaluated op by op in fp32 (no FMA
// contraction) so that floor(z) is bit-identical to the reference given the same cdf.
// ---------------------------------------------------------------------------------------
__global__ void sample_bins_kernel(const float* __restrict__ cdf, int* __restrict__ i0,
                                   float* __restrict__ w1, int n, int t_in) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gq = __fmul_rn(__fsub_rn(cdf[i], 0.5f), 2.0f);
    float z = __fmul_rn(__fmul_rn(__fadd_rn(gq, 1.0f), 0.5f), (float)(t_in - 1));
    float f = floorf(z);
    i0[i] = (int)f;
    w1[i] = __fsub_rn(z, f);
}

// batch-independent bins of F.interpolate(mode='linear', align_corners=True): t_in -> t_out
__global__ void linear_bins_kernel(int* __restrict__ i0, float* __restrict__ w1, int t_in, int t_out) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= t_out) return;
    float scale = (t_out > 1) ? (float)(t_in - 1) / (float)(t_out - 1) : 0.f;
    float src = __fmul_rn(scale, (float)u);
    int j0 = (int)src;
    if (j0 > t_in - 1) j0 = t_in - 1;
    i0[u] = j0;
    w1[u] = __fsub_rn(src, (float)j0);
}

// ---------------------------------------------------------------------------------------
// inverse CDF (Interp1d with y = xnew = mid):  cdf [B,K] -> inv [B,K], ind [B,K]
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float mid_of(int j, int K) { return __fdiv_rn((float)j, (float)(K - 1)); }

__global__ void inverse_cdf_fwd_kernel(const float* __restrict__ cdf, float* __restrict__ inv,
                                       int* __restrict__ ind, int B, int K) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    int b = i / K, j = i - b * K;
    const float* c = cdf + (size_t)b * K;
    float q = mid_of(j, K);
    int cnt = 0;                                     // searchsorted(right=False): #entries < q
    for (int m = 0; m < K; ++m) cnt += (c[m] < q) ? 1 : 0;
    int id = cnt - 1;
    id = id < 0 ? 0 : (id > K - 2 ? K - 2 : id);
    const float eps = 1.1920928955078125e-07f;
    float dy = __fsub_rn(mid_of(id + 1, K), mid_of(id, K));
    float D = __fadd_rn(eps, __fsub_rn(c[id + 1], c[id]));
    float slope = __fdiv_rn(dy, D);
    inv[i] = __fadd_rn(mid_of(id, K), __fmul_rn(slope, __fsub_rn(q, c[id])));
    ind[i] = id;
}

// dinv [B,K] -> dcdf_accum [B,K] (+=).  One thread per row, sequential over j: deterministic.
__global__ void inverse_cdf_bwd_kernel(const float* __restrict__ cdf, const int* __restrict__ ind,
                                       const float* __restrict__ dinv, float* __restrict__ dcdf, int B, int K) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* c = cdf + (size_t)b * K;
    float* dc = dcdf + (size_t)b * K;
    const float eps = 1.1920928955078125e-07f;
    for (int j = 0; j < K; ++j) {
        int id = ind[(size_t)b * K + j];
        float q = mid_of(j, K);
        float dy = mid_of(id + 1, K) - mid_of(id, K);
        float D = eps + (c[id + 1] - c[id]);
        float r = (q - c[id]) / D;
        float go = dinv[(size_t)b * K + j];
        dc[id] += go * dy * (r / D - 1.0f / D);
        dc[id + 1] += go * (-dy * r / D);
    }
}

// ---------------------------------------------------------------------------------------
// general Interp1d (interp1d.py:100-141): D rows, N knots, P queries; a row stride of 0
// broadcasts a single row ("flat" inputs of the reference)
// ---------------------------------------------------------------------------------------
__global__ void interp1d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ xnew, float* __restrict__ ynew, int* __restrict__ ind,
                                    int D, int N, int P, int xrs, int yrs, int qrs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * P) return;
    int d = i / P, j = i - d * P;
    const float* xr = x + (size_t)d * xrs;
    const float* yr = y + (size_t)d * yrs;
    float q = xnew[(size_t)d * qrs + j];
    int lo = 0, hi = N;                               // lower_bound: first index with x >= q
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (xr[mid] < q) lo = mid + 1; else hi = mid;
    }
    int id = lo - 1;
    id = id < 0 ? 0 : (id > N - 2 ? N - 2 : id);
    const float eps = 1.1920928955078125e-07f;
    float slope = __fdiv_rn(__fsub_rn(yr[id + 1], yr[id]), __fadd_rn(eps, __fsub_rn(xr[id + 1], xr[id])));
    ynew[i] = __fadd_rn(yr[id], __fmul_rn(slope, __fsub_rn(q, xr[id])));
    ind[i] = id;
}

__global__ void interp1d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ xnew, const int* __restrict__ ind,
                                    const float* __restrict__ dynew, float* dx, float* dy, float* dq, int D, int N, int P,
                                    int xrs, int yrs, int qrs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * P) return;
    int d = i / P, j = i - d * P;
    const float* xr = x + (size_t)d * xrs;
    const float* yr = y + (size_t)d * yrs;
    int id = ind[i];
    float q = xnew[(size_t)d * qrs + j], go = dynew[i];
    const float eps = 1.1920928955078125e-07f;
    float Dn = eps + (xr[id + 1] - xr[id]);
    float slope = (yr[id + 1] - yr[id]) / Dn;
    float r = (q - xr[id]) / Dn;
    if (dx) {
        atomicAdd(dx + (size_t)d * xrs + id, go * slope * (r - 1.0f));
        atomicAdd(dx + (size_t)d * xrs + id + 1, -go * slope * r);
    }
    if (dy) {
        atomicAdd(dy + (size_t)d * yrs + id, go * (1.0f - r));
        atomicAdd(dy + (size_t)d * yrs + id + 1, go * r);
    }
    if (dq) atomicAdd(dq + (size_t)d * qrs + j, go * slope);
}

// ---------------------------------------------------------------------------------------
// K11 forward: out[o,k,:] = (1-w1) x[o,i0,:] + w1 x[o,i0+1,:]   (zero padding outside [0,T-1])
// ---------------------------------------------------------------------------------------
template <typename V> struct VecOps;
template <> struct VecOps<float4> {
    static __device__ __forceinline__ float4 zero() { return f4_zero(); }
    static __device__ __forceinline__ float4 ld(const float4* p) { return __ldg(p); }
    static __device__ __forceinline__ void st(float4* p, float4 v) { __stcs(p, v); }
    static __device__ __forceinline__ float4 lerp(float a, float4 x, float b, float4 y) { return f4_axpby(a, x, b, y); }
    static __device__ __forceinline__ void fma(float4& acc, float a, float4 x) { f4_fma(acc, a, x); }
    static __device__ __forceinline__ float dotdiff(float4 g, float4 x1, float4 x0) {
        return g.x * (x1.x - x0.x) + g.y * (x1.y - x0.y) + g.z * (x1.z - x0.z) + g.w * (x1.w - x0.w);
    }
};
template <> struct VecOps<float> {
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void st(float* p, float v) { __stcs(p, v); }
    static __device__ __forceinline__ float lerp(float a, float x, float b, float y) { return a * x + b * y; }
    static __device__ __forceinline__ void fma(float& acc, float a, float x) { acc = fmaf(a, x, acc); }
    static __device__ __forceinline__ float dotdiff(float g, float x1, float x0) { return g * (x1 - x0); }
};

template <typename V, int KU>
__global__ void __launch_bounds__(256)
temporal_gather_fwd_kernel(const V* __restrict__ x, const int* __restrict__ i0, const float* __restrict__ w1,
                           V* __restrict__ out, long long outer, long long opb, int T, int K, long long inner_v) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= outer * inner_v) return;
    long long o = idx / inner_v;
    long long iv = idx - o * inner_v;
    long long b = o / opb;
    const V* xs = x + o * (long long)T * inner_v + iv;
    V* os = out + o * (long long)K * inner_v + iv;
    const int* bi = i0 + b * K;
    const float* bw = w1 + b * K;
    for (int k0 = 0; k0 < K; k0 += KU) {
        V a0[KU], a1[KU];
        float w[KU];
#pragma unroll
        for (int u = 0; u < KU; ++u) {                 // issue all loads of the group first (MLP)
            int k = k0 + u;
            a0[u] = VecOps<V>::zero();
            a1[u] = VecOps<V>::zero();
            w[u] = 0.f;
            if (k < K) {
                int a = __ldg(bi + k);
                w[u] = __ldg(bw + k);
                if (a >= 0 && a < T) a0[u] = VecOps<V>::ld(xs + (long long)a * inner_v);
                if (a + 1 >= 0 && a + 1 < T && w[u] != 0.f) a1[u] = VecOps<V>::ld(xs + (long long)(a + 1) * inner_v);
            }
        }
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            int k = k0 + u;
            if (k < K) VecOps<V>::st(os + (long long)k * inner_v, VecOps<V>::lerp(1.0f - w[u], a0[u], w[u], a1[u]));
        }
    }
}

// ---------------------------------------------------------------------------------------
// K11 backward w.r.t. x in gather form.  A tiny prep kernel inverts the (k -> i0,i0+1) map
// into a per-sample CSR over source frames t; the main kernel then writes every dx frame
// exactly once (zeros where no sample point touches it): no atomics, no memset.
// CSR layout per sample b (ints):  row_ptr[T+1] | ent_k[2K] ; floats: ent_w[2K]
// ---------------------------------------------------------------------------------------
__global__ void gather_csr_kernel(const int* __restrict__ i0, const float* __restrict__ w1, int* __restrict__ row_ptr,
                                  int* __restrict__ ent_k, float* __restrict__ ent_w, int nb, int T, int K) {
    // one warp per sample; lanes own source frames t
    int b = blockIdx.x, lane = threadIdx.x;
    const int* bi = i0 + (size_t)b * K;
    const float* bw = w1 + (size_t)b * K;
    int* rp = row_ptr + (size_t)b * (T + 1);
    int* ek = ent_k + (size_t)b * 2 * K;
    float* ew = ent_w + (size_t)b * 2 * K;
    if (lane == 0) rp[0] = 0;
    int carry = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        int t = t0 + lane;
        int cnt = 0;
        if (t < T)
            for (int k = 0; k < K; ++k) {
                int a = __ldg(bi + k);
                cnt += (a == t) ? 1 : 0;
                cnt += (a + 1 == t && __ldg(bw + k) != 0.f) ? 1 : 0;
            }
        int v = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int tmp = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += tmp;
        }
        if (t < T) {
            rp[t + 1] = carry + v;
            int e = carry + v - cnt;
            for (int k = 0; k < K; ++k) {
                int a = __ldg(bi + k);
                float w = __ldg(bw + k);
                if (a == t) { ek[e] = k; ew[e] = 1.0f - w; ++e; }
                if (a + 1 == t && w != 0.f) { ek[e] = k; ew[e] = w; ++e; }
            }
        }
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
}

template <typename V>
__global__ void __launch_bounds__(256)
temporal_gather_bwd_x_kernel(const V* __restrict__ gout, const int* __restrict__ row_ptr, const int* __restrict__ ent_k,
                             const float* __restrict__ ent_w, V* __restrict__ dx, long long outer, long long opb,
                             int T, int K, long long inner_v) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= outer * inner_v) return;
    long long o = idx / inner_v;
    long long iv = idx - o * inner_v;
    long long b = o / opb;
    const V* gs = gout + o * (long long)K * inner_v + iv;
    V* ds = dx + o * (long long)T * inner_v + iv;
    const int* rp = row_ptr + b * (T + 1);
    const int* ek = ent_k + b * 2 * K;
    const float* ew = ent_w + b * 2 * K;
    int e = __ldg(rp);
    for (int t = 0; t < T; ++t) {
        int e1 = __ldg(rp + t + 1);
        V acc = VecOps<V>::zero();
        for (; e < e1; ++e) VecOps<V>::fma(acc, __ldg(ew + e), VecOps<V>::ld(gs + (long long)__ldg(ek + e) * inner_v));
        VecOps<V>::st(ds + (long long)t * inner_v, acc);
    }
}

// ---------------------------------------------------------------------------------------
// K11 backward w.r.t. the sample coordinate:
//   dcoord[b,k] += scale * sum_{o in b, i} gout[o,k,i] * (x[o,i0+1,i] - x[o,i0,i])
// grid = (chunks of inner, outer, K); block reduce; one atomicAdd per CTA.
// ---------------------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(256)
temporal_gather_bwd_z_kernel(const V* __restrict__ gout, const V* __restrict__ x, const int* __restrict__ i0,
                             double* __restrict__ dcoord, long long opb, int T, int K, long long inner_v,
                             long long chunk_v, float scale) {
    long long o = blockIdx.y;
    int k = blockIdx.z;
    long long b = o / opb;
    int a = __ldg(i0 + b * K + k);
    bool v0 = (a >= 0 && a < T), v1 = (a + 1 >= 0 && a + 1 < T);
    const V* gs = gout + (o * (long long)K + k) * inner_v;
    const V* x0 = x + (o * (long long)T + (v0 ? a : 0)) * inner_v;
    const V* x1 = x + (o * (long long)T + (v1 ? a + 1 : 0)) * inner_v;
    long long lo = (long long)blockIdx.x * chunk_v;
    long long hi = lo + chunk_v < inner_v ? lo + chunk_v : inner_v;
    // the terms gout*(x1-x0) have mixed signs and largely cancel: accumulate in double (the kernel is HBM-bound)
    double acc = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        V g = VecOps<V>::ld(gs + i);
        V p0 = v0 ? VecOps<V>::ld(x0 + i) : VecOps<V>::zero();
        V p1 = v1 ? VecOps<V>::ld(x1 + i) : VecOps<V>::zero();
        acc += (double)VecOps<V>::dotdiff(g, p1, p0);
    }
    __shared__ double red[8];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) atomicAdd(dcoord + b * K + k, v * (double)scale);
    }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int cf_gridpool_cdf_fwd(const float* g, float* cdf, int B, int n, cudaStream_t stream) {
    CF_CHECK_ARG(g && cdf && B > 0 && n > 0, "bad argument");
    cdf_fwd_kernel<<<cf_cdiv((long long)B * 32, 128), 128, 0, stream>>>(g, cdf, B, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_gridpool_cdf_bwd(const float* g, const float* dcdf, float* dg, int B, int n, cudaStream_t stream) {
    CF_CHECK_ARG(g && dcdf && dg && B > 0 && n > 0, "bad argument");
    cdf_bwd_kernel<<<cf_cdiv((long long)B * 32, 128), 128, 0, stream>>>(g, dcdf, dg, B, n);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_sample_bins(const float* coord, int32_t* i0, float* w1, int n, int t_in, cudaStream_t stream) {
    CF_CHECK_ARG(coord && i0 && w1 && n > 0 && t_in > 0, "bad argument");
    sample_bins_kernel<<<cf_cdiv(n, 128), 128, 0, stream>>>(coord, i0, w1, n, t_in);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_linear_bins(int32_t* i0, float* w1, int t_in, int t_out, cudaStream_t stream) {
    CF_CHECK_ARG(i0 && w1 && t_in > 0 && t_out > 0, "bad argument");
    linear_bins_kernel<<<cf_cdiv(t_out, 128), 128, 0, stream>>>(i0, w1, t_in, t_out);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_inverse_cdf_fwd(const float* cdf, float* inv, int32_t* ind, int B, int K, cudaStream_t stream) {
    CF_CHECK_ARG(cdf && inv && ind && B > 0 && K > 1, "bad argument");
    inverse_cdf_fwd_kernel<<<cf_cdiv((long long)B * K, 128), 128, 0, stream>>>(cdf, inv, ind, B, K);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_inverse_cdf_bwd(const float* cdf, const int32_t* ind, const float* dinv, float* dcdf_accum, int B, int K,
                       cudaStream_t stream) {
    CF_CHECK_ARG(cdf && ind && dinv && dcdf_accum && B > 0 && K > 1, "bad argument");
    inverse_cdf_bwd_kernel<<<cf_cdiv(B, 64), 64, 0, stream>>>(cdf, ind, dinv, dcdf_accum, B, K);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_interp1d_fwd(const float* x, const float* y, const float* xnew, float* ynew, int32_t* ind, int D, int N, int P,
                    int x_row_stride, int y_row_stride, int xnew_row_stride, cudaStream_t stream) {
    CF_CHECK_ARG(x && y && xnew && ynew && ind && D > 0 && N > 1 && P > 0, "bad argument");
    interp1d_fwd_kernel<<<cf_cdiv((long long)D * P, 128), 128, 0, stream>>>(x, y, xnew, ynew, ind, D, N, P, x_row_stride,
                                                                         y_row_stride, xnew_row_stride);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_interp1d_bwd(const float* x, const float* y, const float* xnew, const int32_t* ind, const float* dynew, float* dx_accum,
                    float* dy_accum, float* dxnew_accum, int D, int N, int P, int x_row_stride, int y_row_stride,
                    int xnew_row_stride, cudaStream_t stream) {
    CF_CHECK_ARG(x && y && xnew && ind && dynew && D > 0 && N > 1 && P > 0, "bad argument");
    interp1d_bwd_kernel<<<cf_cdiv((long long)D * P, 128), 128, 0, stream>>>(x, y, xnew, ind, dynew, dx_accum, dy_accum,
                                                                         dxnew_accum, D, N, P, x_row_stride, y_row_stride,
                                                                         xnew_row_stride);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

int cf_temporal_gather_fwd(const float* x, const int32_t* i0, const float* w1, float* out, int64_t outer,
                           int64_t outer_per_b, int T, int K, int64_t inner, cudaStream_t stream) {
    CF_CHECK_ARG(x && i0 && w1 && out, "null pointer");
    CF_CHECK_ARG(outer > 0 && outer_per_b > 0 && T > 0 && K > 0 && inner > 0, "bad shape");
    if ((inner & 3) == 0 && aligned16(x) && aligned16(out)) {
        long long iv = inner / 4, tot = outer * iv;
        temporal_gather_fwd_kernel<float4, 4><<<(unsigned)cf_cdiv64(tot, 256), 256, 0, stream>>>(
            (const float4*)x, i0, w1, (float4*)out, outer, outer_per_b, T, K, iv);
    } else {
        long long tot = outer * inner;
        temporal_gather_fwd_kernel<float, 4><<<(unsigned)cf_cdiv64(tot, 256), 256, 0, stream>>>(
            x, i0, w1, out, outer, outer_per_b, T, K, inner);
    }
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

size_t cf_temporal_gather_bwd_ws_bytes(int64_t n_batch, int T, int K) {
    return (size_t)n_batch * ((size_t)(T + 1) + 4 * (size_t)K) * 4 + 64;
}

int cf_temporal_gather_bwd_x(const float* gout, const int32_t* i0, const float* w1, float* dx, void* ws,
                             size_t ws_bytes, int64_t outer, int64_t outer_per_b, int T, int K, int64_t inner,
                             cudaStream_t stream) {
    CF_CHECK_ARG(gout && i0 && w1 && dx && ws, "null pointer");
    CF_CHECK_ARG(outer > 0 && outer_per_b > 0 && T > 0 && K > 0 && inner > 0, "bad shape");
    int64_t nb = (outer + outer_per_b - 1) / outer_per_b;
    CF_CHECK_ARG(ws_bytes >= cf_temporal_gather_bwd_ws_bytes(nb, T, K), "workspace too small");
    int* row_ptr = (int*)ws;
    int* ent_k = row_ptr + nb * (T + 1);
    float* ent_w = (float*)(ent_k + nb * 2 * K);
    gather_csr_kernel<<<(unsigned)nb, 32, 0, stream>>>(i0, w1, row_ptr, ent_k, ent_w, (int)nb, T, K);
    if ((inner & 3) == 0 && aligned16(gout) && aligned16(dx)) {
        long long iv = inner / 4, tot = outer * iv;
        temporal_gather_bwd_x_kernel<float4><<<(unsigned)cf_cdiv64(tot, 256), 256, 0, stream>>>(
            (const float4*)gout, row_ptr, ent_k, ent_w, (float4*)dx, outer, outer_per_b, T, K, iv);
    } else {
        long long tot = outer * inner;
        temporal_gather_bwd_x_kernel<float><<<(unsigned)cf_cdiv64(tot, 256), 256, 0, stream>>>(
            gout, row_ptr, ent_k, ent_w, dx, outer, outer_per_b, T, K, inner);
    }
    CF_COUNT_LAUNCH(2);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_temporal_gather_bwd_coord(const float* gout, const float* x, const int32_t* i0, double* dcoord_accum,
                                 int64_t outer, int64_t outer_per_b, int T, int K, int64_t inner, float scale,
                                 cudaStream_t stream) {
    CF_CHECK_ARG(gout && x && i0 && dcoord_accum, "null pointer");
    CF_CHECK_ARG(outer > 0 && outer <= 65535 * 64LL && outer_per_b > 0 && T > 0 && K > 0 && K <= 65535 && inner > 0, "bad shape");
    CF_CHECK_ARG(outer <= 65535, "outer too large for grid.y");
    if ((inner & 3) == 0 && aligned16(gout) && aligned16(x)) {
        long long iv = inner / 4;
        long long chunk = 2048;
        dim3 grid((unsigned)cf_cdiv64(iv, chunk), (unsigned)outer, (unsigned)K);
        temporal_gather_bwd_z_kernel<float4><<<grid, 256, 0, stream>>>((const float4*)gout, (const float4*)x, i0,
                                                                      dcoord_accum, outer_per_b, T, K, iv, chunk, scale);
    } else {
        long long chunk = 8192;
        dim3 grid((unsigned)cf_cdiv64(inner, chunk), (unsigned)outer, (unsigned)K);
        temporal_gather_bwd_z_kernel<float><<<grid, 256, 0, stream>>>(gout, x, i0, dcoord_accum, outer_per_b, T, K,
                                                                     inner, chunk, scale);
    }
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

}  // extern "C"
