// Pointwise-conv GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   y[b,r,n] = epi( sum_k pro(x[b,r,k]) * w[n,k] (+ bias[n]) )        dense rows, forward + data gradient
//
// Same contract as the CUDA-core cf_pw_conv path (x3d_pw.cu): BatchNorm / ReLU / SE gate / Swish /
// BatchNorm-backward are a per-(sample,channel) prologue on the A operand, bias / activation
// derivative / residual add / BatchNorm statistics are the epilogue.  Reference call sites:
// conv1x1x1 (x3d_fine.py:100-105; used :149,166,356,370), nn.Linear fc2 (:380), the k=1 Conv1d layers of the
// fusion block (x3d_coarse.py:216-219,232-246,335-336).
//
// Precision: the reference is fp32 and the parity bar is 1e-3 on train-mode logits, which single-pass
// TF32 misses (3e-3, SURVEY 8(a) finding 2).  Every operand is split on the fly into hi = tf32(x) and
// lo = tf32(x - hi) (round to nearest) and three kind::tf32 MMAs (lo*hi + hi*lo + hi*hi) accumulate in fp32 in TMEM ("3xTF32"):
// ~2^-21 relative error, and the extra tensor work hides under the HBM time of these skinny GEMMs.
//
// One CTA = 128 rows of ONE sample x one tile of <= 128 output channels; K is walked in chunks of 32
// floats (one 128-byte swizzle row).  Per chunk: one thread arms an mbarrier and issues a bulk-async
// copy (TMA unit, UBLKCP) of the pre-packed, pre-swizzled hi/lo weight block into shared memory; all
// 256 threads load their part of the activation rows (coalesced 16-byte loads), apply the prologue,
// split hi/lo and store into the canonical K-major SWIZZLE_128B layout; after a proxy fence one thread
// issues the tcgen05.mma's and commits them to an mbarrier that frees the stage (2 stages: the loads of
// chunk c+1 overlap the MMAs of chunk c).  The epilogue drains TMEM with tcgen05.ld into a padded
// shared-memory tile and writes it out row-major with coalesced vector stores, accumulating the
// BatchNorm statistics (one double atomic per CTA and channel).
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

#define TC_THREADS 256
#define TC_NT_MAX 128
#define TC_A_STAGE_BYTES (2 * TC_BM * TC_KC * 4)      /* hi + lo */

#include "tc_ptx.cuh"

__device__ __forceinline__ float tc_swish(float v) { return v * cf_sigmoid(v); }
__device__ __forceinline__ float tc_dswish(float v) {
    float s = cf_sigmoid(v);
    return s * (1.0f + v * (1.0f - s));
}
__device__ __forceinline__ float tc_pro(int mode, float x, float x2, float a, float b, float c) {
    switch (mode) {
        case CF_PRO_AFFINE: return fmaf(a, x, b);
        case CF_PRO_AFFINE_RELU: return fmaxf(fmaf(a, x, b), 0.f);
        case CF_PRO_AFFINE_SWISH: return tc_swish(fmaf(a, x, b));
        case CF_PRO_AFFINE2: return fmaf(a, x, fmaf(b, x2, c));
        default: return x;
    }
}

// ---------------------------------------------------------------------------------------
// weight packing: w[n*w_sn + k*w_sk] -> per (n-tile, k-chunk) block [hi | lo][NTp rows][32 floats], swizzled
// ---------------------------------------------------------------------------------------
__global__ void pw_tc_pack_kernel(const float* __restrict__ w, long long w_sn, long long w_sk, float* __restrict__ pack, int K,
                                  int N, int NT, int NTp, int ntiles, int nchunks) {
    cf_pdl_enter();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)ntiles * nchunks * NTp * 8;
    if (i >= total) return;
    int q = (int)(i & 7);
    long long t = i >> 3;
    int nl = (int)(t % NTp); t /= NTp;
    int c = (int)(t % nchunks);
    int j = (int)(t / nchunks);
    int n = j * NT + nl;
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int k = c * TC_KC + q * 4 + e;
        float v = (nl < NT && n < N && k < K) ? __ldg(w + (long long)n * w_sn + (long long)k * w_sk) : 0.f;
        tf32_split(v, hi[e], lo[e]);
    }
    char* blk = (char*)pack + ((long long)(j * nchunks + c) * 2 * NTp * 128);
    uint32_t off = sw128_off(nl, q);
    *reinterpret_cast<float4*>(blk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(blk + (size_t)NTp * 128 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// ---------------------------------------------------------------------------------------
// main kernel.  AV / EV = vector width (floats) of the activation loads / output accesses
// ---------------------------------------------------------------------------------------
template <int W> struct VecIO;
template <> struct VecIO<4> {
    static __device__ __forceinline__ void ld(const float* p, float* v) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void ldrw(const float* p, float* v) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct VecIO<2> {
    static __device__ __forceinline__ void ld(const float* p, float* v) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void ldrw(const float* p, float* v) {
        float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <> struct VecIO<1> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void ldrw(const float* p, float* v) { v[0] = *p; }
    static __device__ __forceinline__ void st(float* p, const float* v) { *p = v[0]; }
};

// loads the 4 floats x[k..k+3] of one row (zero beyond K) with AV-wide accesses
template <int AV>
__device__ __forceinline__ void load_chunk(const float* row, int k, int K, float* v) {
#pragma unroll
    for (int e = 0; e < 4; e += AV) {
        if (k + e < K) VecIO<AV>::ld(row + k + e, v + e);
        else {
#pragma unroll
            for (int u = 0; u < AV; ++u) v[e + u] = 0.f;
        }
    }
}

template <int AV, int EV>
__global__ void __launch_bounds__(TC_THREADS) pw_tc_kernel(const cf_pw_args a, const float* __restrict__ pack, int R,
                                                           int tiles_per_sample, int NT, int NTp, int nchunks,
                                                           uint32_t tmem_cols) {
    cf_pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_b[2];
    __shared__ __align__(8) uint64_t mma_done[2];
    __shared__ __align__(8) uint64_t acc_done;
    __shared__ uint32_t tmem_addr_s;
    __shared__ float red[2 * TC_NT_MAX];

    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B tiles: 1024-B aligned
    const uint32_t b_stage_bytes = 2u * (uint32_t)NTp * 128u;
    uint8_t* As = base;                                          // [2 stages][hi|lo][128 rows][128 B]
    uint8_t* Bs = base + 2 * TC_A_STAGE_BYTES;                   // [2 stages][hi|lo][NTp rows][128 B]
    float* tab = reinterpret_cast<float*>(Bs + 2 * b_stage_bytes);   // [3][K]
    float* Cs = reinterpret_cast<float*>(base);                  // epilogue tile, aliases the stages
    const int LDC = NTp + 4;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / tiles_per_sample;
    const int r0 = (blockIdx.x - b * tiles_per_sample) * TC_BM;
    const int j = blockIdx.y, n0 = j * NT;
    const int K = a.K, N = a.N;
    const int nvalid = min(NT, N - n0);
    const int rows_valid = min(TC_BM, R - r0);
    const int pro = a.pro_mode;

    if (warp == 0) {
        tmem_alloc(&tmem_addr_s, tmem_cols);
        tmem_relinquish();
    }
    if (tid == 32) {
        mbar_init(&full_b[0], 1); mbar_init(&full_b[1], 1);
        mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1);
        mbar_init(&acc_done, 1);
        fence_mbar_init();
    }
    if (pro != CF_PRO_NONE) {
        for (int i = tid; i < K; i += TC_THREADS) {
            tab[i] = a.pro_a[(size_t)b * K + i];
            tab[K + i] = a.pro_b ? a.pro_b[(size_t)b * K + i] : 0.f;
            tab[2 * K + i] = a.pro_c ? a.pro_c[(size_t)b * K + i] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_addr_s;

    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7, 10),
    // both K-major (bits 15,16 = 0), N >> 3 @ bit 17, M >> 4 @ bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NTp >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

    const int q = tid & 7, rr = tid >> 3;                        // 16-byte chunk within the 128-byte row; row within a pass
    const float* xb = a.x + (size_t)b * R * K;
    const float* x2b = a.x2 ? a.x2 + (size_t)b * R * K : nullptr;
    const float* pk = pack + (size_t)j * nchunks * (2 * NTp * 32);

    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1, k0 = c * TC_KC;
        if (c >= 2) mbar_wait(&mma_done[s], (uint32_t)(((c >> 1) - 1) & 1));      // the MMAs that read stage s retired
        if (tid == 0) {
            mbar_expect_tx(&full_b[s], b_stage_bytes);
            bulk_g2s(Bs + s * b_stage_bytes, pk + (size_t)c * (2 * NTp * 32), b_stage_bytes, &full_b[s]);
        }
        const int k = k0 + q * 4;
        float v[4][4], v2[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {                            // issue all global loads of the chunk first
            const int row = p * 32 + rr;
            if (row < rows_valid) {
                load_chunk<AV>(xb + (size_t)(r0 + row) * K, k, K, v[p]);
                if (pro == CF_PRO_AFFINE2) load_chunk<AV>(x2b + (size_t)(r0 + row) * K, k, K, v2[p]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[p][e] = 0.f;
            }
        }
        float pa[4], pb[4], pc[4];
        if (pro != CF_PRO_NONE) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool kv = k + e < K;
                pa[e] = kv ? tab[k + e] : 0.f;
                pb[e] = kv ? tab[K + k + e] : 0.f;
                pc[e] = kv ? tab[2 * K + k + e] : 0.f;
            }
        }
        uint8_t* a_hi = As + s * TC_A_STAGE_BYTES;
        uint8_t* a_lo = a_hi + TC_BM * TC_KC * 4;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int row = p * 32 + rr;
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float t = v[p][e];
                if (pro != CF_PRO_NONE) t = (row < rows_valid && k + e < K) ? tc_pro(pro, t, pro == CF_PRO_AFFINE2 ? v2[p][e] : 0.f, pa[e], pb[e], pc[e]) : 0.f;
                tf32_split(t, hi[e], lo[e]);
            }
            const uint32_t off = sw128_off(row, q);
            *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();                                     // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            mbar_wait(&full_b[s], (uint32_t)((c >> 1) & 1));     // weight block landed
            tc_fence_after();
            const int nk8 = min(4, (K - k0 + 7) >> 3);
            const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo);
            const uint32_t b_hi_s = smem_u32(Bs + s * b_stage_bytes), b_lo_s = b_hi_s + (uint32_t)NTp * 128u;
            for (int k8 = 0; k8 < nk8; ++k8) {
                const uint32_t ko = (uint32_t)k8 * 32u;          // 8 tf32 = 32 bytes along K inside the swizzle row
                umma_tf32(tmem, make_desc_sw128(a_lo_s + ko), make_desc_sw128(b_hi_s + ko), idesc, (uint32_t)((c | k8) != 0));
                umma_tf32(tmem, make_desc_sw128(a_hi_s + ko), make_desc_sw128(b_lo_s + ko), idesc, 1u);
                umma_tf32(tmem, make_desc_sw128(a_hi_s + ko), make_desc_sw128(b_hi_s + ko), idesc, 1u);
            }
            umma_commit(&mma_done[s]);
            if (c == nchunks - 1) umma_commit(&acc_done);
        }
    }
    mbar_wait(&acc_done, 0u);
    tc_fence_after();

    // ---- epilogue phase 1: TMEM -> padded shared tile (thread = one accumulator row, 8 columns per load)
    {
        const int q4 = warp & 3, half = warp >> 2;
        const int row = q4 * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16);
        for (int cb = half; cb < (NTp >> 3); cb += 2) {
            float r8[8];
            tmem_ld8(tbase + (uint32_t)(cb * 8), r8);
            float* dst = Cs + (size_t)row * LDC + cb * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(r8[0], r8[1], r8[2], r8[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(r8[4], r8[5], r8[6], r8[7]);
        }
    }
    for (int i = tid; i < 2 * TC_NT_MAX; i += TC_THREADS) red[i] = 0.f;
    tc_fence_before();
    __syncthreads();

    // ---- epilogue phase 2: shared tile -> global, coalesced along N
    const int epi = a.epi_mode, smode = a.stats_mode;
    const int CGn = (nvalid + EV - 1) / EV;
    const int rpp = TC_THREADS / CGn;
    const int cg = tid % CGn, rs = tid / CGn;
    if (rs < rpp) {
        const int nl = cg * EV, n = n0 + nl;
        float bi[EV], ea[EV], eb[EV], s1[EV], s2[EV];
#pragma unroll
        for (int e = 0; e < EV; ++e) {
            const bool nv = nl + e < nvalid;
            bi[e] = (nv && a.bias) ? a.bias[n + e] : 0.f;
            ea[e] = (nv && a.epi_a) ? a.epi_a[(size_t)b * N + n + e] : 1.f;
            eb[e] = (nv && a.epi_b) ? a.epi_b[(size_t)b * N + n + e] : 0.f;
            s1[e] = 0.f;
            s2[e] = 0.f;
        }
        const bool need_aux = (epi >= CF_EPI_DRELU && epi <= CF_EPI_ADD_AUX) || epi == CF_EPI_AFFINE_ADD_RELU || smode == CF_STATS_SUM_AUX;
        for (int r = rs; r < rows_valid; r += rpp) {
            const size_t dense = ((size_t)b * R + r0 + r) * N + n;
            float vv[EV], ax[EV];
#pragma unroll
            for (int e = 0; e < EV; ++e) { vv[e] = Cs[(size_t)r * LDC + nl + e] + bi[e]; ax[e] = 0.f; }
            if (need_aux) VecIO<EV>::ld(a.aux + dense, ax);
#pragma unroll
            for (int e = 0; e < EV; ++e) {
                float t = vv[e];
                if (epi == CF_EPI_RELU) t = fmaxf(t, 0.f);
                else if (epi == CF_EPI_DRELU) t = (fmaf(ea[e], ax[e], eb[e]) > 0.f) ? t : 0.f;
                else if (epi == CF_EPI_DSWISH) t *= tc_dswish(fmaf(ea[e], ax[e], eb[e]));
                else if (epi == CF_EPI_ADD_AUX) t += ax[e];
                else if (epi == CF_EPI_SIGMOID) t = cf_sigmoid(t);
                else if (epi == CF_EPI_AFFINE) t = fmaf(ea[e], t, eb[e]);
                else if (epi == CF_EPI_AFFINE_ADD_RELU) t = fmaxf(fmaf(ea[e], t, eb[e]) + ax[e], 0.f);
                vv[e] = t;
                s1[e] += t;
                s2[e] += (smode == CF_STATS_SUM_AUX) ? t * ax[e] : t * t;
            }
            if (a.accumulate) {
                float old[EV];
                VecIO<EV>::ldrw(a.y + dense, old);
#pragma unroll
                for (int e = 0; e < EV; ++e) vv[e] += old[e];
            }
            VecIO<EV>::st(a.y + dense, vv);
        }
        if (smode != CF_STATS_NONE) {
#pragma unroll
            for (int e = 0; e < EV; ++e)
                if (nl + e < nvalid) { atomicAdd(red + nl + e, s1[e]); atomicAdd(red + TC_NT_MAX + nl + e, s2[e]); }
        }
    }
    __syncthreads();
    if (smode != CF_STATS_NONE && tid < nvalid) {
        double* st = a.stats + ((size_t)b * N + n0 + tid) * 2;
        atomicAdd(st, (double)red[tid]);
        atomicAdd(st + 1, (double)red[TC_NT_MAX + tid]);
    }
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct TcTiling { int ntiles, NT, NTp, nchunks; };

static TcTiling tc_tiling(int K, int N) {
    TcTiling t;
    t.ntiles = (N + TC_NT_MAX - 1) / TC_NT_MAX;
    int nt = (N + t.ntiles - 1) / t.ntiles;
    t.NT = (nt + 7) / 8 * 8;                       // tile starts stay 32-byte aligned
    t.ntiles = (N + t.NT - 1) / t.NT;
    t.NTp = (t.NT + 15) / 16 * 16;                 // UMMA M=128 needs N % 16 == 0
    t.nchunks = (K + TC_KC - 1) / TC_KC;
    return t;
}

static size_t tc_v1_ws_bytes(int K, int N) {
    if (K <= 0 || N <= 0) return 0;
    TcTiling t = tc_tiling(K, N);
    return (size_t)t.ntiles * t.nchunks * 2 * t.NTp * 128;
}

// host launcher of the weight packer, shared with the persistent kernel (x3d_pw_tc2.cu)
void cf_pw_tc_pack_launch(const float* w, long long w_sn, long long w_sk, float* pack, int K, int N, int NT, int NTp, int ntiles,
                          int nchunks, cudaStream_t stream) {
    long long total = (long long)ntiles * nchunks * NTp * 8;
    cf_launch(pw_tc_pack_kernel, (unsigned)cf_cdiv64(total, 256), 256, 0, stream, w, w_sn, w_sk, pack, K, N, NT, NTp, ntiles, nchunks);
}

static size_t tc_smem_bytes(const TcTiling& t, int K) {
    return 1024 + 2 * (size_t)TC_A_STAGE_BYTES + 2 * (size_t)(2 * t.NTp * 128) + 3 * (size_t)K * 4;
}

template <int AV, int EV>
static int launch_tc(const cf_pw_args* a, const TcTiling& t, int R, uint32_t tmem_cols, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024) {     // opt in per launch: the attribute is per device and the call is cheap
        cudaError_t e = cudaFuncSetAttribute(pw_tc_kernel<AV, EV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            cf_set_error("cf_pw_conv_tc: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(e));
            return CF_ERR_CUDA;
        }
    }
    int tps = cf_cdiv(R, TC_BM);
    dim3 grid((unsigned)(tps * a->B), (unsigned)t.ntiles);
    cf_launch(pw_tc_kernel<AV, EV>, grid, TC_THREADS, smem, stream, *a, a->wpack, R, tps, t.NT, t.NTp, t.nchunks, tmem_cols);
    return CF_OK;
}

// the first tensor-core kernel (one tile per CTA), kept behind CFNET_PW_TC_V1=1 for A/B measurements
int cf_pw_conv_tc_v1(const cf_pw_args* a, cudaStream_t stream) {
    const int K = a->K, N = a->N;
    const int R = a->g.T * a->g.H * a->g.W;
    TcTiling t = tc_tiling(K, N);
    CF_CHECK_ARG(a->wpack_bytes >= (int64_t)tc_v1_ws_bytes(K, N), "weight-pack workspace too small");
    CF_CHECK_ARG((((uintptr_t)a->wpack) & 127) == 0, "weight-pack workspace must be 128-byte aligned");
    size_t smem = tc_smem_bytes(t, K);
    CF_CHECK_ARG(smem <= 225 * 1024, "K too large for the tensor-core path");
    CF_CHECK_ARG((long long)cf_cdiv(R, TC_BM) * a->B < (1LL << 31), "too many row tiles");
    {
        long long total = (long long)t.ntiles * t.nchunks * t.NTp * 8;
        cf_launch(pw_tc_pack_kernel, (unsigned)cf_cdiv64(total, 256), 256, 0, stream, a->w, a->w_sn, a->w_sk, a->wpack, K, N, t.NT, t.NTp,
                                                                            t.ntiles, t.nchunks);
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < t.NTp) tmem_cols <<= 1;
    uintptr_t xa = (uintptr_t)a->x | (uintptr_t)(a->x2 ? a->x2 : a->x);
    int av = ((K & 3) == 0 && (xa & 15) == 0) ? 4 : (((K & 1) == 0 && (xa & 7) == 0) ? 2 : 1);
    uintptr_t ya = (uintptr_t)a->y | (uintptr_t)(a->aux ? a->aux : a->y);
    int ev = ((N & 3) == 0 && (ya & 15) == 0) ? 4 : (((N & 1) == 0 && (ya & 7) == 0) ? 2 : 1);
    int rc;
#define CF_TC_CASE(AV_, EV_) if (av == AV_ && ev == EV_) rc = launch_tc<AV_, EV_>(a, t, R, tmem_cols, smem, stream); else
    CF_TC_CASE(4, 4) CF_TC_CASE(4, 2) CF_TC_CASE(4, 1) CF_TC_CASE(2, 4) CF_TC_CASE(2, 2) CF_TC_CASE(2, 1) CF_TC_CASE(1, 4)
    CF_TC_CASE(1, 2) CF_TC_CASE(1, 1) rc = CF_ERR_ARG;
#undef CF_TC_CASE
    if (rc != CF_OK) return rc;
    CF_COUNT_LAUNCH(2);
    CF_CHECK_LAUNCH();
    return rc;
}
