// Device code of the persistent tcgen05 pointwise GEMM (x3d_pw_tc2.cu): 8 producer warps (4 rows per thread and chunk),
// one MMA-issuer warp + one TMA-loader lane in the same warpgroup, 8 epilogue warps; register budgets 120 / 32 / 104.
namespace P2_NS {

// sigmoid from ex2.approx / rcp.approx (2^-22 relative: at the 3xTF32 level, far inside the 1e-3 parity bar): 5 issue
// slots (2 of them MUFU) instead of ~16 for expf + IEEE division, which made the Swish producers instruction-bound
__device__ __forceinline__ float p2_sigmoid(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
__device__ __forceinline__ float p2_swish(float v) { return v * p2_sigmoid(v); }
__device__ __forceinline__ float p2_dswish(float v) {
    float s = p2_sigmoid(v);
    return s * (1.0f + v * (1.0f - s));
}
template <int PRO>
__device__ __forceinline__ float p2_pro(float x, float x2, float a, float b, float c) {
    if (PRO == CF_PRO_AFFINE) return fmaf(a, x, b);
    if (PRO == CF_PRO_AFFINE_RELU) return fmaxf(fmaf(a, x, b), 0.f);
    if (PRO == CF_PRO_AFFINE_SWISH) return p2_swish(fmaf(a, x, b));
    if (PRO == CF_PRO_AFFINE2) return fmaf(a, x, fmaf(b, x2, c));
    return x;
}

template <int W> struct P2Vec;
template <> struct P2Vec<4> {
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
template <> struct P2Vec<2> {
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
template <> struct P2Vec<1> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void ldrw(const float* p, float* v) { v[0] = *p; }
    static __device__ __forceinline__ void st(float* p, const float* v) { *p = v[0]; }
};

// position of one (tile, k-chunk) work item in a CTA's sequence.  tile = (b * tps + rtile) * ntiles + j; a CTA steps by
// gridDim.x tiles at a time, done incrementally (g_j = grid % ntiles, g_rt = grid / ntiles from the host): no divisions
// in the per-item path.
struct P2Item {
    int c, b, rtile, r0, j;
    bool valid;
};
__device__ __forceinline__ void p2_first(P2Item& it, const P2Params& p) {
    const int tile = (int)blockIdx.x;              // blockIdx.x < total_tiles (grid = min(SMs, tiles))
    const int rt = tile / p.ntiles;
    it.j = tile - rt * p.ntiles;
    it.b = rt / p.tps;
    it.rtile = rt - it.b * p.tps;
    it.r0 = it.rtile * TC_BM;
    it.c = 0;
    it.valid = true;
}
__device__ __forceinline__ void p2_next_tile(P2Item& it, const P2Params& p) {
    it.j += p.g_j;
    it.rtile += p.g_rt;
    if (it.j >= p.ntiles) { it.j -= p.ntiles; ++it.rtile; }
    while (it.rtile >= p.tps) { it.rtile -= p.tps; ++it.b; }
    it.r0 = it.rtile * TC_BM;
    it.valid = it.b < p.B;
}
__device__ __forceinline__ void p2_advance(P2Item& it, const P2Params& p) {
    if (++it.c == p.nchunks) {
        it.c = 0;
        p2_next_tile(it, p);
    }
}

// row of the dense side -> row of the strided volume (within the sample)
__device__ __forceinline__ long long p2_map_row(const P2Params& p, int r) {
    const int w = r % p.gW, q = r / p.gW;
    const int h = q % p.gH, t = q / p.gH;
    return ((long long)(t * p.gst) * p.gHi + h * p.gsh) * p.gWi + w * p.gsw;
}

// ---------------------------------------------------------------------------------------
// producers
// ---------------------------------------------------------------------------------------
template <int AV, bool X2>
__device__ __forceinline__ void p2_load_item(const cf_pw_args& a, const P2Params& p, const P2Item& it, int q, int rr,
                                             float (&v)[P2_PROD_PASSES][4], float (&v2)[X2 ? P2_PROD_PASSES : 1][4]) {
    const int K = a.K;
    const int k = it.c * TC_KC + q * 4;
    const int rows_valid = min(TC_BM, p.R - it.r0);
    constexpr int RPP = P2_PROD_WARPS * 4;                        // rows per pass
    if (p.gmode == 1) {                                           // gathered rows (strided 1x1x1 conv): one row map per pass
#pragma unroll
        for (int pp = 0; pp < P2_PROD_PASSES; ++pp) {
            const int row = pp * RPP + rr;
            const bool rv = row < rows_valid;
            const float* xp = a.x + (size_t)it.b * p.g_sample_stride + (rv ? p2_map_row(p, it.r0 + row) : 0) * K + k;
#pragma unroll
            for (int e = 0; e < 4; e += AV) {
                if (rv && k + e < K) P2Vec<AV>::ld(xp + e, &v[pp][e]);
                else {
#pragma unroll
                    for (int u = 0; u < AV; ++u) v[pp][e + u] = 0.f;
                }
            }
        }
        return;
    }
    const size_t off = ((size_t)it.b * p.R + it.r0 + rr) * K + k;
    const float* xp = a.x + off;
    const float* x2p = X2 ? a.x2 + off : nullptr;
    const size_t step = (size_t)RPP * K;
#pragma unroll
    for (int pp = 0; pp < P2_PROD_PASSES; ++pp, xp += step, x2p += X2 ? step : 0) {
        const bool rv = pp * RPP + rr < rows_valid;
#pragma unroll
        for (int e = 0; e < 4; e += AV) {
            if (rv && k + e < K) {
                P2Vec<AV>::ld(xp + e, &v[pp][e]);
                if (X2) P2Vec<AV>::ld(x2p + e, &v2[X2 ? pp : 0][e]);
            } else {
#pragma unroll
                for (int u = 0; u < AV; ++u) {
                    v[pp][e + u] = 0.f;
                    if (X2) v2[X2 ? pp : 0][e + u] = 0.f;
                }
            }
        }
    }
}

template <int AV, int PRO>
__device__ __forceinline__ void p2_producer(const cf_pw_args& a, const P2Params& p, uint8_t* stages, float* tab, uint64_t* full,
                                            uint64_t* empty, const float* __restrict__ pack, int tid) {
    constexpr bool X2 = PRO == CF_PRO_AFFINE2;
    constexpr int NSET = X2 ? 2 : (P2_PROD_PASSES == 4 ? 4 : 3);  // register sets of loads in flight (8 or 16 floats each)
    const int lane = tid & 31;
    const int q = tid & 7, rr = tid >> 3;
    float v[NSET][P2_PROD_PASSES][4];
    float v2[NSET][X2 ? P2_PROD_PASSES : 1][4];
    P2Item ld, pr;
    p2_first(ld, p);
    pr = ld;
#pragma unroll
    for (int u = 0; u < NSET - 1; ++u) {
        if (ld.valid) {
            p2_load_item<AV, X2>(a, p, ld, q, rr, v[u], v2[u]);
            p2_advance(ld, p);
        }
    }
    int cur_b = -1;
    int s = 0;
    uint32_t ph = 0;
#ifdef CFNET_P2_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tt = P2_T0();
    const long long tstart = tt;
#endif
    while (pr.valid) {
#pragma unroll
        for (int u = 0; u < NSET; ++u) {
            if (!pr.valid) break;
            if (ld.valid) {
                p2_load_item<AV, X2>(a, p, ld, q, rr, v[(u + NSET - 1) % NSET], v2[(u + NSET - 1) % NSET]);
                p2_advance(ld, p);
            }
            // ---- prologue tables of this tile's sample (shared by the producer warps only)
            if (PRO != CF_PRO_NONE && pr.c == 0 && pr.b != cur_b) {
                named_bar_sync(1, P2_PROD_THREADS);
                for (int t = tid; t < p.KP; t += P2_PROD_THREADS) {
                    const bool kv = t < a.K;
                    tab[t] = kv ? a.pro_a[(size_t)pr.b * a.K + t] : 0.f;
                    tab[p.KP + t] = (kv && a.pro_b) ? a.pro_b[(size_t)pr.b * a.K + t] : 0.f;
                    tab[2 * p.KP + t] = (kv && a.pro_c) ? a.pro_c[(size_t)pr.b * a.K + t] : 0.f;
                }
                named_bar_sync(1, P2_PROD_THREADS);
                cur_b = pr.b;
            }
            P2_ACC(0, tt);                                   // 0: loads issued + tables
            mbar_wait_b(&empty[s], ph ^ 1u);
            P2_ACC(1, tt);                                   // 1: wait for a free stage
            uint8_t* stage = stages + (size_t)s * p.stage_bytes;
            if (!p.resident && tid == 0) {
                mbar_expect_tx(&full[s], p.b_chunk_bytes);
                bulk_g2s(stage + P2_A_STAGE, pack + ((size_t)pr.j * p.nchunks + pr.c) * (p.b_chunk_bytes / 4), p.b_chunk_bytes,
                         &full[s]);
            }
            const int k = pr.c * TC_KC + q * 4;
            float pa[4], pb[4], pc[4];
            if (PRO != CF_PRO_NONE) {
                const float4 ta = *reinterpret_cast<const float4*>(tab + k);
                const float4 tb = *reinterpret_cast<const float4*>(tab + p.KP + k);
                pa[0] = ta.x; pa[1] = ta.y; pa[2] = ta.z; pa[3] = ta.w;
                pb[0] = tb.x; pb[1] = tb.y; pb[2] = tb.z; pb[3] = tb.w;
                if (X2) {
                    const float4 tc = *reinterpret_cast<const float4*>(tab + 2 * p.KP + k);
                    pc[0] = tc.x; pc[1] = tc.y; pc[2] = tc.z; pc[3] = tc.w;
                }
            }
            const int rows_valid = min(TC_BM, p.R - pr.r0);
            uint8_t* a_hi = stage;
            uint8_t* a_lo = stage + TC_BM * TC_KC * 4;
#pragma unroll
            for (int pp = 0; pp < P2_PROD_PASSES; ++pp) {
                const int row = pp * (P2_PROD_WARPS * 4) + rr;
                float hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float t = v[u][pp][e];
                    if (PRO != CF_PRO_NONE) {
                        // padded k: tables are zero there (every prologue maps 0 with zero tables to 0); padded rows: mask
                        t = p2_pro<PRO>(t, X2 ? v2[u][X2 ? pp : 0][e] : 0.f, pa[e], pb[e], X2 ? pc[e] : 0.f);
                        if (row >= rows_valid) t = 0.f;
                    }
                    tf32_split(t, hi[e], lo[e]);
                }
                const uint32_t off = sw128_off(row, q);
                *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            P2_ACC(2, tt);                                   // 2: prologue + split + stores
            fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            P2_ACC(3, tt);                                   // 3: proxy fence + warp sync
            if (lane == 0) mbar_arrive(&full[s]);
            P2_ACC(4, tt);                                   // 4: arrive
            if (++s == p.nstages) { s = 0; ph ^= 1u; }
            p2_advance(pr, p);
        }
    }
#ifdef CFNET_P2_TIMING
    if (p.timing && blockIdx.x == 0 && tid == 0) {
        for (int i = 0; i < 5; ++i) p2_dbg[i] = tacc[i];
        p2_dbg[7] = clock64() - tstart;
    }
#endif
}

// ---------------------------------------------------------------------------------------
// TMA-fed producers.  One loader lane (in the MMA warpgroup) walks the same item list and keeps p.nraw raw activation
// tiles in flight with cp.async.bulk.tensor (tensor maps of cf_make_row_tmap): the loads no longer live in producer
// registers, so their latency is hidden by the depth of the raw ring instead of by 2-4 register sets per thread.  The
// producers read a landed tile from shared memory, hand the raw stage straight back, and then do what they did before:
// prologue -> hi/lo TF32 split -> SWIZZLE_128B operand stage.
//   FOLD == 1: a raw stage is one [128 rows][32 floats] k-chunk (pitch 128 B);
//   FOLD  > 1: (K not a multiple of 4, K <= 64) a raw stage is the whole [128 rows][K] tile (pitch K*4 B), read chunk by chunk.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void p2_loader(const cf_pw_args& a, const P2Params& p, uint8_t* raw, uint64_t* rfull, uint64_t* rempty,
                                          const CUtensorMap* tmx, const CUtensorMap* tmx2) {
    const bool x2 = a.pro_mode == CF_PRO_AFFINE2;
    const uint32_t bytes = p.raw_in_bytes * (x2 ? 2u : 1u);
    tma_prefetch_desc(tmx);
    if (x2) tma_prefetch_desc(tmx2);
    int rs = 0;
    uint32_t rph = 0;
    P2Item it;
    for (p2_first(it, p); it.valid; p2_advance(it, p)) {
        if (p.fold != 1 && it.c != 0) continue;
        mbar_wait_b(&rempty[rs], rph ^ 1u);                  // every producer warp has read this raw stage
        uint8_t* dst = raw + (size_t)rs * p.raw_stage_bytes;
        mbar_expect_tx(&rfull[rs], bytes);
        const int c0 = p.fold == 1 ? it.c * TC_KC : 0, c1 = it.r0 / p.fold;
        tma_load_3d(dst, tmx, c0, c1, it.b, &rfull[rs]);
        if (x2) tma_load_3d(dst + p.raw_in_bytes, tmx2, c0, c1, it.b, &rfull[rs]);
        if (++rs == p.nraw) { rs = 0; rph ^= 1u; }
    }
}

template <int FOLD>
__device__ __forceinline__ void p2_raw_read(const uint8_t* rx, int row, int q, int k, int K, float* v) {
    if (FOLD == 1) {
        const float4 t = *reinterpret_cast<const float4*>(rx + row * 128 + q * 16);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if (FOLD == 2) {
        const float* src = reinterpret_cast<const float*>(rx) + row * K + k;
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
            if (k + e < K) {
                const float2 t = *reinterpret_cast<const float2*>(src + e);
                v[e] = t.x; v[e + 1] = t.y;
            } else {
                v[e] = 0.f; v[e + 1] = 0.f;
            }
        }
    } else {
        const float* src = reinterpret_cast<const float*>(rx) + row * K + k;
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (k + e < K) ? src[e] : 0.f;
    }
}

template <int FOLD, int PRO>
__device__ __forceinline__ void p2_producer_tma(const cf_pw_args& a, const P2Params& p, uint8_t* stages, float* tab, uint64_t* full,
                                                uint64_t* empty, uint64_t* rfull, uint64_t* rempty, const uint8_t* raw,
                                                const float* __restrict__ pack, int tid) {
    constexpr bool X2 = PRO == CF_PRO_AFFINE2;
    const int lane = tid & 31;
    const int q = tid & 7, rr = tid >> 3;
    const int K = a.K;
    P2Item pr;
    p2_first(pr, p);
    int cur_b = -1, s = 0, rs = 0;
    uint32_t ph = 0, rph = 0;
    while (pr.valid) {
        if (PRO != CF_PRO_NONE && pr.c == 0 && pr.b != cur_b) {  // prologue tables of this tile's sample
            named_bar_sync(1, P2_PROD_THREADS);
            for (int t = tid; t < p.KP; t += P2_PROD_THREADS) {
                const bool kv = t < K;
                tab[t] = kv ? a.pro_a[(size_t)pr.b * K + t] : 0.f;
                tab[p.KP + t] = (kv && a.pro_b) ? a.pro_b[(size_t)pr.b * K + t] : 0.f;
                tab[2 * p.KP + t] = (kv && a.pro_c) ? a.pro_c[(size_t)pr.b * K + t] : 0.f;
            }
            named_bar_sync(1, P2_PROD_THREADS);
            cur_b = pr.b;
        }
        if (FOLD == 1 || pr.c == 0) mbar_wait_b(&rfull[rs], rph);        // the raw tile has landed
        const uint8_t* rx = raw + (size_t)rs * p.raw_stage_bytes;
        const int k = pr.c * TC_KC + q * 4;
        float v[P2_PROD_PASSES][4], v2[X2 ? P2_PROD_PASSES : 1][4];
#pragma unroll
        for (int pp = 0; pp < P2_PROD_PASSES; ++pp) {
            const int row = pp * (P2_PROD_WARPS * 4) + rr;
            p2_raw_read<FOLD>(rx, row, q, k, K, v[pp]);
            if (X2) p2_raw_read<FOLD>(rx + p.raw_in_bytes, row, q, k, K, v2[X2 ? pp : 0]);
        }
        if (FOLD == 1 || pr.c == p.nchunks - 1) {                        // values are in registers: give the raw stage back
            __syncwarp();
            if (lane == 0) mbar_arrive(&rempty[rs]);
            if (++rs == p.nraw) { rs = 0; rph ^= 1u; }
        }
        mbar_wait_b(&empty[s], ph ^ 1u);
        uint8_t* stage = stages + (size_t)s * p.stage_bytes;
        if (!p.resident && tid == 0) {
            mbar_expect_tx(&full[s], p.b_chunk_bytes);
            bulk_g2s(stage + P2_A_STAGE, pack + ((size_t)pr.j * p.nchunks + pr.c) * (p.b_chunk_bytes / 4), p.b_chunk_bytes, &full[s]);
        }
        float pa[4], pb[4], pc[4];
        if (PRO != CF_PRO_NONE) {
            const float4 ta = *reinterpret_cast<const float4*>(tab + k);
            const float4 tb = *reinterpret_cast<const float4*>(tab + p.KP + k);
            pa[0] = ta.x; pa[1] = ta.y; pa[2] = ta.z; pa[3] = ta.w;
            pb[0] = tb.x; pb[1] = tb.y; pb[2] = tb.z; pb[3] = tb.w;
            if (X2) {
                const float4 tc = *reinterpret_cast<const float4*>(tab + 2 * p.KP + k);
                pc[0] = tc.x; pc[1] = tc.y; pc[2] = tc.z; pc[3] = tc.w;
            }
        }
        const int rows_valid = min(TC_BM, p.R - pr.r0);
        uint8_t* a_hi = stage;
        uint8_t* a_lo = stage + TC_BM * TC_KC * 4;
#pragma unroll
        for (int pp = 0; pp < P2_PROD_PASSES; ++pp) {
            const int row = pp * (P2_PROD_WARPS * 4) + rr;
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float t = v[pp][e];
                if (PRO != CF_PRO_NONE) {
                    t = p2_pro<PRO>(t, X2 ? v2[X2 ? pp : 0][e] : 0.f, pa[e], pb[e], X2 ? pc[e] : 0.f);
                    if (row >= rows_valid) t = 0.f;
                }
                tf32_split(t, hi[e], lo[e]);
            }
            const uint32_t off = sw128_off(row, q);
            *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (++s == p.nstages) { s = 0; ph ^= 1u; }
        p2_advance(pr, p);
    }
}

// ---------------------------------------------------------------------------------------
// MMA issuer: one warp runs the loop convergently (descriptor arithmetic stays warp-uniform), one elected lane issues.
// Issuing from a divergent `if (lane == 0)` inside a producer warp measured ~70 cycles per tcgen05.mma (per-lane
// descriptor math moved to uniform registers one MMA at a time) and made that warp the pipeline's slowest stage.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool p2_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void p2_mma_warp(const cf_pw_args& a, const P2Params& p, uint8_t* stages, uint8_t* wres, uint64_t* full,
                                            uint64_t* empty, uint64_t* tfull, uint64_t* tempty, uint64_t* wres_bar, uint32_t tmem) {
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7, 10),
    // both K-major (bits 15,16 = 0), N >> 3 @ bit 17, M >> 4 @ bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NTp >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint64_t lo_off = (uint64_t)((TC_BM * TC_KC * 4) >> 4);          // A lo tile follows A hi (descriptor address units: 16 B)
    const uint64_t blo_off = (uint64_t)(((uint32_t)p.NTp * 128u) >> 4);    // B lo tile follows B hi
    const uint64_t stages_desc = make_desc_sw128(smem_u32(stages));
    const uint64_t wres_desc = make_desc_sw128(smem_u32(wres));
#ifdef CFNET_P2_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tt = P2_T0();
    const long long tstart = tt;
#endif
    if (p.resident) mbar_wait_b(wres_bar, 0u);
    int s = 0;
    uint32_t ph = 0, tcount = 0;
    P2Item it;
    for (p2_first(it, p); it.valid; p2_next_tile(it, p), ++tcount) {
        const int acc = (int)(tcount & 1u);
        P2_ACC(0, tt);
        mbar_wait_b(&tempty[acc], ((tcount >> 1) & 1u) ^ 1u);             // the epilogue drained this accumulator
        P2_ACC(1, tt);                                                     // 1: wait for the accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)(acc * p.acc_stride);
        for (int c = 0; c < p.nchunks; ++c) {
            mbar_wait_b(&full[s], ph);                                     // producers (and the weight block) filled this stage
            P2_ACC(2, tt);                                                 // 2: wait for a full stage
            tc_fence_after();
            const uint64_t a_hi = stages_desc + (uint64_t)(((uint32_t)s * p.stage_bytes) >> 4);
            const uint64_t a_lo = a_hi + lo_off;
            const uint64_t b_hi = p.resident ? wres_desc + (uint64_t)(((uint32_t)c * p.b_chunk_bytes) >> 4) : a_hi + (uint64_t)(P2_A_STAGE >> 4);
            const uint64_t b_lo = b_hi + blo_off;
            const int nk8 = min(4, (a.K - c * TC_KC + 7) >> 3);
            if (p2_elect_one()) {
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8) {
                    if (k8 < nk8) {
                        const uint64_t ko = (uint64_t)(k8 * 2);            // 8 tf32 = 32 bytes along K inside the swizzle row
                        if (!p.dbg_1x) {
                            umma_tf32(d_tmem, a_lo + ko, b_hi + ko, idesc, (uint32_t)((c | k8) != 0));
                            umma_tf32(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                            umma_tf32(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
                        } else {                                           // debug: single-pass TF32 (wrong results, timing only)
                            umma_tf32(d_tmem, a_hi + ko, b_hi + ko, idesc, (uint32_t)((c | k8) != 0));
                        }
                    }
                }
                umma_commit(&empty[s]);                                    // stage reusable once these MMAs retire
                if (c == p.nchunks - 1) umma_commit(&tfull[acc]);          // accumulator complete
            }
            __syncwarp();
            P2_ACC(3, tt);                                                 // 3: issue + commits
            if (++s == p.nstages) { s = 0; ph ^= 1u; }
        }
    }
#ifdef CFNET_P2_TIMING
    if (p.timing && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
        for (int i = 0; i < 4; ++i) p2_dbg[16 + i] = tacc[i];
        p2_dbg[20] = clock64() - tstart;
    }
#endif
}

// ---------------------------------------------------------------------------------------
// epilogue: one 32-column slab, shared tile -> global (coalesced along N)
// ---------------------------------------------------------------------------------------

// packed fp32x2 arithmetic (sm_100 FADD2 / FFMA2): halves the issue slots of the bias add and the statistics
__device__ __forceinline__ void p2_add2(float& a0, float& a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
        : "+f"(a0), "+f"(a1)
        : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void p2_fma2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(c0), "+f"(c1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// FULL: all 128 rows of the tile are valid (every tile but the last one of a sample): no per-pass row checks
template <int EV, int EPI, int SMODE, bool FULL>
__device__ __forceinline__ void p2_store_rows(const cf_pw_args& a, const P2Params& p, int b, int r0, int ncol, const float* __restrict__ cp,
                                              float* __restrict__ dp, const float* __restrict__ ap, size_t gstep, int rs, int rows_valid,
                                              const float* bi, const float* ea, const float* eb, bool has_bias, float* s1, float* s2) {
    constexpr int CPR = 32 / EV, RPP = 32 / CPR, NPASS = 32 / RPP;      // a warp stores its own 32 rows, RPP rows per pass
    constexpr bool EPI_AUX = EPI == CF_EPI_DRELU || EPI == CF_EPI_DSWISH || EPI == CF_EPI_ADD_AUX || EPI == CF_EPI_AFFINE_ADD_RELU;
    constexpr bool NEED_AUX = EPI_AUX || SMODE == CF_STATS_SUM_AUX;
    float ax[NEED_AUX ? NPASS : 1][EV];
    if (NEED_AUX) {
#pragma unroll
        for (int i = 0; i < NPASS; ++i, ap += gstep) {
            if (FULL || rs + i * RPP < rows_valid) P2Vec<EV>::ld(ap, ax[NEED_AUX ? i : 0]);
            else {
#pragma unroll
                for (int e = 0; e < EV; ++e) ax[NEED_AUX ? i : 0][e] = 0.f;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NPASS; ++i, dp += gstep, cp += RPP * P2_CS_LD) {
        if (!FULL && rs + i * RPP >= rows_valid) break;
        float vv[EV];
        P2Vec<EV>::ldrw(cp, vv);
        if (has_bias) {
#pragma unroll
            for (int e = 0; e < EV; e += 2) p2_add2(vv[e], vv[e + 1], bi[e], bi[e + 1]);
        }
        const float* axe = ax[NEED_AUX ? i : 0];
#pragma unroll
        for (int e = 0; e < EV; ++e) {
            float t = vv[e];
            if (EPI == CF_EPI_RELU) t = fmaxf(t, 0.f);
            else if (EPI == CF_EPI_DRELU) t = (fmaf(ea[e], axe[e], eb[e]) > 0.f) ? t : 0.f;
            else if (EPI == CF_EPI_DSWISH) t *= p2_dswish(fmaf(ea[e], axe[e], eb[e]));
            else if (EPI == CF_EPI_ADD_AUX) t += axe[e];
            else if (EPI == CF_EPI_SIGMOID) t = p2_sigmoid(t);
            else if (EPI == CF_EPI_AFFINE) t = fmaf(ea[e], t, eb[e]);
            else if (EPI == CF_EPI_AFFINE_ADD_RELU) t = fmaxf(fmaf(ea[e], t, eb[e]) + axe[e], 0.f);
            vv[e] = t;
        }
        if (SMODE != CF_STATS_NONE) {
#pragma unroll
            for (int e = 0; e < EV; e += 2) {
                p2_add2(s1[e], s1[e + 1], vv[e], vv[e + 1]);
                if (SMODE == CF_STATS_SUM_AUX) p2_fma2(s2[e], s2[e + 1], vv[e], vv[e + 1], axe[e], axe[e + 1]);
                else p2_fma2(s2[e], s2[e + 1], vv[e], vv[e + 1], vv[e], vv[e + 1]);
            }
        }
        if (p.gmode == 2) {                                      // strided 1x1x1 conv, data gradient: y[map(row)] += result
            float* sp = a.y + (size_t)b * p.g_sample_stride + p2_map_row(p, r0 + rs + i * RPP) * a.N + ncol;
            float old[EV];
            P2Vec<EV>::ldrw(sp, old);
#pragma unroll
            for (int e = 0; e < EV; ++e) vv[e] += old[e];
            P2Vec<EV>::st(sp, vv);
        } else {
            P2Vec<EV>::st(dp, vv);
        }
    }
}

template <int EV, int EPI, int SMODE>
__device__ __forceinline__ void p2_store_slab(const cf_pw_args& a, const P2Params& p, const float* __restrict__ Cs, float* __restrict__ redw,
                                              int b, int r0, int rows_valid, int R, int n0, int col0, int nvalid, int gt) {
    constexpr int CPR = 32 / EV;             // column groups per row
    constexpr int RPP = 32 / CPR;            // rows per pass of one warp
    const int N = a.N;
    const int lane = gt & 31;
    const int cg = lane % CPR, rs = (gt & ~31) + lane / CPR;     // the warp stores the 32 rows it brought from TMEM itself
    const int nl = col0 + cg * EV;           // column within the channel tile
    const bool active = nl < nvalid;         // nvalid and nl are multiples of EV: a column group is all-valid or all-padding
    float s1[EV], s2[EV];
#pragma unroll
    for (int e = 0; e < EV; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
    if (active) {
        const int n = n0 + nl;
        float bi[EV], ea[EV], eb[EV];
#pragma unroll
        for (int e = 0; e < EV; ++e) {
            bi[e] = a.bias ? a.bias[n + e] : 0.f;
            constexpr bool EPI_TAB = EPI == CF_EPI_DRELU || EPI == CF_EPI_DSWISH || EPI == CF_EPI_AFFINE || EPI == CF_EPI_AFFINE_ADD_RELU;
            ea[e] = EPI_TAB ? a.epi_a[(size_t)b * N + n + e] : 1.f;
            eb[e] = EPI_TAB ? a.epi_b[(size_t)b * N + n + e] : 0.f;
        }
        const size_t g0 = ((size_t)b * R + r0 + rs) * N + n;
        const size_t gstep = (size_t)RPP * N;
        const float* cp = Cs + rs * P2_CS_LD + cg * EV;
        const float* ap = a.aux ? a.aux + g0 : nullptr;
        if (rows_valid == TC_BM)
            p2_store_rows<EV, EPI, SMODE, true>(a, p, b, r0, n, cp, a.y + g0, ap, gstep, rs, rows_valid, bi, ea, eb, a.bias != nullptr, s1, s2);
        else
            p2_store_rows<EV, EPI, SMODE, false>(a, p, b, r0, n, cp, a.y + g0, ap, gstep, rs, rows_valid, bi, ea, eb, a.bias != nullptr, s1, s2);
    }
    if (SMODE != CF_STATS_NONE) {
        // column sums: the 32 / CPR row groups of a warp meet by shuffle, then the first CPR lanes add into THIS WARP's
        // row of the per-CTA table (a slab belongs to one group, a table row to one warp: no atomics, no barrier)
#pragma unroll
        for (int o = CPR; o < 32; o <<= 1)
#pragma unroll
            for (int e = 0; e < EV; ++e) {
                s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
                s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
            }
        if (active && (gt & 31) < CPR) {
            float* r1 = redw + n0 + nl;
#pragma unroll
            for (int e = 0; e < EV; ++e) {
                r1[e] += s1[e];
                r1[P2_RED_N + e] += s2[e];
            }
        }
    }
}

__device__ __forceinline__ void p2_flush_stats(const cf_pw_args& a, float* red, int b, int et) {
    named_bar_sync(4, P2_EPI_THREADS);
    for (int i = et; i < a.N; i += P2_EPI_THREADS) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            t1 += red[w * 2 * P2_RED_N + i];
            t2 += red[w * 2 * P2_RED_N + P2_RED_N + i];
            red[w * 2 * P2_RED_N + i] = 0.f;
            red[w * 2 * P2_RED_N + P2_RED_N + i] = 0.f;
        }
        double* st = a.stats + ((size_t)b * a.N + i) * 2;
        atomicAdd(st, (double)t1);
        atomicAdd(st + 1, (double)t2);
    }
    named_bar_sync(4, P2_EPI_THREADS);
}

template <int EV, int EPI, int SMODE>
__device__ __forceinline__ void p2_epilogue(const cf_pw_args& a, const P2Params& p, float* Cs_all, float* red,
                                            uint64_t* tfull, uint64_t* tempty, uint32_t tmem, int warp, int lane) {
    constexpr int RPP = 128 / (32 / EV);
    const int ew = warp - P2_EPI_WARP0;
    const int grp = ew >> 2;
    const int qd = warp & 3;                                 // TMEM lane quadrant this warp may read
    const int gt = qd * 32 + lane;                           // row of the tile this thread brings from TMEM; its warp stores rows qd*32..+31
    const int et = ew * 32 + lane;
    float* Cs = Cs_all + grp * P2_CS_FLOATS;
    float* redw = red + (ew & 3) * 2 * P2_RED_N;                // this warp's row of the statistics table
    const int row_own = qd * 32 + lane;
    const int nslabs = (p.NTp + 31) >> 5;
    int cur_b = -1;
    uint32_t tcount = 0;
#ifdef CFNET_P2_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tt = P2_T0();
    const long long tstart = tt;
#endif
    P2Item it;
    for (p2_first(it, p); it.valid; p2_next_tile(it, p), ++tcount) {
        if (SMODE != CF_STATS_NONE && it.b != cur_b) {
            if (cur_b >= 0) p2_flush_stats(a, red, cur_b, et);
            cur_b = it.b;
        }
        const int acc = (int)(tcount & 1u);
        const uint32_t aph = (tcount >> 1) & 1u;
        const int rows_valid = min(TC_BM, p.R - it.r0);
        const int n0 = it.j * p.NT;
        const int nvalid = min(p.NT, a.N - n0);
        P2_ACC(0, tt);                                       // 0: tile bookkeeping (+ statistics flush)
        mbar_wait_b(&tfull[acc], aph);
        P2_ACC(1, tt);                                       // 1: wait for the accumulator
        tc_fence_after();
        const uint32_t tbase = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * p.acc_stride);
        for (int slab = grp; slab < nslabs; slab += 2) {
            float r32[32];
            tmem_ld32(tbase + (uint32_t)(slab * 32), r32);
            if (slab + 2 >= nslabs) {                        // last TMEM read of this warp for this tile: free the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            }
            P2_ACC(2, tt);                                   // 2: tcgen05.ld + release
            __syncwarp();                                    // this warp's rows of Cs: its previous slab has been read
            P2_ACC(3, tt);                                   // 3: group barriers
            float* dst = Cs + row_own * P2_CS_LD;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(r32[4 * i], r32[4 * i + 1], r32[4 * i + 2], r32[4 * i + 3]);
            P2_ACC(4, tt);                                   // 4: accumulator rows -> shared slab
            __syncwarp();
            P2_ACC(3, tt);
            p2_store_slab<EV, EPI, SMODE>(a, p, Cs, redw, it.b, it.r0, rows_valid, p.R, n0, slab * 32, nvalid, gt);
            P2_ACC(5, tt);                                   // 5: slab -> global (+ aux, activation, statistics partials)
        }
        if (grp >= nslabs) {                                 // a group without slabs still releases the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    if (SMODE != CF_STATS_NONE && cur_b >= 0) p2_flush_stats(a, red, cur_b, et);
#ifdef CFNET_P2_TIMING
    if (p.timing && blockIdx.x == 0 && et == 0) {
        for (int i = 0; i < 7; ++i) p2_dbg[8 + i] = tacc[i];
        p2_dbg[15] = clock64() - tstart;
    }
#endif
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(P2_THREADS, 1) pw_tc2_kernel(const cf_pw_args a, const float* __restrict__ pack, const P2Params p,
                                                               int av, int ev, const __grid_constant__ CUtensorMap tmx,
                                                               const __grid_constant__ CUtensorMap tmx2) {
    cf_pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[P2_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty[P2_MAX_STAGES];
    __shared__ __align__(8) uint64_t tfull[2];
    __shared__ __align__(8) uint64_t tempty[2];
    __shared__ __align__(8) uint64_t wres_bar;
    __shared__ __align__(8) uint64_t rfull[P2_MAX_RAW];
    __shared__ __align__(8) uint64_t rempty[P2_MAX_RAW];
    __shared__ uint32_t tmem_addr_s;
    __shared__ float red[4 * 2 * P2_RED_N];                          // [epilogue warp in group][sum, sum2][channel]

    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B tiles: 1024-B aligned
    uint8_t* stages = base;
    uint8_t* wres = stages + (size_t)p.nstages * p.stage_bytes;
    float* Cs = reinterpret_cast<float*>(wres + (p.resident ? (size_t)p.nchunks * p.b_chunk_bytes : 0));
    float* tab = Cs + 2 * P2_CS_FLOATS;
    uint8_t* raw = base + p.raw_off;                                                  // TMA landing ring (128-B aligned)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == P2_PROD_WARPS) {
        tmem_alloc(&tmem_addr_s, p.tmem_cols);
        tmem_relinquish();
    }
    if (tid == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            mbar_init(&full[s], P2_PROD_WARPS + (p.resident ? 0 : 1));   // + the streamed weight block's expect_tx arrival
            mbar_init(&empty[s], 1);
        }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], P2_EPI_WARPS); mbar_init(&tempty[1], P2_EPI_WARPS);
        mbar_init(&wres_bar, 1);
        for (int s = 0; s < p.nraw; ++s) {
            mbar_init(&rfull[s], 1);                                    // the loader's expect_tx arrival
            mbar_init(&rempty[s], P2_PROD_WARPS);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < 4 * 2 * P2_RED_N; i += P2_THREADS) red[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_addr_s;

    if (warp < P2_PROD_WARPS) {
        // ================= producers =================
#if P2_PROD_INC
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(P2_REGS_PROD));
#endif
#define P2_PARGS a, p, stages, tab, full, empty, pack, tid
        if (p.resident && tid == 0) {
            mbar_expect_tx(&wres_bar, (uint32_t)p.nchunks * p.b_chunk_bytes);
            for (int c = 0; c < p.nchunks; ++c)
                bulk_g2s(wres + (size_t)c * p.b_chunk_bytes, pack + (size_t)c * (p.b_chunk_bytes / 4), p.b_chunk_bytes, &wres_bar);
        }
#define P2_PROD(AV_)                                                                                              \
    switch (a.pro_mode) {                                                                                         \
        case CF_PRO_AFFINE: p2_producer<AV_, CF_PRO_AFFINE>(P2_PARGS); break;    \
        case CF_PRO_AFFINE_RELU: p2_producer<AV_, CF_PRO_AFFINE_RELU>(P2_PARGS); break;   \
        case CF_PRO_AFFINE_SWISH: p2_producer<AV_, CF_PRO_AFFINE_SWISH>(P2_PARGS); break; \
        case CF_PRO_AFFINE2: p2_producer<AV_, CF_PRO_AFFINE2>(P2_PARGS); break;  \
        default: p2_producer<AV_, CF_PRO_NONE>(P2_PARGS); break;                 \
    }
#define P2_TARGS a, p, stages, tab, full, empty, rfull, rempty, raw, pack, tid
#define P2_PROD_T(F_)                                                                                            \
    switch (a.pro_mode) {                                                                                        \
        case CF_PRO_AFFINE: p2_producer_tma<F_, CF_PRO_AFFINE>(P2_TARGS); break;                                 \
        case CF_PRO_AFFINE_RELU: p2_producer_tma<F_, CF_PRO_AFFINE_RELU>(P2_TARGS); break;                       \
        case CF_PRO_AFFINE_SWISH: p2_producer_tma<F_, CF_PRO_AFFINE_SWISH>(P2_TARGS); break;                     \
        case CF_PRO_AFFINE2: p2_producer_tma<F_, CF_PRO_AFFINE2>(P2_TARGS); break;                               \
        default: p2_producer_tma<F_, CF_PRO_NONE>(P2_TARGS); break;                                              \
    }
        if (p.tma) {
            if (p.fold == 1) { P2_PROD_T(1) } else if (p.fold == 2) { P2_PROD_T(2) } else { P2_PROD_T(4) }
        } else if (av == 4) { P2_PROD(4) } else { P2_PROD(2) }
#undef P2_PROD_T
#undef P2_TARGS
#undef P2_PROD
#undef P2_PARGS
    } else if (warp < P2_EPI_WARP0) {
        // ================= MMA issuer (first warp of its warpgroup; the other three only give up their registers) =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(P2_REGS_MMA));
        if (warp == P2_MMA_WARP) p2_mma_warp(a, p, stages, wres, full, empty, tfull, tempty, &wres_bar, tmem);
        else if (p.tma && warp == P2_MMA_WARP + 1 && lane == 0) p2_loader(a, p, raw, rfull, rempty, &tmx, &tmx2);
    } else {
        // ================= epilogue =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(P2_REGS_EPI));
#define P2_EARGS a, p, Cs, red, tfull, tempty, tmem, warp, lane
#define P2_EPI_S(EV_, EPI_)                                                                  \
    switch (a.stats_mode) {                                                                  \
        case CF_STATS_SUM_SQ: p2_epilogue<EV_, EPI_, CF_STATS_SUM_SQ>(P2_EARGS); break;      \
        case CF_STATS_SUM_AUX: p2_epilogue<EV_, EPI_, CF_STATS_SUM_AUX>(P2_EARGS); break;    \
        default: p2_epilogue<EV_, EPI_, CF_STATS_NONE>(P2_EARGS); break;                     \
    }
#define P2_EPI(EV_)                                                      \
    switch (a.epi_mode) {                                                \
        case CF_EPI_RELU: P2_EPI_S(EV_, CF_EPI_RELU) break;              \
        case CF_EPI_DRELU: P2_EPI_S(EV_, CF_EPI_DRELU) break;            \
        case CF_EPI_DSWISH: P2_EPI_S(EV_, CF_EPI_DSWISH) break;          \
        case CF_EPI_ADD_AUX: P2_EPI_S(EV_, CF_EPI_ADD_AUX) break;        \
        case CF_EPI_SIGMOID: P2_EPI_S(EV_, CF_EPI_SIGMOID) break;        \
        case CF_EPI_AFFINE: P2_EPI_S(EV_, CF_EPI_AFFINE) break;          \
        case CF_EPI_AFFINE_ADD_RELU: P2_EPI_S(EV_, CF_EPI_AFFINE_ADD_RELU) break; \
        default: P2_EPI_S(EV_, CF_EPI_NONE) break;                       \
    }
        if (ev == 4) { P2_EPI(4) } else { P2_EPI(2) }
#undef P2_EPI
#undef P2_EPI_S
#undef P2_EARGS
    }
    tc_fence_before();
    __syncthreads();
    if (warp == P2_PROD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem, p.tmem_cols);
    }
}


}  // namespace P2_NS
