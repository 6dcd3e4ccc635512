// Persistent, warp-specialised pointwise-conv GEMM on the tcgen05 tensor cores (sm_100a).
//
//   y[b,r,n] = epi( sum_k pro(x[b,r,k]) * w[n,k] (+ bias[n]) )        dense rows, forward + data gradient
//
// Reference call sites: conv1x1x1 (x3d_fine.py:100-105; used :149,166,356,370), nn.Linear fc2 (:380), the k=1
// Conv1d layers of the fusion block (x3d_coarse.py:216-219,232-246,335-336).  Same contract and the same
// 3xTF32 arithmetic as the one-tile-per-CTA kernel in x3d_pw_tc.cu (which measured 16-21 % of HBM peak: one
// CTA did load -> MMA -> epilogue strictly in sequence with 12 KB of loads in flight per SM).  Here:
//
//   * one CTA per SM walks the (row tile, channel tile) list; 16 warps in two roles that only meet at mbarriers:
//       warps 0-7   producers: global loads of the next 3 (1 with a second input) activation chunks are in
//                   flight in registers while the current one gets its BatchNorm/ReLU/Swish/BN-backward
//                   prologue, the hi/lo TF32 split and the SWIZZLE_128B store into a ring of A stages;
//                   the warp that completes a stage has one lane issue its tcgen05.mma's (3 per 8 k) into one of two
//                   TMEM accumulators and commit the stage-free / accumulator-full barriers;
//       warps 8-15  epilogue: two groups of four warps (one per TMEM lane quadrant) take alternate 32-column
//                   slabs: tcgen05.ld -> padded shared slab -> coalesced row-major stores with bias /
//                   activation derivative / residual add and the BatchNorm statistics;
//     so the loads of tile i+1, the MMAs of tile i and the stores of tile i-1 overlap;
//   * weights (pre-split, pre-swizzled by pw_tc_pack_kernel) stay resident in shared memory when they fit
//     next to >= 3 A stages (layers 1-2 and the small fusion layers), otherwise their k-chunk travels with
//     each A stage as one bulk-async copy (TMA unit) from L2;
//   * BatchNorm statistics are kept per CTA in shared memory across tiles and flushed (fp64 atomics) only
//     when the sample changes: 148 x B flushes per launch instead of one per tile.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include "tc_ptx.cuh"
#include <stdlib.h>

#define P2_PROD_WARPS 8
#define P2_EPI_WARPS 8
#define P2_PROD_THREADS (P2_PROD_WARPS * 32)
#define P2_EPI_THREADS (P2_EPI_WARPS * 32)
#define P2_THREADS ((P2_PROD_WARPS + P2_EPI_WARPS) * 32)
#define P2_A_STAGE (2 * TC_BM * TC_KC * 4) /* hi + lo: 32 KB */
#define P2_CS_LD 36
#define P2_CS_FLOATS (TC_BM * P2_CS_LD)
#define P2_MAX_STAGES 4
#define P2_RED_N 512
#define P2_NT_MAX 224
#define P2_SMEM_MAX (222 * 1024) /* dynamic; + ~4.2 KB static stays under the 227 KB per-CTA limit */

struct P2Params {
    int B, R, tps, ntiles, NT, NTp, nchunks, nstages, resident, acc_stride, KP, g_j, g_rt;
    uint32_t tmem_cols, b_chunk_bytes, stage_bytes;
    long long total_tiles;
};

// sigmoid from ex2.approx / rcp.approx (2^-22 relative: at the 3xTF32 level, far inside the 1e-3 parity bar); the
// IEEE expf + division cost 4x more issue slots and made the Swish producers instruction-bound
__device__ __forceinline__ float p2_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
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

// ---------------------------------------------------------------------------------------
// producers
// ---------------------------------------------------------------------------------------
template <int AV, bool X2>
__device__ __forceinline__ void p2_load_item(const cf_pw_args& a, const P2Params& p, const P2Item& it, int q, int rr,
                                             float (&v)[4][4], float (&v2)[X2 ? 4 : 1][4]) {
    const int K = a.K;
    const int k = it.c * TC_KC + q * 4;
    const int rows_valid = min(TC_BM, p.R - it.r0);
    const size_t base = ((size_t)it.b * p.R + it.r0) * K;
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
        const int row = pp * 32 + rr;
        const bool rv = row < rows_valid;
#pragma unroll
        for (int e = 0; e < 4; e += AV) {
            if (rv && k + e < K) {
                P2Vec<AV>::ld(a.x + base + (size_t)row * K + k + e, &v[pp][e]);
                if (X2) P2Vec<AV>::ld(a.x2 + base + (size_t)row * K + k + e, &v2[X2 ? pp : 0][e]);
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
__device__ __forceinline__ void p2_producer(const cf_pw_args& a, const P2Params& p, uint8_t* stages, uint8_t* wres, float* tab,
                                            uint64_t* full, uint64_t* empty, uint64_t* tfull, uint64_t* tempty, uint64_t* wres_bar,
                                            uint32_t* fill_cnt, uint32_t tmem, const float* __restrict__ pack, int tid) {
    constexpr bool X2 = PRO == CF_PRO_AFFINE2;
    constexpr int NSET = X2 ? 2 : 4;
    const int lane = tid & 31;
    const int q = tid & 7, rr = tid >> 3;
    float v[NSET][4][4];
    float v2[NSET][X2 ? 4 : 1][4];
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
    uint32_t ph = 0, tcount = 0;
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
            mbar_wait_b(&empty[s], ph ^ 1u);
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
            for (int pp = 0; pp < 4; ++pp) {
                const int row = pp * 32 + rr;
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
            fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                // The producer warp that completes a stage issues its MMAs (no dedicated MMA warp: 16 warps keep the
                // 128-register budget).  Stages complete in item order because every warp finishes item i -- including
                // the MMA issue when it was the last one in -- before it touches item i+1.
                __threadfence_block();
                const uint32_t old = atomicAdd(&fill_cnt[s], 1u);
                if (old == P2_PROD_WARPS - 1) {
                    __threadfence_block();
                    fill_cnt[s] = 0;                                     // nobody counts on this stage again before empty[s] fires
                    const int acc = (int)(tcount & 1u);
                    if (pr.c == 0) mbar_wait_b(&tempty[acc], ((tcount >> 1) & 1u) ^ 1u);   // the epilogue drained this accumulator
                    if (p.resident) mbar_wait_b(wres_bar, 0u);
                    else mbar_wait_b(&full[s], ph);                      // this chunk's weight block landed
                    tc_fence_after();
                    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7, 10),
                    // both K-major (bits 15,16 = 0), N >> 3 @ bit 17, M >> 4 @ bit 24
                    const uint32_t idesc =
                        (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NTp >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
                    const uint32_t d_tmem = tmem + (uint32_t)(acc * p.acc_stride);
                    const uint32_t a_hi_s = smem_u32(stage), a_lo_s = a_hi_s + TC_BM * TC_KC * 4;
                    const uint32_t b_hi_s = p.resident ? smem_u32(wres + (size_t)pr.c * p.b_chunk_bytes) : a_hi_s + P2_A_STAGE;
                    const uint32_t b_lo_s = b_hi_s + (uint32_t)p.NTp * 128u;
                    const int nk8 = min(4, (a.K - pr.c * TC_KC + 7) >> 3);
                    for (int k8 = 0; k8 < nk8; ++k8) {
                        const uint32_t ko = (uint32_t)k8 * 32u;          // 8 tf32 = 32 bytes along K inside the swizzle row
                        umma_tf32(d_tmem, make_desc_sw128(a_lo_s + ko), make_desc_sw128(b_hi_s + ko), idesc, (uint32_t)((pr.c | k8) != 0));
                        umma_tf32(d_tmem, make_desc_sw128(a_hi_s + ko), make_desc_sw128(b_lo_s + ko), idesc, 1u);
                        umma_tf32(d_tmem, make_desc_sw128(a_hi_s + ko), make_desc_sw128(b_hi_s + ko), idesc, 1u);
                    }
                    umma_commit(&empty[s]);                              // stage reusable once these MMAs retire
                    if (pr.c == p.nchunks - 1) umma_commit(&tfull[acc]); // accumulator complete
                }
            }
            __syncwarp();
            if (++s == p.nstages) { s = 0; ph ^= 1u; }
            if (pr.c == p.nchunks - 1) ++tcount;
            p2_advance(pr, p);
        }
    }
}

// ---------------------------------------------------------------------------------------
// epilogue: one 32-column slab, shared tile -> global (coalesced along N)
// ---------------------------------------------------------------------------------------
template <int EV, int EPI>
__device__ __forceinline__ void p2_store_slab(const cf_pw_args& a, const float* __restrict__ Cs, float* red, int b, int r0,
                                              int rows_valid, int R, int n0, int col0, int nvalid, int gt) {
    constexpr int CPR = 32 / EV;             // column groups per row
    constexpr int RPP = 128 / CPR;           // rows per pass
    constexpr int NPASS = TC_BM / RPP;
    const int N = a.N;
    const int cg = gt % CPR, rs = gt / CPR;
    const int nl = col0 + cg * EV;           // column within the channel tile
    if (nl >= nvalid) return;
    const int n = n0 + nl;
    const int smode = a.stats_mode;
    constexpr bool EPI_AUX = EPI == CF_EPI_DRELU || EPI == CF_EPI_DSWISH || EPI == CF_EPI_ADD_AUX;
    const bool need_aux = EPI_AUX || smode == CF_STATS_SUM_AUX;
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
    const size_t gbase = ((size_t)b * R + r0) * N + n;
    float ax[NPASS][EV];
    if (need_aux) {
#pragma unroll
        for (int i = 0; i < NPASS; ++i) {
            const int r = rs + i * RPP;
            if (r < rows_valid) P2Vec<EV>::ld(a.aux + gbase + (size_t)r * N, ax[i]);
            else {
#pragma unroll
                for (int e = 0; e < EV; ++e) ax[i][e] = 0.f;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NPASS; ++i) {
        const int r = rs + i * RPP;
        if (r >= rows_valid) break;
        float vv[EV];
        P2Vec<EV>::ldrw(Cs + r * P2_CS_LD + cg * EV, vv);
#pragma unroll
        for (int e = 0; e < EV; ++e) {
            float t = vv[e] + bi[e];
            const float axe = need_aux ? ax[i][e] : 0.f;
            if (EPI == CF_EPI_RELU) t = fmaxf(t, 0.f);
            else if (EPI == CF_EPI_DRELU) t = (fmaf(ea[e], axe, eb[e]) > 0.f) ? t : 0.f;
            else if (EPI == CF_EPI_DSWISH) t *= p2_dswish(fmaf(ea[e], axe, eb[e]));
            else if (EPI == CF_EPI_ADD_AUX) t += axe;
            else if (EPI == CF_EPI_SIGMOID) t = p2_sigmoid(t);
            vv[e] = t;
            s1[e] += t;
            s2[e] += (smode == CF_STATS_SUM_AUX) ? t * axe : t * t;
        }
        float* dst = a.y + gbase + (size_t)r * N;
        if (a.accumulate) {
            float old[EV];
            P2Vec<EV>::ldrw(dst, old);
#pragma unroll
            for (int e = 0; e < EV; ++e) vv[e] += old[e];
        }
        P2Vec<EV>::st(dst, vv);
    }
    if (smode != CF_STATS_NONE) {
#pragma unroll
        for (int e = 0; e < EV; ++e)
            if (nl + e < nvalid) {
                atomicAdd(red + n + e, s1[e]);
                atomicAdd(red + P2_RED_N + n + e, s2[e]);
            }
    }
}

__device__ __forceinline__ void p2_flush_stats(const cf_pw_args& a, float* red, int b, int et) {
    named_bar_sync(4, P2_EPI_THREADS);
    for (int i = et; i < a.N; i += P2_EPI_THREADS) {
        double* st = a.stats + ((size_t)b * a.N + i) * 2;
        atomicAdd(st, (double)red[i]);
        atomicAdd(st + 1, (double)red[P2_RED_N + i]);
        red[i] = 0.f;
        red[P2_RED_N + i] = 0.f;
    }
    named_bar_sync(4, P2_EPI_THREADS);
}

template <int EV, int EPI>
__device__ __forceinline__ void p2_epilogue(const cf_pw_args& a, const P2Params& p, float* Cs_all, float* red, uint64_t* tfull,
                                            uint64_t* tempty, uint32_t tmem, int warp, int lane) {
    const int ew = warp - P2_PROD_WARPS;
    const int grp = ew >> 2;
    const int qd = warp & 3;                                 // TMEM lane quadrant this warp may read
    const int gt = (ew & 3) * 32 + lane;                     // thread index within the group (phase 2)
    const int et = ew * 32 + lane;
    float* Cs = Cs_all + grp * P2_CS_FLOATS;
    const int row_own = qd * 32 + lane;
    const int nslabs = (p.NTp + 31) >> 5;
    const bool do_stats = a.stats_mode != CF_STATS_NONE;
    int cur_b = -1;
    uint32_t tcount = 0;
    P2Item it;
    for (p2_first(it, p); it.valid; p2_next_tile(it, p), ++tcount) {
        if (do_stats && it.b != cur_b) {
            if (cur_b >= 0) p2_flush_stats(a, red, cur_b, et);
            cur_b = it.b;
        }
        const int acc = (int)(tcount & 1u);
        const uint32_t aph = (tcount >> 1) & 1u;
        const int rows_valid = min(TC_BM, p.R - it.r0);
        const int n0 = it.j * p.NT;
        const int nvalid = min(p.NT, a.N - n0);
        mbar_wait_b(&tfull[acc], aph);
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
            named_bar_sync(2 + grp, 128);                    // the previous slab's readers are done with Cs
            float* dst = Cs + row_own * P2_CS_LD;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(r32[4 * i], r32[4 * i + 1], r32[4 * i + 2], r32[4 * i + 3]);
            named_bar_sync(2 + grp, 128);
            p2_store_slab<EV, EPI>(a, Cs, red, it.b, it.r0, rows_valid, p.R, n0, slab * 32, nvalid, gt);
        }
        if (grp >= nslabs) {                                 // a group without slabs still releases the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    if (do_stats && cur_b >= 0) p2_flush_stats(a, red, cur_b, et);
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(P2_THREADS, 1) pw_tc2_kernel(const cf_pw_args a, const float* __restrict__ pack, const P2Params p,
                                                               int av, int ev) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[P2_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty[P2_MAX_STAGES];
    __shared__ __align__(8) uint64_t tfull[2];
    __shared__ __align__(8) uint64_t tempty[2];
    __shared__ __align__(8) uint64_t wres_bar;
    __shared__ uint32_t tmem_addr_s;
    __shared__ uint32_t fill_cnt[P2_MAX_STAGES];
    __shared__ float red[2 * P2_RED_N];

    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B tiles: 1024-B aligned
    uint8_t* stages = base;
    uint8_t* wres = stages + (size_t)p.nstages * p.stage_bytes;
    float* Cs = reinterpret_cast<float*>(wres + (p.resident ? (size_t)p.nchunks * p.b_chunk_bytes : 0));
    float* tab = Cs + 2 * P2_CS_FLOATS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == P2_PROD_WARPS) {
        tmem_alloc(&tmem_addr_s, p.tmem_cols);
        tmem_relinquish();
    }
    if (tid == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            mbar_init(&full[s], 1);                          // streamed weight block (expect_tx arrival)
            fill_cnt[s] = 0;
            mbar_init(&empty[s], 1);
        }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], P2_EPI_WARPS); mbar_init(&tempty[1], P2_EPI_WARPS);
        mbar_init(&wres_bar, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 2 * P2_RED_N; i += P2_THREADS) red[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_addr_s;

    if (warp < P2_PROD_WARPS) {
        // ================= producers (+ MMA issue by the warp that completes a stage) =================
#define P2_PARGS a, p, stages, wres, tab, full, empty, tfull, tempty, &wres_bar, fill_cnt, tmem, pack, tid
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
        if (av == 4) { P2_PROD(4) } else if (av == 2) { P2_PROD(2) } else { P2_PROD(1) }
#undef P2_PROD
#undef P2_PARGS
    } else {
        // ================= epilogue =================
#define P2_EPI(EV_)                                                                                                    \
    switch (a.epi_mode) {                                                                                              \
        case CF_EPI_RELU: p2_epilogue<EV_, CF_EPI_RELU>(a, p, Cs, red, tfull, tempty, tmem, warp, lane); break;        \
        case CF_EPI_DRELU: p2_epilogue<EV_, CF_EPI_DRELU>(a, p, Cs, red, tfull, tempty, tmem, warp, lane); break;      \
        case CF_EPI_DSWISH: p2_epilogue<EV_, CF_EPI_DSWISH>(a, p, Cs, red, tfull, tempty, tmem, warp, lane); break;    \
        case CF_EPI_ADD_AUX: p2_epilogue<EV_, CF_EPI_ADD_AUX>(a, p, Cs, red, tfull, tempty, tmem, warp, lane); break;  \
        case CF_EPI_SIGMOID: p2_epilogue<EV_, CF_EPI_SIGMOID>(a, p, Cs, red, tfull, tempty, tmem, warp, lane); break;  \
        default: p2_epilogue<EV_, CF_EPI_NONE>(a, p, Cs, red, tfull, tempty, tmem, warp, lane); break;                 \
    }
        if (ev == 4) { P2_EPI(4) } else if (ev == 2) { P2_EPI(2) } else { P2_EPI(1) }
#undef P2_EPI
    }
    tc_fence_before();
    __syncthreads();
    if (warp == P2_PROD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem, p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
void cf_pw_tc_pack_launch(const float* w, long long w_sn, long long w_sk, float* pack, int K, int N, int NT, int NTp, int ntiles,
                          int nchunks, cudaStream_t stream);                             // x3d_pw_tc.cu
int cf_pw_conv_tc_v1(const cf_pw_args* a, cudaStream_t stream);                          // x3d_pw_tc.cu

static void p2_tiling(int K, int N, P2Params& p, int nt_max = P2_NT_MAX) {
    p.ntiles = (N + nt_max - 1) / nt_max;
    int nt = (N + p.ntiles - 1) / p.ntiles;
    p.NT = (nt + 7) / 8 * 8;                       // tile starts stay 32-byte aligned
    p.ntiles = (N + p.NT - 1) / p.NT;
    p.NTp = (p.NT + 15) / 16 * 16;                 // UMMA M=128 needs N % 16 == 0
    p.nchunks = (K + TC_KC - 1) / TC_KC;
    p.KP = p.nchunks * TC_KC;
    p.b_chunk_bytes = 2u * (uint32_t)p.NTp * 128u;
}

// bytes of the packed weight workspace; the layout is shared with the v1 kernel's packer but the tiling differs
extern "C" size_t cf_pw_tc_ws_bytes(int K, int N) {
    if (K <= 0 || N <= 0) return 0;
    P2Params p;
    size_t v2 = 0;
    static const int nt_try[] = {P2_NT_MAX, 160, 128, 96, 64, 32};        // every tiling the launcher may pick
    for (int ti = 0; ti < 6; ++ti) {
        p2_tiling(K, N, p, nt_try[ti]);
        size_t b = (size_t)p.ntiles * p.nchunks * p.b_chunk_bytes;
        if (b > v2) v2 = b;
    }
    // v1 tiling (<= 128 channels per tile), kept for the A/B switch CFNET_PW_TC_V1
    int nt1 = (N + 127) / 128;
    int NT1 = ((N + nt1 - 1) / nt1 + 7) / 8 * 8;
    nt1 = (N + NT1 - 1) / NT1;
    int NTp1 = (NT1 + 15) / 16 * 16;
    size_t v1 = (size_t)nt1 * p.nchunks * 2 * NTp1 * 128;
    return v2 > v1 ? v2 : v1;
}

static int p2_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static bool p2_use_v1() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CFNET_PW_TC_V1");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// called by cf_pw_conv (x3d_pw.cu) for dense problems when the caller supplied a weight-pack workspace
int cf_pw_conv_tc(const cf_pw_args* a, cudaStream_t stream) {
    if (p2_use_v1()) return cf_pw_conv_tc_v1(a, stream);
    const int K = a->K, N = a->N;
    P2Params p;
    p2_tiling(K, N, p);
    p.R = a->g.T * a->g.H * a->g.W;
    p.tps = cf_cdiv(p.R, TC_BM);
    p.total_tiles = (long long)a->B * p.tps * p.ntiles;
    p.B = a->B;
    CF_CHECK_ARG(p.total_tiles < (1LL << 31), "too many tiles");
    CF_CHECK_ARG(a->wpack_bytes >= (int64_t)cf_pw_tc_ws_bytes(K, N), "weight-pack workspace too small");
    CF_CHECK_ARG((((uintptr_t)a->wpack) & 127) == 0, "weight-pack workspace must be 128-byte aligned");
    if (a->stats_mode != CF_STATS_NONE && N > P2_RED_N) return cf_pw_conv_tc_v1(a, stream);   // per-CTA statistics table holds 512 channels
    // shared-memory plan: weights resident next to >= 3 A stages, else streamed with each stage; if even two stages
    // of the widest channel tile do not fit (very long K: big prologue tables), narrow the channel tile
    size_t smem = 0;
    static const int nt_try[] = {P2_NT_MAX, 160, 128, 96, 64, 32};
    bool planned = false;
    for (int ti = 0; ti < 6 && !planned; ++ti) {
        if (ti > 0) {
            p2_tiling(K, N, p, nt_try[ti]);
            p.total_tiles = (long long)a->B * p.tps * p.ntiles;
        }
        const size_t fixed = 1024 + 2 * (size_t)P2_CS_FLOATS * 4 + 3 * (size_t)p.KP * 4;
        if (fixed + 2 * (size_t)P2_A_STAGE >= P2_SMEM_MAX) break;
        const size_t avail = P2_SMEM_MAX - fixed;
        const size_t wres_bytes = (size_t)p.nchunks * p.b_chunk_bytes;
        p.resident = (p.ntiles == 1 && wres_bytes + 3 * (size_t)P2_A_STAGE <= avail) ? 1 : 0;
        if (p.resident) {
            p.stage_bytes = P2_A_STAGE;
            p.nstages = (int)((avail - wres_bytes) / P2_A_STAGE);
            if (p.nstages > P2_MAX_STAGES) p.nstages = P2_MAX_STAGES;
            smem = fixed + wres_bytes + (size_t)p.nstages * p.stage_bytes;
            planned = true;
        } else {
            p.stage_bytes = P2_A_STAGE + p.b_chunk_bytes;
            p.nstages = (int)(avail / p.stage_bytes);
            if (p.nstages > P2_MAX_STAGES) p.nstages = P2_MAX_STAGES;
            if (p.nstages >= 2) {
                smem = fixed + (size_t)p.nstages * p.stage_bytes;
                planned = true;
            }
        }
    }
    CF_CHECK_ARG(planned, "K too large for the tensor-core path");
    CF_CHECK_ARG(a->wpack_bytes >= (int64_t)((size_t)p.ntiles * p.nchunks * p.b_chunk_bytes), "weight-pack workspace too small");
    CF_CHECK_ARG(p.total_tiles < (1LL << 31), "too many tiles");
    p.acc_stride = (p.NTp + 31) / 32 * 32;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < 2 * p.acc_stride) p.tmem_cols <<= 1;
    cf_pw_tc_pack_launch(a->w, a->w_sn, a->w_sk, a->wpack, K, N, p.NT, p.NTp, p.ntiles, p.nchunks, stream);
    uintptr_t xa = (uintptr_t)a->x | (uintptr_t)(a->x2 ? a->x2 : a->x);
    int av = ((K & 3) == 0 && (xa & 15) == 0) ? 4 : (((K & 1) == 0 && (xa & 7) == 0) ? 2 : 1);
    uintptr_t ya = (uintptr_t)a->y | (uintptr_t)(a->aux ? a->aux : a->y);
    int ev = ((N & 3) == 0 && (ya & 15) == 0) ? 4 : (((N & 1) == 0 && (ya & 7) == 0) ? 2 : 1);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(pw_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_MAX);
        if (e != cudaSuccess) {
            cf_set_error("cf_pw_conv_tc: cannot opt in to %d B of shared memory: %s", P2_SMEM_MAX, cudaGetErrorString(e));
            return CF_ERR_CUDA;
        }
        attr_done = true;
    }
    long long grid = p.total_tiles < p2_sm_count() ? p.total_tiles : p2_sm_count();
    p.g_j = (int)(grid % p.ntiles);
    p.g_rt = (int)(grid / p.ntiles);
    pw_tc2_kernel<<<(unsigned)grid, P2_THREADS, smem, stream>>>(*a, a->wpack, p, av, ev);
    CF_COUNT_LAUNCH(2);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
