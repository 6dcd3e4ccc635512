// Persistent, warp-specialised pointwise-conv GEMM on the tcgen05 tensor cores (sm_100a).
//
//   y[b,r,n] = epi( sum_k pro(x[b,r,k]) * w[n,k] (+ bias[n]) )        dense rows, forward + data gradient
//
// Reference call sites: conv1x1x1 (x3d_fine.py:100-105; used :149,166,356,370), nn.Linear fc2 (:380), the k=1
// Conv1d layers of the fusion block (x3d_coarse.py:216-219,232-246,335-336).  Same contract and the same
// 3xTF32 arithmetic as the one-tile-per-CTA kernel in x3d_pw_tc.cu (which measured 16-21 % of HBM peak: one
// CTA did load -> MMA -> epilogue strictly in sequence with 12 KB of loads in flight per SM).  Here:
//
//   * one CTA per SM walks the (row tile, channel tile) list; 28 warps in three roles that only meet at mbarriers (register budgets
//     rebalanced per role with setmaxnreg: 72 / 24 / 96):
//       warps 0-15  producers: global loads of the next 3 (1 with a second input) activation chunks are in
//                   flight in registers while the current one gets its BatchNorm/ReLU/Swish/BN-backward
//                   prologue, the hi/lo TF32 split and the SWIZZLE_128B store into a ring of A stages;
//       warp  16    MMA issuer: the warp walks the stages convergently, one elected lane issues the tcgen05.mma's
//                   (3 per 8 k) into one of two TMEM accumulators and commits stage-free / accumulator-full barriers;
//       warps 20-27 epilogue: two groups of four warps (one per TMEM lane quadrant) take alternate 32-column
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
#include "tma_host.cuh"
#include <stdlib.h>
#include <string.h>

#define P2_A_STAGE (2 * TC_BM * TC_KC * 4) /* hi + lo: 32 KB */
#define P2_CS_LD 36
#define P2_CS_FLOATS (TC_BM * P2_CS_LD)
#define P2_MAX_STAGES 4
#define P2_MAX_RAW 8
#define P2_RED_N 512
#define P2_NT_MAX 224
#define P2_SMEM_MAX (210 * 1024) /* dynamic; + ~16.5 KB static (statistics table) stays under the 227 KB per-CTA limit */

// cycle counters of CTA 0 (CFNET_PW_TC_TIMING=1; read back with cf_pw_tc_debug_read): where each role's time goes
__device__ long long p2_dbg[32];
// (compiled in only with -DCFNET_P2_TIMING: the 8 counters cost 16 registers per thread, which the 72-register producers
// do not have)
#ifdef CFNET_P2_TIMING
#define P2_T0() (p.timing ? clock64() : 0)
#define P2_ACC(slot, t0) do { if (p.timing) { long long t1__ = clock64(); tacc[slot] += t1__ - (t0); (t0) = t1__; } } while (0)
#define P2_TIMING_ON 1
#else
#define P2_T0() 0
#define P2_ACC(slot, t0) do { } while (0)
#define P2_TIMING_ON 0
#endif

struct P2Params {
    int B, R, tps, ntiles, NT, NTp, nchunks, nstages, resident, acc_stride, KP, g_j, g_rt, timing, dbg_1x;
    // strided 1x1x1 convs (the downsample branch, x3d_fine.py:277-288): gmode 1 = the A rows are gathered from a
    // [Ti,Hi,Wi] volume at (t*st, h*sh, w*sw); gmode 2 = the output rows are scattered to it and accumulated (y +=)
    int gmode, gH, gW, gHi, gWi, gst, gsh, gsw;
    long long g_sample_stride;
    uint32_t tmem_cols, b_chunk_bytes, stage_bytes;
    long long total_tiles;
    // TMA-fed producers: raw activation tiles land in a ring of nraw stages (raw_stage_bytes each = raw_in_bytes per input
    // tensor) at raw_off from the 1024-aligned base; fold = rows presented as one TMA row (tma_host.cuh)
    int tma, fold, nraw;
    uint32_t raw_in_bytes, raw_stage_bytes, raw_off;
};

// ---- role split: 8 producer warps (the 16-warp split of round 1 lost every same-box A/B and is gone)
#define P2_NS p2w8
#define P2_PROD_WARPS 8
#define P2_REGS_PROD 120
#define P2_REGS_MMA 32
#define P2_REGS_EPI 104
#define P2_PROD_INC 1                  /* producers raise their register count (launch: 640 threads x 96) */
#include "x3d_pw_tc2_roles.cuh"
#include "x3d_pw_tc2_kernel.cuh"
#include "x3d_pw_tc2_unroles.cuh"

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


// ---- packing many weights in one launch (the channel tile of the default tiling; cf_pw_conv_tc re-packs if it narrows it)
extern "C" int cf_pw_pack_nt(int K, int N) {
    if (K <= 0 || N <= 0) return 0;
    P2Params p;
    p2_tiling(K, N, p);
    return p.NT;
}
extern "C" size_t cf_sizeof_pack_item(void) { return sizeof(cf_pack_item); }

__global__ void __launch_bounds__(256) pw_pack_many_kernel(const cf_pack_item* __restrict__ items) {
    cf_pdl_enter();
    const cf_pack_item it = items[blockIdx.x];
    const int NT = it.nt, NTp = (NT + 15) / 16 * 16;
    const int ntiles = (it.N + NT - 1) / NT, nchunks = (it.K + TC_KC - 1) / TC_KC;
    const long long total = (long long)ntiles * nchunks * NTp * 8;
    for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < total; i += (long long)gridDim.y * 256) {
        const int q = (int)(i & 7);
        long long t = i >> 3;
        const int nl = (int)(t % NTp); t /= NTp;
        const int c = (int)(t % nchunks);
        const int j = (int)(t / nchunks);
        const int n = j * NT + nl;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = c * TC_KC + q * 4 + e;
            const float v = (nl < NT && n < it.N && k < it.K) ? __ldg(it.w + (long long)n * it.w_sn + (long long)k * it.w_sk) : 0.f;
            tf32_split(v, hi[e], lo[e]);
        }
        char* blk = (char*)it.pack + ((long long)(j * nchunks + c) * 2 * NTp * 128);
        const uint32_t off = sw128_off(nl, q);
        *reinterpret_cast<float4*>(blk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(blk + (size_t)NTp * 128 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

extern "C" int cf_pw_pack_many(const cf_pack_item* items, int n, cudaStream_t stream) {
    CF_CHECK_ARG(items && n > 0 && n <= 65535, "bad argument");
    cf_launch(pw_pack_many_kernel, dim3((unsigned)n, 8), 256, 0, stream, items);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

// debug: cycle counters of CTA 0 of the last persistent launch (built with -DCFNET_P2_TIMING -DCFNET_AB, CFNET_PW_TC_TIMING=1)
extern "C" int cf_pw_tc_debug_read(long long* out16) {
    long long zero[32] = {0};
    if (cudaMemcpyFromSymbol(out16, p2_dbg, 24 * sizeof(long long)) != cudaSuccess) return CF_ERR_CUDA;
    cudaMemcpyToSymbol(p2_dbg, zero, sizeof(zero));
    return CF_OK;
}

// called by cf_pw_conv (x3d_pw.cu) for dense problems when the caller supplied a weight-pack workspace
static int p2_run(const cf_pw_args* a, cudaStream_t stream, int* plan_nt);

int cf_pw_conv_tc(const cf_pw_args* a, cudaStream_t stream) { return p2_run(a, stream, nullptr); }

// The channel tile the persistent kernel WOULD use for this call (its shared-memory plan may narrow the default tile):
// what a persistent pack of the weights has to be made with; 0 when the call does not take the persistent kernel.
extern "C" int cf_pw_plan_nt(const cf_pw_args* a) {
    if (!a || !a->wpack) return 0;
    int nt = 0;
    const int rc = p2_run(a, nullptr, &nt);
    return rc == CF_OK ? nt : 0;
}

// plan_nt != NULL: plan only (no launch, no packing), *plan_nt = the channel tile
static int p2_run(const cf_pw_args* a, cudaStream_t stream, int* plan_nt) {
    if (plan_nt) *plan_nt = 0;
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
    // rare shapes stay on the one-tile-per-CTA kernel: odd channel counts (the 157-class head), in-place accumulation
    // (only the scattered CUDA-core path of the strided convs uses it), statistics over more than 512 channels
    uintptr_t xa = (uintptr_t)a->x | (uintptr_t)(a->x2 ? a->x2 : a->x);
    int av = ((K & 3) == 0 && (xa & 15) == 0) ? 4 : (((K & 1) == 0 && (xa & 7) == 0) ? 2 : 1);
    uintptr_t ya = (uintptr_t)a->y | (uintptr_t)(a->aux ? a->aux : a->y);
    int ev = ((N & 3) == 0 && (ya & 15) == 0) ? 4 : (((N & 1) == 0 && (ya & 7) == 0) ? 2 : 1);
    const int taps1 = a->g.kt == 1 && a->g.kh == 1 && a->g.kw == 1 && a->g.pt == 0 && a->g.ph == 0 && a->g.pw == 0 && a->g.ch_stride == 1;
    p.gmode = a->gather_in ? 1 : (a->scatter_out ? 2 : 0);
    if (p.gmode) {
        CF_CHECK_ARG(taps1 && a->g.pos_stride == (p.gmode == 1 ? K : N), "tensor-core path: strided 1x1x1 channels-last only");
        CF_CHECK_ARG(p.gmode == 1 ? a->pro_mode != CF_PRO_AFFINE2 : (a->accumulate && a->stats_mode == CF_STATS_NONE && !a->aux),
                     "tensor-core path: unsupported strided combination");
        if (((a->g.sample_stride * 4) & 15) != 0) return -1;
    }
    p.gH = a->g.H; p.gW = a->g.W; p.gHi = a->g.Hi; p.gWi = a->g.Wi; p.gst = a->g.st; p.gsh = a->g.sh; p.gsw = a->g.sw;
    p.g_sample_stride = a->g.sample_stride;
    if (av == 1 || ev == 1 || (a->accumulate && p.gmode != 2) || (a->stats_mode != CF_STATS_NONE && N > P2_RED_N))
        return (p.gmode || plan_nt) ? -1 : cf_pw_conv_tc_v1(a, stream);      // (-1: the caller falls back to the CUDA-core gather kernel)

    // ---- TMA-fed producers: dense rows (gathered rows keep the register-load producers), describable by a tensor map
    const bool x2 = a->pro_mode == CF_PRO_AFFINE2;
    CUtensorMap tmx, tmx2;
    memset(&tmx, 0, sizeof(tmx));
    memset(&tmx2, 0, sizeof(tmx2));
    p.tma = 0; p.fold = 1; p.nraw = 0; p.raw_in_bytes = p.raw_stage_bytes = p.raw_off = 0;
    if (p.gmode == 0 && cf_env("CFNET_P2_TMA", 1)) {
        const int fold = (K % 4 == 0) ? 1 : ((K % 2 == 0) ? 2 : 4);
        if ((fold == 1 || K <= 64) && cf_make_row_tmap(&tmx, a->x, a->B, p.R, K, fold) &&
            (!x2 || cf_make_row_tmap(&tmx2, a->x2, a->B, p.R, K, fold))) {
            p.tma = 1;
            p.fold = fold;
            p.raw_in_bytes = fold == 1 ? (uint32_t)(TC_BM * TC_KC * 4) : (uint32_t)(((size_t)TC_BM * K * 4 + 127) / 128 * 128);
            p.raw_stage_bytes = p.raw_in_bytes * (x2 ? 2u : 1u);
        }
    }

    // shared-memory plan.  Register-load producers (round 1): weights resident next to >= 3 A stages when they fit, else
    // streamed with each stage (>= 2 stages); if even that does not fit (very long K), narrow the channel tile.
    // TMA-fed producers: the operand stages only decouple producers from the MMA issuer (2-3 are enough) and the rest of the
    // budget is raw stages (= bytes in flight); taken when the weights stay resident and >= 3 raw stages (2 for one input
    // tensor) fit -- a second input (BatchNorm-backward prologue) doubles the raw stage and usually does not.
    // Keep some L1 when the epilogue moves 8-byte vectors (N % 4 != 0): the unified L1/shared array is carved in steps
    // (... 164, 196, 228 KB); above 196 KB per CTA (dynamic + 16.5 KB static + 1 KB reserved) no L1 is left and those
    // loads / stores (two per 32-byte sector) go to L2 twice: measured -18 % .. -43 % on the layer-1 shapes.
    const size_t l1_friendly = 196 * 1024 - 1024 - 17 * 1024;
    size_t smem = 0;
    static const int nt_try[] = {P2_NT_MAX, 160, 128, 96, 64, 32};
    bool planned = false;
    for (int ti = 0; ti < 6 && !planned; ++ti) {
        if (ti > 0) {
            p2_tiling(K, N, p, nt_try[ti]);
            p.total_tiles = (long long)a->B * p.tps * p.ntiles;
        }
        const size_t fixed = 1024 + 2 * (size_t)P2_CS_FLOATS * 4 + 3 * (size_t)p.KP * 4 + 128;
        if (fixed + 2 * (size_t)P2_A_STAGE >= P2_SMEM_MAX) break;
        const size_t avail = P2_SMEM_MAX - fixed;
        const size_t wres_bytes = (size_t)p.nchunks * p.b_chunk_bytes;
        p.resident = (p.ntiles == 1 && wres_bytes + 3 * (size_t)P2_A_STAGE <= avail) ? 1 : 0;
        if (p.tma && p.resident) {
            const size_t cap = ev == 4 ? (size_t)P2_SMEM_MAX : l1_friendly;
            const int raw_need = x2 ? 3 : 2;
            size_t used = fixed + wres_bytes + 2 * (size_t)P2_A_STAGE;
            if (used + raw_need * (size_t)p.raw_stage_bytes <= cap) {
                p.stage_bytes = P2_A_STAGE;
                p.nstages = 2;
                if (used + P2_A_STAGE + 4 * (size_t)p.raw_stage_bytes <= cap) {      // room for a third operand stage
                    p.nstages = 3;
                    used += P2_A_STAGE;
                }
                p.nraw = (int)((cap - used) / p.raw_stage_bytes);
                if (p.nraw > P2_MAX_RAW) p.nraw = P2_MAX_RAW;
                p.raw_off = (uint32_t)((used - 1024 - 128 + 127) / 128 * 128);   // content ends 128 B (the slack in `fixed`) before `used`
                smem = 1024 + p.raw_off + (size_t)p.nraw * p.raw_stage_bytes;
                planned = true;
                break;
            }
        }
        p.tma = 0;
        p.nraw = 0;
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
        if (planned) {
            const int min_stages = p.resident ? 3 : 2;
            while (smem > l1_friendly && p.nstages > min_stages) {
                --p.nstages;
                smem -= p.stage_bytes;
            }
        }
    }
    CF_CHECK_ARG(planned, "K too large for the tensor-core path");
    if (plan_nt) {
        *plan_nt = p.NT;
        return CF_OK;
    }
    CF_CHECK_ARG(a->wpack_bytes >= (int64_t)((size_t)p.ntiles * p.nchunks * p.b_chunk_bytes), "weight-pack workspace too small");
    CF_CHECK_ARG(p.total_tiles < (1LL << 31), "too many tiles");
    p.acc_stride = (p.NTp + 31) / 32 * 32;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < 2 * p.acc_stride) p.tmem_cols <<= 1;
    const bool prepacked = a->wpack_nt > 0 && a->wpack_nt == p.NT;   // packed once per optimizer step by cf_pw_pack_many
    if (!prepacked) cf_pw_tc_pack_launch(a->w, a->w_sn, a->w_sk, a->wpack, K, N, p.NT, p.NTp, p.ntiles, p.nchunks, stream);
    static CfOncePerDevice attr_done;
    if (attr_done.need()) {
        cudaError_t e = cudaFuncSetAttribute(p2w8::pw_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_MAX);
        if (e != cudaSuccess) {
            cf_set_error("cf_pw_conv_tc: cannot opt in to %d B of shared memory: %s", P2_SMEM_MAX, cudaGetErrorString(e));
            return CF_ERR_CUDA;
        }
        attr_done.mark();
    }
    long long grid = p.total_tiles < p2_sm_count() ? p.total_tiles : p2_sm_count();
    p.timing = cf_env("CFNET_PW_TC_TIMING", 0);
    p.dbg_1x = cf_env("CFNET_PW_TC_1X", 0);
    p.g_j = (int)(grid % p.ntiles);
    p.g_rt = (int)(grid / p.ntiles);
    cf_launch(p2w8::pw_tc2_kernel, (unsigned)grid, (8 + 4 + 8) * 32, smem, stream, *a, a->wpack, p, av, ev, tmx, tmx2);
    CF_COUNT_LAUNCH(prepacked ? 1 : 2);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
