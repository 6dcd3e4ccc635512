// Clip input pipeline: decoded RGB frames (uint8, HWC) -> normalised fp32 clip [3,T,S,S] in one kernel
// (SURVEY 8(f) next-4).  Replaces, per frame, the reference's CPU chain
//   charades_fine.py:170-172                    [spatial_transform(img) for img in imgs]; stack; permute -> [3,T,S,S]
//   transforms/spatial_transforms.py:488-503    MultiScaleRandomCropMultigrid  (crop box, img.resize(BILINEAR))
//   transforms/spatial_transforms.py:216-230    CenterCropScaled               (centre box, img.resize(BILINEAR))
//   transforms/spatial_transforms.py:342-354    RandomHorizontalFlip
//   transforms/spatial_transforms.py:46-87      ToTensor(255)
//   transforms/spatial_transforms.py:108-118    Normalize(mean, std)
//   charades_fine.py:215-226                    mt_collate_fn's zero padding of the shorter clips
//
// img.resize is Pillow's ImagingResample (Resample.c): separable triangle filter, 22-bit fixed-point coefficients,
// horizontal pass then vertical pass, each rounded and clipped to uint8.  The integer arithmetic below is the same,
// so the uint8 image and therefore the fp32 clip are bit-identical to the reference's.
//
// Byte/integer work bound by HBM: per frame 3*crop^2 bytes in, 12*S^2 bytes out.  CTA = (band of BAND output rows,
// frame).  Horizontal pass: the input rows the band needs are resampled straight from global memory (adjacent output
// columns read adjacent bytes -> sector-coalesced through L1) into a planar uint8 tile in shared memory; vertical pass:
// a thread owns 4 adjacent output columns, reads uchar4 per tap from the tile, maps the three uint8 results through a
// 768-entry shared-memory table (the only place the fp32 divisions of ToTensor/Normalize are evaluated: IEEE-exact
// __fdiv_rn once per table entry instead of 6 divisions per pixel, which would bound the kernel by the FP32 pipe) and
// streams three float4 out (st.global.cs, reversed when flipped).
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include <math.h>

#define CLIP_PRECISION_BITS 22
#define CLIP_BAND 16

// ------------------------------------------------------------------------------------------------------------
// Host: Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0), whole axis.
// Plain double arithmetic on the host, statement by statement as Resample.c evaluates it.
// ------------------------------------------------------------------------------------------------------------
extern "C" int cf_resample_ksize(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return 0;
    double scale = (double)in_size / (double)out_size;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    double support = 1.0 * filterscale;
    return (int)ceil(support) * 2 + 1;
}

extern "C" int cf_resample_coeffs(int in_size, int out_size, int* bounds, int* kk) {
    CF_CHECK_ARG(in_size > 0 && out_size > 0 && bounds && kk, "bad arguments");
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    const double ss = 1.0 / filterscale;
    double* w = new double[ksize];
    for (int xx = 0; xx < out_size; ++xx) {
        volatile double center = (xx + 0.5) * scale;      // volatile: no contraction / excess precision across statements
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        volatile double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            volatile double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            w[x] = a < 1.0 ? 1.0 - a : 0.0;
            ww = ww + w[x];
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) w[x] = w[x] / ww;
        int* k = kk + (size_t)xx * ksize;
        for (int x = 0; x < ksize; ++x) {
            if (x >= xmax) { k[x] = 0; continue; }
            volatile double v = w[x] * (double)(1 << CLIP_PRECISION_BITS);
            k[x] = w[x] < 0.0 ? (int)(-0.5 + v) : (int)(0.5 + v);
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    delete[] w;
    return CF_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Device
// ------------------------------------------------------------------------------------------------------------
__global__ void normalize_lut_kernel(float* __restrict__ lut, float m0, float m1, float m2, float s0, float s1, float s2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 768) return;
    const int c = i >> 8, v = i & 255;
    const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), s = c == 0 ? s0 : (c == 1 ? s1 : s2);
    lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), m), s);     // img.float().div(255); t.sub_(m).div_(s)
}

// Bilinear (triangle) weights are >= 0 and sum to 2^22 +- ksize, so acc >> 22 is already in [0,255]:
// acc <= 2^21 + 255 * (2^22 + ksize) < 256 * 2^22.  Pillow's clip8() is the identity here.
__device__ __forceinline__ int round8(int acc) { return acc >> CLIP_PRECISION_BITS; }

struct ClipArgs {
    const uint8_t* frames;     // [T,H,W,3]
    float* out;                // sample base; channel stride out_stride_c, frame stride S*S
    const int* bounds_h;       // [S,2]
    const int* kk_h;           // [S,ksize_h]
    const int* bounds_v;
    const int* kk_v;
    const float* lut;          // [3,256]
    int T, Tout, H, W, x1, y1, S, Sp, ksize_h, ksize_v, flip, rows_max;
    long long out_stride_c;
};

// A CTA keeps its band and walks over the frames t = blockIdx.y, blockIdx.y + gridDim.y, ...: tables, bounds and the
// per-thread set-up are amortised over them, and the grid is sized to ONE resident wave (no tail wave).

// byte i of a 32-bit word as an int (one PRMT)
__device__ __forceinline__ int byte_of(unsigned u, int i) { return (int)__byte_perm(u, 0u, 0x4440u + (unsigned)i); }
// table[(acc >> 22)] with the index scaled in one shift + mask
__device__ __forceinline__ float lut_at(const float* l, int acc) {
    return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(l) + ((acc >> (CLIP_PRECISION_BITS - 2)) & 0x3fc));
}

// KS3: ksize == 3 (every up-scale and the identity): horizontal tap weights live in registers.
// SPLIT: the CTA is two thread groups of blockDim/2 -- group 0 runs the horizontal pass of frame k+1 into one half of a double
// buffered tile while group 1 runs the vertical pass + stores of frame k from the other half (one barrier per frame), so the
// global-load latency of one pass hides behind the arithmetic and stores of the other.  !SPLIT: one group does both in turn.
template <bool KS3, bool FLIP, bool SPLIT>
__global__ void __launch_bounds__(1024) clip_preprocess_kernel(const ClipArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut_s = reinterpret_cast<float*>(smem_raw);                 // 768 floats
    int* vb = reinterpret_cast<int*>(smem_raw + 768 * sizeof(float));  // [CLIP_BAND][2]  (first row relative to r0, taps)
    int* vk = vb + 2 * CLIP_BAND;                                      // [CLIP_BAND][ksize_v]
    unsigned char* hs = reinterpret_cast<unsigned char*>(vk + CLIP_BAND * a.ksize_v);   // [rows_max][3][Sp], 16-byte aligned by the host
    const int nthr = SPLIT ? blockDim.x >> 1 : blockDim.x;              // threads per role
    const int role = SPLIT ? (threadIdx.x >= nthr ? 1 : 0) : 0;
    const int tid = threadIdx.x - role * nthr, S = a.S, Sp = a.Sp;
    const int oy0 = blockIdx.x * CLIP_BAND;
    const int nrow_out = min(CLIP_BAND, S - oy0);
    const int quads = S >> 2;

    // input rows (relative to the crop) this band needs: bounds are monotone in oy
    const int r0 = __ldg(a.bounds_v + 2 * oy0);
    const int last = oy0 + nrow_out - 1;
    const int nrows = __ldg(a.bounds_v + 2 * last) + __ldg(a.bounds_v + 2 * last + 1) - r0;
    for (int i = tid; i < 768; i += nthr) lut_s[i] = __ldg(a.lut + i);
    for (int i = tid; i < nrow_out; i += nthr) {
        vb[2 * i] = __ldg(a.bounds_v + 2 * (oy0 + i)) - r0;
        vb[2 * i + 1] = __ldg(a.bounds_v + 2 * (oy0 + i) + 1);
    }
    for (int i = tid; i < nrow_out * a.ksize_v; i += nthr) vk[i] = __ldg(a.kk_v + oy0 * a.ksize_v + i);

    // vertical-pass role of this thread: 4 adjacent columns (quad q) of every rgs-th row of the band
    const int rgs = max(nthr / quads, 1);
    const int rg = tid / quads;
    const int q_first = rg < rgs ? tid - rg * quads : quads;            // threads beyond rgs*quads idle in the vertical pass
    const int q_step = nthr < quads ? nthr : quads;
    const size_t row_bytes = (size_t)a.W * 3;

    const int hs_bytes = a.rows_max * 3 * Sp;
    const int nf = a.Tout > (int)blockIdx.y ? (a.Tout - 1 - (int)blockIdx.y) / (int)gridDim.y + 1 : 0;   // frames of this CTA
    for (int it = 0; it < nf + (SPLIT ? 1 : 0); ++it) {                // uniform over the CTA
        // frame of the horizontal pass (th) and of the vertical pass (tv) in this iteration
        const int th = blockIdx.y + it * gridDim.y;
        const int tv = SPLIT ? th - (int)gridDim.y : th;
        unsigned char* hs_h = SPLIT ? hs + (it & 1) * hs_bytes : hs;
        const unsigned char* hs_v = SPLIT ? hs + ((it + 1) & 1) * hs_bytes : hs;
        const bool do_h = (!SPLIT || role == 0) && it < nf && th < a.T;
        const bool do_v = (!SPLIT || role == 1) && (!SPLIT || it > 0);
        const bool live = tv < a.T;
        float* outf = a.out + (size_t)(tv > 0 ? tv : 0) * S * S;
        const int t = th;
        if (do_h) {
            // ---- horizontal pass: a thread owns output column ox and walks down the crop rows [r0, r0+nrows) -> hs[r][c][ox] ----
            const uint8_t* fr = a.frames + ((size_t)t * a.H + (a.y1 + r0)) * row_bytes + (size_t)a.x1 * 3;
            for (int ox = tid; ox < S; ox += nthr) {
                const int xmin = __ldg(a.bounds_h + 2 * ox), n = __ldg(a.bounds_h + 2 * ox + 1);
                const uint8_t* src = fr + xmin * 3;
                unsigned char* d = hs_h + ox;
                const int half = 1 << (CLIP_PRECISION_BITS - 1);
                if (KS3 && n == 2) {
                    // support 1: the window holds exactly 2 pixels away from the borders -> 6 adjacent bytes, immediate offsets,
                    // one global and three shared pointers advanced per row
                    const int k0 = __ldg(a.kk_h + ox * 3), k1 = __ldg(a.kk_h + ox * 3 + 1);
                    const uint8_t* p = src;
                    unsigned char *d0 = d, *d1 = d + Sp, *d2 = d + 2 * Sp;
                    const int dstep = 3 * Sp;
#pragma unroll 4
                    for (int r = 0; r < nrows; ++r) {
                        const int s0 = half + (int)__ldg(p) * k0 + (int)__ldg(p + 3) * k1;
                        const int s1 = half + (int)__ldg(p + 1) * k0 + (int)__ldg(p + 4) * k1;
                        const int s2 = half + (int)__ldg(p + 2) * k0 + (int)__ldg(p + 5) * k1;
                        *d0 = (unsigned char)round8(s0);
                        *d1 = (unsigned char)round8(s1);
                        *d2 = (unsigned char)round8(s2);
                        p += row_bytes; d0 += dstep; d1 += dstep; d2 += dstep;
                    }
                } else {
                    const int* k = a.kk_h + ox * a.ksize_h;
                    const uint8_t* p = src;
                    unsigned char *d0 = d, *d1 = d + Sp, *d2 = d + 2 * Sp;
                    const int dstep = 3 * Sp;
                    if (n <= 4) {                                       // down-scales below 1.5x and border columns: weights in registers
                        int kr[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) kr[j] = j < n ? __ldg(k + j) : 0;
#pragma unroll 2
                        for (int r = 0; r < nrows; ++r) {
                            int s0 = half, s1 = half, s2 = half;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (j < n) {
                                    s0 += (int)__ldg(p + 3 * j) * kr[j];
                                    s1 += (int)__ldg(p + 3 * j + 1) * kr[j];
                                    s2 += (int)__ldg(p + 3 * j + 2) * kr[j];
                                }
                            }
                            *d0 = (unsigned char)round8(s0);
                            *d1 = (unsigned char)round8(s1);
                            *d2 = (unsigned char)round8(s2);
                            p += row_bytes; d0 += dstep; d1 += dstep; d2 += dstep;
                        }
                    } else {
                        for (int r = 0; r < nrows; ++r) {
                            int s0 = half, s1 = half, s2 = half;
                            for (int j = 0; j < n; ++j) {
                                const int kj = __ldg(k + j);
                                s0 += (int)__ldg(p + 3 * j) * kj;
                                s1 += (int)__ldg(p + 3 * j + 1) * kj;
                                s2 += (int)__ldg(p + 3 * j + 2) * kj;
                            }
                            *d0 = (unsigned char)round8(s0);
                            *d1 = (unsigned char)round8(s1);
                            *d2 = (unsigned char)round8(s2);
                            p += row_bytes; d0 += dstep; d1 += dstep; d2 += dstep;
                        }
                    }
                }
            }
        }
        if (!SPLIT) __syncthreads();                                    // hs (and, first time, the tables) complete

        // ---- vertical pass + normalisation table + store ----
        for (int q = do_v ? q_first : quads; q < quads; q += q_step) {
            const int qo = FLIP ? quads - 1 - q : q;
            for (int i = rg; i < nrow_out; i += rgs) {
                float* orow = outf + (size_t)(oy0 + i) * S + 4 * qo;
                if (!live) {                                            // collate padding: literal zeros
#pragma unroll
                    for (int c = 0; c < 3; ++c) stcs4(reinterpret_cast<float4*>(orow + (size_t)c * a.out_stride_c), f4_zero());
                    continue;
                }
                const int lo = vb[2 * i], n = vb[2 * i + 1];
                const int* k = vk + i * a.ksize_v;
                int acc[3][4];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[c][e] = 1 << (CLIP_PRECISION_BITS - 1);
                const unsigned char* row = hs_v + lo * 3 * Sp + 4 * q;
                for (int j = 0; j < n; ++j, row += 3 * Sp) {
                    const int kj = k[j];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const unsigned u = *reinterpret_cast<const unsigned*>(row + c * Sp);
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[c][e] += byte_of(u, e) * kj;
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* l = lut_s + c * 256;
                    float4 v;
                    if (FLIP) v = make_float4(lut_at(l, acc[c][3]), lut_at(l, acc[c][2]), lut_at(l, acc[c][1]), lut_at(l, acc[c][0]));
                    else      v = make_float4(lut_at(l, acc[c][0]), lut_at(l, acc[c][1]), lut_at(l, acc[c][2]), lut_at(l, acc[c][3]));
                    stcs4(reinterpret_cast<float4*>(orow + (size_t)c * a.out_stride_c), v);
                }
            }
        }
        __syncthreads();                                                // SPLIT: hand the tile over; else hs is rewritten by the next frame
    }
}

extern "C" int cf_normalize_lut(float* lut, float mean0, float mean1, float mean2, float std0, float std1, float std2,
                                cudaStream_t stream) {
    CF_CHECK_ARG(lut != nullptr, "lut is NULL");
    normalize_lut_kernel<<<3, 256, 0, stream>>>(lut, mean0, mean1, mean2, std0, std1, std2);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

static size_t clip_smem_bytes(int size, int rows_max, int ksize_v, int buffers) {
    const int Sp = (size + 15) & ~15;
    size_t tables = 768 * sizeof(float) + (size_t)CLIP_BAND * (2 + ksize_v) * sizeof(int);
    tables = (tables + 15) & ~(size_t)15;
    return tables + (size_t)buffers * rows_max * 3 * Sp;
}

extern "C" size_t cf_clip_preprocess_smem_bytes(int size, int rows_max, int ksize) {
    return clip_smem_bytes(size, rows_max, ksize, 2);
}

extern "C" int cf_clip_preprocess(const uint8_t* frames, float* out, const int* bounds_h, const int* kk_h,
                                  const int* bounds_v, const int* kk_v, const float* lut, int T, int H, int W, int x1,
                                  int y1, int crop, int size, int ksize_h, int ksize_v, int rows_max, int flip,
                                  int t_out, int64_t out_stride_c, cudaStream_t stream) {
    CF_CHECK_ARG(out && bounds_h && kk_h && bounds_v && kk_v && lut, "NULL pointer");
    CF_CHECK_ARG(T >= 0 && t_out >= T && t_out > 0, "need 0 <= T <= t_out");
    CF_CHECK_ARG(T == 0 || frames != nullptr, "frames is NULL");
    CF_CHECK_ARG(size > 0 && size % 4 == 0, "output size must be a positive multiple of 4");
    CF_CHECK_ARG(crop > 0 && x1 >= 0 && y1 >= 0 && x1 + crop <= W && y1 + crop <= H, "crop box outside the frame");
    CF_CHECK_ARG(ksize_h == cf_resample_ksize(crop, size) && ksize_v == ksize_h, "coefficient tables do not match crop/size");
    CF_CHECK_ARG(rows_max > 0, "rows_max must be positive");
    CF_CHECK_ARG(((uintptr_t)out & 15) == 0 && (out_stride_c % 4) == 0, "out must be 16-byte aligned with a channel stride % 4 == 0");
    CF_CHECK_ARG(out_stride_c >= (int64_t)t_out * size * size, "channel stride smaller than t_out frames");
    CF_CHECK_ARG(t_out <= 65535, "t_out above the grid limit");
    ClipArgs a;
    a.frames = frames; a.out = out; a.bounds_h = bounds_h; a.kk_h = kk_h; a.bounds_v = bounds_v; a.kk_v = kk_v; a.lut = lut;
    a.T = T; a.Tout = t_out; a.H = H; a.W = W; a.x1 = x1; a.y1 = y1; a.S = size; a.Sp = (size + 15) & ~15;
    a.ksize_h = ksize_h; a.ksize_v = ksize_v; a.flip = flip ? 1 : 0; a.rows_max = rows_max; a.out_stride_c = out_stride_c;
    // CFNET_CLIP_SPLIT=1 selects the two-group variant (A/B only: it lost the same-box comparison, 313 vs 272 us -- the kernel is
    // bound by the L1/LSU data pipe (byte loads, table look-ups), not by exposed load latency)
    const int want_split = cf_env("CFNET_CLIP_SPLIT", 0);
    size_t smem = clip_smem_bytes(size, rows_max, ksize_v, 2);
    const bool split = want_split && smem <= 200 * 1024;               // two thread groups + double-buffered tile
    if (!split) smem = clip_smem_bytes(size, rows_max, ksize_v, 1);
    CF_CHECK_ARG(smem <= 200 * 1024, "band of input rows does not fit in shared memory (down-scale factor too large)");
    const bool ks3 = ksize_h == 3;
    void (*kern)(const ClipArgs);
    if (split) kern = ks3 ? (flip ? clip_preprocess_kernel<true, true, true> : clip_preprocess_kernel<true, false, true>)
                          : (flip ? clip_preprocess_kernel<false, true, true> : clip_preprocess_kernel<false, false, true>);
    else       kern = ks3 ? (flip ? clip_preprocess_kernel<true, true, false> : clip_preprocess_kernel<true, false, false>)
                          : (flip ? clip_preprocess_kernel<false, true, false> : clip_preprocess_kernel<false, false, false>);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { cf_set_error("cf_clip_preprocess: smem opt-in failed: %s", cudaGetErrorString(e)); return CF_ERR_CUDA; }
    }
    // one thread per output column in the horizontal pass; (size/4) x row-groups in the vertical pass
    int threads = ((size + 31) / 32) * 32;
    threads = threads < 64 ? 64 : (threads > 512 ? 512 : threads);
    if (split) threads *= 2;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int sms = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int bands = cf_cdiv(size, CLIP_BAND);
    int groups = (per_sm * sms) / bands;                                // frame groups that fit in one resident wave
    groups = groups < 1 ? 1 : (groups > t_out ? t_out : groups);
    dim3 grid(bands, groups);
    kern<<<grid, threads, smem, stream>>>(a);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
