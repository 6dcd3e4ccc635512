/* cfnet_b200 -- C ABI of the B200-native Coarse-Fine X3D hot path.
 *
 * The reference (kkahatapitiya/Coarse-Fine-Networks) is pure PyTorch and has no FFI of its
 * own: its boundary is the nn.Module surface of x3d_fine.py / x3d_coarse.py / interp1d.py.
 * This header is the C-ABI that sits *under* that surface.  Every entry point names the
 * reference lines it replaces.  Conventions:
 *   - plain pointers and sizes only (no torch types); all pointers are DEVICE pointers to
 *     fp32 (or int32 where stated), dense, 16-byte aligned when the fast path is wanted;
 *   - the caller owns every buffer, including workspaces (`*_ws_bytes` tells the size);
 *   - nothing allocates, synchronises or keeps global mutable state besides the launch
 *     counter and the thread-local error string; kernels are enqueued on `stream`;
 *   - return 0 on success, non-zero on error; `cf_last_error()` describes the failure.
 * Big activations are viewed as [outer, T, inner] (inner contiguous): NCTHW tensors use
 * outer=B*C, inner=H*W, outer_per_b=C; channels-last (NTHWC) tensors use outer=B,
 * inner=H*W*C, outer_per_b=1.
 */
#ifndef CFNET_B200_H
#define CFNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* ---- library ------------------------------------------------------------------------ */
const char* cf_last_error(void);
int cf_abi_version(void);
/* number of kernel launches issued through this library since load (bench "gpu_launches") */
unsigned long long cf_launch_count(void);

/* ---- Grid Pool: confidence -> CDF (x3d_coarse.py:384-392) ---------------------------- */
/* g [B,n] pre-sigmoid per-interval confidences -> cdf [B,n+1]:
 * sigma(0.5 g); p = 1 - sigma; p / (sum p + 1e-16); warp prefix sum; cdf[:,0] = 0. */
int cf_gridpool_cdf_fwd(const float* g, float* cdf, int B, int n, cudaStream_t stream);
/* dcdf [B,n+1] -> dg [B,n] (autograd of the line above) */
int cf_gridpool_cdf_bwd(const float* g, const float* dcdf, float* dg, int B, int n, cudaStream_t stream);

/* ---- CDF -> frame-index bins (x3d_coarse.py:394,440 + grid_sampler_unnormalize) ------- */
/* coord [n] in [0,1] -> i0 [n] = floor(z) (int32, bit-exact), w1 [n] = z - i0,
 * z = (((coord-0.5)*2 + 1)/2)*(t_in-1) evaluated op-by-op in fp32. */
int cf_sample_bins(const float* coord, int32_t* i0, float* w1, int n, int t_in, cudaStream_t stream);
/* bins of F.interpolate(mode='linear'/'trilinear', align_corners=True) along T
 * (x3d_coarse.py:449,725): i0 [t_out], w1 [t_out]. */
int cf_linear_bins(int32_t* i0, float* w1, int t_in, int t_out, cudaStream_t stream);

/* ---- inverse CDF = Interp1d()(cdf, mid, mid) (interp1d.py:100-141; x3d_coarse.py:435-438) */
int cf_inverse_cdf_fwd(const float* cdf, float* inv, int32_t* ind, int B, int K, cudaStream_t stream);
/* dcdf_accum [B,K] += d inv / d cdf ^T dinv  (ind is not differentiable) */
int cf_inverse_cdf_bwd(const float* cdf, const int32_t* ind, const float* dinv, float* dcdf_accum, int B, int K,
                       cudaStream_t stream);

/* ---- general Interp1d()(x, y, xnew) (interp1d.py:4-141): D rows of N knots, P queries per row;
 * ind = clamp(searchsorted(x,xnew)-1, 0, N-2); ynew = y[ind] + slope[ind]*(xnew - x[ind]),
 * slope = (y[1:]-y[:-1])/(eps + x[1:]-x[:-1]).  A row stride of 0 broadcasts one row. */
int cf_interp1d_fwd(const float* x, const float* y, const float* xnew, float* ynew, int32_t* ind, int D, int N, int P,
                    int x_row_stride, int y_row_stride, int xnew_row_stride, cudaStream_t stream);
/* gradients accumulate (+=) into zero-filled buffers laid out like the inputs; NULL skips one */
int cf_interp1d_bwd(const float* x, const float* y, const float* xnew, const int32_t* ind, const float* dynew, float* dx_accum,
                    float* dy_accum, float* dxnew_accum, int D, int N, int P, int x_row_stride, int y_row_stride,
                    int xnew_row_stride, cudaStream_t stream);

/* ---- temporal lerp gather = F.grid_sample along T (x3d_coarse.py:396-403, 442-445) ---- */
/* out[o,k,:] = (1-w1[b,k]) x[o,i0[b,k],:] + w1[b,k] x[o,i0[b,k]+1,:],  b = o / outer_per_b,
 * frames outside [0,T-1] contribute zero.  x [outer,T,inner] -> out [outer,K,inner]. */
int cf_temporal_gather_fwd(const float* x, const int32_t* i0, const float* w1, float* out, int64_t outer,
                           int64_t outer_per_b, int T, int K, int64_t inner, cudaStream_t stream);
size_t cf_temporal_gather_bwd_ws_bytes(int64_t n_batch, int T, int K);
/* dx [outer,T,inner] written completely (gather form, zeros where untouched). */
int cf_temporal_gather_bwd_x(const float* gout, const int32_t* i0, const float* w1, float* dx, void* ws,
                             size_t ws_bytes, int64_t outer, int64_t outer_per_b, int T, int K, int64_t inner,
                             cudaStream_t stream);
/* dcoord_accum[b,k] += scale * sum_{o in b,i} gout[o,k,i] (x[o,i0+1,i] - x[o,i0,i]);
 * scale = T-1 turns d/dz into d/dcdf.  The terms cancel heavily, so the sum is carried in fp64
 * (dcoord_accum is a zero-filled double [B,K] buffer; the caller rounds it to fp32 once). */
int cf_temporal_gather_bwd_coord(const float* gout, const float* x, const int32_t* i0, double* dcoord_accum,
                                 int64_t outer, int64_t outer_per_b, int T, int K, int64_t inner, float scale,
                                 cudaStream_t stream);

/* ====================================================================================== */
/* X3D conv stacks.  Activations are channels-last fp32: a [B,C,T,H,W] tensor is stored as   */
/* [B, R, C] rows with R = T*H*W ("rows per sample").  Train-mode BatchNorm / SE / FiLM are   */
/* never materialised: every producer kernel accumulates per-(sample,channel) statistics in  */
/* its epilogue and every consumer applies a per-(sample,channel) affine + activation in its  */
/* prologue (tables of shape [B,C]).                                                         */
/* ====================================================================================== */

/* prologue applied to an input element x (and x2) of channel k of sample b */
#define CF_PRO_NONE 0          /* x                                   */
#define CF_PRO_AFFINE 1        /* a[b,k]*x + b[b,k]                   */
#define CF_PRO_AFFINE_RELU 2   /* relu(a*x + b)     bn+ReLU  (x3d_fine.py:150-151) */
#define CF_PRO_AFFINE_SWISH 3  /* swish(a*x + b)    bn+SE+Swish (x3d_fine.py:154-164) */
#define CF_PRO_AFFINE2 4       /* a*x + b*x2 + c    BatchNorm backward as an affine map of (dz, y) */
/* epilogue applied to an output element acc of channel n of sample b */
#define CF_EPI_NONE 0          /* acc (+bias)                         */
#define CF_EPI_RELU 1          /* relu(acc + bias)                    */
#define CF_EPI_DRELU 2         /* acc * [ea*aux + eb > 0]             */
#define CF_EPI_DSWISH 3        /* acc * swish'(ea*aux + eb)           */
#define CF_EPI_ADD_AUX 4       /* acc + aux                           */
#define CF_EPI_SIGMOID 5       /* sigmoid(acc + bias)                 */
#define CF_EPI_AFFINE 6        /* ea*acc + eb        folded (eval) BatchNorm of a shortcut conv (x3d_fine.py:286-288)       */
#define CF_EPI_AFFINE_ADD_RELU 7 /* relu(ea*acc + eb + aux)  folded bn3 + residual + ReLU in conv3's epilogue (:167-173, eval) */
/* statistics accumulated (double atomics) into stats[b][n][0..1] */
#define CF_STATS_NONE 0
#define CF_STATS_SUM_SQ 1      /* sum y, sum y^2      (forward: BatchNorm + SE pooling) */
#define CF_STATS_SUM_AUX 2     /* sum y, sum y*aux    (backward: BatchNorm/SE reductions) */

/* Geometry of the row <-> position map shared by the conv kernels.  The "dense" side has one
 * row per position (t,h,w) of a [T,H,W] volume (rows per sample R = T*H*W).  The "gathered"
 * side is a [Ti,Hi,Wi] volume addressed at (t*st-pt+dt, h*sh-ph+dh, w*sw-pw+dw) for tap
 * (dt,dh,dw) of a kt x kh x kw window; positions outside the volume read as zero (after the
 * prologue) / are skipped on scatter.  Element (position p, channel c) of sample b lives at
 * base + b*sample_stride + p*pos_stride + c*ch_stride, so both channels-last activations
 * (pos_stride=C, ch_stride=1) and the NCTHW network input (pos_stride=1, ch_stride=Ti*Hi*Wi)
 * can be gathered.  With taps > 1 the GEMM reduction index is k = c*taps + tap, which is the
 * native [Cout][Cin][kt][kh][kw] weight layout of nn.Conv3d.  1x1x1 stride-1: all taps 1,
 * strides 1, pads 0, (Ti,Hi,Wi) = (T,H,W). */
typedef struct {
    int T, H, W;
    int Ti, Hi, Wi;
    int kt, kh, kw;
    int st, sh, sw;
    int pt, ph, pw;
    int64_t pos_stride, ch_stride, sample_stride;
} cf_geom;

/* Convolution as a GEMM over rows (pointwise 1x1x1 convs, linear layers, and -- through the
 * tap gather -- the small dense convs of the stem and of the Grid Pool confidence branch):
 *   y[b,r,n] = epi( sum_k pro(x[b, gather(r,k)]) * w[n*w_sn + k*w_sk] (+ bias[n]) )
 * forward  : conv1/conv3/downsample/conv5/fc1/fc2 (x3d_fine.py:149,166,286,356,370,380),
 *            conv1_s (x3d_fine.py:210-215), pool_1.conv1-3 (x3d_coarse.py:362-366)
 * backward : data gradient = the same GEMM with (w_sn,w_sk) swapped; for strided / windowed
 *            convs the result is scattered (scatter_out) to the gathered side. */
typedef struct {
    const float* x;      /* dense [B,R,K], or the gathered tensor when gather_in */
    const float* x2;     /* second input for CF_PRO_AFFINE2 (dense only), else NULL */
    const float* w;
    const float* bias;   /* [N] or NULL */
    float* y;            /* dense [B,R,N], or the gathered-side tensor when scatter_out */
    const float* pro_a;  /* [B,Cin] tables (NULL when pro_mode == NONE) */
    const float* pro_b;
    const float* pro_c;
    const float* aux;    /* dense [B,R,N] for DRELU/DSWISH/ADD_AUX/STATS_SUM_AUX */
    const float* epi_a;  /* [B,N] tables for DRELU/DSWISH */
    const float* epi_b;
    double* stats;       /* [B,N,2] or NULL */
    int64_t w_sn, w_sk;
    int B, K, N;         /* K (or N when scatter_out) = channels * taps of the gathered side */
    cf_geom g;
    int gather_in;       /* 1: x is addressed through g (strided and/or windowed conv forward) */
    int scatter_out;     /* 1: y is addressed through g; taps > 1 => atomic accumulation */
    int accumulate;      /* 1: y += result */
    int pro_mode, epi_mode, stats_mode;
    float* wpack;        /* workspace of cf_pw_tc_ws_bytes(K,N) bytes (128-B aligned) or NULL.  When given and the
                          * problem is dense (no gather / scatter) the GEMM runs on the tcgen05 tensor cores
                          * (3xTF32, fp32 accumulation in TMEM); NULL selects the fp32 CUDA-core kernel. */
    int64_t wpack_bytes;
    int wpack_nt;        /* 0: wpack is scratch, this call packs w into it.  > 0: wpack already holds w packed by
                          * cf_pw_pack_many with channel tile cf_pw_pack_nt(K,N) == wpack_nt (weights that did not
                          * change since): the persistent kernel skips its packing launch when its tiling matches. */
} cf_pw_args;
int cf_pw_conv(const cf_pw_args* a, cudaStream_t stream);
/* bytes of the packed (hi/lo split, swizzled) weight workspace of the tensor-core path */
size_t cf_pw_tc_ws_bytes(int K, int N);
/* Packing the weights of MANY GEMMs in one launch (once per optimizer step instead of once per conv call):
 * item i packs w (strides w_sn, w_sk; [N,K] logical) into pack (cf_pw_tc_ws_bytes(K,N) bytes) with the channel tile
 * nt = cf_pw_pack_nt(K,N).  `items` is a DEVICE array of n items. */
typedef struct {
    const float* w;
    float* pack;
    int64_t w_sn, w_sk;
    int K, N, nt, pad;
} cf_pack_item;
int cf_pw_pack_nt(int K, int N);                 /* the default channel tile for (K,N) */
int cf_pw_plan_nt(const cf_pw_args* a);          /* the channel tile THIS call will use (its shared-memory plan may narrow the
                                                  * default), 0 if it does not take the persistent tensor-core kernel */
int cf_pw_pack_many(const cf_pack_item* items, int n, cudaStream_t stream);
size_t cf_sizeof_pack_item(void);
/* debug: per-role cycle counters of CTA 0 of the last tensor-core launch made with CFNET_PW_TC_TIMING=1 (24 values) */
int cf_pw_tc_debug_read(long long* out16);

/* weight gradient:
 *   dw[n*K + k] += sum_{b,r} pro_dy(dy[b,r,n], dy2[b,r,n]) * pro_x(x[b, gather(r,k)]);
 *   dbias[n]    += sum_{b,r} pro_dy(...) */
typedef struct {
    const float* dy;     /* dense [B,R,N] */
    const float* dy2;    /* second input for CF_PRO_AFFINE2 or NULL */
    const float* dy_a;   /* [B,N] tables */
    const float* dy_b;
    const float* dy_c;
    const float* x;      /* dense [B,R,K] or gathered tensor */
    const float* x_a;    /* [B,Cin] tables */
    const float* x_b;
    float* dw;           /* [N,K] (+=, caller zero-fills once per step) */
    float* dbias;        /* [N] or NULL (+=) */
    int B, K, N;
    cf_geom g;
    int gather_in;
    int dy_mode, x_mode;
} cf_pw_wgrad_args;
int cf_pw_wgrad(const cf_pw_wgrad_args* a, cudaStream_t stream);
size_t cf_sizeof_pw_args(void);
size_t cf_sizeof_pw_wgrad_args(void);

/* depthwise conv (channels-last).  g.(T,H,W) = output volume, g.(Ti,Hi,Wi) = input volume.
 *   fwd  : y[B,T,H,W,C]   = sum_taps pro(x[B,Ti,Hi,Wi,C]) * w[C,taps]        (x3d_fine.py:89-97,153; conv1_t :216-222)
 *   dgrad: y[B,Ti,Hi,Wi,C] = epi( sum_taps pro(x[B,T,H,W,C], x2) * w[C,taps] ),  aux/epi tables at the
 *          input positions (the pre-activation the forward prologue consumed)
 *   wgrad: y[C,taps] += sum_pos pro(x[pos], x2[pos]) * act(aux[pos_in(tap)]),  act = relu(epi_a*aux+epi_b)
 *          when epi_a != NULL (x = output gradient, aux = forward input)
 *   dgrad with dw_out != NULL (and epi_mode DRELU): the same call also accumulates the weight gradient
 *          dw_out[C,taps] += ... (exactly what wgrad would add to y) -- for the 3x3x3 stride-1 convs in ONE pass over
 *          (x, x2, aux) (x3d_dw3.cu), otherwise as the two kernels back to back. */
typedef struct {
    const float* x;
    const float* x2;
    const float* w;
    float* y;
    const float* pro_a;
    const float* pro_b;
    const float* pro_c;
    const float* aux;
    const float* epi_a;
    const float* epi_b;
    double* stats;       /* [B,C,2] */
    float* dw_out;       /* [C,taps] += (dgrad only) or NULL */
    int B, C;
    cf_geom g;
    int pro_mode, epi_mode, stats_mode;
} cf_dw_args;
int cf_dw_conv_fwd(const cf_dw_args* a, cudaStream_t stream);
int cf_dw_conv_dgrad(const cf_dw_args* a, cudaStream_t stream);
int cf_dw_conv_wgrad(const cf_dw_args* a, cudaStream_t stream);
size_t cf_sizeof_dw_args(void);

/* SubBatchNorm3d (x3d_fine.py:13-62): per-sample statistics -> per-(sample,channel) affine
 * tables tab_a*y + tab_b == weight*BN(y)+bias.  training: batch statistics per split group
 * (sample b belongs to group b % splits, channel index g*C+c of split_bn), running stats of
 * split_bn updated with `momentum` (unbiased variance); eval: running stats of `bn`. */
typedef struct {
    const double* stats;     /* [B,C,2] (sum y, sum y^2) per sample; unused in eval */
    const float* gamma;      /* [C] */
    const float* beta;       /* [C] */
    float* running_mean;     /* training: [splits*C] (updated); eval: [C] (read) */
    float* running_var;
    float* tab_a;            /* [B,C] out */
    float* tab_b;            /* [B,C] out */
    float* mean;             /* [splits,C] out (saved for backward) */
    float* invstd;           /* [splits,C] out */
    int B, C, splits;
    int64_t rows_per_sample;
    float momentum, eps;
    int training;
} cf_bn_args;
int cf_bn_finalize(const cf_bn_args* a, cudaStream_t stream);

/* BatchNorm backward as an affine map: dy = P*d + Q*y + R per (sample,channel), where d is the
 * tensor the sums were taken of.  sums[b,c] = (sum dz, sum dz*y) of the gradient w.r.t. the BN
 * output.  Optional SE coupling: dz = gate*d + cst (see cf_se_bwd) => P = c1*gate, R += c1*cst.
 * dgamma/dbeta accumulate (+=).  training==0: running-stat BN, dy = a*dz. */
typedef struct {
    const double* sums;      /* [B,C,2] */
    const float* gamma;      /* [C] */
    const float* mean;       /* [splits,C] */
    const float* invstd;     /* [splits,C] */
    const float* gate;       /* [B,C] or NULL */
    const float* cst;        /* [B,C] or NULL */
    float* dgamma;           /* [C] += */
    float* dbeta;            /* [C] += */
    float* tab_p;            /* [B,C] out */
    float* tab_q;
    float* tab_r;
    int B, C, splits;
    int64_t rows_per_sample;
    int training;
} cf_bn_bwd_args;
int cf_bn_bwd_coeffs(const cf_bn_bwd_args* a, cudaStream_t stream);

/* Squeeze-and-Excitation (x3d_fine.py:157-163): pooled = mean_thw(bn2(y2)) from the statistics,
 * hidden = relu(fc1 pooled), gate = sigmoid(fc2 hidden); the gate is folded into the tables
 * (out_a = gate*tab_a, out_b = gate*tab_b) that the Swish prologue of conv3 consumes. */
typedef struct {
    const double* stats;     /* [B,C,2] of y2 */
    const float* tab_a;      /* [B,C] BN2 tables */
    const float* tab_b;
    const float* w1;         /* fc1.weight [Wd,C] */
    const float* b1;         /* [Wd] */
    const float* w2;         /* fc2.weight [C,Wd] */
    const float* b2;         /* [C] */
    float* pooled;           /* [B,C]  saved */
    float* hidden;           /* [B,Wd] saved */
    float* gate;             /* [B,C]  saved */
    float* out_a;            /* [B,C] */
    float* out_b;            /* [B,C] */
    int B, C, Wd;
    int64_t rows_per_sample;
} cf_se_args;
int cf_se_fwd(const cf_se_args* a, cudaStream_t stream);

/* SE backward.  In: sums[b,c] = (sum dU, sum dU*y2) with dU = dL/d(gate*bn2(y2)).  Out (in
 * place): sums of dz = gate*dU + cst w.r.t. bn2's output, cst[b,c] = dL/dpooled / R; fc grads +=. */
typedef struct {
    double* sums;            /* [B,C,2] in/out */
    const double* stats_y;   /* [B,C,2] forward statistics of y2 */
    const float* tab_a;      /* BN2 tables [B,C] (before the gate) */
    const float* tab_b;
    const float* w1;
    const float* w2;
    const float* pooled;
    const float* hidden;
    const float* gate;
    float* dw1;              /* += */
    float* db1;
    float* dw2;
    float* db2;
    float* cst;              /* [B,C] out */
    int B, C, Wd;
    int64_t rows_per_sample;
} cf_se_bwd_args;
int cf_se_bwd(const cf_se_bwd_args* a, cudaStream_t stream);

/* residual join (x3d_fine.py:169-173): out = relu(a3*y3 + b3 + res'),
 * res' = ar*res + br (downsample branch, tables) or res (identity) or 0 (res == NULL).
 * Stage-final blocks of the fine stream's global tower (x3d_fine.py:345-354) also emit the (H/rh, W/rw) block average
 * of `out` -- F.adaptive_avg_pool3d(x, (None,7,7)) -- from the same pass: pooled [B,T,H/rh,W/rw,C] != NULL with
 * rows_per_sample == T*H*W and H % rh == W % rw == 0 (one thread owns a pooling block: no atomics). */
typedef struct {
    const float* y;          /* [B,R,C] */
    const float* tab_a;      /* [B,C] */
    const float* tab_b;
    const float* res;        /* [B,R,C] or NULL */
    const float* res_a;      /* [B,C] or NULL */
    const float* res_b;
    float* out;              /* [B,R,C] */
    float* pooled;           /* [B,T,H/rh,W/rw,C] or NULL */
    int B, C;
    int T, H, W, rh, rw;     /* only read when pooled != NULL */
    int64_t rows_per_sample;
} cf_residual_args;
int cf_residual_fwd(const cf_residual_args* a, cudaStream_t stream);

/* backward of the join: dz = dout' * [out > 0]; sums_y[b,c] = (sum dz, sum dz*y);
 * sums_res likewise against `res` (downsample branch pre-BN tensor) when given.
 * dout' = dout + dpool[block of the row] / (rh*rw) when the forward also emitted the pooled features
 * (dpool [B,T,H/rh,W/rw,C] != NULL; dout may then be NULL = no gradient from the next stage). */
typedef struct {
    const float* dout;       /* [B,R,C] (or NULL with dpool) */
    const float* out;
    const float* y;
    const float* res;        /* or NULL */
    const float* dpool;      /* [B,T,H/rh,W/rw,C] or NULL */
    float* dz;
    double* sums_y;          /* [B,C,2] */
    double* sums_res;        /* [B,C,2] or NULL */
    int B, C;
    int T, H, W, rh, rw;     /* only read when dpool != NULL */
    int64_t rows_per_sample;
} cf_residual_bwd_args;
int cf_residual_bwd(const cf_residual_bwd_args* a, cudaStream_t stream);

/* block average pooling over (H,W) of a channels-last tensor with an optional bn+relu
 * prologue: AdaptiveAvgPool3d((None,1,1)) (x3d_fine.py:255,366) and
 * adaptive_avg_pool3d(x,(None,7,7)) (x3d_fine.py:345-360) when H,W are multiples of the output.
 * x [B,T,H,W,C] -> y [B,T,H/rh,W/rw,C]. */
typedef struct {
    const float* x;
    float* y;
    const float* tab_a;      /* [B,C] or NULL (no prologue) */
    const float* tab_b;
    int B, C, T, H, W, rh, rw;
} cf_pool_args;
int cf_block_avgpool_fwd(const cf_pool_args* a, cudaStream_t stream);
/* backward: dz[B,T,H,W,C] = dy/(rh*rw) * [tab_a*x+tab_b > 0] (x = pre-activation, when tables given);
 * sums[b,c] = (sum dz, sum dz*x).  `accumulate`: dz += ... */
typedef struct {
    const float* dy;         /* [B,T,H/rh,W/rw,C] */
    const float* x;          /* pre-activation [B,T,H,W,C] or NULL */
    const float* tab_a;
    const float* tab_b;
    float* dz;               /* [B,T,H,W,C] */
    double* sums;            /* [B,C,2] or NULL */
    int B, C, T, H, W, rh, rw;
    int accumulate;
} cf_pool_bwd_args;
int cf_block_avgpool_bwd(const cf_pool_bwd_args* a, cudaStream_t stream);

/* out = dy * [y > 0]  (backward of a materialised ReLU, e.g. after fc1: x3d_fine.py:371) */
int cf_relu_bwd(const float* dy, const float* y, float* out, int64_t n, cudaStream_t stream);

/* standalone module surfaces (SubBatchNorm3d.forward x3d_fine.py:51-62, Swish :65-86) */
/* stats[b,c] += (sum x, sum x*y) (y == NULL: sum x^2) over the rows of sample b */
int cf_channel_stats(const float* x, const float* y, double* stats, int B, int C, int64_t rows_per_sample,
                     cudaStream_t stream);
/* out = pro(x, x2) with per-(sample,channel) tables, mode = CF_PRO_* */
typedef struct {
    const float* x;
    const float* x2;
    const float* tab_a;
    const float* tab_b;
    const float* tab_c;
    float* out;
    int B, C;
    int64_t rows_per_sample;
    int mode;
} cf_affine_args;
int cf_affine_apply(const cf_affine_args* a, cudaStream_t stream);
int cf_swish_fwd(const float* x, float* out, int64_t n, cudaStream_t stream);
int cf_swish_bwd(const float* x, const float* dy, float* dx, int64_t n, cudaStream_t stream);

/* out = dy * y * (1 - y)  (backward of a materialised sigmoid: x3d_coarse.py:219,245,336) */
int cf_sigmoid_bwd(const float* dy, const float* y, float* out, int64_t n, cudaStream_t stream);

/* ====================================================================================== */
/* Multi-stage Fusion (x3d_coarse.py:175-351).  Fine-stream features are channels-last     */
/* [B,Tf,P,C] row tensors with P = 7*7 pixels; everything is evaluated at that 7x7 base     */
/* resolution (the reference's adaptive_max_pool2d up-sampling at :214,315,322 is an exact   */
/* nearest replication, so every later op is constant over the replicated blocks) and only   */
/* the final scale/shift (FiLM) kernel reads the base maps through an index map.             */
/* ====================================================================================== */

/* Gaussian.forward with tx given (x3d_coarse.py:256-286):
 *   mu[b,k] = (cdf[b,k]*tx + start[b]) / ratio; sigma_b = sum_t mask[b,t] / 8;
 *   f = exp(-(t-mu)^2 / (2 sigma^2 + 1e-16)); gx[b,t,k] = f / (max_t f + 1e-16). */
int cf_gaussian_fwd(const float* cdf, const float* start, const float* mask, float* gx, int B, int Tf, int Tl,
                    float tx, float ratio, cudaStream_t stream);
/* dcdf_accum[b,k] += d gx / d cdf ^T dgx (through mu and through the max) */
int cf_gaussian_bwd(const float* cdf, const float* start, const float* mask, const float* dgx, float* dcdf_accum,
                    int B, int Tf, int Tl, float tx, float ratio, cudaStream_t stream);

/* RewightLayer aligned aggregation (x3d_coarse.py:221-225):
 *   A[t,k,p] = att[t,p]*gx[t,k];  den[k,p] = sum_t A*mask + 1e-6;
 *   agg[k,p,c] = sum_t x[t,p,c] * A[t,k,p] * mask[t] / den[k,p]          (per sample b) */
typedef struct {
    const float* x;          /* [B,Tf,P,C] */
    const float* att;        /* [B,Tf,P]  sigmoid attention */
    const float* gx;         /* [B,Tf,Tl] */
    const float* mask;       /* [B,Tf] */
    float* agg;              /* [B,Tl,P,C] out */
    float* den;              /* [B,Tl,P] out (saved for backward) */
    int B, C, Tf, Tl, P;
} cf_rewight_args;
int cf_rewight_agg_fwd(const cf_rewight_args* a, cudaStream_t stream);
typedef struct {
    const float* x;
    const float* att;
    const float* gx;
    const float* mask;
    const float* agg;
    const float* den;
    const float* dagg;       /* [B,Tl,P,C] */
    float* dx;               /* [B,Tf,P,C] out, or NULL (fine features detached) */
    float* datt;             /* [B,Tf,P] out */
    float* dgx;              /* [B,Tf,Tl] += (caller zero-fills) */
    int B, C, Tf, Tl, P;
} cf_rewight_bwd_args;
int cf_rewight_agg_bwd(const cf_rewight_bwd_args* a, cudaStream_t stream);
size_t cf_sizeof_rewight_args(void);
size_t cf_sizeof_rewight_bwd_args(void);

/* scale/shift modulation x*m + c (x3d_coarse.py:664,669,674,679,721) with m, c given at a
 * base resolution (Hb,Wb) dividing (H,W): out[b,t,y,x,ch] = x * scale[b,t,y/rh,x/rw,ch] + shift[...]. */
typedef struct {
    const float* x;          /* [B,T,H,W,C] */
    const float* scale;      /* [B,T,Hb,Wb,C] */
    const float* shift;      /* [B,T,Hb,Wb,C] */
    float* out;              /* [B,T,H,W,C] */
    int B, C, T, H, W, Hb, Wb;
} cf_film_args;
int cf_film_fwd(const cf_film_args* a, cudaStream_t stream);
typedef struct {
    const float* dout;       /* [B,T,H,W,C] */
    const float* x;
    const float* scale;
    float* dx;               /* [B,T,H,W,C] out or NULL */
    float* dscale;           /* [B,T,Hb,Wb,C] out */
    float* dshift;           /* [B,T,Hb,Wb,C] out */
    int B, C, T, H, W, Hb, Wb;
} cf_film_bwd_args;
int cf_film_bwd(const cf_film_bwd_args* a, cudaStream_t stream);
size_t cf_sizeof_film_args(void);
size_t cf_sizeof_film_bwd_args(void);

/* nearest replication of a base map to (H,W) (what F.adaptive_max_pool2d computes at
 * x3d_coarse.py:214,315,322 when H,W are multiples of Hb,Wb); module-surface outputs only. */
int cf_nearest_up(const float* x, float* out, int B, int T, int Hb, int Wb, int H, int W, int C, cudaStream_t stream);
/* backward: dx[b,t,yb,xb,c] = sum over the replicated block of dout */
int cf_nearest_up_bwd(const float* dout, float* dx, int B, int T, int Hb, int Wb, int H, int W, int C,
                      cudaStream_t stream);

/* block max pooling over (H,W) of a channels-last map: F.adaptive_max_pool2d as a down-sampler
 * (x3d_coarse.py:315,322, H,W multiples of the target).  idx = position of the (first) maximum
 * inside the block, saved for the backward scatter.  x [B,T,H,W,C] -> out [B,T,H/rh,W/rw,C]. */
int cf_block_maxpool_fwd(const float* x, float* out, int32_t* idx, int B, int T, int H, int W, int C, int rh, int rw,
                         cudaStream_t stream);
int cf_block_maxpool_bwd(const float* dout, const int32_t* idx, float* dx, int B, int T, int H, int W, int C, int rh, int rw,
                         cudaStream_t stream);

/* ====================================================================================== */
/* GPU JPEG decode (loader input)                                        */
/* ====================================================================================== */
/* Replaces the per-frame host decode of the reference's loaders (charades_fine.py:22-27 pil_loader, :46-56
 * video_loader, :78-101 load_rgb_frames): the JPEG streams of a clip are decoded by nvJPEG's batched decoder straight
 * into the uint8 [n,H,W,3] device tensor cf_clip_preprocess reads.  The decoder object owns the nvJPEG handle / state
 * (nvJPEG is dlopen'ed at creation; it allocates its own scratch memory); one decoder per host thread.
 * `data[i]` / `lengths[i]` are HOST pointers to the i-th JPEG stream; every stream must decode to H x W.
 * Decoded pixels follow the JPEG standard but are not bit-identical to libjpeg-turbo's (see tests/test_jpeg_gpu.py). */
int cf_jpeg_create(void** decoder);
int cf_jpeg_destroy(void* decoder);
int cf_jpeg_image_info(void* decoder, const unsigned char* data, size_t length, int* height, int* width, int* components);
int cf_jpeg_decode_batch(void* decoder, const unsigned char* const* data, const size_t* lengths, int n, unsigned char* out, int H,
                         int W, cudaStream_t stream);

/* ====================================================================================== */
/* Training-step glue                                                                       */
/* ====================================================================================== */
/* Charades localisation loss of the scripts (train_fine.py:199-212,226; train_coarse_fineFEAT.py:226-247):
 * logits [B,C,T] are linearly interpolated to the label length TL -- align_corners != 0: the
 * align_corners=True grid of train_fine.py:199; align_corners == 0: F.interpolate's default half-pixel grid
 * (src = max((u+0.5)*T/TL - 0.5, 0)), which is what train_coarse_fineFEAT.py:226 calls --,
 * probs = sigmoid * mask; loss2[0] += BCE_mean(max_t probs, max_t labels), loss2[1] += BCE_sum(probs,
 * labels)/(sum(mask)*C) (caller zero-fills loss2); dlogits [B,C,T] = d(scale*(loss2[0]+loss2[1]))/dlogits,
 * i.e. scale = 1/(2*num_steps_per_update) reproduces the scripts.  dlogits may be NULL (evaluation). */
int cf_charades_loss(const float* logits, const float* labels, const float* masks, float* loss2, float* dlogits, int B, int C,
                     int T, int TL, float scale, int align_corners, cudaStream_t stream);

/* The data-parallel collective: one NCCL communicator per process / GPU behind an opaque handle and one
 * in-place fp32 SUM all-reduce of the flat gradient buffer per step over NVLink / NVSwitch -- replaces nn.DataParallel's
 * per-step broadcast / gather / reduce-add (train_fine.py:122-123, train_coarse_fineFEAT.py:129-130).  NCCL is dlopen'ed
 * (libnccl.so.2).  cf_comm_unique_id fills 128 bytes on ONE rank; the caller hands them to the other ranks (host side),
 * then every rank calls cf_comm_init with its current CUDA device set.  The 1/world factor lives in cf_sgd_flat. */
int cf_comm_unique_id(void* id128);
int cf_comm_init(void** comm, int world, int rank, const void* id128);
int cf_comm_allreduce(void* comm, float* buf, int64_t n, cudaStream_t stream);
int cf_comm_destroy(void* comm);

/* fused SGD with momentum over flat fp32 buffers (optim.SGD, train_fine.py:130): g = grad_scale*g + wd*p;
 * v = momentum*v + g; p -= lr*v; g = 0.  Elements [0,n_split) use lr0, the rest lr1 (the 'rw'/'mix'
 * parameter group at 10x, train_coarse_fineFEAT.py:137-141).  grad_scale = 1/world folds the
 * data-parallel mean into the update. */
int cf_sgd_flat(float* p, float* g, float* v, int64_t n, int64_t n_split, float lr0, float lr1, float momentum,
                float weight_decay, float grad_scale, cudaStream_t stream);

/* ---- evaluation: average precision per class (apmeter.py:98-136) ---------------------------------------- */
/* truth_sorted [K,N]: the 0/1 targets of class k ordered by descending score of that class (the caller sorts);
 * weight_sorted [K,N] per-sample weights in the same order, or NULL; ap [K] out:
 * ap[k] = sum_{i: truth} (tp_i / rg_i) / max(sum truth, 1), tp = cumsum(truth*weight), rg = 1..N or cumsum(weight). */
int cf_ap_sorted(const float* truth_sorted, const float* weight_sorted, float* ap, int N, int K, cudaStream_t stream);

/* ---- input pipeline: decoded RGB frames -> normalised clip ---------------------------
 * Replaces the per-frame CPU chain of charades_fine.py:170-172 with the transforms of train_fine.py:74-80:
 * MultiScaleRandomCropMultigrid / CenterCropScaled (transforms/spatial_transforms.py:488-503, 216-230: crop box +
 * PIL img.resize(BILINEAR)), RandomHorizontalFlip (342-354), ToTensor(255) (46-87), Normalize (108-118), and the
 * zero padding of mt_collate_fn (charades_fine.py:215-226).  Bit-exact with Pillow's 8-bit ImagingResample. */

/* HOST functions (no GPU): Pillow's bilinear coefficient table for resampling a whole axis in_size -> out_size.
 * ksize = 2*ceil(max(in/out,1)) + 1 taps per output; bounds [out,2] = (first input index, tap count);
 * kk [out,ksize] 22-bit fixed-point weights (host pointers, caller-allocated). */
int cf_resample_ksize(int in_size, int out_size);
int cf_resample_coeffs(int in_size, int out_size, int* bounds, int* kk);

/* lut [3,256] (device) = ((v/255) - mean_c) / std_c with every step rounded to fp32 (ToTensor + Normalize). */
int cf_normalize_lut(float* lut, float mean0, float mean1, float mean2, float std0, float std1, float std2,
                     cudaStream_t stream);

/* frames [T,H,W,3] uint8 (device, RGB interleaved = PIL tobytes()); crop box (x1,y1,crop,crop); bounds_h, kk_h, bounds_v, kk_v: the
 * device copies of cf_resample_coeffs(crop, size) (horizontal and vertical tables may be the same buffers);
 * rows_max >= the largest number of crop rows any band of 16 output rows touches; flip != 0 mirrors left-right.
 * out: base of one sample of a [B,3,t_out,size,size] fp32 batch (channel stride out_stride_c elements >=
 * t_out*size*size); frames [T,t_out) are written as zeros (collate padding).  size % 4 == 0. */
size_t cf_clip_preprocess_smem_bytes(int size, int rows_max, int ksize);   /* dynamic shared memory of the launch (<= 200 KB) */
int cf_clip_preprocess(const uint8_t* frames, float* out, const int* bounds_h, const int* kk_h,
                       const int* bounds_v, const int* kk_v, const float* lut, int T, int H, int W, int x1,
                       int y1, int crop, int size, int ksize_h, int ksize_v, int rows_max, int flip,
                       int t_out, int64_t out_stride_c, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CFNET_B200_H */
