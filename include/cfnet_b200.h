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
 * scale = T-1 turns d/dz into d/dcdf. */
int cf_temporal_gather_bwd_coord(const float* gout, const float* x, const int32_t* i0, float* dcoord_accum,
                                 int64_t outer, int64_t outer_per_b, int T, int K, int64_t inner, float scale,
                                 cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CFNET_B200_H */
