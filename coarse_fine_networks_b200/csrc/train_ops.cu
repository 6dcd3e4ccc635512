// Training-step glue kernels (sm_100a): the Charades localisation loss of the reference scripts
// and a fused SGD-momentum update over the flat parameter / gradient buffers.
//
// Reference semantics:
//   loss ........ train_fine.py:199-212,226: F.interpolate(logits, TL, 'linear', align_corners=True);
//                 train_coarse_fineFEAT.py:226-247: F.interpolate(logits, TL, 'linear') -- the DEFAULT
//                 align_corners=False grid (half-pixel centres); the caller chooses with `align_corners`;
//                 probs = sigmoid * mask;
//                 cls = BCE_mean(max_t probs, max_t labels); loc = BCE_sum(probs, labels) / (sum(mask) * C);
//                 loss = (cls + loc) / (2 * num_steps_per_update)
//   optimiser ... optim.SGD(momentum=0.9, weight_decay=1e-5) (train_fine.py:130), fusion parameters
//                 ('rw' / 'mix') at 10x the learning rate (train_coarse_fineFEAT.py:137-141);
//                 PyTorch update order: g += wd*p; v = mu*v + g; p -= lr*v.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

// nn.BCELoss clamps its log terms at -100 and its backward uses (p - y) / max(p (1 - p), 1e-12)
__device__ __forceinline__ float bce_term(float p, float y) {
    float lp = fmaxf(logf(p), -100.0f), lq = fmaxf(logf(1.0f - p), -100.0f);
    return -(y * lp + (1.0f - y) * lq);
}
__device__ __forceinline__ float bce_grad(float p, float y) { return (p - y) / fmaxf((1.0f - p) * p, 1e-12f); }

// one CTA (128 threads) per (b,c) row.  smem: dl[T]
__global__ void __launch_bounds__(128) charades_loss_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                            const float* __restrict__ masks, float* __restrict__ loss,
                                                            float* __restrict__ dlogits, int B, int C, int T, int TL,
                                                            float scale, int align_corners) {
    extern __shared__ float dl[];
    __shared__ float red[4];
    __shared__ int redi[4];
    __shared__ float bc[4];
    const int row = blockIdx.x, b = row / C, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* lg = logits + (size_t)row * T;
    const float* lb = labels + (size_t)row * TL;
    const float* mk = masks + (size_t)b * TL;
    for (int i = tid; i < T; i += 128) dl[i] = 0.f;
    // sum of all masks (denominator of the localisation term)
    float ms = 0.f;
    for (int i = tid; i < B * TL; i += 128) ms += masks[i];
    ms = warp_sum(ms);
    if (lane == 0) red[warp] = ms;
    __syncthreads();
    const float msum = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    // ATen's area_pixel_compute_scale / _source_index (UpSample.cuh), evaluated in fp32 like upsample_linear1d
    const float step = align_corners ? ((TL > 1) ? (float)(T - 1) / (float)(TL - 1) : 0.f) : (float)T / (float)TL;
    auto source = [&](int u) -> float {
        if (align_corners) return __fmul_rn(step, (float)u);
        const float sidx = __fsub_rn(__fmul_rn(step, __fadd_rn((float)u, 0.5f)), 0.5f);
        return sidx < 0.f ? 0.f : sidx;
    };
    float loc = 0.f, pmax = -1.f, ymax = -INFINITY;
    int amax = 0;
    for (int u = tid; u < TL; u += 128) {
        float src = source(u);
        int j0 = min((int)src, T - 1), j1 = min(j0 + 1, T - 1);
        float lam = src - (float)j0;
        float z = (1.0f - lam) * lg[j0] + lam * lg[j1];
        float p = cf_sigmoid(z) * mk[u];
        float y = lb[u];
        loc += bce_term(p, y);
        if (p > pmax) { pmax = p; amax = u; }
        ymax = fmaxf(ymax, y);
    }
    loc = warp_sum(loc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float op = __shfl_xor_sync(0xffffffffu, pmax, o);
        int oa = __shfl_xor_sync(0xffffffffu, amax, o);
        if (op > pmax || (op == pmax && oa < amax)) { pmax = op; amax = oa; }
        ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (lane == 0) { red[warp] = loc; bc[warp] = pmax; redi[warp] = amax; }
    __syncthreads();
    float locs = red[0] + red[1] + red[2] + red[3];
    float pm = bc[0];
    int am = redi[0];
    for (int w = 1; w < 4; ++w)
        if (bc[w] > pm || (bc[w] == pm && redi[w] < am)) { pm = bc[w]; am = redi[w]; }
    __syncthreads();
    if (lane == 0) red[warp] = ymax;
    __syncthreads();
    const float ym = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    const float loc_den = msum * (float)C, cls_den = (float)B * (float)C;
    if (tid == 0) {
        atomicAdd(loss + 0, bce_term(pm, ym) / cls_den);
        atomicAdd(loss + 1, locs / loc_den);
    }
    if (dlogits == nullptr) return;
    const float gcls = scale * bce_grad(pm, ym) / cls_den;
    for (int u = tid; u < TL; u += 128) {
        float src = source(u);
        int j0 = min((int)src, T - 1), j1 = min(j0 + 1, T - 1);
        float lam = src - (float)j0;
        float z = (1.0f - lam) * lg[j0] + lam * lg[j1];
        float s = cf_sigmoid(z), m = mk[u];
        float p = s * m;
        float dp = scale * bce_grad(p, lb[u]) / loc_den + (u == am ? gcls : 0.f);
        float dz = dp * m * s * (1.0f - s);
        atomicAdd(dl + j0, (1.0f - lam) * dz);
        atomicAdd(dl + j1, lam * dz);
    }
    __syncthreads();
    for (int i = tid; i < T; i += 128) dlogits[(size_t)row * T + i] = dl[i];
}

// fused SGD-momentum over the flat buffers; elements [0, n_split) use lr0, the rest lr1.
// g is consumed and zeroed (the next step's weight-gradient kernels accumulate into it).
__global__ void __launch_bounds__(256) sgd_flat_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ v,
                                                       long long n4, long long split4, float lr0, float lr1, float mu, float wd,
                                                       float gscale) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    const float lr = i < split4 ? lr0 : lr1;
    float4 pp = p[i], gg = g[i], vv = v[i];
    gg.x = fmaf(wd, pp.x, gg.x * gscale); gg.y = fmaf(wd, pp.y, gg.y * gscale);
    gg.z = fmaf(wd, pp.z, gg.z * gscale); gg.w = fmaf(wd, pp.w, gg.w * gscale);
    vv.x = fmaf(mu, vv.x, gg.x); vv.y = fmaf(mu, vv.y, gg.y); vv.z = fmaf(mu, vv.z, gg.z); vv.w = fmaf(mu, vv.w, gg.w);
    pp.x -= lr * vv.x; pp.y -= lr * vv.y; pp.z -= lr * vv.z; pp.w -= lr * vv.w;
    p[i] = pp;
    v[i] = vv;
    g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

extern "C" {

int cf_charades_loss(const float* logits, const float* labels, const float* masks, float* loss2, float* dlogits, int B, int C,
                     int T, int TL, float scale, int align_corners, cudaStream_t stream) {
    CF_CHECK_ARG(logits && labels && masks && loss2, "null pointer");
    CF_CHECK_ARG(B > 0 && C > 0 && T > 0 && TL > 0 && T <= 8192, "bad shape");
    charades_loss_kernel<<<(unsigned)(B * C), 128, (size_t)T * sizeof(float), stream>>>(logits, labels, masks, loss2, dlogits, B,
                                                                                      C, T, TL, scale, align_corners);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

int cf_sgd_flat(float* p, float* g, float* v, int64_t n, int64_t n_split, float lr0, float lr1, float momentum,
                float weight_decay, float grad_scale, cudaStream_t stream) {
    CF_CHECK_ARG(p && g && v && n > 0, "bad argument");
    CF_CHECK_ARG((n & 3) == 0 && (n_split & 3) == 0 && n_split >= 0 && n_split <= n, "n and n_split must be multiples of 4");
    CF_CHECK_ARG(((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)v)) & 15) == 0, "buffers must be 16-byte aligned");
    sgd_flat_kernel<<<(unsigned)cf_cdiv64(n / 4, 256), 256, 0, stream>>>((float4*)p, (float4*)g, (float4*)v, n / 4, n_split / 4,
                                                                       lr0, lr1, momentum, weight_decay, grad_scale);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}

}  // extern "C"
