// Average precision per class from score-sorted targets (apmeter.py:98-136 of the reference, the metric of
// train_coarse_fineFEAT.py:241-263 / extract_fineFEAT.py).
//
// The reference loops over the K classes on the CPU: sort the N scores of the class (descending), gather the targets,
// tp = cumsum(truth [* weight]), rg = 1..N (or cumsum(weight)), precision = tp / rg, ap = sum(precision[truth]) /
// max(sum(truth), 1).  Here the sort (all classes at once, class-major [K,N]) is done by the caller; this kernel does the
// scan / divide / masked sum with one CTA per class: coalesced reads, a block-wide prefix sum per 1024-element chunk
// carried across chunks in fp64 (exact for the unweighted counts), the same fp32 division per element as the reference.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

#define AP_THREADS 256
#define AP_ITEMS 4

__global__ void __launch_bounds__(AP_THREADS) ap_sorted_kernel(const float* __restrict__ truth, const float* __restrict__ weight,
                                                              float* __restrict__ ap, int N) {
    __shared__ double wsum[2][AP_THREADS / 32];
    __shared__ double carry[2];
    __shared__ double red[AP_THREADS / 32];
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* tr = truth + (size_t)k * N;
    const float* wr = weight ? weight + (size_t)k * N : nullptr;
    if (tid == 0) { carry[0] = 0.0; carry[1] = 0.0; }
    __syncthreads();
    double acc = 0.0, npos = 0.0;
    for (int base = 0; base < N; base += AP_THREADS * AP_ITEMS) {
        const int i0 = base + tid * AP_ITEMS;
        float t[AP_ITEMS], w[AP_ITEMS];
        double ptp[AP_ITEMS], prg[AP_ITEMS];
        double stp = 0.0, srg = 0.0;
#pragma unroll
        for (int j = 0; j < AP_ITEMS; ++j) {
            const bool v = i0 + j < N;
            t[j] = v ? tr[i0 + j] : 0.f;
            w[j] = v ? (wr ? wr[i0 + j] : 1.f) : 0.f;
            stp += (double)(t[j] * w[j]);                      // weighted_truth = truth * weight (fp32 product, as the reference)
            srg += (double)w[j];
            ptp[j] = stp;
            prg[j] = srg;
        }
        // exclusive prefix of the per-thread sums over the block
        double xtp = stp, xrg = srg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double a = __shfl_up_sync(0xffffffffu, xtp, o), b = __shfl_up_sync(0xffffffffu, xrg, o);
            if (lane >= o) { xtp += a; xrg += b; }
        }
        if (lane == 31) { wsum[0][warp] = xtp; wsum[1][warp] = xrg; }
        __syncthreads();
        double otp = carry[0], org = carry[1];
        for (int q = 0; q < warp; ++q) { otp += wsum[0][q]; org += wsum[1][q]; }
        otp += xtp - stp;
        org += xrg - srg;
#pragma unroll
        for (int j = 0; j < AP_ITEMS; ++j) {
            if (i0 + j < N && t[j] != 0.f) {
                const float tp = (float)(otp + ptp[j]);
                const float rg = wr ? (float)(org + prg[j]) : (float)(i0 + j + 1);
                acc += (double)(tp / rg);
                npos += (double)t[j];
            }
        }
        __syncthreads();
        if (tid == AP_THREADS - 1) { carry[0] = otp + stp; carry[1] = org + srg; }
        __syncthreads();
    }
    // block sums of acc and npos
    acc = warp_sum_d(acc);
    npos = warp_sum_d(npos);
    if (lane == 0) { red[warp] = acc; wsum[0][warp] = npos; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, n = 0.0;
        for (int q = 0; q < AP_THREADS / 32; ++q) { a += red[q]; n += wsum[0][q]; }
        ap[k] = (float)(a / (n > 1.0 ? n : 1.0));
    }
}

extern "C" int cf_ap_sorted(const float* truth_sorted, const float* weight_sorted, float* ap, int N, int K, cudaStream_t stream) {
    CF_CHECK_ARG(truth_sorted && ap && N > 0 && K > 0, "bad arguments");
    ap_sorted_kernel<<<K, AP_THREADS, 0, stream>>>(truth_sorted, weight_sorted, ap, N);
    CF_COUNT_LAUNCH(1);
    CF_CHECK_LAUNCH();
    return CF_OK;
}
