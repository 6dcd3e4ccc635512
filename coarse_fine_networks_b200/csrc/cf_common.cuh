// Shared device/host helpers for the cfnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CF_OK 0
#define CF_ERR_ARG 1
#define CF_ERR_CUDA 2

extern "C" const char* cf_last_error(void);
void cf_set_error(const char* fmt, ...);

#define CF_CHECK_ARG(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            cf_set_error("%s: %s", __func__, msg);                \
            return CF_ERR_ARG;                                    \
        }                                                         \
    } while (0)

#define CF_CHECK_LAUNCH()                                                        \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            cf_set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e__)); \
            return CF_ERR_CUDA;                                                  \
        }                                                                        \
    } while (0)

// launch counter (read by bench.py -> "gpu_launches")
extern unsigned long long g_cf_launches;
#define CF_COUNT_LAUNCH(n) (g_cf_launches += (unsigned long long)(n))

// "done once" flag for cudaFuncSetAttribute opt-ins: the attribute belongs to the device / context, so the flag is kept per
// device (a process that touches a second GPU opts in there too) and is safe to race on from several host threads.
struct CfOncePerDevice {
    unsigned long long mask = 0;
    bool need() const {
        int d = 0;
        cudaGetDevice(&d);
        return !((__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> (d & 63)) & 1ull);
    }
    void mark() {
        int d = 0;
        cudaGetDevice(&d);
        __atomic_fetch_or(&mask, 1ull << (d & 63), __ATOMIC_RELEASE);
    }
};

// Experiment switches (environment variables selecting a replaced kernel variant for same-box A/B runs) exist only in the
// -DCFNET_AB build (python -m coarse_fine_networks_b200.build --ab); the shipped library reads no environment.
#ifdef CFNET_AB
#include <stdlib.h>
static inline int cf_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
#else
static inline int cf_env(const char*, int dflt) { return dflt; }
#endif

static inline int cf_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long cf_cdiv64(long long a, long long b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// Programmatic dependent launch.  Every kernel launched through cf_launch() starts with cf_pdl_enter(): it waits until the
// grid before it in the stream has completed and its writes are visible (griddepcontrol.wait: a no-op without a programmatic
// edge), so the data dependence is the stream's usual one; the launch itself travels the programmatic edge (recorded as such
// by stream capture), which is cheaper than a full dependency between the ~1 200 kernels of a step.  The one-wave table kernels
// (BatchNorm tables, SE) additionally let the grid AFTER them be scheduled while they run (cf_pdl_enter_early): its CTAs are
// resident and parked at their own wait when the tables are ready.  Doing the same in the large kernels LOST 3.6 ms per step
// (profiles/r02_ab_same_box.md section 14), so they do not trigger early.
__device__ __forceinline__ void cf_pdl_enter() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void cf_pdl_enter_early() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifndef CFNET_PDL_NOTRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
static inline cudaError_t cf_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cf_env("CFNET_PDL", 1) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float cf_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Packed fp32x2 FMA (sm_100 FFMA2): c += a * b on both halves in ONE issue slot.  On Blackwell a scalar FFMA warp
// instruction occupies the FMA pipe for two cycles per SM sub-partition (64 FMA/clk/SM); only the packed form reaches the
// full 128 FMA/clk/SM, and the depthwise 3x3x3 stencils (27 FMA per output) are FMA-issue-bound at the scalar rate.
__device__ __forceinline__ void ffma2(float2& c, const float2 a, const float2 b) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(c.x), "+f"(c.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

// 128-bit streaming accessors: inputs read once go through the read-only path, outputs
// written once bypass L1 so they do not evict the small tables / reused frames.
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ void stcs4(float4* p, float4 v) { __stcs(p, v); }

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_axpby(float a, float4 x, float b, float4 y) {
    return make_float4(a * x.x + b * y.x, a * x.y + b * y.y, a * x.z + b * y.z, a * x.w + b * y.w);
}
__device__ __forceinline__ void f4_fma(float4& acc, float a, float4 x) {
    acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y);
    acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
}
#endif
