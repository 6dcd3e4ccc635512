// GPU JPEG decode for the loaders: the frames of a clip go from JPEG bytes to ONE uint8 [T,H,W,3]
// device tensor -- exactly what cf_clip_preprocess reads -- through nvJPEG's batched decoder, replacing the reference's
// per-frame PIL decode on the host (charades_fine.py:22-27 pil_loader, :46-56 video_loader, :78-101 load_rgb_frames).
//
// nvJPEG is a CUDA-toolkit library (Huffman decode on its hybrid GPU backend, IDCT / upsampling / colour conversion on
// the GPU); it is loaded with dlopen at decoder creation so that libcfnet_b200.so carries no link-time dependency on it:
// a box without libnvjpeg gets a clear error from cf_jpeg_create (the loaders then keep decoding on the host), never a
// library that fails to load.  The decoder object owns the nvJPEG handle + state; it is NOT thread-safe (one per loader
// thread) and, unlike the rest of this ABI, nvJPEG allocates its own device / pinned scratch memory.
//
// Pixel parity: nvJPEG and PIL's libjpeg-turbo implement the same baseline JPEG standard with different (both conforming)
// IDCT and chroma up-sampling arithmetic, so decoded pixels are NOT bit-identical: measured mean |d| 0.5 grey levels on 4:4:4
// streams and 2-3 on 4:2:0 streams (libjpeg-turbo interpolates the sub-sampled chroma, nvJPEG replicates it);
// tests/test_jpeg_gpu.py states and checks the tolerance against PIL.
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include <dlfcn.h>
#include <nvjpeg.h>
#include <stdlib.h>
#include <string.h>

namespace {

struct NvjpegApi {
    void* so = nullptr;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t*, nvjpegPinnedAllocator_t*, unsigned int, nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*DecodeBatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
    nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char* const*, const size_t*, nvjpegImage_t*,
                                    cudaStream_t) = nullptr;
};

NvjpegApi* load_api() {
    static NvjpegApi api;
    static int state = 0;                 // 0 = not tried, 1 = ok, -1 = unavailable
    if (state == 0) {
        static const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
        for (int i = 0; i < 3 && !api.so; ++i) api.so = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
        state = -1;
        if (api.so) {
#define CF_SYM(field, sym) *(void**)(&api.field) = dlsym(api.so, sym)
            CF_SYM(CreateSimple, "nvjpegCreateSimple");
            CF_SYM(CreateEx, "nvjpegCreateEx");
            CF_SYM(Destroy, "nvjpegDestroy");
            CF_SYM(JpegStateCreate, "nvjpegJpegStateCreate");
            CF_SYM(JpegStateDestroy, "nvjpegJpegStateDestroy");
            CF_SYM(GetImageInfo, "nvjpegGetImageInfo");
            CF_SYM(DecodeBatchedInitialize, "nvjpegDecodeBatchedInitialize");
            CF_SYM(DecodeBatched, "nvjpegDecodeBatched");
#undef CF_SYM
            if (api.CreateSimple && api.Destroy && api.JpegStateCreate && api.JpegStateDestroy && api.GetImageInfo &&
                api.DecodeBatchedInitialize && api.DecodeBatched)
                state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}

struct Decoder {
    NvjpegApi* api;
    nvjpegHandle_t handle;
    nvjpegJpegState_t state;
    int batch;                            // batch size the state is initialised for
};

}  // namespace

extern "C" {

int cf_jpeg_create(void** decoder) {
    CF_CHECK_ARG(decoder, "null pointer");
    *decoder = nullptr;
    NvjpegApi* api = load_api();
    if (!api) {
        cf_set_error("cf_jpeg_create: libnvjpeg.so.12 not found (or incomplete): decode on the host instead");
        return CF_ERR_CUDA;
    }
    Decoder* d = (Decoder*)calloc(1, sizeof(Decoder));
    CF_CHECK_ARG(d, "out of host memory");
    d->api = api;
    nvjpegStatus_t st = NVJPEG_STATUS_NOT_INITIALIZED;
    if (api->CreateEx) st = api->CreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, 0, &d->handle);   // GPU-assisted Huffman decode
    if (st != NVJPEG_STATUS_SUCCESS) st = api->CreateSimple(&d->handle);
    if (st == NVJPEG_STATUS_SUCCESS) st = api->JpegStateCreate(d->handle, &d->state);
    if (st != NVJPEG_STATUS_SUCCESS) {
        cf_set_error("cf_jpeg_create: nvJPEG initialisation failed (status %d)", (int)st);
        free(d);
        return CF_ERR_CUDA;
    }
    *decoder = d;
    return CF_OK;
}

int cf_jpeg_destroy(void* decoder) {
    Decoder* d = (Decoder*)decoder;
    if (!d) return CF_OK;
    d->api->JpegStateDestroy(d->state);
    d->api->Destroy(d->handle);
    free(d);
    return CF_OK;
}

int cf_jpeg_image_info(void* decoder, const unsigned char* data, size_t length, int* height, int* width, int* components) {
    Decoder* d = (Decoder*)decoder;
    CF_CHECK_ARG(d && data && length > 0 && height && width, "bad argument");
    int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    nvjpegStatus_t st = d->api->GetImageInfo(d->handle, data, length, &nc, &ss, ws, hs);
    if (st != NVJPEG_STATUS_SUCCESS) {
        cf_set_error("cf_jpeg_image_info: not a decodable JPEG stream (status %d)", (int)st);
        return CF_ERR_ARG;
    }
    *height = hs[0];
    *width = ws[0];
    if (components) *components = nc;
    return CF_OK;
}

int cf_jpeg_decode_batch(void* decoder, const unsigned char* const* data, const size_t* lengths, int n, unsigned char* out, int H,
                         int W, cudaStream_t stream) {
    Decoder* d = (Decoder*)decoder;
    CF_CHECK_ARG(d && data && lengths && out, "null pointer");
    CF_CHECK_ARG(n > 0 && n <= 65536 && H > 0 && W > 0, "bad shape");
    for (int i = 0; i < n; ++i) {          // every frame must have the size of the output slot it is decoded into
        int h = 0, w = 0;
        if (cf_jpeg_image_info(d, data[i], lengths[i], &h, &w, nullptr) != CF_OK) return CF_ERR_ARG;
        if (h != H || w != W) {
            cf_set_error("cf_jpeg_decode_batch: frame %d is %dx%d, expected %dx%d", i, h, w, H, W);
            return CF_ERR_ARG;
        }
    }
    nvjpegStatus_t st = NVJPEG_STATUS_SUCCESS;
    if (d->batch != n) {
        st = d->api->DecodeBatchedInitialize(d->handle, d->state, n, 1, NVJPEG_OUTPUT_RGBI);
        if (st != NVJPEG_STATUS_SUCCESS) {
            cf_set_error("cf_jpeg_decode_batch: nvjpegDecodeBatchedInitialize failed (status %d)", (int)st);
            return CF_ERR_CUDA;
        }
        d->batch = n;
    }
    nvjpegImage_t* dst = (nvjpegImage_t*)calloc((size_t)n, sizeof(nvjpegImage_t));
    CF_CHECK_ARG(dst, "out of host memory");
    for (int i = 0; i < n; ++i) {          // interleaved RGB: one plane, pitch W*3 -> [n,H,W,3] contiguous
        dst[i].channel[0] = out + (size_t)i * H * W * 3;
        dst[i].pitch[0] = (size_t)W * 3;
    }
    st = d->api->DecodeBatched(d->handle, d->state, data, lengths, dst, stream);
    free(dst);
    if (st != NVJPEG_STATUS_SUCCESS) {
        cf_set_error("cf_jpeg_decode_batch: nvjpegDecodeBatched failed (status %d)", (int)st);
        return CF_ERR_CUDA;
    }
    CF_COUNT_LAUNCH(1);
    return CF_OK;
}

}  // extern "C"
