// N1: JPEG decode on the device through nvJPEG, replacing `imageio.imread` in the loader
// (voc12/dataloader.py:189) for throughput runs: file bytes in, uint8 [n,H,W,3] (RGB, interleaved) in HBM out -- the layout
// irn_resize_forward / PseudoLabelPipeline.pyramids consume.  nvJPEG is a LIBRARY (like cuBLAS); it is loaded with dlopen at
// decoder creation, so libirn_b200.so itself does not depend on it.  Not bit-identical to libjpeg-turbo (PIL / imageio): IDCT
// and chroma up-sampling differ by +-1..2 levels, so parity runs keep the host decoder.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <vector>

#include "common.h"

namespace irn {

struct NvJpegApi {
    void* so = nullptr;
    nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t*, nvjpegPinnedAllocator_t*, unsigned int, nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*DecodeBatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
    nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char* const*, const size_t*, nvjpegImage_t*, cudaStream_t) = nullptr;
};

static int load_api(NvJpegApi& api) {
    static const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"};
    for (const char* n : names)
        if ((api.so = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!api.so) return fail(kUnsupported, "nvJPEG library not found (libnvjpeg.so.12): %s", dlerror());
#define IRN_SYM(field, name)                                                                  \
    *(void**)(&api.field) = dlsym(api.so, name);                                              \
    if (!api.field) return fail(kUnsupported, "nvJPEG symbol %s missing", name);
    IRN_SYM(CreateEx, "nvjpegCreateEx")
    IRN_SYM(Destroy, "nvjpegDestroy")
    IRN_SYM(JpegStateCreate, "nvjpegJpegStateCreate")
    IRN_SYM(JpegStateDestroy, "nvjpegJpegStateDestroy")
    IRN_SYM(GetImageInfo, "nvjpegGetImageInfo")
    IRN_SYM(DecodeBatchedInitialize, "nvjpegDecodeBatchedInitialize")
    IRN_SYM(DecodeBatched, "nvjpegDecodeBatched")
#undef IRN_SYM
    return kOk;
}

}  // namespace irn

struct irn_jpeg {
    irn::NvJpegApi api;
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    int backend = 0;        // the backend actually in use: 0 = nvJPEG default (hybrid CPU Huffman + GPU IDCT), 1 = hardware engine
    int batch = 0;          // batch size the state is initialised for
};

using namespace irn;

extern "C" int irn_jpeg_decoder_create(int backend, irn_jpeg** out) {
    if (!out || backend < 0 || backend > 1) return fail(kBadArg, "irn_jpeg_decoder_create: backend must be 0 (default) or 1 (hardware engine)");
    irn_jpeg* d = new irn_jpeg();
    int rc = load_api(d->api);
    if (rc) {
        delete d;
        return rc;
    }
    nvjpegStatus_t st = NVJPEG_STATUS_NOT_INITIALIZED;
    if (backend == 1) {
        st = d->api.CreateEx(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, 0, &d->handle);
        d->backend = 1;
    }
    if (st != NVJPEG_STATUS_SUCCESS) {     // no hardware engine (or not asked for): the default GPU-assisted decoder
        st = d->api.CreateEx(NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, 0, &d->handle);
        d->backend = 0;
    }
    if (st != NVJPEG_STATUS_SUCCESS) {
        dlclose(d->api.so);
        delete d;
        return fail(kCudaError, "nvjpegCreateEx failed (status %d)", (int)st);
    }
    st = d->api.JpegStateCreate(d->handle, &d->state);
    if (st != NVJPEG_STATUS_SUCCESS) {
        d->api.Destroy(d->handle);
        dlclose(d->api.so);
        delete d;
        return fail(kCudaError, "nvjpegJpegStateCreate failed (status %d)", (int)st);
    }
    *out = d;
    return kOk;
}

extern "C" int irn_jpeg_decoder_backend(const irn_jpeg* d) { return d ? d->backend : -1; }

extern "C" void irn_jpeg_decoder_destroy(irn_jpeg* d) {
    if (!d) return;
    if (d->state) d->api.JpegStateDestroy(d->state);
    if (d->handle) d->api.Destroy(d->handle);
    if (d->api.so) dlclose(d->api.so);
    delete d;
}

// Header parse on the host: image height / width (component 0) and component count
extern "C" int irn_jpeg_image_size(irn_jpeg* d, const uint8_t* data, size_t length, int* H, int* W, int* n_components) {
    if (!d || !data || !H || !W) return fail(kBadArg, "irn_jpeg_image_size: bad argument");
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    nvjpegStatus_t st = d->api.GetImageInfo(d->handle, data, length, &nc, &ss, ws, hs);
    if (st != NVJPEG_STATUS_SUCCESS) return fail(kBadArg, "nvjpegGetImageInfo failed (status %d): not a JPEG stream?", (int)st);
    *H = hs[0];
    *W = ws[0];
    if (n_components) *n_components = nc;
    return kOk;
}

// n JPEG streams (host pointers) of identical size H x W -> out_dev uint8 [n,H,W,3] RGB interleaved, enqueued on `stream`
// (nvJPEG's Huffman stage runs on the calling thread before the call returns).
extern "C" int irn_jpeg_decode_batch(irn_jpeg* d, const uint8_t* const* data, const size_t* lengths, int n, uint8_t* out_dev, int H, int W,
                                     irn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d || !data || !lengths || !out_dev || n <= 0 || H <= 0 || W <= 0) return fail(kBadArg, "irn_jpeg_decode_batch: bad argument");
    for (int i = 0; i < n; ++i) {
        int h = 0, w = 0, nc = 0;
        int rc = irn_jpeg_image_size(d, data[i], lengths[i], &h, &w, &nc);
        if (rc) return rc;
        if (h != H || w != W) return fail(kBadArg, "irn_jpeg_decode_batch: image %d is %dx%d, the batch is %dx%d", i, h, w, H, W);
    }
    if (d->batch != n) {
        nvjpegStatus_t st = d->api.DecodeBatchedInitialize(d->handle, d->state, n, 1, NVJPEG_OUTPUT_RGBI);
        if (st != NVJPEG_STATUS_SUCCESS) return fail(kCudaError, "nvjpegDecodeBatchedInitialize(%d) failed (status %d)", n, (int)st);
        d->batch = n;
    }
    std::vector<nvjpegImage_t> dst(n);
    for (int i = 0; i < n; ++i) {
        for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) {
            dst[i].channel[c] = nullptr;
            dst[i].pitch[c] = 0;
        }
        dst[i].channel[0] = out_dev + (size_t)i * H * W * 3;
        dst[i].pitch[0] = (size_t)W * 3;
    }
    nvjpegStatus_t st = d->api.DecodeBatched(d->handle, d->state, data, lengths, dst.data(), stream);
    if (st != NVJPEG_STATUS_SUCCESS && d->backend == 1) {
        // the hardware engine takes baseline single-scan streams only (no 4:1:0 / 4:1:1): retry this and all later batches on the
        // default backend
        d->api.JpegStateDestroy(d->state);
        d->api.Destroy(d->handle);
        d->state = nullptr;
        d->handle = nullptr;
        d->batch = 0;
        d->backend = 0;
        if (d->api.CreateEx(NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, 0, &d->handle) != NVJPEG_STATUS_SUCCESS ||
            d->api.JpegStateCreate(d->handle, &d->state) != NVJPEG_STATUS_SUCCESS)
            return fail(kCudaError, "nvJPEG: hardware decode failed (status %d) and the default backend could not be created", (int)st);
        st = d->api.DecodeBatchedInitialize(d->handle, d->state, n, 1, NVJPEG_OUTPUT_RGBI);
        if (st == NVJPEG_STATUS_SUCCESS) {
            d->batch = n;
            st = d->api.DecodeBatched(d->handle, d->state, data, lengths, dst.data(), stream);
        }
    }
    if (st != NVJPEG_STATUS_SUCCESS) return fail(kCudaError, "nvjpegDecodeBatched failed (status %d)", (int)st);
    return kOk;
}
