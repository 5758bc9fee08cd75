// hb_ingest.cu — region ingest: JPEG-compressed 4096 x 4096 tiles -> planar uint8 regions in HBM, the
// layout hb_vit256_forward_u8 reads.  Replaces, for the accelerated path, Whole_Slide_Bag_FP.__getitem__
// (datasets/dataset_h5.py:194-207: OpenSlide read_region -> PIL RGB -> ToTensor/Normalize on the CPU, one worker process per
// decode) + collate_features (utils/utils.py:58-61) + the fp32 host -> device copy (hipt_4k.py:69): the compressed bytes cross
// PCIe (a few MB per region instead of 201 MB of fp32), the decoder is nvJPEG's batched GPU decode, and its RGB planes land
// directly in the [R, 3, H, W] uint8 staging buffer of the slide pipeline (ToTensor + Normalize are already folded into the
// patch-embed weights).  No kernel of ours is involved: nvJPEG is library code, bound at run time with dlopen so that
// libhipt_b200.so keeps no link-time dependency beyond libcuda / libstdc++.
#include <dlfcn.h>
#include <new>
#include <stdlib.h>
#include <string.h>

#include <nvjpeg.h>

#include "../../include/hipt_b200.h"
#include "hb_internal.h"

namespace hb {

struct NvjpegApi {
    void* lib = nullptr;
    nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t*, nvjpegPinnedAllocator_t*, unsigned int, nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*DecodeBatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
    nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char* const*, const size_t*, nvjpegImage_t*, cudaStream_t) = nullptr;
};

static NvjpegApi* nvjpeg_api() {
    static NvjpegApi api;
    static int state = 0;                      // 0 = not tried, 1 = ok, -1 = unavailable
    if (state != 0) return state > 0 ? &api : nullptr;
    const char* names[] = {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) { state = -1; return nullptr; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) ok = false; return p; };
    api.CreateEx = reinterpret_cast<decltype(api.CreateEx)>(sym("nvjpegCreateEx"));
    api.Destroy = reinterpret_cast<decltype(api.Destroy)>(sym("nvjpegDestroy"));
    api.JpegStateCreate = reinterpret_cast<decltype(api.JpegStateCreate)>(sym("nvjpegJpegStateCreate"));
    api.JpegStateDestroy = reinterpret_cast<decltype(api.JpegStateDestroy)>(sym("nvjpegJpegStateDestroy"));
    api.GetImageInfo = reinterpret_cast<decltype(api.GetImageInfo)>(sym("nvjpegGetImageInfo"));
    api.DecodeBatchedInitialize = reinterpret_cast<decltype(api.DecodeBatchedInitialize)>(sym("nvjpegDecodeBatchedInitialize"));
    api.DecodeBatched = reinterpret_cast<decltype(api.DecodeBatched)>(sym("nvjpegDecodeBatched"));
    state = ok ? 1 : -1;
    return ok ? &api : nullptr;
}

}  // namespace hb

struct hb_jpeg_decoder {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    int backend = 0;
    int batch = 0;                              // batch size nvjpegDecodeBatchedInitialize was last called with
    int max_batch = 0;
    int cpu_threads = 1;                        // host threads nvJPEG may use for the stages it keeps on the CPU
};

extern "C" {

using namespace hb;

static const char* hb_jpeg_backend_name(int b) {
    switch (b) {
        case NVJPEG_BACKEND_DEFAULT: return "default";
        case NVJPEG_BACKEND_HYBRID: return "hybrid (CPU Huffman)";
        case NVJPEG_BACKEND_GPU_HYBRID: return "gpu_hybrid (GPU Huffman)";
        case NVJPEG_BACKEND_HARDWARE: return "hardware (NVJPG engine)";
    }
    return "unknown";
}

int hb_jpeg_decoder_create(hb_jpeg_decoder** out, int max_batch, int backend) {
    if (!out || max_batch < 1) return set_error("hb_jpeg_decoder_create: bad argument");
    NvjpegApi* api = nvjpeg_api();
    if (!api) return set_error("hb_jpeg_decoder_create: libnvjpeg.so.12 is not loadable (%s)", dlerror());
    hb_jpeg_decoder* d = new (std::nothrow) hb_jpeg_decoder();
    if (!d) return set_error("hb_jpeg_decoder_create: out of host memory");
    // backend < 0: GPU-assisted Huffman first (the recommended backend for large images), then the library default
    const int order_auto[2] = {NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_DEFAULT};
    const int order_one[1] = {backend};
    const int* order = backend < 0 ? order_auto : order_one;
    const int n_order = backend < 0 ? 2 : 1;
    nvjpegStatus_t st = NVJPEG_STATUS_NOT_INITIALIZED;
    for (int i = 0; i < n_order; ++i) {
        st = api->CreateEx(static_cast<nvjpegBackend_t>(order[i]), nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &d->handle);
        if (st == NVJPEG_STATUS_SUCCESS) { d->backend = order[i]; break; }
        d->handle = nullptr;
    }
    if (st != NVJPEG_STATUS_SUCCESS) { delete d; return set_error("hb_jpeg_decoder_create: nvjpegCreateEx failed with status %d", static_cast<int>(st)); }
    st = api->JpegStateCreate(d->handle, &d->state);
    if (st != NVJPEG_STATUS_SUCCESS) {
        api->Destroy(d->handle);
        delete d;
        return set_error("hb_jpeg_decoder_create: nvjpegJpegStateCreate failed with status %d", static_cast<int>(st));
    }
    d->max_batch = max_batch;
    {
        const char* e = getenv("HB_JPEG_CPU_THREADS");
        const int v = e ? atoi(e) : 4;
        d->cpu_threads = v < 1 ? 1 : (v > 64 ? 64 : v);
    }
    *out = d;
    return 0;
}

const char* hb_jpeg_decoder_backend(const hb_jpeg_decoder* d) { return d ? hb_jpeg_backend_name(d->backend) : "none"; }

void hb_jpeg_decoder_destroy(hb_jpeg_decoder* d) {
    if (!d) return;
    NvjpegApi* api = nvjpeg_api();
    if (api) {
        if (d->state) api->JpegStateDestroy(d->state);
        if (d->handle) api->Destroy(d->handle);
    }
    delete d;
}

int hb_jpeg_probe(hb_jpeg_decoder* d, const unsigned char* jpeg, size_t length, int* width, int* height, int* components,
                  int* subsampling) {
    NvjpegApi* api = nvjpeg_api();
    if (!d || !api || !jpeg) return set_error("hb_jpeg_probe: bad argument");
    int nc = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    const nvjpegStatus_t st = api->GetImageInfo(d->handle, jpeg, length, &nc, &ss, w, h);
    if (st != NVJPEG_STATUS_SUCCESS) return set_error("hb_jpeg_probe: nvjpegGetImageInfo failed with status %d", static_cast<int>(st));
    if (width) *width = w[0];
    if (height) *height = h[0];
    if (components) *components = nc;
    if (subsampling) *subsampling = static_cast<int>(ss);
    return 0;
}

int hb_jpeg_decode_tiles(hb_jpeg_decoder* d, const unsigned char* const* jpeg_host, const size_t* lengths, int n,
                         void* regions_u8, int height, int width, int tile_h, int tile_w, void* stream_v) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    NvjpegApi* api = nvjpeg_api();
    if (!d || !api) return set_error("hb_jpeg_decode_tiles: no decoder");
    if (n <= 0) return 0;
    if (n > d->max_batch) return set_error("hb_jpeg_decode_tiles: %d images > decoder batch %d", n, d->max_batch);
    if (!jpeg_host || !lengths || !regions_u8 || height <= 0 || width <= 0 || tile_h <= 0 || tile_w <= 0)
        return set_error("hb_jpeg_decode_tiles: bad argument");
    if (height % tile_h != 0 || width % tile_w != 0)
        return set_error("hb_jpeg_decode_tiles: the %d x %d region is not a whole number of %d x %d tiles", height, width, tile_h, tile_w);
    const int tiles_x = width / tile_w, per_region = (height / tile_h) * tiles_x;
    if (n % per_region != 0) return set_error("hb_jpeg_decode_tiles: %d images are not a whole number of regions (%d tiles each)", n, per_region);
    for (int i = 0; i < n; ++i) {               // every tile must be an RGB image of exactly the tile size
        int nc = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
        nvjpegChromaSubsampling_t ss;
        const nvjpegStatus_t st = api->GetImageInfo(d->handle, jpeg_host[i], lengths[i], &nc, &ss, w, h);
        if (st != NVJPEG_STATUS_SUCCESS) return set_error("hb_jpeg_decode_tiles: image %d is not a decodable JPEG (status %d)", i, static_cast<int>(st));
        if (nc != 3 || w[0] != tile_w || h[0] != tile_h)
            return set_error("hb_jpeg_decode_tiles: image %d is %d x %d with %d components, expected %d x %d RGB", i, w[0], h[0], nc, tile_w, tile_h);
    }
    if (d->batch != n) {
        const nvjpegStatus_t st = api->DecodeBatchedInitialize(d->handle, d->state, n, d->cpu_threads, NVJPEG_OUTPUT_RGB);
        if (st != NVJPEG_STATUS_SUCCESS) return set_error("hb_jpeg_decode_tiles: nvjpegDecodeBatchedInitialize failed with status %d", static_cast<int>(st));
        d->batch = n;
    }
    nvjpegImage_t* dst = static_cast<nvjpegImage_t*>(calloc(static_cast<size_t>(n), sizeof(nvjpegImage_t)));
    if (!dst) return set_error("hb_jpeg_decode_tiles: out of host memory");
    const size_t plane = static_cast<size_t>(height) * width;
    for (int i = 0; i < n; ++i) {
        const int r = i / per_region, t = i - r * per_region, ty = t / tiles_x, tx = t - ty * tiles_x;
        const size_t off = static_cast<size_t>(ty) * tile_h * width + static_cast<size_t>(tx) * tile_w;
        for (int c = 0; c < 3; ++c) {           // planar R, G, B = channels 0..2 of region r of the [R, 3, H, W] batch
            dst[i].channel[c] = static_cast<unsigned char*>(regions_u8) + (static_cast<size_t>(r) * 3 + c) * plane + off;
            dst[i].pitch[c] = static_cast<size_t>(width);
        }
    }
    const nvjpegStatus_t st = api->DecodeBatched(d->handle, d->state, jpeg_host, lengths, dst, stream);
    free(dst);
    if (st != NVJPEG_STATUS_SUCCESS) return set_error("hb_jpeg_decode_tiles: nvjpegDecodeBatched failed with status %d", static_cast<int>(st));
    return 0;
}

}  // extern "C"
