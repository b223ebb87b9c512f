// JPEG frames decoded straight into device memory ("GPU preprocessing / decode").
//
// The reference's tracker path turns a video into temporary JPEG files and re-reads every frame with PIL on the host
// (yolo_seg/utils/video_reader.py:57-99); its image-directory mode reads JPEG / PNG files the same way.  Here the
// bitstream is the only thing that crosses PCIe (~0.1-0.3 MB instead of 1.2-6 MB of pixels): nvJPEG decodes it into an
// interleaved BGR device buffer that the device LetterBox / the stem kernel consume directly.  nvJPEG is a library call
// on the I/O side of the path (like cv2.imread was), not one of the hot kernels; it is loaded lazily with dlopen so that
// libypb200.so itself has no link-time dependency on it (a box without libnvjpeg keeps every other entry point).
// NVDEC (H.264/HEVC video) would slot in at the same place; this image ships neither libnvcuvid nor the Video Codec SDK
// headers, so it cannot be built or tested here.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>

#include <cstdio>
#include <cstring>
#include <mutex>

namespace {

struct Api {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  bool ok = false;
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) return;
    a.CreateSimple = reinterpret_cast<decltype(a.CreateSimple)>(dlsym(a.lib, "nvjpegCreateSimple"));
    a.JpegStateCreate = reinterpret_cast<decltype(a.JpegStateCreate)>(dlsym(a.lib, "nvjpegJpegStateCreate"));
    a.GetImageInfo = reinterpret_cast<decltype(a.GetImageInfo)>(dlsym(a.lib, "nvjpegGetImageInfo"));
    a.Decode = reinterpret_cast<decltype(a.Decode)>(dlsym(a.lib, "nvjpegDecode"));
    a.ok = a.CreateSimple && a.JpegStateCreate && a.GetImageInfo && a.Decode;
  });
  return a;
}

struct Decoder {  // one per host thread (nvJPEG states are not thread-safe) and device
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  int device = -1;
};

bool decoder(Decoder** out, char* err, int errlen) {
  static thread_local Decoder d;
  int dev = 0;
  cudaGetDevice(&dev);
  if (d.handle == nullptr || d.device != dev) {
    Api& a = api();
    if (!a.ok) { snprintf(err, errlen, "libnvjpeg not available on this machine"); return false; }
    if (a.CreateSimple(&d.handle) != NVJPEG_STATUS_SUCCESS || a.JpegStateCreate(d.handle, &d.state) != NVJPEG_STATUS_SUCCESS) {
      d.handle = nullptr;
      snprintf(err, errlen, "nvjpegCreateSimple / nvjpegJpegStateCreate failed");
      return false;
    }
    d.device = dev;
  }
  *out = &d;
  return true;
}

}  // namespace

extern "C" int ypb_host_jpeg_info(const unsigned char* data, size_t n, int* h, int* w, char* err, int errlen) {
  Decoder* d;
  if (!decoder(&d, err, errlen)) return -1;
  int comps = 0, widths[NVJPEG_MAX_COMPONENT] = {0}, heights[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t ss;
  const nvjpegStatus_t st = api().GetImageInfo(d->handle, data, n, &comps, &ss, widths, heights);
  if (st != NVJPEG_STATUS_SUCCESS) { snprintf(err, errlen, "nvjpegGetImageInfo failed (%d): not a JPEG bitstream?", (int)st); return -1; }
  *h = heights[0];
  *w = widths[0];
  return 0;
}

// dst_dev: device (H, W, 3) uint8, interleaved BGR (what cv2.imread / cap.read() hand the reference).
extern "C" int ypb_host_jpeg_decode(void* stream, const unsigned char* data, size_t n, unsigned char* dst_dev, int H, int W, char* err,
                                    int errlen) {
  Decoder* d;
  if (!decoder(&d, err, errlen)) return -1;
  int h = 0, w = 0;
  if (ypb_host_jpeg_info(data, n, &h, &w, err, errlen) != 0) return -1;
  if (h != H || w != W) { snprintf(err, errlen, "JPEG is %dx%d, destination is %dx%d", w, h, W, H); return -1; }
  nvjpegImage_t img;
  memset(&img, 0, sizeof img);
  img.channel[0] = dst_dev;
  img.pitch[0] = (size_t)W * 3;
  const nvjpegStatus_t st = api().Decode(d->handle, d->state, data, n, NVJPEG_OUTPUT_BGRI, &img, reinterpret_cast<cudaStream_t>(stream));
  if (st != NVJPEG_STATUS_SUCCESS) { snprintf(err, errlen, "nvjpegDecode failed (%d)", (int)st); return -1; }
  return 0;
}
