// Handle lifecycle, error reporting and workspace management of the C ABI
// (include/lyft_voxel.h).
#include <stdarg.h>

#include "lv_common.cuh"

static thread_local char g_err[512] = "";

void lv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int lv_buffer::ensure(size_t bytes, cudaStream_t stream, int fill_byte, bool* grew) {
  if (grew) *grew = false;
  if (bytes <= cap && ptr != nullptr) return LV_OK;
  if (bytes == 0) bytes = 256;
  // growing: wait for anything that may still use the old allocation
  if (ptr) {
    LV_CHECK_CUDA(cudaStreamSynchronize(stream));
    LV_CHECK_CUDA(cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
  }
  size_t want = lv_align_up(bytes, 256);
  cudaError_t e = cudaMalloc(&ptr, want);
  if (e != cudaSuccess) {
    ptr = nullptr;
    lv_set_error("workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    return LV_E_NOMEM;
  }
  cap = want;
  if (fill_byte >= 0) LV_CHECK_CUDA(cudaMemsetAsync(ptr, fill_byte, want, stream));
  if (grew) *grew = true;
  return LV_OK;
}

void lv_buffer::release() {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
}

int lv_mirror::sync(const void* src, size_t bytes, cudaStream_t stream, const void** out) {
  bool same = host.size() == bytes && dev.ptr != nullptr && memcmp(host.data(), src, bytes) == 0;
  if (!same) {
    // `host` is pageable: cudaMemcpyAsync stages it before returning, so the
    // vector may be overwritten by the next call without draining the stream.
    LV_CHECK(dev.ensure(bytes, stream));
    host.assign(static_cast<const unsigned char*>(src), static_cast<const unsigned char*>(src) + bytes);
    LV_CHECK_CUDA(cudaMemcpyAsync(dev.ptr, host.data(), bytes, cudaMemcpyHostToDevice, stream));
  }
  *out = dev.ptr;
  return LV_OK;
}

extern "C" {

int lv_abi_version(void) { return LV_ABI_VERSION; }

const char* lv_version_string(void) { return "lyftvoxel_b200 0.1.0 (sm_100a)"; }

int lv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* lv_last_error(lv_handle*) { return g_err; }

int lv_create(int device, lv_handle** out) {
  if (!out) {
    lv_set_error("lv_create: out is NULL");
    return LV_E_INVALID;
  }
  *out = nullptr;
  int n = lv_device_count();
  if (n <= 0) {
    lv_set_error("lv_create: no CUDA device visible (this library has no CPU fallback)");
    return LV_E_NODEVICE;
  }
  if (device < 0 || device >= n) {
    lv_set_error("lv_create: device %d out of range [0,%d)", device, n);
    return LV_E_INVALID;
  }
  cudaDeviceProp prop;
  LV_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    lv_set_error("lv_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                 device, prop.major, prop.minor);
    return LV_E_NODEVICE;
  }
  LV_CHECK_CUDA(cudaSetDevice(device));
  lv_handle* h = new lv_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    lv_set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
    return LV_E_CUDA;
  }
  *out = h;
  return LV_OK;
}

int lv_destroy(lv_handle* h) {
  if (!h) return LV_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (lv_buffer* b : lv_all_buffers(h)) b->release();
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return LV_OK;
}

int64_t lv_workspace_bytes(lv_handle* h) {
  if (!h) return 0;
  int64_t total = 0;
  for (lv_buffer* b : lv_all_buffers(h)) total += (int64_t)b->cap;
  return total;
}

int64_t lv_launch_count(lv_handle* h) { return h ? h->launches : 0; }

int lv_host_alloc(size_t bytes, void** out) {
  if (!out) {
    lv_set_error("lv_host_alloc: out is NULL");
    return LV_E_INVALID;
  }
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) {
    lv_set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return LV_E_NOMEM;
  }
  return LV_OK;
}

int lv_host_free(void* p) {
  if (!p) return LV_OK;
  LV_CHECK_CUDA(cudaFreeHost(p));
  return LV_OK;
}

int lv_set_option(lv_handle* h, const char* name, int64_t value) {
  if (!h || !name) {
    lv_set_error("lv_set_option: null argument");
    return LV_E_INVALID;
  }
  if (strcmp(name, "bev_frames_in_flight") == 0) {
    h->bev_frames_in_flight = value;
    return LV_OK;
  }
  if (strcmp(name, "bev_tma") == 0) {
    h->bev_tma = value;
    return LV_OK;
  }
  if (strcmp(name, "bev_u16") == 0) {
    h->bev_u16 = value;
    return LV_OK;
  }
  if (strcmp(name, "bev_fused_zero") == 0) {
    h->bev_fused_zero = value;
    return LV_OK;
  }
  if (strcmp(name, "disable_tma") == 0) {
    h->disable_tma = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_frame_kernel") == 0) {
    h->vox_frame_kernel = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_list_path") == 0) {
    h->vox_list_path = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_fused_prologue") == 0) {
    h->vox_fused_prologue = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_generic_rows") == 0) {
    h->vox_generic_rows = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_small_bins") == 0) {
    h->vox_small_bins = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_two_level_scan") == 0) {
    h->vox_two_level_scan = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_hash_map") == 0) {
    h->vox_hash_map = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_rows_waves") == 0) {
    h->vox_rows_waves = value;
    return LV_OK;
  }
  if (strcmp(name, "canvas_variant") == 0) {
    h->canvas_variant = value;
    return LV_OK;
  }
  if (strcmp(name, "vox_dense_map_limit_bytes") == 0) {
    h->vox_dense_map_limit_bytes = value;
    return LV_OK;
  }
  lv_set_error("lv_set_option: unknown option '%s'", name);
  return LV_E_INVALID;
}

}  // extern "C"
