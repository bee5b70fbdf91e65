// Library plumbing: error string, device info, workspace, copies, casts.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace bf {

static thread_local char g_err[1024] = "";
static thread_local int g_last_path = 0;
static void* g_ws_ptr = nullptr;
static size_t g_ws_bytes = 0;
static unsigned long long g_launches = 0;

void count_launch() { ++g_launches; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

void set_last_path(int p) { g_last_path = p; }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("BF_PDL"); v = (e && e[0] == '1') ? 1 : 0; }   // measured: off is faster (see common.cuh)
  return v != 0;
}
void* workspace_ptr() { return g_ws_ptr; }
size_t workspace_bytes() { return g_ws_bytes; }

int workspace_require(size_t bytes, void** out) {
  if (bytes > g_ws_bytes || g_ws_ptr == nullptr)
    return fail(BF_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu (call bf_set_workspace)", bytes,
                g_ws_bytes);
  *out = g_ws_ptr;
  return BF_OK;
}

__global__ void cast_f64_f32_kernel(const double* __restrict__ s, float* __restrict__ d, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = (float)s[i];
}
__global__ void cast_f32_f64_kernel(const float* __restrict__ s, double* __restrict__ d, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = (double)s[i];
}

// (d0,d1,d2) -> (d1,d0,d2); rows of d2 contiguous floats are moved whole, coalesced on both sides.
__global__ void transpose01_kernel(const float* __restrict__ s, float* __restrict__ d, int d0, int d1, int d2) {
  size_t total = (size_t)d0 * d1 * d2;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c = (int)(i % d2);
    size_t r = i / d2;
    int b = (int)(r % d0);  // destination is (d1, d0, d2): r = a*d0 + b
    int a = (int)(r / d0);
    d[i] = s[((size_t)b * d1 + a) * d2 + c];
  }
}

__global__ void affine_kernel(const float* __restrict__ x, float* __restrict__ d, size_t n, float a, float b) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = fmaf(a, x[i], b);
}

}  // namespace bf

using namespace bf;

extern "C" {

const char* bf_last_error(void) { return g_err; }
int bf_version(void) { return 100; }
int bf_last_path(void) { return g_last_path; }
unsigned long long bf_launch_count(void) { return g_launches; }

int bf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  BF_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  BF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  BF_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  BF_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return BF_OK;
}

int bf_set_workspace(void* d_ptr, size_t bytes) {
  g_ws_ptr = d_ptr;
  g_ws_bytes = d_ptr ? bytes : 0;
  return BF_OK;
}

int bf_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
  BF_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
  return BF_OK;
}
int bf_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
  BF_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  return BF_OK;
}
int bf_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream) {
  BF_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return BF_OK;
}
int bf_stream_sync(void* stream) {
  BF_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return BF_OK;
}

int bf_cast_f64_to_f32(const double* s, float* d, size_t n, void* stream) {
  if (n == 0) return BF_OK;
  cast_f64_f32_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(s, d, n);
  BF_LAUNCH_CHECK();
  return BF_OK;
}
int bf_cast_f32_to_f64(const float* s, double* d, size_t n, void* stream) {
  if (n == 0) return BF_OK;
  cast_f32_f64_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(s, d, n);
  BF_LAUNCH_CHECK();
  return BF_OK;
}
int bf_transpose01(const float* src, float* dst, int d0, int d1, int d2, void* stream) {
  BF_REQUIRE(d0 >= 0 && d1 >= 0 && d2 >= 0, BF_ERR_ARG, "bf_transpose01: negative dimension");
  size_t n = (size_t)d0 * d1 * d2;
  if (n == 0) return BF_OK;
  transpose01_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(src, dst, d0, d1, d2);
  BF_LAUNCH_CHECK();
  return BF_OK;
}
int bf_affine(const float* x, float* dst, size_t n, float a, float b, void* stream) {
  if (n == 0) return BF_OK;
  affine_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(x, dst, n, a, b);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

}  // extern "C"
