// Library-level entry points of the C ABI: version, error string, device info.
#include <cstdlib>
#include <cstdarg>
#include <cstdio>

#include "pp_common.cuh"

namespace {
thread_local char g_error[512] = "";
}

void pp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

namespace {
struct DevInfo {
  int sm_count = 0, major = 0, minor = 0;
  int64_t smem_optin = 0;
  bool ok = false;
};
// one slot per device ordinal; written once (benign race: same values)
DevInfo g_dev[64];

const DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  DevInfo& d = g_dev[dev];
  if (!d.ok) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
    d.sm_count = v;
    cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    d.smem_optin = v;
    d.ok = true;
  }
  return &d;
}
}  // namespace

int pp_sm_count() {
  const DevInfo* d = dev_info();
  return d ? d->sm_count : 148;
}
int64_t pp_smem_optin() {
  const DevInfo* d = dev_info();
  return d ? d->smem_optin : 227 * 1024;
}

int pp_env_int(const char* name, int fallback) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : fallback;
}

int pp_configure_kernel(const void* kernel, int threads, size_t smem, int* ctas_per_sm) {
  struct Entry { const void* fn; int dev, threads; size_t smem; int per_sm; };
  static thread_local Entry cache[32];
  static thread_local int used = 0;
  int dev = 0;
  PP_CUDA_OK(cudaGetDevice(&dev));
  for (int i = 0; i < used; ++i) {
    const Entry& e = cache[i];
    if (e.fn == kernel && e.dev == dev && e.threads == threads && e.smem == smem) {
      *ctas_per_sm = e.per_sm;
      return PP_OK;
    }
  }
  // always opt in to the device maximum: a smaller later request must not shrink the limit that an
  // earlier (cached) configuration of the same kernel relies on
  cudaFuncAttributes fa{};
  PP_CUDA_OK(cudaFuncGetAttributes(&fa, kernel));
  const int64_t dyn_max = pp_smem_optin() - static_cast<int64_t>(fa.sharedSizeBytes);
  PP_REQUIRE(static_cast<int64_t>(smem) <= dyn_max, PP_ERR_UNSUPPORTED_SHAPE,
             "kernel needs %zu bytes of dynamic shared memory, only %lld available", smem, static_cast<long long>(dyn_max));
  PP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn_max)));
  // ask for the largest shared-memory carve-out: the occupancy computed below assumes it, the driver's default
  // heuristic may pick a smaller one (capture Q: 4 resident CTAs where 6 fit)
  PP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  PP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  PP_REQUIRE(per_sm >= 1, PP_ERR_UNSUPPORTED_SHAPE, "kernel does not fit on an SM (threads=%d, smem=%zu)", threads, smem);
  if (used < 32) cache[used++] = Entry{kernel, dev, threads, smem, per_sm};
  *ctas_per_sm = per_sm;
  return PP_OK;
}

extern "C" {

int pp_version(void) { return PP_ABI_VERSION; }

#ifndef PP_SOURCE_HASH
#define PP_SOURCE_HASH "unknown"
#endif
// the tag lets the build script read the hash out of the file without loading the library (an already loaded copy of a
// library cannot be replaced within a process)
static const char kSourceHashTag[] = "pp_source_hash=" PP_SOURCE_HASH ";";
const char* pp_source_hash(void) {
  static thread_local char hash[sizeof(kSourceHashTag)];
  size_t n = 0;
  for (const char* c = kSourceHashTag + 15; *c && *c != ';'; ++c) hash[n++] = *c;
  hash[n] = 0;
  return hash;
}

const char* pp_last_error_string(void) { return g_error; }

int pp_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* smem_optin_bytes) {
  const DevInfo* d = dev_info();
  PP_REQUIRE(d != nullptr, PP_ERR_CUDA, "pp_device_info: no usable CUDA device");
  if (sm_count) *sm_count = d->sm_count;
  if (cc_major) *cc_major = d->major;
  if (cc_minor) *cc_minor = d->minor;
  if (smem_optin_bytes) *smem_optin_bytes = d->smem_optin;
  return PP_OK;
}

}  // extern "C"
