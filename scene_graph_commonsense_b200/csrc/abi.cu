// Process-wide pieces of the C ABI: error text, device gate, the host-side commonsense bitmap builder.
#include <mutex>

#include "hc_common.cuh"

namespace hc {
thread_local char g_last_error[512] = {0};

static int g_check_result[HC_MAX_DEVICES];      // 0 = not checked yet, 1 = sm_100, 2 = other architecture
static int g_sms[HC_MAX_DEVICES];
static std::mutex g_mu;

// SM count of the CURRENT device (grid sizing of the persistent kernels); 148 before the first hc_device_check on it
int num_sms() {
  const int dev = current_device();
  return (dev >= 0 && dev < HC_MAX_DEVICES && g_sms[dev] > 0) ? g_sms[dev] : 148;
}
}  // namespace hc

using namespace hc;

extern "C" const char* hc_last_error(void) { return g_last_error; }
extern "C" int hc_abi_version(void) { return HC_ABI_VERSION; }

extern "C" int hc_device_check(void) {
  const int dev = current_device();
  if (dev < 0 || dev >= HC_MAX_DEVICES) return fail(HC_E_CUDA, "hc_device_check: no CUDA device (this library has no CPU fallback)");
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_check_result[dev] == 0) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return cuda_status("cudaGetDeviceProperties");
    g_sms[dev] = prop.multiProcessorCount;
    g_check_result[dev] = (prop.major == 10) ? 1 : 2;
  }
  if (g_check_result[dev] != 1) return fail(HC_E_ARCH, "hc_device_check: device is not sm_100 (compute capability 10.x required)");
  return HC_OK;
}

extern "C" int hc_cs_bitmap_build(const int64_t* aligned_keys, int64_t n_aligned, const int64_t* violated_keys, int64_t n_violated,
                                  uint32_t* bitmap_out) {
  HC_REQUIRE(bitmap_out != nullptr, HC_E_NULL, "hc_cs_bitmap_build: bitmap_out is NULL");
  HC_REQUIRE((n_aligned == 0 || aligned_keys) && (n_violated == 0 || violated_keys), HC_E_NULL, "hc_cs_bitmap_build: key array is NULL");
  memset(bitmap_out, 0, sizeof(uint32_t) * HC_BITMAP_WORDS);
  for (int64_t i = 0; i < n_aligned; ++i) {
    int64_t k = aligned_keys[i];
    HC_REQUIRE(k >= 0 && k < HC_TRIPLET_SPACE, HC_E_SHAPE, "hc_cs_bitmap_build: aligned key out of range");
    bitmap_out[k >> 5] |= 1u << (k & 31);
  }
  for (int64_t i = 0; i < n_violated; ++i) {
    int64_t k = violated_keys[i];
    HC_REQUIRE(k >= 0 && k < HC_TRIPLET_SPACE, HC_E_SHAPE, "hc_cs_bitmap_build: violated key out of range");
    bitmap_out[k >> 5] &= ~(1u << (k & 31));
  }
  return HC_OK;
}
