// Process-wide pieces of the C ABI: error text, device gate, the host-side commonsense bitmap builder.
#include <mutex>

#include "hc_common.cuh"

namespace hc {
thread_local char g_last_error[512] = {0};

static int g_checked_device = -1;
static int g_check_result = HC_E_ARCH;
static int g_num_sms = 0;
static std::mutex g_mu;

int num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }
}  // namespace hc

using namespace hc;

extern "C" const char* hc_last_error(void) { return g_last_error; }
extern "C" int hc_abi_version(void) { return HC_ABI_VERSION; }

extern "C" int hc_device_check(void) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(HC_E_CUDA, "hc_device_check: no CUDA device (this library has no CPU fallback)");
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (dev == g_checked_device) {
      if (g_check_result != HC_OK) fail(g_check_result, "hc_device_check: device is not sm_100 (compute capability 10.x required)");
      return g_check_result;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return cuda_status("cudaGetDeviceProperties");
    g_checked_device = dev;
    g_num_sms = prop.multiProcessorCount;
    g_check_result = (prop.major == 10) ? HC_OK : HC_E_ARCH;
    if (g_check_result != HC_OK) fail(HC_E_ARCH, "hc_device_check: device is not sm_100 (compute capability 10.x required)");
    return g_check_result;
  }
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
