// Shared host/device helpers for the hiercom_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hiercom_b200.h"

namespace hc {

// ---- error plumbing: no exceptions cross the C ABI ------------------------------------
extern thread_local char g_last_error[512];

inline int fail(int code, const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}

inline int cuda_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return HC_E_CUDA;
  }
  return HC_OK;
}

#define HC_REQUIRE(cond, code, msg)                      \
  do {                                                   \
    if (!(cond)) return ::hc::fail((code), (msg));       \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms();

// Per-device opt-ins (cudaFuncSetAttribute for > 48 KB of dynamic shared memory) are keyed on the current device: the attribute
// belongs to the device's context, so a process that uses a second GPU has to set it there too.  One process per GPU is the
// supported deployment; this only keeps a multi-device process from failing with an invalid-argument launch.
constexpr int HC_MAX_DEVICES = 64;
inline int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  return dev;
}

// ---- device helpers -----------------------------------------------------------------------------------
// Python slice-bound semantics of `mask[int(lo):int(hi)]` on an axis of length `size`
// (evaluate.py:115, evaluator.py:86): a negative bound wraps once, then both clamp to [0,size].
__device__ __forceinline__ int slice_bound(int v, int size) {
  if (v < 0) { v += size; if (v < 0) v = 0; }
  return v > size ? size : v;
}

struct Rect { int x0, x1, y0, y1; };

// box = (xmin, xmax, ymin, ymax) already truncated toward zero (int())
__device__ __forceinline__ Rect rect_of(int4 b, int fs) {
  Rect r;
  r.x0 = slice_bound(b.x, fs); r.x1 = slice_bound(b.y, fs);
  r.y0 = slice_bound(b.z, fs); r.y1 = slice_bound(b.w, fs);
  if (r.x1 < r.x0) r.x1 = r.x0;
  if (r.y1 < r.y0) r.y1 = r.y0;
  return r;
}

__device__ __forceinline__ int rect_area(const Rect& r) { return (r.x1 - r.x0) * (r.y1 - r.y0); }

__device__ __forceinline__ int rect_inter(const Rect& a, const Rect& b) {
  int w = min(a.x1, b.x1) - max(a.x0, b.x0);
  int h = min(a.y1, b.y1) - max(a.y0, b.y0);
  return (w > 0 && h > 0) ? w * h : 0;
}

// The pixels of a box's conv2_1 half maps (U / V) that hc_conv2_box_blocks lists - the union of its 8 x bh-pixel blocks, a rectangle;
// everywhere else the map equals the background map bit for bit (3x3 convolution of a map that is tanh(bias) outside the box).
__device__ __forceinline__ Rect conv2_valid_rect(int4 box, int fs, int bh) {
  const Rect r = rect_of(box, fs);
  Rect o = {0, 0, 0, 0};
  if (r.x1 <= r.x0 || r.y1 <= r.y0) return o;
  const int xlo = max(0, r.x0 - 1) & ~1, xhi = min(fs, r.x1 + 1), ylo = max(0, r.y0 - 1) & ~1, yhi = min(fs, r.y1 + 1);
  o.x0 = min(xlo, fs - 8);  o.x1 = min(fs, xlo + 8 * ((xhi - xlo + 7) / 8));          // (a block past the edge is shifted back inside)
  o.y0 = min(ylo, fs - bh); o.y1 = min(fs, ylo + bh * ((yhi - ylo + bh - 1) / bh));
  return o;
}
__device__ __forceinline__ bool rect_has(const Rect& r, int x, int y) { return x >= r.x0 && x < r.x1 && y >= r.y0 && y < r.y1; }

// evaluator.py:84-94: float(intersect)/float(union) >= thresh in Python doubles; 0 when union == 0
__device__ __forceinline__ bool grid_iou_ge(const Rect& a, const Rect& b, double thresh) {
  int inter = rect_inter(a, b);
  int uni = rect_area(a) + rect_area(b) - inter;
  if (uni == 0) return 0.0 >= thresh;
  return (double)inter / (double)uni >= thresh;
}

__device__ __forceinline__ bool bitmap_test(const uint32_t* __restrict__ bm, int key) {
  return (__ldg(bm + (key >> 5)) >> (key & 31)) & 1u;
}

}  // namespace hc
