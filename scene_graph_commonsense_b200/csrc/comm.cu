// Cross-rank reduction of the counter vector and the buffer-sizing queries of the C ABI.
//
// hc_counts_allreduce: ONE ncclAllReduce(sum, int64) of the 765-slot counter vector per evaluation window, in place, on the
// caller's stream - the only collective of the path (images shard across GPUs; R@K / mR@K are then computed on every rank from
// identical integers).  The reference has no such step: every rank writes its own JSON (utils.py:486).
// NCCL is bound at run time (dlopen): the library already loaded into the process is reused (the communicator handed in was
// created by it), so libhiercom_b200.so has no link-time NCCL dependency and loads on a box without NCCL.
#include <dlfcn.h>

#include <mutex>

#include "hc_common.cuh"

namespace hc {
namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: ncclResult_t 0 == success, ncclInt64 == 4, ncclSum == 0, 128-byte id)
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*GetUniqueIdFn)(void*);
typedef int (*CommInitRankFn)(void**, int, HcNcclId, int);
typedef int (*CommDestroyFn)(void*);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
  void* lib = nullptr;
  AllReduceFn all_reduce = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
};

Nccl g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return HC_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy the host framework already loaded, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(HC_E_CUDA, "hc_counts_allreduce: libnccl.so.2 is not loadable in this process");
  Nccl n;
  n.lib = h;
  n.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
  n.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
  n.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
  n.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
  n.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
  if (!n.all_reduce || !n.get_unique_id || !n.comm_init_rank || !n.comm_destroy)
    return fail(HC_E_CUDA, "hc_counts_allreduce: libnccl.so.2 lacks ncclAllReduce / ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy");
  g_nccl = n;
  return HC_OK;
}

int nccl_status(int r, const char* what) {
  if (r == 0) return HC_OK;
  snprintf(g_last_error, sizeof(g_last_error), "%s: NCCL error %d (%s)", what, r, g_nccl.error_string ? g_nccl.error_string(r) : "?");
  return HC_E_CUDA;
}

}  // namespace
}  // namespace hc

using namespace hc;

extern "C" int hc_counts_allreduce(void* nccl_comm, int64_t* counters, int64_t n, hc_stream_t stream_) {
  HC_REQUIRE(nccl_comm && counters, HC_E_NULL, "hc_counts_allreduce: communicator and counters must be non-NULL");
  HC_REQUIRE(n > 0, HC_E_SHAPE, "hc_counts_allreduce: n must be positive");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  rc = load_nccl();
  if (rc != HC_OK) return rc;
  return nccl_status(g_nccl.all_reduce(counters, counters, (size_t)n, /*ncclInt64*/ 4, /*ncclSum*/ 0, nccl_comm,
                                       reinterpret_cast<cudaStream_t>(stream_)),
                     "hc_counts_allreduce");
}

extern "C" int hc_nccl_unique_id(HcNcclId* id_out) {
  HC_REQUIRE(id_out, HC_E_NULL, "hc_nccl_unique_id: id_out is NULL");
  int rc = load_nccl();
  if (rc != HC_OK) return rc;
  return nccl_status(g_nccl.get_unique_id(id_out), "hc_nccl_unique_id");
}

extern "C" int hc_nccl_comm_create(const HcNcclId* id, int32_t n_ranks, int32_t rank, void** comm_out) {
  HC_REQUIRE(id && comm_out, HC_E_NULL, "hc_nccl_comm_create: NULL pointer");
  HC_REQUIRE(n_ranks > 0 && rank >= 0 && rank < n_ranks, HC_E_SHAPE, "hc_nccl_comm_create: bad rank / n_ranks");
  int rc = hc_device_check();
  if (rc != HC_OK) return rc;
  rc = load_nccl();
  if (rc != HC_OK) return rc;
  return nccl_status(g_nccl.comm_init_rank(comm_out, n_ranks, *id, rank), "hc_nccl_comm_create");
}

extern "C" int hc_nccl_comm_destroy(void* comm) {
  if (!comm) return HC_OK;
  int rc = load_nccl();
  if (rc != HC_OK) return rc;
  return nccl_status(g_nccl.comm_destroy(comm), "hc_nccl_comm_destroy");
}

// ------------------------------------------------------------------------------------------------ sizing queries (host, no GPU)
extern "C" int64_t hc_pairs_enumerate_workspace_bytes(int64_t sum_tri, int32_t n_images, int32_t n_groups, int32_t max_tri,
                                                      int64_t* ws_ov_bytes, int64_t* ws_any_bytes, int64_t* ws_counts_bytes) {
  if (sum_tri < 0 || n_images < 0 || n_groups < 0 || max_tri < 0) return fail(HC_E_SHAPE, "hc_pairs_enumerate_workspace_bytes: negative size");
  const int64_t a = sum_tri > 0 ? sum_tri : 1, b = (int64_t)n_groups * max_tri > 0 ? (int64_t)n_groups * max_tri : 1,
                c = (int64_t)(n_images > 0 ? n_images : 1) * 4;
  if (ws_ov_bytes) *ws_ov_bytes = a;
  if (ws_any_bytes) *ws_any_bytes = n_groups > 0 ? b : 0;
  if (ws_counts_bytes) *ws_counts_bytes = c;
  return a + (n_groups > 0 ? b : 0) + c;
}

extern "C" int64_t hc_conv3_blocks_capacity(int64_t n_pairs, int32_t block_rows, int32_t block_cols) {
  const int bc = block_cols ? block_cols : 8;
  if (n_pairs < 0 || !(block_rows == 8 || block_rows == 4) || !(bc == 8 || (bc == 4 && block_rows == 4)))
    return fail(HC_E_SHAPE, "hc_conv3_blocks_capacity: blocks are 8x8, 8x4 or 4x4 conv3 pixels");
  const int64_t c = n_pairs * (256 / (block_rows * bc));
  return c > 0 ? c : 1;
}

extern "C" int64_t hc_conv2_box_blocks_capacity(int64_t n_box, int32_t block_rows) {
  if (n_box < 0 || !(block_rows == 8 || block_rows == 4)) return fail(HC_E_SHAPE, "hc_conv2_box_blocks_capacity: block_rows must be 8 or 4");
  const int64_t c = n_box * 4 * (32 / block_rows);
  return c > 0 ? c : 1;
}

// Device bytes of every buffer the batched relation path allocates for one window (the sizes RelationPipeline uses), so a host that
// owns the allocator can reserve them up front.  All activations are 16-bit; feature_size 32, 256 conv1 channels, 512 / 1024
// conv2 / conv3 channels, fc1 4096, fc2 512 (model.py:116-133).
extern "C" int hc_relation_workspace_bytes(int32_t n_images, int64_t n_box, int64_t n_pairs, int64_t chunk_pairs, int32_t shared_fc1,
                                           hc_relation_workspace* out) {
  HC_REQUIRE(out, HC_E_NULL, "hc_relation_workspace_bytes: out is NULL");
  HC_REQUIRE(n_images >= 0 && n_box >= 0 && n_pairs >= 0 && chunk_pairs > 0, HC_E_SHAPE, "hc_relation_workspace_bytes: bad sizes");
  const int64_t fs = 32, px = fs * fs;
  const int64_t nbx = n_box + 1;                                       // + the empty box of the shared-footprint maps
  memset(out, 0, sizeof(*out));
  out->pixels_packed = (int64_t)n_images * px * 320 * 2;               // [B*1024, 320] 16-bit conv1 operand
  out->conv1_out = (int64_t)n_images * px * 256 * 2;                   // T [B,32,32,256]
  out->box_select = nbx * px * 256 * 2;                                // Abox [nbox+1,32,32,256]
  out->conv2_halves = 2 * nbx * px * 512 * 2;                          // U, V [nbox+1,32,32,512]
  const int64_t chunk = n_pairs < chunk_pairs ? (n_pairs > 0 ? n_pairs : 1) : chunk_pairs;
  const int64_t n_buf = n_pairs > chunk ? 2 : 1;                       // pooling of chunk k+1 under the GEMMs of chunk k
  out->pooled_conv2 = n_buf * chunk * 256 * 512 * 2;                   // P2 [chunk,16,16,512]
  out->work_lists = n_buf * (chunk * 16 * 4 + chunk * 8) + 2 * nbx * 16 * 4;
  if (shared_fc1) {
    // (box, empty), (empty, box), background maps + the per-box difference operand map - background of the K-cell-sparse fc1 rows
    // (sorted rows, only visited cells touched) + the pooled conv2 buffer of the per-box pass
    out->box_maps = (2 * n_box + 1) * 64 * 1024 * 2 + 2 * n_box * 64 * 1024 * 2 + (2 * n_box < chunk ? 2 * n_box : chunk) * 256 * 512 * 2;
    out->box_fc1_rows = (2 * n_box + 1) * 4096 * 4 + 2 * n_box * (4 + 4 + 4 + 8) + ((2 * n_box + 255) / 256) * 8;
    out->fc1_operand = n_pairs * 64 * 1024 * 2;                        // D [P, 64 cells, 1024]: only visited cells are touched
    out->row_maps = n_pairs * (4 + 4 + 4 + 4 + 8) + ((n_pairs + 255) / 256) * 8;
  } else {
    out->pooled_conv3 = n_buf * chunk * 64 * 1024 * 2;                 // P3 [chunk,8,8,1024]
  }
  out->fc1_out = (shared_fc1 ? n_pairs : chunk) * 4096 * 2;            // H1
  out->fc2_raw = n_pairs * 512 * 4;
  out->head_out = n_pairs * (50 + 3 + 1 + 1) * 4 + n_box * 2 * 512 * 4;
  out->candidates = n_pairs * 3 * (4 + 4) + n_pairs * (4 + 4);
  out->total = out->pixels_packed + out->conv1_out + out->box_select + out->conv2_halves + out->pooled_conv2 + out->work_lists +
               out->box_maps + out->box_fc1_rows + out->fc1_operand + out->row_maps + out->pooled_conv3 + out->fc1_out + out->fc2_raw +
               out->head_out + out->candidates;
  return HC_OK;
}
