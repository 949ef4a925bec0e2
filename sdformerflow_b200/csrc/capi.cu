// capi.cu — host-side plumbing of the C-ABI: thread-local error text, launch accounting, neuron
// parameter validation and the shared row-tiling policy.
#include <atomic>
#include <cmath>
#include "sdf_common.cuh"
#include "tc_ptx.cuh"

namespace sdf {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return SDF_ERR_CUDA;
  }
  count_launch();
  return SDF_OK;
}

int validate_neuron(const sdf_neuron_cfg& c) {
  SDF_REQUIRE(c.kind >= SDF_NEURON_LIF && c.kind <= SDF_NEURON_PLIF, "neuron: unknown kind %d", c.kind);
  SDF_REQUIRE(c.surrogate == SDF_SG_ATAN || c.surrogate == SDF_SG_SIGMOID, "neuron: unknown surrogate %d", c.surrogate);
  if (c.kind != SDF_NEURON_IF) SDF_REQUIRE(c.tau >= 1.0, "neuron: tau=%g must be >= 1", c.tau);
  SDF_REQUIRE(std::isfinite(c.v_th), "neuron: v_th must be finite");
  return SDF_OK;
}

NeuronP make_neuron(const sdf_neuron_cfg& c) {
  NeuronP p;
  p.v_th = (float)c.v_th;
  p.v_reset = c.hard_reset ? (float)c.v_reset : 0.f;
  p.tau = (float)c.tau;
  p.inv_tau = 1.f / p.tau;
  p.sg_alpha = (float)c.sg_alpha;
  p.kind = c.kind;
  p.hard = c.hard_reset ? 1 : 0;
  p.detach = c.detach_reset ? 1 : 0;
  p.sg = c.surrogate;
  int e;
  float m = frexpf(p.tau, &e);
  p.tau_pow2 = (m == 0.5f) ? 1 : 0;
  return p;
}

bool make_row_tiling(int64_t rows, int64_t C, int V, int target_threads, int max_blocks, RowTiling* rt) {
  if (C <= 0 || C % V != 0) return false;
  const int64_t vpr = C / V;  // vectors per row
  int64_t ncol = (vpr + target_threads - 1) / target_threads;
  while (ncol <= vpr && (vpr % ncol != 0 || vpr / ncol > target_threads)) ++ncol;
  if (ncol > vpr) return false;
  rt->ncol = (int)ncol;
  rt->R = (int)(vpr / ncol);
  rt->k = target_threads / rt->R;
  if (rt->k < 1) rt->k = 1;
  if ((int64_t)rt->k > rows && rows > 0) rt->k = (int)rows;
  rt->threads = rt->R * rt->k;
  rt->tile_w = (int64_t)rt->R * V;
  int64_t need = (rows + rt->k - 1) / rt->k;
  int64_t cap = max_blocks / ncol;
  if (cap < 1) cap = 1;
  rt->blocks = (int)(need < cap ? (need > 0 ? need : 1) : cap);
  return true;
}

// ---- TMA tensor maps -----------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is fetched through the runtime (cudaGetDriverEntryPoint), so the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

int make_tmap(CUtensorMap* out, int dtype, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("TMA: cuTensorMapEncodeTiled is not available from this driver");
    return SDF_ERR_CUDA;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) gs[i] = strides_bytes[i];
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 12832 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                : swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = fn(out, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                        const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("TMA: cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] stride0 %llu", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return SDF_ERR_CUDA;
  }
  return SDF_OK;
}

int num_sms() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = kNumSMs;
    return v;
  }();
  return n;
}

}  // namespace sdf

extern "C" int sdf_version(void) { return SDF_VERSION_MAJOR * 100 + SDF_VERSION_MINOR; }
extern "C" const char* sdf_last_error(void) { return sdf::g_err; }
extern "C" int64_t sdf_launch_count(void) { return sdf::g_launches.load(std::memory_order_relaxed); }
