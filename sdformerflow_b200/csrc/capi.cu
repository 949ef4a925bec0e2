// capi.cu — host-side plumbing of the C-ABI: thread-local error text, launch accounting, neuron
// parameter validation and the shared row-tiling policy.
#include <atomic>
#include <cmath>
#include "sdf_common.cuh"

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

}  // namespace sdf

extern "C" int sdf_version(void) { return SDF_VERSION_MAJOR * 100 + SDF_VERSION_MINOR; }
extern "C" const char* sdf_last_error(void) { return sdf::g_err; }
extern "C" int64_t sdf_launch_count(void) { return sdf::g_launches.load(std::memory_order_relaxed); }
