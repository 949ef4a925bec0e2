// sdf_common.cuh — shared device/host helpers for libsdf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/sdf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsdf_b200 is written for sm_100a (B200) only"
#endif

namespace sdf {

constexpr int kNumSMs = 148;           // B200: 2 dies x 74 SMs
constexpr int kMaxPartialBlocks = 148 * 4;

// ---- host side: error reporting and launch accounting -------------------------------------
void set_error(const char* fmt, ...);
int finish_launch(const char* what);   // cudaGetLastError -> status, bumps the launch counter
void count_launch();

#define SDF_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) {                                  \
      ::sdf::set_error(__VA_ARGS__);                \
      return SDF_ERR_INVALID_ARG;                   \
    }                                               \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device side: streaming 128-bit access --------------------------------------------------
// ld.global.cs / st.global.cs (evict-first): every hot tensor here is touched once per kernel.
// Intrinsics (not asm volatile) so the compiler is free to hoist independent loads.
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ld_stream2(const float* p) { return __ldcs(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float ld_stream1(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_stream2(float* p, float2 v) { __stcs(reinterpret_cast<float2*>(p), v); }
__device__ __forceinline__ void st_stream1(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) { __stcs(reinterpret_cast<unsigned int*>(p), v); }
__device__ __forceinline__ void st_stream_u64(void* p, uint64_t v) {
  __stcs(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

// ---- spike stores: 4 consecutive spikes (as 0/1 floats in a float4) in the three ABI dtypes -
template <int DT>
__device__ __forceinline__ void store_spike4(void* base, int64_t off, float4 s) {
  if (DT == SDF_SPIKE_F32) {
    st_stream4(reinterpret_cast<float*>(base) + off, s);
  } else if (DT == SDF_SPIKE_U8) {
    uint32_t w = (s.x != 0.f ? 1u : 0u) | (s.y != 0.f ? 0x100u : 0u) | (s.z != 0.f ? 0x10000u : 0u) |
                 (s.w != 0.f ? 0x1000000u : 0u);
    st_stream_u32(reinterpret_cast<uint8_t*>(base) + off, w);
  } else {  // bf16 1.0 = 0x3F80
    uint64_t w = (s.x != 0.f ? 0x3F80ull : 0ull) | (s.y != 0.f ? 0x3F80ull << 16 : 0ull) |
                 (s.z != 0.f ? 0x3F80ull << 32 : 0ull) | (s.w != 0.f ? 0x3F80ull << 48 : 0ull);
    st_stream_u64(reinterpret_cast<uint16_t*>(base) + off, w);
  }
}
template <int DT>
__device__ __forceinline__ void store_spike1(void* base, int64_t off, float s) {
  if (DT == SDF_SPIKE_F32) {
    reinterpret_cast<float*>(base)[off] = s;
  } else if (DT == SDF_SPIKE_U8) {
    reinterpret_cast<uint8_t*>(base)[off] = s != 0.f ? 1 : 0;
  } else {
    reinterpret_cast<uint16_t*>(base)[off] = s != 0.f ? 0x3F80 : 0;
  }
}

// ---- neuron parameters in device form -------------------------------------------------------
struct NeuronP {
  float v_th, v_reset, tau, inv_tau;
  float sg_alpha;
  int kind, hard, detach, sg;
  int tau_pow2;  // 1: x/tau == x*inv_tau exactly
};

NeuronP make_neuron(const sdf_neuron_cfg& c);
int validate_neuron(const sdf_neuron_cfg& c);

// charge: h from (v, x).  Exact op order of spikingjelly: no FMA contraction.
// SIMPLE = LIF, soft reset (v_reset None), tau a power of two, detach_reset, ATan: the configuration of every
// shipped/smoke config; same arithmetic as the generic path with the uniform branches compiled out.
__host__ __device__ __forceinline__ bool neuron_is_simple(const NeuronP& p) {
  return p.kind == SDF_NEURON_LIF && !p.hard && p.tau_pow2 && p.detach && p.sg == SDF_SG_ATAN;
}
template <bool SIMPLE>
__device__ __forceinline__ float neuron_charge_t(const NeuronP& p, float v, float x);
template <bool SIMPLE>
__device__ __forceinline__ float neuron_reset_t(const NeuronP& p, float h, float s);
template <bool SIMPLE>
__device__ __forceinline__ float neuron_grad_h_t(const NeuronP& p, float h, float gs, float gv);

__device__ __forceinline__ float neuron_charge(const NeuronP& p, float v, float x) {
  if (p.kind == SDF_NEURON_IF) return __fadd_rn(v, x);
  float d = (p.hard && p.v_reset != 0.f) ? __fsub_rn(x, __fsub_rn(v, p.v_reset)) : __fsub_rn(x, v);
  if (p.kind == SDF_NEURON_PLIF) return __fadd_rn(v, __fmul_rn(d, p.inv_tau));
  float q = p.tau_pow2 ? __fmul_rn(d, p.inv_tau) : __fdiv_rn(d, p.tau);
  return __fadd_rn(v, q);
}
__device__ __forceinline__ float neuron_fire(const NeuronP& p, float h) {
  return (__fsub_rn(h, p.v_th) >= 0.f) ? 1.f : 0.f;
}
__device__ __forceinline__ float neuron_reset(const NeuronP& p, float h, float s) {
  if (p.hard) return s != 0.f ? p.v_reset : h;
  return __fsub_rn(h, s * p.v_th);
}
// surrogate derivative at z = h - v_th
__device__ __forceinline__ float surrogate_grad(const NeuronP& p, float z) {
  if (p.sg == SDF_SG_ATAN) {
    float c = 1.5707963267948966f * p.sg_alpha;
    float cz = c * z;
    return __fdividef(p.sg_alpha * 0.5f, 1.f + cz * cz);   // gradient path: 2-ulp division is plenty
  }
  float sgax = 1.f / (1.f + __expf(-p.sg_alpha * z));
  return (1.f - sgax) * sgax * p.sg_alpha;
}
// dL/dh given upstream spike grad gs, carried dL/dv (gv); returns gh
__device__ __forceinline__ float neuron_grad_h(const NeuronP& p, float h, float gs, float gv) {
  float z = __fsub_rn(h, p.v_th);
  float sg = surrogate_grad(p, z);
  float s = z >= 0.f ? 1.f : 0.f;
  float dv_dh;
  if (p.hard) {
    dv_dh = (1.f - s);
    if (!p.detach) dv_dh += (p.v_reset - h) * sg;
  } else {
    dv_dh = 1.f;
    if (!p.detach) dv_dh -= p.v_th * sg;
  }
  return gs * sg + gv * dv_dh;
}
template <> __device__ __forceinline__ float neuron_charge_t<false>(const NeuronP& p, float v, float x) { return neuron_charge(p, v, x); }
template <> __device__ __forceinline__ float neuron_charge_t<true>(const NeuronP& p, float v, float x) {
  return __fadd_rn(v, __fmul_rn(__fsub_rn(x, v), p.inv_tau));
}
template <> __device__ __forceinline__ float neuron_reset_t<false>(const NeuronP& p, float h, float s) { return neuron_reset(p, h, s); }
template <> __device__ __forceinline__ float neuron_reset_t<true>(const NeuronP& p, float h, float s) { return __fsub_rn(h, s * p.v_th); }
template <> __device__ __forceinline__ float neuron_grad_h_t<false>(const NeuronP& p, float h, float gs, float gv) {
  return neuron_grad_h(p, h, gs, gv);
}
template <> __device__ __forceinline__ float neuron_grad_h_t<true>(const NeuronP& p, float h, float gs, float gv) {
  const float cz = 1.5707963267948966f * p.sg_alpha * __fsub_rn(h, p.v_th);
  return fmaf(gs, __fdividef(p.sg_alpha * 0.5f, fmaf(cz, cz, 1.f)), gv);
}

__device__ __forceinline__ float neuron_dh_dx(const NeuronP& p) {
  return p.kind == SDF_NEURON_IF ? 1.f : p.inv_tau;
}
__device__ __forceinline__ float neuron_dh_dv(const NeuronP& p) {
  return p.kind == SDF_NEURON_IF ? 1.f : 1.f - p.inv_tau;
}

// ---- 32-bit division by a runtime constant (multiply-high + shift), valid for n < 2^31 ----------
struct FastDiv {
  uint32_t d, mul, sh;
  __host__ void init(uint32_t div) {
    d = div;
    if (div <= 1) { mul = 0; sh = 0; return; }
    uint32_t lg = 0;
    while ((1ull << lg) < div) ++lg;                 // ceil(log2 div)
    const uint32_t pw = 31 + lg;
    mul = (uint32_t)(((1ull << pw) + div - 1) / div);
    sh = pw - 32;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d <= 1 ? n : (__umulhi(n, mul) >> sh); }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

// ---- float4 helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float& f4(float4& v, int i) { return reinterpret_cast<float*>(&v)[i]; }
__device__ __forceinline__ const float& f4(const float4& v, int i) { return reinterpret_cast<const float*>(&v)[i]; }

// ---- row-tiled launch geometry for channels-last [rows, C] problems ------------------------
// A block covers k rows x tile_w channels; thread (ry, rx) owns V consecutive channels
// col0 + rx*V of row ry for the kernel's lifetime, so per-channel quantities (BN scale/shift,
// partial sums) stay in registers.  gridDim.y = ncol column tiles when C is too wide for one
// block; gridDim.x blocks stride over row groups.
struct RowTiling {
  int R;        // threads per row (per column tile)
  int k;        // rows per block iteration
  int ncol;     // column tiles (gridDim.y)
  int threads;  // R * k
  int blocks;   // gridDim.x
  int64_t tile_w;
};
// returns false when C cannot be tiled (C % (V*ncol) != 0)
bool make_row_tiling(int64_t rows, int64_t C, int V, int target_threads, int max_blocks, RowTiling* rt);

// Block-level reduction of per-thread channel accumulators across the k rows of a block; result
// written (not accumulated) to partials[blockIdx.x][slot][col0 + rx*4 ..].  smem: float[threads*4].
template <int NV, int V = 4>
__device__ __forceinline__ void block_reduce_rows_to_partials(float (*acc)[4], float* smem, float* partials,
                                                              int R, int k, int64_t C, int64_t col0) {
  const int rx = threadIdx.x % R;
  const int ry = threadIdx.x / R;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) smem[(ry * R + rx) * V + i] = acc[v][i];
    __syncthreads();
    if (ry == 0) {
      float t[V];
#pragma unroll
      for (int i = 0; i < V; ++i) t[i] = smem[rx * V + i];
      for (int j = 1; j < k; ++j) {
#pragma unroll
        for (int i = 0; i < V; ++i) t[i] += smem[(j * R + rx) * V + i];
      }
      float* dst = partials + ((int64_t)blockIdx.x * NV + v) * C + col0 + rx * V;
#pragma unroll
      for (int i = 0; i < V; ++i) dst[i] = t[i];
    }
  }
}

}  // namespace sdf
