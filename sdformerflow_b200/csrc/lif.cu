// lif.cu — K1/K2: multi-step LIF / IF / PLIF forward and surrogate-gradient backward, and the
// PSN forward/backward (K1p), each with the preceding BatchNorm-apply folded into the prologue.
//
// Replaces spikingjelly's LIFNode.multi_step_forward (python loop of ~10*T ATen kernels, or the
// cupy LIFNodeFPTT/BPTT kernels) reached through Spiking_neuron.forward
// (reference models/STSwinNet_SNN/Spiking_modules.py:26-99) and PSN.forward
// (models/STSwinNet_SNN/Spiking_submodules.py:207-211).
//
// Design: HBM-bound streaming kernels.  One thread owns V adjacent neurons (V = 4 -> 128-bit
// accesses), issues all T loads up front (T independent requests in flight per thread), walks the
// T time bins with the membrane potential in registers and streams the spikes out.  Nothing but
// the input is saved for backward: K2 re-reads u, recomputes h_t and walks the adjoint
// recurrence in registers, emitting the per-channel BatchNorm partial sums on the way.
// Algorithmic traffic per neuron-timestep: fwd 4 B in + {4,1,2} B out; bwd 12 B.
#include <cstdlib>
#include "sdf_common.cuh"

namespace sdf {

// ---- V-wide vector access -------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&r)[V]) {
  if (V == 4) {
    float4 t = ld_stream4(p);
    r[0] = t.x; r[V > 1 ? 1 : 0] = t.y; r[V > 2 ? 2 : 0] = t.z; r[V > 3 ? 3 : 0] = t.w;
  } else if (V == 2) {
    float2 t = ld_stream2(p);
    r[0] = t.x; r[V > 1 ? 1 : 0] = t.y;
  } else {
    r[0] = ld_stream1(p);
  }
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float (&r)[V]) {
  if (V == 4) {
    st_stream4(p, make_float4(r[0], r[V > 1 ? 1 : 0], r[V > 2 ? 2 : 0], r[V > 3 ? 3 : 0]));
  } else if (V == 2) {
    st_stream2(p, make_float2(r[0], r[V > 1 ? 1 : 0]));
  } else {
    st_stream1(p, r[0]);
  }
}
template <int DT, int V>
__device__ __forceinline__ void store_spikes(void* base, int64_t off, const float (&s)[V]) {
  if (V == 4) {
    store_spike4<DT>(base, off, make_float4(s[0], s[V > 1 ? 1 : 0], s[V > 2 ? 2 : 0], s[V > 3 ? 3 : 0]));
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) store_spike1<DT>(base, off + i, s[i]);
  }
}

// ---- sequence addressing + tiling (device form of sdf_seq_layout) --------------------------
struct SeqP {
  int64_t n_neurons, inner, stride_b, stride_t;
  int64_t row_w;    // neurons per tile row: C in channel-fixed mode, tile_w otherwise
  int64_t tile_w;   // neurons per block-row: R * V
  int64_t n_rows;   // ceil(n_neurons / row_w)
  int64_t C, hw;
  int R, k;         // threads per row, rows per block
  int chan_mode;    // 0: no affine, 1: channel fixed per thread (registers), 2: channels-last per-iteration, 3: NCHW per-iteration
  int T;            // runtime T (generic kernels)
};

__device__ __forceinline__ int64_t seq_base(const SeqP& s, int64_t n) {
  if (s.stride_b == 0) return n;
  int64_t b = n / s.inner;
  return b * s.stride_b + (n - b * s.inner);
}

template <int V>
__device__ __forceinline__ void load_affine(const SeqP& s, const float* scale, const float* shift, int64_t n,
                                            float (&sc)[V], float (&sh)[V]) {
  if (s.chan_mode == 2) {
    int64_t c = n % s.C;
#pragma unroll
    for (int i = 0; i < V; ++i) { sc[i] = __ldg(scale + c + i); sh[i] = __ldg(shift + c + i); }
  } else if (s.chan_mode == 3) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      int64_t c = ((n + i) / s.hw) % s.C;
      sc[i] = __ldg(scale + c); sh[i] = __ldg(shift + c);
    }
  }
}

template <int V>
__device__ __forceinline__ void init_affine(const SeqP& s, const float* scale, const float* shift, int64_t col,
                                            float (&sc)[V], float (&sh)[V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) { sc[i] = 1.f; sh[i] = 0.f; }
  if (s.chan_mode == 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) { sc[i] = scale[col + i]; sh[i] = shift[col + i]; }
  }
}

// ---- K1: forward ------------------------------------------------------------------------------
struct LifFwdP {
  const float* u; void* spike; float* h_seq; const float* v_init; float* v_final;
  const float* scale; const float* shift;
  SeqP s; NeuronP nrn;
};

// T > 0: compile-time step count (fully unrolled).  T == 0: runtime T <= 32.
template <int T, int V, int DT, bool SIMPLE>
__global__ void __launch_bounds__(512, (T > 0 && T <= 10 && DT != SDF_SPIKE_F32) ? 2 : 1) lif_fwd_kernel(const LifFwdP p) {
  constexpr int TM = T > 0 ? T : 32;
  const SeqP& s = p.s;
  const NeuronP nrn = p.nrn;
  const int Tn = T > 0 ? T : s.T;
  const int rx = threadIdx.x % s.R, ry = threadIdx.x / s.R;
  const int64_t col = (int64_t)blockIdx.y * s.tile_w + (int64_t)rx * V;
  float sc[V], sh[V];
  init_affine<V>(s, p.scale, p.shift, col, sc, sh);
  for (int64_t row = (int64_t)blockIdx.x * s.k + ry; row < s.n_rows; row += (int64_t)gridDim.x * s.k) {
    const int64_t n = row * s.row_w + col;
    if (n >= s.n_neurons) continue;
    const int64_t off = seq_base(s, n);
    float x[TM][V];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) ldv<V>(p.u + off + t * s.stride_t, x[t]);
    if (s.chan_mode >= 2) load_affine<V>(s, p.scale, p.shift, n, sc, sh);
    float v[V];
    if (p.v_init) {
      ldv<V>(p.v_init + n, v);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = nrn.hard ? nrn.v_reset : 0.f;
    }
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (t < Tn) {
        float h[V], sp[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float xx = fmaf(x[t][i], sc[i], sh[i]);
          h[i] = neuron_charge_t<SIMPLE>(nrn, v[i], xx);
          sp[i] = neuron_fire(nrn, h[i]);
          v[i] = neuron_reset_t<SIMPLE>(nrn, h[i], sp[i]);
        }
        store_spikes<DT, V>(p.spike, off + t * s.stride_t, sp);
        if (p.h_seq) stv<V>(p.h_seq + off + t * s.stride_t, h);
      }
    }
    if (p.v_final) stv<V>(p.v_final + n, v);
  }
}

// ---- K2: backward -----------------------------------------------------------------------------
struct LifBwdP {
  const float* u; const float* gs; float* gu; float* gx; const float* v_init;
  const float* scale; const float* shift;
  const float* coef;      // optional [3, C]: BatchNorm backward folded in, gu = a[c]*dx + b[c]*u + c0[c] (channels-last only)
  float* bn_partials; float* plif_partials;
  SeqP s; NeuronP nrn;
};

// PRE = 1: all grad loads issued up front (max loads in flight, ~190 regs, 1 CTA/SM);
// PRE = 0: grad loads streamed inside the adjoint loop (128 regs, 2 CTAs/SM).
template <int T, int V, int PRE, bool SIMPLE>
__global__ void __launch_bounds__(288, (T >= 16) ? 1 : ((V == 2) ? (PRE ? 2 : 3) : (PRE ? 1 : 2))) lif_bwd_kernel(const LifBwdP p) {
  constexpr int TM = T > 0 ? T : 32;
  extern __shared__ float smem[];
  const SeqP& s = p.s;
  const NeuronP nrn = p.nrn;
  const int Tn = T > 0 ? T : s.T;
  const int rx = threadIdx.x % s.R, ry = threadIdx.x / s.R;
  const int64_t col = (int64_t)blockIdx.y * s.tile_w + (int64_t)rx * V;
  float sc[V], sh[V];
  init_affine<V>(s, p.scale, p.shift, col, sc, sh);
  float ca[V], cb[V], cc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { ca[i] = 0.f; cb[i] = 0.f; cc[i] = 0.f; }
  if (p.coef) {
#pragma unroll
    for (int i = 0; i < V; ++i) { ca[i] = p.coef[col + i]; cb[i] = p.coef[s.C + col + i]; cc[i] = p.coef[2 * s.C + col + i]; }
  }
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float plif_acc = 0.f;
  const float dh_dx = neuron_dh_dx(nrn), dh_dv = neuron_dh_dv(nrn);
  const float vr = nrn.hard ? nrn.v_reset : 0.f;
  for (int64_t row = (int64_t)blockIdx.x * s.k + ry; row < s.n_rows; row += (int64_t)gridDim.x * s.k) {
    const int64_t n = row * s.row_w + col;
    if (n >= s.n_neurons) continue;
    const int64_t off = seq_base(s, n);
    float u[TM][V], h[TM][V], gpre[PRE ? TM : 1][V];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) ldv<V>(p.u + off + t * s.stride_t, u[t]);
    if (PRE) {
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (t < Tn) ldv<V>(p.gs + off + t * s.stride_t, gpre[PRE ? t : 0]);
    }
    if (s.chan_mode >= 2) load_affine<V>(s, p.scale, p.shift, n, sc, sh);
    float v0[V], v[V];
    if (p.v_init) {
      ldv<V>(p.v_init + n, v0);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) v0[i] = vr;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = v0[i];
    // forward recompute of h_t (exactly the forward arithmetic)
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (t < Tn) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float xx = fmaf(u[t][i], sc[i], sh[i]);
          h[t][i] = neuron_charge_t<SIMPLE>(nrn, v[i], xx);
          v[i] = neuron_reset_t<SIMPLE>(nrn, h[t][i], neuron_fire(nrn, h[t][i]));
        }
      }
    }
    // adjoint recurrence, t = T-1 .. 0
    float gv[V];
#pragma unroll
    for (int i = 0; i < V; ++i) gv[i] = 0.f;
#pragma unroll
    for (int t = TM - 1; t >= 0; --t) {
      if (t < Tn) {
        float dx[V], du[V], g[V];
        if (PRE) {
#pragma unroll
          for (int i = 0; i < V; ++i) g[i] = gpre[PRE ? t : 0][i];
        } else {
          ldv<V>(p.gs + off + t * s.stride_t, g);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float gh = neuron_grad_h_t<SIMPLE>(nrn, h[t][i], g[i], gv[i]);
          dx[i] = gh * dh_dx;
          gv[i] = gh * dh_dv;
          du[i] = p.coef ? fmaf(ca[i], dx[i], fmaf(cb[i], u[t][i], cc[i])) : dx[i] * sc[i];
          if (s.chan_mode == 1) {
            acc[0][i] += dx[i];
            acc[1][i] += dx[i] * u[t][i];
          }
          if (nrn.kind == SDF_NEURON_PLIF) {
            float vp = t > 0 ? neuron_reset(nrn, h[t > 0 ? t - 1 : 0][i], neuron_fire(nrn, h[t > 0 ? t - 1 : 0][i])) : v0[i];
            float xx = fmaf(u[t][i], sc[i], sh[i]);
            plif_acc += gh * (xx - (vp - vr));
          }
        }
        if (p.gu) stv<V>(p.gu + off + t * s.stride_t, du);
        if (p.gx) stv<V>(p.gx + off + t * s.stride_t, dx);
      }
    }
  }
  if (p.bn_partials && s.chan_mode == 1 && V > 1)
    block_reduce_rows_to_partials<2, V>(acc, smem, p.bn_partials, s.R, s.k, s.C, (int64_t)blockIdx.y * s.tile_w);
  if (p.plif_partials) {
    __syncthreads();
    float w = plif_acc;
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)((blockDim.x + 31) / 32); ++i) t += smem[i];
      p.plif_partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
  }
}

// ---- K2s: backward with the operands staged through shared memory by cp.async ------------------
// The register-resident K2 above loads a row block (u and dL/ds of T steps), then computes for ~1 us with nothing in flight,
// then stores: with one 288-thread CTA per SM (T = 10 needs ~220 registers) the memory pipe idles during every compute
// phase (0.58-0.66 of the HBM roofline).  Here each thread's 2T 16-byte operands of the NEXT row block are copied into a
// second shared-memory stage by cp.async while the current block is processed from the first (a thread only ever reads its
// own slots: no block barrier in the loop), so 92 KB per SM are in flight during the compute phases; only h stays in
// registers (u and dL/ds are re-read from the stage where needed).  Persistent grid (one CTA per SM).  Same arithmetic and
// operation order per neuron as lif_bwd_kernel: identical gradients; the BN partial sums group the rows differently.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int T, bool SIMPLE>
__global__ void __launch_bounds__(288, 1) lif_bwd_stream_kernel(const LifBwdP p) {
  constexpr int V = 4;
  extern __shared__ float4 stg4[];                 // [2 stages][2T slots][blockDim.x]
  const SeqP& s = p.s;
  const NeuronP nrn = p.nrn;
  const int nthr = blockDim.x, tid = threadIdx.x;
  const int rx = tid % s.R, ry = tid / s.R;
  const int64_t col = (int64_t)blockIdx.y * s.tile_w + (int64_t)rx * V;
  float sc[V], sh[V];
  init_affine<V>(s, p.scale, p.shift, col, sc, sh);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const float dh_dx = neuron_dh_dx(nrn), dh_dv = neuron_dh_dv(nrn);
  const float vr = nrn.hard ? nrn.v_reset : 0.f;
  const int64_t gstride = (int64_t)gridDim.x * s.k;
  auto issue = [&](int64_t row, int stage) {
    const int64_t n = row * s.row_w + col;
    if (row < s.n_rows && n < s.n_neurons) {
      const int64_t off = seq_base(s, n);
      float4* dst = stg4 + (size_t)stage * 2 * T * nthr + tid;
#pragma unroll
      for (int t = 0; t < T; ++t) cp_async16(dst + t * nthr, p.u + off + t * s.stride_t);
#pragma unroll
      for (int t = 0; t < T; ++t) cp_async16(dst + (T + t) * nthr, p.gs + off + t * s.stride_t);
    }
    cp_async_commit();
  };
  int64_t row = (int64_t)blockIdx.x * s.k + ry;
  issue(row, 0);
  int stage = 0;
  for (; row < s.n_rows; row += gstride, stage ^= 1) {
    issue(row + gstride, stage ^ 1);
    cp_async_wait<1>();                            // everything but the group just committed has landed
    const int64_t n = row * s.row_w + col;
    if (n >= s.n_neurons) continue;
    const int64_t off = seq_base(s, n);
    const float4* src = stg4 + (size_t)stage * 2 * T * nthr + tid;
    if (s.chan_mode >= 2) load_affine<V>(s, p.scale, p.shift, n, sc, sh);
    float h[T][V], v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = vr;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float4 u4 = src[t * nthr];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xx = fmaf(f4(u4, i), sc[i], sh[i]);
        h[t][i] = neuron_charge_t<SIMPLE>(nrn, v[i], xx);
        v[i] = neuron_reset_t<SIMPLE>(nrn, h[t][i], neuron_fire(nrn, h[t][i]));
      }
    }
    float gv[V];
#pragma unroll
    for (int i = 0; i < V; ++i) gv[i] = 0.f;
#pragma unroll
    for (int t = T - 1; t >= 0; --t) {
      const float4 g4 = src[(T + t) * nthr];
      const float4 u4 = src[t * nthr];
      float dx[V], du[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float gh = neuron_grad_h_t<SIMPLE>(nrn, h[t][i], f4(g4, i), gv[i]);
        dx[i] = gh * dh_dx;
        gv[i] = gh * dh_dv;
        du[i] = dx[i] * sc[i];
        if (s.chan_mode == 1) {
          acc[0][i] += dx[i];
          acc[1][i] += dx[i] * f4(u4, i);
        }
      }
      if (p.gu) stv<V>(p.gu + off + t * s.stride_t, du);
      if (p.gx) stv<V>(p.gx + off + t * s.stride_t, dx);
    }
  }
  cp_async_wait<0>();
  if (p.bn_partials && s.chan_mode == 1)
    block_reduce_rows_to_partials<2, V>(acc, reinterpret_cast<float*>(stg4), p.bn_partials, s.R, s.k, s.C, (int64_t)blockIdx.y * s.tile_w);
}

// ---- K1p: PSN ---------------------------------------------------------------------------------
struct PsnP {
  const float* u; void* spike; float* h_seq;
  const float* gs; float* gu; float* gx; float* gh; float* x_out;
  const float* weight; const float* bias;
  const float* scale; const float* shift;
  float* bn_partials;
  float* wg_partials;   // backward, optional: per-block partial sums of dW [T*T] and db [T] (see psn_bwd_kernel<.., WG = true>)
  SeqP s; NeuronP nrn;  // nrn: only sg / sg_alpha used (threshold is 0)
};

template <int T, int V, int DT>
__global__ void __launch_bounds__(512) psn_fwd_kernel(const PsnP p) {
  constexpr int TM = T > 0 ? T : 32;
  __shared__ float sw[32 * 32 + 32];
  const SeqP& s = p.s;
  const int Tn = T > 0 ? T : s.T;
  for (int i = threadIdx.x; i < Tn * Tn; i += blockDim.x) sw[i] = p.weight[i];
  for (int i = threadIdx.x; i < Tn; i += blockDim.x) sw[1024 + i] = p.bias[i];
  __syncthreads();
  const int rx = threadIdx.x % s.R, ry = threadIdx.x / s.R;
  const int64_t col = (int64_t)blockIdx.y * s.tile_w + (int64_t)rx * V;
  float sc[V], sh[V];
  init_affine<V>(s, p.scale, p.shift, col, sc, sh);
  for (int64_t row = (int64_t)blockIdx.x * s.k + ry; row < s.n_rows; row += (int64_t)gridDim.x * s.k) {
    const int64_t n = row * s.row_w + col;
    if (n >= s.n_neurons) continue;
    const int64_t off = seq_base(s, n);
    float x[TM][V];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) ldv<V>(p.u + off + t * s.stride_t, x[t]);
    if (s.chan_mode >= 2) load_affine<V>(s, p.scale, p.shift, n, sc, sh);
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) {
#pragma unroll
        for (int i = 0; i < V; ++i) x[t][i] = fmaf(x[t][i], sc[i], sh[i]);
      }
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (t < Tn) {
        float h[V], sp[V];
#pragma unroll
        for (int i = 0; i < V; ++i) h[i] = 0.f;
#pragma unroll
        for (int k = 0; k < TM; ++k)
          if (k < Tn) {
            const float w = sw[t * Tn + k];
#pragma unroll
            for (int i = 0; i < V; ++i) h[i] = fmaf(w, x[k][i], h[i]);
          }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          h[i] += sw[1024 + t];
          sp[i] = h[i] >= 0.f ? 1.f : 0.f;
        }
        store_spikes<DT, V>(p.spike, off + t * s.stride_t, sp);
        if (p.h_seq) stv<V>(p.h_seq + off + t * s.stride_t, h);
      }
    }
  }
}

// WG = true: the parameter gradients dW[t][k] = sum_n dh[t,n] * x[k,n], db[t] = sum_n dh[t,n] are accumulated in registers
// over the thread's neurons and reduced per block — dh and x never go to HBM (the separate sdf_psn_wgrad pass re-read
// 8 B per neuron-timestep that this kernel had to write first: 28 B -> 12 B per neuron-timestep).
template <int T, int V, bool WG = false>
__global__ void __launch_bounds__(256) psn_bwd_kernel(const PsnP p) {
  constexpr int TM = T > 0 ? T : 32;
  constexpr int TW = WG ? TM : 1;
  extern __shared__ float smem[];
  __shared__ float sw[32 * 32 + 32];
  const SeqP& s = p.s;
  const NeuronP nrn = p.nrn;
  const int Tn = T > 0 ? T : s.T;
  for (int i = threadIdx.x; i < Tn * Tn; i += blockDim.x) sw[i] = p.weight[i];
  for (int i = threadIdx.x; i < Tn; i += blockDim.x) sw[1024 + i] = p.bias[i];
  __syncthreads();
  const int rx = threadIdx.x % s.R, ry = threadIdx.x / s.R;
  const int64_t col = (int64_t)blockIdx.y * s.tile_w + (int64_t)rx * V;
  float sc[V], sh[V];
  init_affine<V>(s, p.scale, p.shift, col, sc, sh);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float wacc[TW][TW], wb[TW];
#pragma unroll
  for (int t = 0; t < TW; ++t) {
    wb[t] = 0.f;
#pragma unroll
    for (int k = 0; k < TW; ++k) wacc[t][k] = 0.f;
  }
  for (int64_t row = (int64_t)blockIdx.x * s.k + ry; row < s.n_rows; row += (int64_t)gridDim.x * s.k) {
    const int64_t n = row * s.row_w + col;
    if (n >= s.n_neurons) continue;
    const int64_t off = seq_base(s, n);
    float u[TM][V], dh[TM][V];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) ldv<V>(p.u + off + t * s.stride_t, u[t]);
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) ldv<V>(p.gs + off + t * s.stride_t, dh[t]);
    if (s.chan_mode >= 2) load_affine<V>(s, p.scale, p.shift, n, sc, sh);
    if (p.x_out) {
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (t < Tn) {
          float x[V];
#pragma unroll
          for (int i = 0; i < V; ++i) x[i] = fmaf(u[t][i], sc[i], sh[i]);
          stv<V>(p.x_out + (int64_t)t * s.n_neurons + n, x);
        }
    }
    // dh_t = g_t * sg(h_t), h_t recomputed from x = u*scale + shift
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (t < Tn) {
        float h[V];
#pragma unroll
        for (int i = 0; i < V; ++i) h[i] = 0.f;
#pragma unroll
        for (int k = 0; k < TM; ++k)
          if (k < Tn) {
            const float w = sw[t * Tn + k];
#pragma unroll
            for (int i = 0; i < V; ++i) h[i] = fmaf(w, fmaf(u[k][i], sc[i], sh[i]), h[i]);
          }
#pragma unroll
        for (int i = 0; i < V; ++i) dh[t][i] *= surrogate_grad(nrn, h[i] + sw[1024 + t]);
        if (p.gh) stv<V>(p.gh + (int64_t)t * s.n_neurons + n, dh[t]);
      }
    }
    if (WG) {
#pragma unroll
      for (int k = 0; k < TW; ++k) {
        float xk[V];
#pragma unroll
        for (int i = 0; i < V; ++i) xk[i] = fmaf(u[k][i], sc[i], sh[i]);
#pragma unroll
        for (int t = 0; t < TW; ++t)
#pragma unroll
          for (int i = 0; i < V; ++i) wacc[t][k] = fmaf(dh[t][i], xk[i], wacc[t][k]);
      }
#pragma unroll
      for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int i = 0; i < V; ++i) wb[t] += dh[t][i];
    }
    // dx_k = sum_t W[t][k] dh_t
#pragma unroll
    for (int k = 0; k < TM; ++k) {
      if (k < Tn) {
        float dx[V], du[V];
#pragma unroll
        for (int i = 0; i < V; ++i) dx[i] = 0.f;
#pragma unroll
        for (int t = 0; t < TM; ++t)
          if (t < Tn) {
            const float w = sw[t * Tn + k];
#pragma unroll
            for (int i = 0; i < V; ++i) dx[i] = fmaf(w, dh[t][i], dx[i]);
          }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          du[i] = dx[i] * sc[i];
          if (s.chan_mode == 1) {
            acc[0][i] += dx[i];
            acc[1][i] += dx[i] * u[k][i];
          }
        }
        if (p.gu) stv<V>(p.gu + off + k * s.stride_t, du);
        if (p.gx) stv<V>(p.gx + off + k * s.stride_t, dx);
      }
    }
  }
  if (p.bn_partials && s.chan_mode == 1 && V == 4)
    block_reduce_rows_to_partials<2>(acc, smem, p.bn_partials, s.R, s.k, s.C, (int64_t)blockIdx.y * s.tile_w);
  if (WG) {
    // block sum of the T*T + T accumulators: warp shuffles, then one pass over the 8 warp rows (fixed order)
    __shared__ float wred[8][TW * TW + TW];
    __syncthreads();
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < TW * TW + TW; ++e) {
      float v = e < TW * TW ? wacc[e / TW][e % TW] : wb[e - TW * TW];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) wred[wp][e] = v;
    }
    __syncthreads();
    float* out = p.wg_partials + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (TW * TW + TW);
    const int nwarps = blockDim.x >> 5;        // host-checked: blockDim.x is a multiple of 32, <= 256
    for (int e = threadIdx.x; e < TW * TW + TW; e += blockDim.x) {
      float sum = 0.f;
      for (int w = 0; w < nwarps; ++w) sum += wred[w][e];
      out[e] = sum;
    }
  }
}

// ---- K1p backward, streamed: same staging as lif_bwd_stream_kernel -----------------------------------------------------
// psn_bwd_kernel<.., WG = true> holds u, dL/ds -> dh and the T x T + T parameter-gradient accumulators in registers (255,
// one 256-thread CTA per SM) and has nothing in flight while it computes: 0.43 of the HBM roofline.  Here the operands of the
// next row block arrive by cp.async while the current one is processed; dh is written back over the thread's own dL/ds slots,
// so the live registers are x (T x 4) + the accumulators in the parameter-gradient phase and dh + the accumulators in the
// dx phase.  Per accumulator the additions happen in the same order as in psn_bwd_kernel.
template <int T>
__global__ void __launch_bounds__(256, 1) psn_bwd_stream_kernel(const PsnP p) {
  constexpr int V = 4;
  extern __shared__ float4 stg4[];                 // [2 stages][2T slots][blockDim.x]
  __shared__ float sw[T * T + T];
  __shared__ float wred[8][T * T + T];
  const SeqP& s = p.s;
  const NeuronP nrn = p.nrn;
  const int nthr = blockDim.x, tid = threadIdx.x;
  for (int i = tid; i < T * T; i += nthr) sw[i] = p.weight[i];
  for (int i = tid; i < T; i += nthr) sw[T * T + i] = p.bias[i];
  __syncthreads();
  const int rx = tid % s.R, ry = tid / s.R;
  const int64_t col = (int64_t)blockIdx.y * s.tile_w + (int64_t)rx * V;
  float sc[V], sh[V];
  init_affine<V>(s, p.scale, p.shift, col, sc, sh);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float wacc[T][T], wb[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    wb[t] = 0.f;
#pragma unroll
    for (int k = 0; k < T; ++k) wacc[t][k] = 0.f;
  }
  const int64_t gstride = (int64_t)gridDim.x * s.k;
  auto issue = [&](int64_t row, int stage) {
    const int64_t n = row * s.row_w + col;
    if (row < s.n_rows && n < s.n_neurons) {
      const int64_t off = seq_base(s, n);
      float4* dst = stg4 + (size_t)stage * 2 * T * nthr + tid;
#pragma unroll
      for (int t = 0; t < T; ++t) cp_async16(dst + t * nthr, p.u + off + t * s.stride_t);
#pragma unroll
      for (int t = 0; t < T; ++t) cp_async16(dst + (T + t) * nthr, p.gs + off + t * s.stride_t);
    }
    cp_async_commit();
  };
  int64_t row = (int64_t)blockIdx.x * s.k + ry;
  issue(row, 0);
  int stage = 0;
  for (; row < s.n_rows; row += gstride, stage ^= 1) {
    issue(row + gstride, stage ^ 1);
    cp_async_wait<1>();
    const int64_t n = row * s.row_w + col;
    if (n >= s.n_neurons) continue;
    const int64_t off = seq_base(s, n);
    float4* src = stg4 + (size_t)stage * 2 * T * nthr + tid;
    if (s.chan_mode >= 2) load_affine<V>(s, p.scale, p.shift, n, sc, sh);
    {
      float x[T][V];
#pragma unroll
      for (int k = 0; k < T; ++k) {
        const float4 u4 = src[k * nthr];
#pragma unroll
        for (int i = 0; i < V; ++i) x[k][i] = fmaf(f4(u4, i), sc[i], sh[i]);
        if (p.x_out) stv<V>(p.x_out + (int64_t)k * s.n_neurons + n, x[k]);
      }
      // dh_t = g_t * sg(h_t), h_t = sum_k W[t][k] x_k + b_t; written back over the thread's own dL/ds slot
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float h[V], dh[V];
#pragma unroll
        for (int i = 0; i < V; ++i) h[i] = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const float w = sw[t * T + k];
#pragma unroll
          for (int i = 0; i < V; ++i) h[i] = fmaf(w, x[k][i], h[i]);
        }
        const float4 g4 = src[(T + t) * nthr];
#pragma unroll
        for (int i = 0; i < V; ++i) dh[i] = f4(g4, i) * surrogate_grad(nrn, h[i] + sw[T * T + t]);
        src[(T + t) * nthr] = make_float4(dh[0], dh[1], dh[2], dh[3]);
        if (p.gh) stv<V>(p.gh + (int64_t)t * s.n_neurons + n, dh);
      }
      // parameter gradients: dW[t][k] += dh_t . x_k, db[t] += sum dh_t
      if (p.wg_partials) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float4 d4 = src[(T + t) * nthr];
#pragma unroll
          for (int k = 0; k < T; ++k)
#pragma unroll
            for (int i = 0; i < V; ++i) wacc[t][k] = fmaf(f4(d4, i), x[k][i], wacc[t][k]);
#pragma unroll
          for (int i = 0; i < V; ++i) wb[t] += f4(d4, i);
        }
      }
    }
    // dx_k = sum_t W[t][k] dh_t
    float dh[T][V];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float4 d4 = src[(T + t) * nthr];
#pragma unroll
      for (int i = 0; i < V; ++i) dh[t][i] = f4(d4, i);
    }
#pragma unroll
    for (int k = 0; k < T; ++k) {
      float dx[V], du[V];
#pragma unroll
      for (int i = 0; i < V; ++i) dx[i] = 0.f;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float w = sw[t * T + k];
#pragma unroll
        for (int i = 0; i < V; ++i) dx[i] = fmaf(w, dh[t][i], dx[i]);
      }
      const float4 u4 = src[k * nthr];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        du[i] = dx[i] * sc[i];
        if (s.chan_mode == 1) {
          acc[0][i] += dx[i];
          acc[1][i] += dx[i] * f4(u4, i);
        }
      }
      if (p.gu) stv<V>(p.gu + off + k * s.stride_t, du);
      if (p.gx) stv<V>(p.gx + off + k * s.stride_t, dx);
    }
  }
  cp_async_wait<0>();
  if (p.bn_partials && s.chan_mode == 1)
    block_reduce_rows_to_partials<2>(acc, reinterpret_cast<float*>(stg4), p.bn_partials, s.R, s.k, s.C, (int64_t)blockIdx.y * s.tile_w);
  if (p.wg_partials) {
    __syncthreads();
    const int lane = tid & 31, wp = tid >> 5;
#pragma unroll
    for (int e = 0; e < T * T + T; ++e) {
      float v = e < T * T ? wacc[e / T][e % T] : wb[e - T * T];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) wred[wp][e] = v;
    }
    __syncthreads();
    float* out = p.wg_partials + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (T * T + T);
    const int nwarps = nthr >> 5;
    for (int e = tid; e < T * T + T; e += nthr) {
      float sum = 0.f;
      for (int w = 0; w < nwarps; ++w) sum += wred[w][e];
      out[e] = sum;
    }
  }
}

// ---- host: tiling and dispatch --------------------------------------------------------------
struct SeqLaunch {
  SeqP s;
  int V;
  dim3 grid;
  int threads;
};

// V_req: widest vector the kernel family supports; target: threads per block.
static int build_seq(const sdf_seq_layout& lay, int64_t C, int64_t hw, bool affine, bool want_partials,
                     const void* const* ptrs, int nptrs, int V_req, int target, int64_t max_blocks,
                     SeqLaunch* L) {
  SDF_REQUIRE(lay.T >= 1 && lay.T <= 32, "neuron kernel: T=%lld out of range [1,32]", (long long)lay.T);
  SDF_REQUIRE(lay.n_neurons > 0 && lay.inner > 0, "neuron kernel: bad layout");
  SDF_REQUIRE(lay.stride_b == 0 || lay.n_neurons % lay.inner == 0, "neuron kernel: n_neurons %% inner != 0");
  if (affine) SDF_REQUIRE(C > 0 && hw > 0, "neuron kernel: affine prologue needs C, hw > 0");
  SeqP* s = &L->s;
  int V = V_req;
  bool al = (lay.n_neurons % 4 == 0) && (lay.inner % 4 == 0) && (lay.stride_b % 4 == 0) && (lay.stride_t % 4 == 0);
  for (int i = 0; i < nptrs; ++i)
    if (ptrs[i] && !aligned16(ptrs[i])) al = false;
  if (affine && hw == 1 && C % 4 != 0) al = false;
  if (affine && hw > 1 && hw % 4 != 0) al = false;
  if (!al) V = 1;
  s->n_neurons = lay.n_neurons; s->inner = lay.inner; s->stride_b = lay.stride_b; s->stride_t = lay.stride_t;
  s->C = C > 0 ? C : 1; s->hw = hw > 0 ? hw : 1; s->T = (int)lay.T;
  if (lay.stride_b == 0) s->inner = lay.n_neurons;
  RowTiling rt;
  bool chan_fixed = affine && hw == 1 && V > 1 && (lay.stride_b == 0 || lay.inner % C == 0) &&
                    (lay.n_neurons % C == 0) && make_row_tiling(lay.n_neurons / C, C, V, target, (int)max_blocks, &rt);
  if (want_partials)
    SDF_REQUIRE(chan_fixed, "neuron backward: bn_partials need 16B-aligned channels-last input with C %% 4 == 0");
  if (chan_fixed) {
    s->chan_mode = 1; s->row_w = C; s->tile_w = rt.tile_w; s->R = rt.R; s->k = rt.k;
    s->n_rows = lay.n_neurons / C;
    L->grid = dim3(rt.blocks, rt.ncol, 1);
    L->threads = rt.threads;
  } else {
    s->chan_mode = !affine ? 0 : (hw == 1 ? 2 : 3);
    s->R = target / 2; s->k = 2; s->tile_w = (int64_t)s->R * V; s->row_w = s->tile_w;
    s->n_rows = (lay.n_neurons + s->row_w - 1) / s->row_w;
    int64_t need = (s->n_rows + s->k - 1) / s->k;
    L->grid = dim3((unsigned)(need < max_blocks ? need : max_blocks), 1, 1);
    L->threads = s->R * s->k;
  }
  L->V = V;
  return SDF_OK;
}

}  // namespace sdf

using namespace sdf;

extern "C" int64_t sdf_partial_blocks(int64_t rows, int64_t C) {
  (void)rows; (void)C;
  return (int64_t)kNumSMs * 3;
}

#define SDF_DISPATCH_DT(KERNEL, TT, VV, DT, ...)                                       \
  do {                                                                                 \
    if (DT == SDF_SPIKE_F32) KERNEL<TT, VV, SDF_SPIKE_F32> __VA_ARGS__;                \
    else if (DT == SDF_SPIKE_U8) KERNEL<TT, VV, SDF_SPIKE_U8> __VA_ARGS__;             \
    else KERNEL<TT, VV, SDF_SPIKE_BF16> __VA_ARGS__;                                   \
  } while (0)
// LIF forward: + SIMPLE fast path for the hot step counts
#define SDF_DISPATCH_LIF(TT, VV, DT, SIMPLE, ...)                                                   \
  do {                                                                                              \
    if (SIMPLE) {                                                                                   \
      if (DT == SDF_SPIKE_F32) lif_fwd_kernel<TT, VV, SDF_SPIKE_F32, true> __VA_ARGS__;             \
      else if (DT == SDF_SPIKE_U8) lif_fwd_kernel<TT, VV, SDF_SPIKE_U8, true> __VA_ARGS__;          \
      else lif_fwd_kernel<TT, VV, SDF_SPIKE_BF16, true> __VA_ARGS__;                                \
    } else {                                                                                        \
      if (DT == SDF_SPIKE_F32) lif_fwd_kernel<TT, VV, SDF_SPIKE_F32, false> __VA_ARGS__;            \
      else if (DT == SDF_SPIKE_U8) lif_fwd_kernel<TT, VV, SDF_SPIKE_U8, false> __VA_ARGS__;         \
      else lif_fwd_kernel<TT, VV, SDF_SPIKE_BF16, false> __VA_ARGS__;                               \
    }                                                                                               \
  } while (0)

extern "C" int sdf_lif_fwd(const sdf_lif_fwd_args* a) {
  SDF_REQUIRE(a && a->u && a->spike, "sdf_lif_fwd: null argument");
  int st = validate_neuron(a->neuron);
  if (st) return st;
  const bool affine = a->scale != nullptr;
  SDF_REQUIRE(!affine || a->shift, "sdf_lif_fwd: scale without shift");
  if (a->lay.n_neurons == 0) return SDF_OK;
  const int DT = a->spike_dtype;
  SDF_REQUIRE(DT >= 0 && DT <= 2, "sdf_lif_fwd: bad spike_dtype %d", DT);
  LifFwdP p;
  p.u = a->u; p.spike = a->spike; p.h_seq = a->h_seq; p.v_init = a->v_init; p.v_final = a->v_final;
  p.scale = a->scale; p.shift = a->shift; p.nrn = make_neuron(a->neuron);
  const void* ptrs[] = {a->u, a->spike, a->h_seq, a->v_init, a->v_final};
  const int T = (int)a->lay.T;
  const bool fastT = (T == 2 || T == 4 || T == 5 || T == 10 || T == 20);
  SeqLaunch L;
  st = build_seq(a->lay, a->C, a->hw, affine, false, ptrs, 5, fastT ? 4 : 1, 512, (int64_t)kNumSMs * 2, &L);
  if (st) return st;
  p.s = L.s;
  cudaStream_t stream = (cudaStream_t)a->stream;
  const bool simple = neuron_is_simple(p.nrn);
  if (L.V == 4) {
    switch (T) {
      case 2: SDF_DISPATCH_LIF(2, 4, DT, simple, <<<L.grid, L.threads, 0, stream>>>(p)); break;
      case 4: SDF_DISPATCH_LIF(4, 4, DT, false, <<<L.grid, L.threads, 0, stream>>>(p)); break;
      case 5: SDF_DISPATCH_LIF(5, 4, DT, simple, <<<L.grid, L.threads, 0, stream>>>(p)); break;
      case 10: SDF_DISPATCH_LIF(10, 4, DT, simple, <<<L.grid, L.threads, 0, stream>>>(p)); break;
      default: SDF_DISPATCH_LIF(20, 4, DT, false, <<<L.grid, L.threads, 0, stream>>>(p)); break;
    }
  } else {
    SDF_DISPATCH_LIF(0, 1, DT, false, <<<L.grid, L.threads, 0, stream>>>(p));
  }
  return finish_launch("sdf_lif_fwd");
}

extern "C" int sdf_lif_bwd(const sdf_lif_bwd_args* a) {
  SDF_REQUIRE(a && a->u && a->grad_spike && (a->grad_u || a->grad_x || a->bn_partials), "sdf_lif_bwd: null argument");
  SDF_REQUIRE(!a->bn_coef || (a->grad_u && a->scale && a->hw == 1), "sdf_lif_bwd: bn_coef needs grad_u and a channels-last BN site");
  int st = validate_neuron(a->neuron);
  if (st) return st;
  const bool affine = a->scale != nullptr;
  SDF_REQUIRE(!affine || a->shift, "sdf_lif_bwd: scale without shift");
  cudaStream_t stream = (cudaStream_t)a->stream;
  if (a->lay.n_neurons == 0) return SDF_OK;
  LifBwdP p;
  p.u = a->u; p.gs = (const float*)a->grad_spike; p.gu = a->grad_u; p.gx = a->grad_x; p.v_init = a->v_init;
  p.scale = a->scale; p.shift = a->shift; p.bn_partials = a->bn_partials; p.plif_partials = a->plif_partials;
  p.coef = a->bn_coef;
  p.nrn = make_neuron(a->neuron);
  const void* ptrs[] = {a->u, a->grad_spike, a->grad_u, a->grad_x, a->v_init};
  const int T = (int)a->lay.T;
  const bool fastT = (T == 2 || T == 4 || T == 5 || T == 10 || T == 20);
  const bool parts = a->bn_partials || a->plif_partials;
  int64_t max_blocks = (int64_t)kNumSMs * (T == 10 ? 3 : 2);
  if (parts) {
    SDF_REQUIRE(a->n_partial_blocks >= 1, "sdf_lif_bwd: n_partial_blocks < 1");
    if (a->n_partial_blocks < max_blocks) max_blocks = a->n_partial_blocks;
  }
  if (a->bn_partials) SDF_REQUIRE(fastT, "sdf_lif_bwd: bn_partials supported for T in {2,4,5,10,20}");
  SeqLaunch L;
  // T = 10, the common sites (LIF / IF, zero initial state, no folded BN backward): operands staged through shared memory by
  // cp.async, persistent grid (lif_bwd_stream_kernel)
  static const int stream_mode = [] { const char* e = getenv("SDF_LIF_BWD_STREAM"); return e ? atoi(e) : 1; }();
  if (stream_mode && (T == 10 || T == 20) && !a->v_init && !a->bn_coef && !a->plif_partials && a->neuron.kind != SDF_NEURON_PLIF) {
    int64_t mb = kNumSMs;
    if (parts && a->n_partial_blocks < mb) mb = a->n_partial_blocks;
    // two stages of 2T 16-byte slots per thread: 288 threads at T = 10 (180 KB), 128 at T = 20 (160 KB)
    st = build_seq(a->lay, a->C, a->hw, affine, a->bn_partials != nullptr, ptrs, 5, 4, T == 10 ? 288 : 128, mb, &L);
    if (st) return st;
    if (L.V == 4) {
      p.s = L.s;
      if (a->bn_partials && a->n_partial_blocks > (int64_t)L.grid.x)
        cudaMemsetAsync(a->bn_partials + (int64_t)L.grid.x * 2 * a->C, 0,
                        sizeof(float) * (a->n_partial_blocks - L.grid.x) * 2 * a->C, stream);
      const size_t smem_s = (size_t)2 * 2 * T * L.threads * sizeof(float4);
      static bool attr_done = false;
      if (!attr_done) {
        cudaFuncSetAttribute(lif_bwd_stream_kernel<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(lif_bwd_stream_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(lif_bwd_stream_kernel<20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(lif_bwd_stream_kernel<20, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
      }
      const bool simple_s = neuron_is_simple(p.nrn);
      if (T == 10) {
        if (simple_s) lif_bwd_stream_kernel<10, true><<<L.grid, L.threads, smem_s, stream>>>(p);
        else lif_bwd_stream_kernel<10, false><<<L.grid, L.threads, smem_s, stream>>>(p);
      } else {
        if (simple_s) lif_bwd_stream_kernel<20, true><<<L.grid, L.threads, smem_s, stream>>>(p);
        else lif_bwd_stream_kernel<20, false><<<L.grid, L.threads, smem_s, stream>>>(p);
      }
      return finish_launch("sdf_lif_bwd");
    }
  }
  // T = 10 keeps u, h (and the preloaded grads) in registers: 2 neurons per thread there, 4 otherwise
  static const int v4_t10 = [] { const char* e = getenv("SDF_LIF_BWD_V4"); return e ? atoi(e) : 1; }();
  // T = 10 with 4 neurons/thread keeps u, h and the preloaded grads in ~220 registers: 288-thread CTAs, one per SM
  st = build_seq(a->lay, a->C, a->hw, affine, a->bn_partials != nullptr, ptrs, 5, fastT ? (((T == 10 && !v4_t10) || T == 20) ? 2 : 4) : 1,
                 (T == 10 && v4_t10) ? 288 : 256, max_blocks, &L);
  if (st) return st;
  p.s = L.s;
  if (a->plif_partials) SDF_REQUIRE(L.grid.y == 1 || L.grid.x * L.grid.y <= a->n_partial_blocks, "sdf_lif_bwd: plif partial capacity");
  if (a->bn_partials && a->n_partial_blocks > (int64_t)L.grid.x)
    cudaMemsetAsync(a->bn_partials + (int64_t)L.grid.x * 2 * a->C, 0,
                    sizeof(float) * (a->n_partial_blocks - L.grid.x) * 2 * a->C, stream);
  if (a->plif_partials) cudaMemsetAsync(a->plif_partials, 0, sizeof(float) * a->n_partial_blocks, stream);
  const size_t smem = sizeof(float) * 4 * (size_t)L.threads;
  const bool simple = neuron_is_simple(p.nrn);
  if (L.V == 4) {
    switch (T) {
      case 2:
        if (simple) lif_bwd_kernel<2, 4, 1, true><<<L.grid, L.threads, smem, stream>>>(p);
        else lif_bwd_kernel<2, 4, 1, false><<<L.grid, L.threads, smem, stream>>>(p);
        break;
      case 4: lif_bwd_kernel<4, 4, 1, false><<<L.grid, L.threads, smem, stream>>>(p); break;
      case 10:
        if (simple) lif_bwd_kernel<10, 4, 1, true><<<L.grid, L.threads, smem, stream>>>(p);
        else lif_bwd_kernel<10, 4, 1, false><<<L.grid, L.threads, smem, stream>>>(p);
        break;
      default:
        if (simple) lif_bwd_kernel<5, 4, 1, true><<<L.grid, L.threads, smem, stream>>>(p);
        else lif_bwd_kernel<5, 4, 1, false><<<L.grid, L.threads, smem, stream>>>(p);
        break;
    }
  } else if (L.V == 2 && T == 20) {
    // T = 20 (20-bin inputs): u and h of 2 neurons x 20 steps stay in registers (vector path; grads streamed), one CTA per SM
    if (simple) lif_bwd_kernel<20, 2, 0, true><<<L.grid, L.threads, smem, stream>>>(p);
    else lif_bwd_kernel<20, 2, 0, false><<<L.grid, L.threads, smem, stream>>>(p);
  } else if (L.V == 2) {
    static const int pre_mode = [] { const char* e = getenv("SDF_LIF_BWD_PRELOAD"); return e ? atoi(e) : 1; }();
    if (simple && pre_mode) lif_bwd_kernel<10, 2, 1, true><<<L.grid, L.threads, smem, stream>>>(p);
    else if (simple) lif_bwd_kernel<10, 2, 0, true><<<L.grid, L.threads, smem, stream>>>(p);
    else lif_bwd_kernel<10, 2, 1, false><<<L.grid, L.threads, smem, stream>>>(p);
  } else {
    lif_bwd_kernel<0, 1, 0, false><<<L.grid, L.threads, smem, stream>>>(p);
  }
  return finish_launch("sdf_lif_bwd");
}

extern "C" int sdf_psn_fwd(const sdf_psn_fwd_args* a) {
  SDF_REQUIRE(a && a->u && a->spike && a->weight && a->bias, "sdf_psn_fwd: null argument");
  const bool affine = a->scale != nullptr;
  SDF_REQUIRE(!affine || a->shift, "sdf_psn_fwd: scale without shift");
  if (a->lay.n_neurons == 0) return SDF_OK;
  const int DT = a->spike_dtype;
  SDF_REQUIRE(DT >= 0 && DT <= 2, "sdf_psn_fwd: bad spike_dtype %d", DT);
  PsnP p = {};
  p.u = a->u; p.spike = a->spike; p.h_seq = a->h_seq; p.weight = a->weight; p.bias = a->bias;
  p.scale = a->scale; p.shift = a->shift;
  const void* ptrs[] = {a->u, a->spike, a->h_seq};
  const int T = (int)a->lay.T;
  const bool fastT = (T == 2 || T == 4 || T == 5 || T == 10);
  SeqLaunch L;
  int st = build_seq(a->lay, a->C, a->hw, affine, false, ptrs, 3, fastT ? 4 : 1, 512, (int64_t)kNumSMs * 2, &L);
  if (st) return st;
  p.s = L.s;
  cudaStream_t stream = (cudaStream_t)a->stream;
  if (L.V == 4) {
    switch (T) {
      case 2: SDF_DISPATCH_DT(psn_fwd_kernel, 2, 4, DT, <<<L.grid, L.threads, 0, stream>>>(p)); break;
      case 4: SDF_DISPATCH_DT(psn_fwd_kernel, 4, 4, DT, <<<L.grid, L.threads, 0, stream>>>(p)); break;
      case 5: SDF_DISPATCH_DT(psn_fwd_kernel, 5, 4, DT, <<<L.grid, L.threads, 0, stream>>>(p)); break;
      default: SDF_DISPATCH_DT(psn_fwd_kernel, 10, 4, DT, <<<L.grid, L.threads, 0, stream>>>(p)); break;
    }
  } else {
    SDF_DISPATCH_DT(psn_fwd_kernel, 0, 1, DT, <<<L.grid, L.threads, 0, stream>>>(p));
  }
  return finish_launch("sdf_psn_fwd");
}

extern "C" int sdf_psn_bwd(const sdf_psn_bwd_args* a) {
  SDF_REQUIRE(a && a->u && a->grad_spike && (a->grad_u || a->grad_x) && (a->grad_h || a->wgrad_partials) && a->weight && a->bias,
              "sdf_psn_bwd: null argument");
  const bool affine = a->scale != nullptr;
  SDF_REQUIRE(!affine || a->shift, "sdf_psn_bwd: scale without shift");
  if (a->lay.n_neurons == 0) return SDF_OK;
  PsnP p = {};
  p.u = a->u; p.gs = a->grad_spike; p.gu = a->grad_u; p.gx = a->grad_x; p.gh = a->grad_h; p.x_out = a->x_out;
  p.weight = a->weight; p.bias = a->bias; p.scale = a->scale; p.shift = a->shift; p.bn_partials = a->bn_partials;
  p.wg_partials = a->wgrad_partials;
  sdf_neuron_cfg nc = {};
  nc.kind = SDF_NEURON_IF; nc.surrogate = a->surrogate; nc.sg_alpha = a->sg_alpha; nc.tau = 2.0; nc.v_th = 0.0;
  p.nrn = make_neuron(nc);
  const void* ptrs[] = {a->u, a->grad_spike, a->grad_u, a->grad_h, a->x_out, a->grad_x};
  const int T = (int)a->lay.T;
  const bool fastT = (T == 2 || T == 4 || T == 5 || T == 10);
  int64_t max_blocks = (int64_t)kNumSMs * 2;
  if (a->bn_partials) {
    SDF_REQUIRE(a->n_partial_blocks >= 1 && fastT, "sdf_psn_bwd: bn_partials need T in {2,4,5,10}");
    if (a->n_partial_blocks < max_blocks) max_blocks = a->n_partial_blocks;
  }
  SeqLaunch L;
  int st;
  cudaStream_t stream = (cudaStream_t)a->stream;
  // T = 10 with the parameter gradients in the same pass (the training path of the shipped configuration): operands staged
  // through shared memory by cp.async, persistent grid (psn_bwd_stream_kernel)
  static const int stream_mode = [] { const char* e = getenv("SDF_PSN_BWD_STREAM"); return e ? atoi(e) : 1; }();
  if (stream_mode && T == 10 && a->wgrad_partials) {
    int64_t mb = kNumSMs;
    if (a->bn_partials && a->n_partial_blocks < mb) mb = a->n_partial_blocks;
    st = build_seq(a->lay, a->C, a->hw, affine, a->bn_partials != nullptr, ptrs, 6, 4, 256, mb, &L);
    if (st) return st;
    const int64_t nb = (int64_t)L.grid.x * L.grid.y, per = (int64_t)T * T + T;
    if (L.V == 4 && L.threads % 32 == 0 && L.threads <= 256 && a->n_wgrad_blocks >= nb) {
      p.s = L.s;
      if (a->bn_partials && a->n_partial_blocks > (int64_t)L.grid.x)
        cudaMemsetAsync(a->bn_partials + (int64_t)L.grid.x * 2 * a->C, 0,
                        sizeof(float) * (a->n_partial_blocks - L.grid.x) * 2 * a->C, stream);
      if (a->n_wgrad_blocks > nb) cudaMemsetAsync(a->wgrad_partials + nb * per, 0, sizeof(float) * (a->n_wgrad_blocks - nb) * per, stream);
      static bool attr_done = false;
      if (!attr_done) {
        cudaFuncSetAttribute(psn_bwd_stream_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
        attr_done = true;
      }
      const size_t smem_s = (size_t)2 * 2 * 10 * L.threads * sizeof(float4);
      psn_bwd_stream_kernel<10><<<L.grid, L.threads, smem_s, stream>>>(p);
      return finish_launch("sdf_psn_bwd");
    }
  }
  st = build_seq(a->lay, a->C, a->hw, affine, a->bn_partials != nullptr, ptrs, 6, fastT ? 4 : 1,
                 256, max_blocks, &L);
  if (st) return st;
  p.s = L.s;
  if (a->bn_partials && a->n_partial_blocks > (int64_t)L.grid.x)
    cudaMemsetAsync(a->bn_partials + (int64_t)L.grid.x * 2 * a->C, 0,
                    sizeof(float) * (a->n_partial_blocks - L.grid.x) * 2 * a->C, stream);
  const size_t smem = sizeof(float) * 4 * (size_t)L.threads;
  if (a->wgrad_partials) {
    // parameter gradients accumulated in the same pass: vector path only, 256-thread blocks (8 warps in the block reduce)
    const int64_t nb = (int64_t)L.grid.x * L.grid.y, per = (int64_t)T * T + T;
    SDF_REQUIRE(L.V == 4 && fastT && L.threads % 32 == 0 && L.threads <= 256,
                "sdf_psn_bwd: wgrad_partials need the vector path (T in {2,4,5,10}, 16-byte aligned layout, whole warps)");
    SDF_REQUIRE(a->n_wgrad_blocks >= nb, "sdf_psn_bwd: wgrad_partials too small (%lld blocks needed)", (long long)nb);
    if (a->n_wgrad_blocks > nb) cudaMemsetAsync(a->wgrad_partials + nb * per, 0, sizeof(float) * (a->n_wgrad_blocks - nb) * per, stream);
    switch (T) {
      case 2: psn_bwd_kernel<2, 4, true><<<L.grid, L.threads, smem, stream>>>(p); break;
      case 4: psn_bwd_kernel<4, 4, true><<<L.grid, L.threads, smem, stream>>>(p); break;
      case 5: psn_bwd_kernel<5, 4, true><<<L.grid, L.threads, smem, stream>>>(p); break;
      default: psn_bwd_kernel<10, 4, true><<<L.grid, L.threads, smem, stream>>>(p); break;
    }
  } else if (L.V == 4) {
    switch (T) {
      case 2: psn_bwd_kernel<2, 4><<<L.grid, L.threads, smem, stream>>>(p); break;
      case 4: psn_bwd_kernel<4, 4><<<L.grid, L.threads, smem, stream>>>(p); break;
      case 5: psn_bwd_kernel<5, 4><<<L.grid, L.threads, smem, stream>>>(p); break;
      default: psn_bwd_kernel<10, 4><<<L.grid, L.threads, smem, stream>>>(p); break;
    }
  } else {
    psn_bwd_kernel<0, 1><<<L.grid, L.threads, smem, stream>>>(p);
  }
  return finish_launch("sdf_psn_bwd");
}

// ---- PSN parameter gradients: dW[t,k] = sum_n gh[t,n] * x[k,n],  db[t] = sum_n gh[t,n] ------------------------------------
// The reference gets them from autograd through addmm (Spiking_submodules.py:207-211): a [T, n] x [n, T] product with T <= 10
// and n ~ 1e8, which library GEMMs handle badly (cuBLAS picks a large-K SIMT / unsplit kernel: 55-72 ms per training step).
// Here: one thread walks a strided set of neurons with the T x T outer product in registers; per-block partials, summed by
// the caller.  HBM-bound: 8*T bytes per neuron.
namespace sdf {
template <int T>
__global__ void __launch_bounds__(256) psn_wgrad_kernel(const float* __restrict__ gh, const float* __restrict__ x, int64_t n,
                                                        float* __restrict__ partial) {
  float acc[T][T], accb[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    accb[t] = 0.f;
#pragma unroll
    for (int k = 0; k < T; ++k) acc[t][k] = 0.f;
  }
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    float g[T], xv[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { g[t] = ld_stream1(gh + (int64_t)t * n + j); xv[t] = ld_stream1(x + (int64_t)t * n + j); }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      accb[t] += g[t];
#pragma unroll
      for (int k = 0; k < T; ++k) acc[t][k] = fmaf(g[t], xv[k], acc[t][k]);
    }
  }
  __shared__ float red[8];
  float* out = partial + (int64_t)blockIdx.x * (T * T + T);
#pragma unroll
  for (int e = 0; e < T * T + T; ++e) {
    float v = e < T * T ? acc[e / T][e % T] : accb[e - T * T];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      out[e] = s;
    }
    __syncthreads();
  }
}
}  // namespace sdf

extern "C" int sdf_psn_wgrad(const sdf_psn_wgrad_args* a) {
  SDF_REQUIRE(a && a->grad_h && a->x && a->partials, "sdf_psn_wgrad: null argument");
  SDF_REQUIRE(a->n_neurons > 0 && a->n_partial_blocks >= 1, "sdf_psn_wgrad: empty problem");
  int64_t blocks = (a->n_neurons + 255) / 256;
  if (blocks > a->n_partial_blocks) blocks = a->n_partial_blocks;
  if (blocks > sdf::kNumSMs * 4) blocks = sdf::kNumSMs * 4;
  cudaStream_t st = (cudaStream_t)a->stream;
  const int64_t per = a->T * a->T + a->T;
  if (a->n_partial_blocks > blocks) cudaMemsetAsync(a->partials + blocks * per, 0, sizeof(float) * (a->n_partial_blocks - blocks) * per, st);
  switch (a->T) {
    case 2: sdf::psn_wgrad_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(a->grad_h, a->x, a->n_neurons, a->partials); break;
    case 4: sdf::psn_wgrad_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(a->grad_h, a->x, a->n_neurons, a->partials); break;
    case 5: sdf::psn_wgrad_kernel<5><<<(unsigned)blocks, 256, 0, st>>>(a->grad_h, a->x, a->n_neurons, a->partials); break;
    case 10: sdf::psn_wgrad_kernel<10><<<(unsigned)blocks, 256, 0, st>>>(a->grad_h, a->x, a->n_neurons, a->partials); break;
    default: sdf::set_error("sdf_psn_wgrad: T=%lld not built (2, 4, 5, 10)", (long long)a->T); return SDF_ERR_UNSUPPORTED;
  }
  return sdf::finish_launch("sdf_psn_wgrad");
}
