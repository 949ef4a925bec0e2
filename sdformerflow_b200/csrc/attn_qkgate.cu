// attn_qkgate.cu — K5: the QK token-gate spiking window attention core, forward and
// surrogate-gradient backward, as ONE kernel each.
//
// Replaces reference models/STSwinNet_SNN/Spiking_swin_transformer3D.py:671-710
// (Spiking_QK_WindowAttention3D.forward), i.e. per token and real head:
//     q  = LIF_fakeT( bn_q(q_pre) )                       (:672-674)
//     a  = LIF_fakeT( sum_{d<32} q )                      (:687,692-693  sn2_q)
//     k  = LIF_fakeT( bn_k(k_pre) + positional_encoding ) (:676-680)
//     g  = k * a                                          (:694)
//     x[t'', m', pos'', h''*32 + d] = g[((m'*nH + h'')*wd + t'')*P + pos''][d]   (:709-710)
// with neuron time = the window "fake time" axis.  The reference
// runs 3 multi-step neurons (~60 ATen launches), a reduction, a product and two reshape/permute
// copies; here a group of 8 lanes owns one (token, head) = 32 channels as 8 float4, the head sum
// is three shuffles, and g is written straight to its permuted destination (128 B segments).
// HBM bound: reads q_pre,k_pre (8 B) and writes g (4/1/2 B) per token-channel.
#include "sdf_common.cuh"

namespace sdf {

struct QkP {
  const float* q_pre; const float* k_pre; int64_t ld;
  const float* q_scale; const float* q_shift; const float* k_scale; const float* k_shift;
  const float* pos;
  void* gate; float* q_h; float* k_h; float* a_h;
  const float* grad_gate; float* grad_q; float* grad_k; float* part_q; float* part_k;
  int64_t M, P, C, nH, MP, tile_w;
  int R, k, wd;
  FastDiv dP, dwd, dnH;
  NeuronP nrn;
};

__device__ __forceinline__ float head_sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 2, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 1, 8);
  return v;
}

// destination row/col of source group j = ((t*M + m)*P + pos)*nH + ch   (Appendix B.3)
__device__ __forceinline__ int64_t gate_dest(const QkP& p, int64_t t, int64_t nr, int64_t ch) {
  const uint32_t j = (uint32_t)((t * p.MP + nr) * p.nH + ch);   // < 2^31, checked on the host
  uint32_t r, pos2, t2, m2, h2;
  p.dP.divmod(j, r, pos2);
  p.dwd.divmod(r, r, t2);
  p.dnH.divmod(r, m2, h2);
  return (((int64_t)t2 * p.M + m2) * p.P + pos2) * p.C + h2 * 32;
}

template <int T, int DT, bool BWD, bool SIMPLE>
__global__ void __launch_bounds__(512) qkgate_kernel(const QkP p) {
  constexpr int TM = T > 0 ? T : 8;
  extern __shared__ float smem[];
  const int Tn = T > 0 ? T : p.wd;
  const NeuronP nrn = p.nrn;
  const int rx = threadIdx.x % p.R, ry = threadIdx.x / p.R;
  const int64_t col = (int64_t)blockIdx.y * p.tile_w + (int64_t)rx * 4;
  const int64_t ch = col >> 5;          // real head of this thread's 4 channels
  const int64_t cin = col & 31;         // channel offset inside the head
  const float4 qsc = *reinterpret_cast<const float4*>(p.q_scale + col);
  const float4 qsh = *reinterpret_cast<const float4*>(p.q_shift + col);
  const float4 ksc = *reinterpret_cast<const float4*>(p.k_scale + col);
  const float4 ksh = *reinterpret_cast<const float4*>(p.k_shift + col);
  const float dh_dx = neuron_dh_dx(nrn), dh_dv = neuron_dh_dv(nrn);
  const float v0 = nrn.hard ? nrn.v_reset : 0.f;
  float accq[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float acck[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const int64_t stride = (int64_t)gridDim.x * p.k;
  // block-uniform trip count: every lane of a warp must reach the shuffles
  for (int64_t nr0 = (int64_t)blockIdx.x * p.k; nr0 < p.MP; nr0 += stride) {
    const int64_t nr = nr0 + ry;
    const bool valid = nr < p.MP;
    const int64_t nrc = valid ? nr : 0;
    uint32_t mq, posq;
    p.dP.divmod((uint32_t)nrc, mq, posq);
    const int64_t pcoff = (int64_t)posq * p.C + col;  // offset inside one (wh,ww,C) slab of pos
    float4 qp[TM], kp[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) {
        const int64_t o = (t * p.MP + nrc) * p.ld + col;
        qp[t] = ld_stream4(p.q_pre + o);
        kp[t] = ld_stream4(p.k_pre + o);
      }
    // ---- forward (also the recompute half of backward) ----
    float4 hq[TM], hk[TM];   // membrane potentials
    float ha[TM];            // sn2_q membrane
    float av[TM];            // gate a_t
    float vq[4], vk[4], va = v0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { vq[i] = v0; vk[i] = v0; }
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (t < Tn) {
        const float4 pe = __ldg(reinterpret_cast<const float4*>(p.pos + t * p.P * p.C + pcoff));
        float cnt = 0.f;
        float4 sk;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xq = fmaf(f4(qp[t], i), f4(qsc, i), f4(qsh, i));
          f4(hq[t], i) = neuron_charge_t<SIMPLE>(nrn, vq[i], xq);
          const float sq = neuron_fire(nrn, f4(hq[t], i));
          vq[i] = neuron_reset_t<SIMPLE>(nrn, f4(hq[t], i), sq);
          cnt += sq;
          const float xk = __fadd_rn(fmaf(f4(kp[t], i), f4(ksc, i), f4(ksh, i)), f4(pe, i));
          f4(hk[t], i) = neuron_charge_t<SIMPLE>(nrn, vk[i], xk);
          f4(sk, i) = neuron_fire(nrn, f4(hk[t], i));
          vk[i] = neuron_reset_t<SIMPLE>(nrn, f4(hk[t], i), f4(sk, i));
        }
        cnt = head_sum8(cnt);               // exact small integer in fp32
        ha[t] = neuron_charge_t<SIMPLE>(nrn, va, cnt);
        av[t] = neuron_fire(nrn, ha[t]);
        va = neuron_reset_t<SIMPLE>(nrn, ha[t], av[t]);
        if (!BWD && valid) {
          float4 g;
#pragma unroll
          for (int i = 0; i < 4; ++i) f4(g, i) = f4(sk, i) * av[t];
          store_spike4<DT>(p.gate, gate_dest(p, t, nr, ch) + cin, g);
          const int64_t o = (t * p.MP + nr) * p.C + col;
          if (p.q_h) st_stream4(p.q_h + o, hq[t]);
          if (p.k_h) st_stream4(p.k_h + o, hk[t]);
          if (p.a_h && cin == 0) p.a_h[(t * p.MP + nr) * p.nH + ch] = ha[t];
        }
      }
    }
    if (BWD) {
      // ---- adjoint ----
      float4 dg[TM];
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (t < Tn) dg[t] = valid ? ld_stream4(p.grad_gate + gate_dest(p, t, nrc, ch) + cin) : make_float4(0.f, 0.f, 0.f, 0.f);
      float gva = 0.f, gvq[4] = {0.f, 0.f, 0.f, 0.f}, gvk[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = TM - 1; t >= 0; --t) {
        if (t < Tn) {
          // da_t = sum_d dg * sk
          float da = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) da += f4(dg[t], i) * neuron_fire(nrn, f4(hk[t], i));
          da = head_sum8(da);
          const float gha = neuron_grad_h_t<SIMPLE>(nrn, ha[t], da, gva);
          const float dcnt = gha * dh_dx;   // d/d(sum q), broadcast to the 32 q spikes
          gva = gha * dh_dv;
          float4 dq, dk;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float ghq = neuron_grad_h_t<SIMPLE>(nrn, f4(hq[t], i), dcnt, gvq[i]);
            f4(dq, i) = ghq * dh_dx;
            gvq[i] = ghq * dh_dv;
            const float ghk = neuron_grad_h_t<SIMPLE>(nrn, f4(hk[t], i), f4(dg[t], i) * av[t], gvk[i]);
            f4(dk, i) = ghk * dh_dx;
            gvk[i] = ghk * dh_dv;
            if (valid) {
              accq[0][i] += f4(dq, i); accq[1][i] += f4(dq, i) * f4(qp[t], i);
              acck[0][i] += f4(dk, i); acck[1][i] += f4(dk, i) * f4(kp[t], i);
            }
          }
          if (valid) {
            const int64_t o = (t * p.MP + nr) * p.C + col;
            st_stream4(p.grad_q + o, dq);
            st_stream4(p.grad_k + o, dk);
          }
        }
      }
    }
  }
  if (BWD && p.part_q) {
    block_reduce_rows_to_partials<2>(accq, smem, p.part_q, p.R, p.k, p.C, (int64_t)blockIdx.y * p.tile_w);
    block_reduce_rows_to_partials<2>(acck, smem, p.part_k, p.R, p.k, p.C, (int64_t)blockIdx.y * p.tile_w);
  }
}

// dpos[t*PC + j] = sum_m gk[(t*M + m)*PC + j]; blockIdx.y splits M, partial sums meet in atomics
__global__ void pos_grad_kernel(const float* __restrict__ gk, float* __restrict__ gp, int64_t wd, int64_t M, int64_t PC,
                                int64_t m_chunk) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= wd * PC) return;
  const int64_t t = i4 / PC, j = i4 - t * PC;
  const int64_t m0 = (int64_t)blockIdx.y * m_chunk;
  const int64_t m1 = m0 + m_chunk < M ? m0 + m_chunk : M;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* base = gk + t * M * PC + j;
  int64_t m = m0;
  for (; m + 3 < m1; m += 4) {
    float4 a = ld_stream4(base + m * PC), b = ld_stream4(base + (m + 1) * PC);
    float4 c = ld_stream4(base + (m + 2) * PC), d = ld_stream4(base + (m + 3) * PC);
    acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
    acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
  }
  for (; m < m1; ++m) {
    float4 a = ld_stream4(base + m * PC);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  atomicAdd(gp + i4, acc.x); atomicAdd(gp + i4 + 1, acc.y); atomicAdd(gp + i4 + 2, acc.z); atomicAdd(gp + i4 + 3, acc.w);
}

// tiling with 8-lane head groups kept inside warps: R % 8 == 0 and (R*k) % 32 == 0
static int qk_tiling(int64_t MP, int64_t C, int target, int64_t max_blocks, RowTiling* rt) {
  SDF_REQUIRE(C % 32 == 0, "qkgate: C=%lld must be a multiple of head_dim 32", (long long)C);
  SDF_REQUIRE(make_row_tiling(MP, C, 4, target, (int)max_blocks, rt), "qkgate: cannot tile C=%lld", (long long)C);
  SDF_REQUIRE(rt->R % 8 == 0, "qkgate: tile width not a multiple of a head");
  int k0 = 1;
  while ((rt->R * k0) % 32 != 0) ++k0;
  int k = (target / rt->R) / k0 * k0;
  if (k < k0) k = k0;
  rt->k = k;
  rt->threads = rt->R * k;
  SDF_REQUIRE(rt->threads <= 512, "qkgate: block too large");
  int64_t need = (MP + k - 1) / k;
  int64_t cap = max_blocks / rt->ncol;
  if (cap < 1) cap = 1;
  rt->blocks = (int)(need < cap ? need : cap);
  return SDF_OK;
}

}  // namespace sdf

using namespace sdf;

static int qk_common(const float* q_pre, const float* k_pre, int64_t ld, const float* qs, const float* qh,
                     const float* ks, const float* kh, const float* pos, int64_t wd, int64_t M, int64_t P,
                     int64_t C, int64_t nH, const sdf_neuron_cfg& nc, QkP* p) {
  SDF_REQUIRE(q_pre && k_pre && qs && qh && ks && kh && pos, "qkgate: null argument");
  SDF_REQUIRE(aligned16(q_pre) && aligned16(k_pre) && aligned16(pos) && ld % 4 == 0 && ld >= C, "qkgate: alignment");
  SDF_REQUIRE(wd >= 1 && wd <= 8 && M > 0 && P > 0, "qkgate: bad window dims");
  SDF_REQUIRE(C == nH * 32, "qkgate: C=%lld must equal num_heads*32", (long long)C);
  int st = validate_neuron(nc);
  if (st) return st;
  p->q_pre = q_pre; p->k_pre = k_pre; p->ld = ld; p->q_scale = qs; p->q_shift = qh; p->k_scale = ks; p->k_shift = kh;
  SDF_REQUIRE(wd * M * P * nH < (int64_t)1 << 31, "qkgate: token*head count exceeds 2^31");
  p->pos = pos; p->M = M; p->P = P; p->C = C; p->nH = nH; p->MP = M * P; p->wd = (int)wd; p->nrn = make_neuron(nc);
  p->dP.init((uint32_t)P); p->dwd.init((uint32_t)wd); p->dnH.init((uint32_t)nH);
  return SDF_OK;
}

extern "C" int sdf_attn_qkgate_fwd(const sdf_attn_qkgate_fwd_args* a) {
  SDF_REQUIRE(a && a->gate && aligned16(a->gate), "sdf_attn_qkgate_fwd: null/unaligned output");
  QkP p = {};
  int st = qk_common(a->q_pre, a->k_pre, a->ld, a->q_scale, a->q_shift, a->k_scale, a->k_shift, a->pos, a->wd, a->M,
                     a->P, a->C, a->nH, a->neuron, &p);
  if (st) return st;
  const int DT = a->spike_dtype;
  SDF_REQUIRE(DT >= 0 && DT <= 2, "sdf_attn_qkgate_fwd: bad spike_dtype");
  p.gate = a->gate; p.q_h = a->q_h; p.k_h = a->k_h; p.a_h = a->a_h;
  RowTiling rt;
  st = qk_tiling(p.MP, p.C, 256, kNumSMs * 3, &rt);
  if (st) return st;
  p.tile_w = rt.tile_w; p.R = rt.R; p.k = rt.k;
  dim3 grid(rt.blocks, rt.ncol, 1);
  cudaStream_t stream = (cudaStream_t)a->stream;
#define QK_FWD(TT, SM)                                                                                        \
  do {                                                                                                        \
    if (DT == SDF_SPIKE_F32) qkgate_kernel<TT, SDF_SPIKE_F32, false, SM><<<grid, rt.threads, 0, stream>>>(p); \
    else if (DT == SDF_SPIKE_U8) qkgate_kernel<TT, SDF_SPIKE_U8, false, SM><<<grid, rt.threads, 0, stream>>>(p); \
    else qkgate_kernel<TT, SDF_SPIKE_BF16, false, SM><<<grid, rt.threads, 0, stream>>>(p);                    \
  } while (0)
  const bool simple = neuron_is_simple(p.nrn);
  if (a->wd == 2 && simple) QK_FWD(2, true);
  else if (a->wd == 2) QK_FWD(2, false);
  else if (a->wd == 4) QK_FWD(4, false);
  else QK_FWD(0, false);
#undef QK_FWD
  return finish_launch("sdf_attn_qkgate_fwd");
}

extern "C" int sdf_attn_qkgate_bwd(const sdf_attn_qkgate_bwd_args* a) {
  SDF_REQUIRE(a && a->grad_gate && a->grad_q && a->grad_k && aligned16(a->grad_gate) && aligned16(a->grad_q) && aligned16(a->grad_k),
              "sdf_attn_qkgate_bwd: null/unaligned argument");
  SDF_REQUIRE((a->bn_partials_q == nullptr) == (a->bn_partials_k == nullptr), "sdf_attn_qkgate_bwd: q/k partials go together");
  QkP p = {};
  int st = qk_common(a->q_pre, a->k_pre, a->ld, a->q_scale, a->q_shift, a->k_scale, a->k_shift, a->pos, a->wd, a->M,
                     a->P, a->C, a->nH, a->neuron, &p);
  if (st) return st;
  p.grad_gate = a->grad_gate; p.grad_q = a->grad_q; p.grad_k = a->grad_k; p.part_q = a->bn_partials_q; p.part_k = a->bn_partials_k;
  int64_t cap = kNumSMs * 2;
  if (p.part_q) {
    SDF_REQUIRE(a->n_partial_blocks >= 1, "sdf_attn_qkgate_bwd: n_partial_blocks < 1");
    if (a->n_partial_blocks < cap) cap = a->n_partial_blocks;
  }
  RowTiling rt;
  st = qk_tiling(p.MP, p.C, 256, cap, &rt);
  if (st) return st;
  p.tile_w = rt.tile_w; p.R = rt.R; p.k = rt.k;
  dim3 grid(rt.blocks, rt.ncol, 1);
  cudaStream_t stream = (cudaStream_t)a->stream;
  if (p.part_q && a->n_partial_blocks > rt.blocks) {
    const size_t off = (size_t)rt.blocks * 2 * a->C, n = sizeof(float) * (a->n_partial_blocks - rt.blocks) * 2 * a->C;
    cudaMemsetAsync(a->bn_partials_q + off, 0, n, stream);
    cudaMemsetAsync(a->bn_partials_k + off, 0, n, stream);
  }
  const size_t smem = sizeof(float) * 4 * rt.threads;
  if (a->wd == 2 && neuron_is_simple(p.nrn)) qkgate_kernel<2, SDF_SPIKE_F32, true, true><<<grid, rt.threads, smem, stream>>>(p);
  else if (a->wd == 2) qkgate_kernel<2, SDF_SPIKE_F32, true, false><<<grid, rt.threads, smem, stream>>>(p);
  else if (a->wd == 4) qkgate_kernel<4, SDF_SPIKE_F32, true, false><<<grid, rt.threads, smem, stream>>>(p);
  else qkgate_kernel<0, SDF_SPIKE_F32, true, false><<<grid, rt.threads, smem, stream>>>(p);
  return finish_launch("sdf_attn_qkgate_bwd");
}

extern "C" int sdf_pos_grad(const sdf_pos_grad_args* a) {
  SDF_REQUIRE(a && a->grad_k && a->grad_pos && aligned16(a->grad_k) && aligned16(a->grad_pos) && a->PC % 4 == 0 && a->wd > 0 && a->M > 0,
              "sdf_pos_grad: bad argument");
  const int threads = 128;
  const int64_t n4 = a->wd * a->PC / 4;
  cudaStream_t stream = (cudaStream_t)a->stream;
  cudaMemsetAsync(a->grad_pos, 0, sizeof(float) * a->wd * a->PC, stream);
  const int64_t bx = (n4 + threads - 1) / threads;
  int64_t ny = (kNumSMs * 8 + bx - 1) / bx;
  if (ny > a->M) ny = a->M;
  if (ny < 1) ny = 1;
  const int64_t m_chunk = (a->M + ny - 1) / ny;
  ny = (a->M + m_chunk - 1) / m_chunk;
  pos_grad_kernel<<<dim3((unsigned)bx, (unsigned)ny, 1), threads, 0, stream>>>(a->grad_k, a->grad_pos, a->wd, a->M, a->PC, m_chunk);
  return finish_launch("sdf_pos_grad");
}
