// attn_qktv.cu — K3/K4: Q K^T V spiking window attention on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Replaces reference models/STSwinNet_SNN/Spiking_swin_transformer3D.py:320-363 / :438-485
// (Spiking_BN_WindowAttention3D / SDSA_WindowAttention3D):
//     attn = (q*scale) @ k^T + relative_position_bias[h'] + mask[w]        (no softmax, :356-358)
//     x    = (attn @ v).reshape(B_,nH,T,H,W,hd).permute(2,0,3,4,1,5)        (:362-363)
// on the raw [M*nH, N, 32] reinterpretation of the (wd,M,wh,ww,C) spike buffers (Appendix B.4):
// the rows of one (window m', pseudo-head h') pair are 32 contiguous bytes each, N*32 B in all.
//
// Two kernel families live here.  v1 (qktv_kernel, below) handles every window size and the debug outputs; v2
// (qktv2_kernel / qktv2_bwd_kernel, second half of the file) is the warp-specialised pipeline for windows of up to
// 176 tokens, where the per-head bias fits in tensor memory — see the comment block above it.
//
// v1: one persistent CTA per SM walks pairs with a fixed pseudo-head.  Per pair:
//   stage  Q, K, V (row-major bytes -> UMMA canonical SWIZZLE_64B layout: 64-B rows, 16-B chunk index XOR
//          (row/2)%4) into shared memory, u8 {0,1} -> fp16 (exact); Q, K are read K-major, V MN-major (no transpose);
//   MMA 1  S = Q K^T           tcgen05.mma kind::f16, M=128, N<=192, K=16 x2, fp32 accum in TMEM
//                              (S are exact integer counts 0..32);
//   epi 1  T = scale*S + bias[lin_i - lin_j + off] + (-100)[region_i != region_j], in registers
//          (tcgen05.ld), split T = hi + lo in fp16 (22 significant bits) and written back IN PLACE
//          over S (tcgen05.st) as the A operand of the second contraction — the N x N matrix never
//          leaves the SM;
//   MMA 2  O += T_hi V + T_lo V   tcgen05.mma kind::f16 with A from TMEM, B = V (MN-major) from smem;
//   epi 2  O rows -> global in the proj-input order of :362-363 (128 B per row).
// The backward (K4) is three more passes of the same two-contraction structure:
//   dA = dO V^T  -> dQ = scale*dA K        (+ d(bias table) accumulated in shared memory)
//   dA^T = V dO^T -> dK = scale*dA^T Q
//   T^T = (K Q^T ...) -> dV = T^T dO
// with bf16 operands (gradients need range, not 22 bits).
//
// Arithmetic intensity: 4*N^2*32 algorithmic FLOP per pair against 3*N*32 B in and
// N*128 B out — HBM-bound at N=162, tensor-bound from N~576; bench reports both.
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "sdf_common.cuh"
#include "tc_ptx.cuh"   // TMA (cp.async.bulk.tensor) wrappers + the host tensor-map encoder, used by the v2 forward

namespace sdf {

constexpr int kSlotThreads = 256;  // 8 warps per 128-lane TMEM slot: four lane quarters x two column halves
constexpr int kKTmax = 192;        // largest key tile (TMEM columns of S per slot); O accumulator sits right after it

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a converged warp; keeps the surrounding code warp-uniform so that descriptors stay in uniform registers
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A * B with A from shared memory (mma_ss_lh) or from TMEM (mma_ts_lh), B from shared memory.
// Descriptors are given as (lo word, shared hi word) so that per-MMA operand changes are one 32-bit add;
// ACC is compile time (enable-input-d)
template <bool ACC>
__device__ __forceinline__ void mma_ss_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}
// u8 x u8 -> s32 (kind::i8), overwrite D
__device__ __forceinline__ void mma_ss_i8_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 0, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc)
      : "memory");
}
template <bool ACC>
__device__ __forceinline__ void mma_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(hi), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// SWIZZLE_64B canonical layout (both majors, 16-bit elements, 32 elements = 64 B per row): rows of 64 B at a
// 64-B pitch, 8-row groups 512 B apart (SBO), and inside the 1024-B-aligned buffer the 16-B chunk index is XORed
// with address bits [7,8] = (row / 2) % 4.  K-major: a K step of 16 elements advances the start address by 32 B;
// MN-major (rows = K index): a K step of 16 rows advances it by 1024 B.  LBO is not used (one swizzle span).
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// the same descriptor as two 32-bit words: per-MMA address changes are a 32-bit add on the low word
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | (1u << 16); }
constexpr uint32_t kDescHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512 B, version 1, SWIZZLE_64B
// instruction descriptor for kind::f16: fp32 accumulate, A/B both K-major, format 0 = F16, 1 = BF16
// b_mn = 1: B operand is MN-major (its N index is the contiguous one)
__device__ __forceinline__ uint32_t make_idesc(int fmt, int M, int N, int b_mn = 0) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- 16-bit operand helpers ---------------------------------------------------------------------
template <int BF>
__device__ __forceinline__ uint16_t one16() { return BF ? 0x3F80 : 0x3C00; }
template <int BF>
__device__ __forceinline__ uint16_t f2h(float x) {
  if (BF) return __bfloat16_as_ushort(__float2bfloat16_rn(x));
  return __half_as_ushort(__float2half_rn(x));
}
template <int BF>
__device__ __forceinline__ float h2f(uint16_t h) {
  if (BF) return __uint_as_float((uint32_t)h << 16);
  return __half2float(__ushort_as_half(h));
}

// two fp32 -> packed 16-bit pair (lo = a, hi = b) in one instruction, and back
template <int BF>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  if (BF) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
template <int BF>
__device__ __forceinline__ void unpack2(uint32_t r, float& a, float& b) {
  if (BF) {
    a = __uint_as_float(r << 16);
    b = __uint_as_float(r & 0xFFFF0000u);
  } else {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r));
    a = f.x; b = f.y;
  }
}

// ---- kernel parameters ----------------------------------------------------------------------------
struct QktvP {
  const uint8_t* q; const uint8_t* k; const uint8_t* v;
  const float* bias_table; const uint8_t* region;
  float* out; int32_t* s_dbg; float* attn_dbg;
  const float* grad_out; float* grad_q; float* grad_k; float* grad_v; float* grad_table;
  int64_t M, nH, nW, P;
  int N, wd, wh, ww, Rpad, n_mt, n_kt, tab, kt, slot_cols, tmem_cols;
  float scale;
  int has_mask;
};

// shared-memory carve-up (bytes), all operand arrays 16-bit
struct SmemPlan {
  uint32_t a, b, bt, lin, reg, tab, dtab, bars, tmem_slot, total;
};
__host__ __device__ inline SmemPlan plan_smem(int Rpad, int tab, bool bwd) {
  SmemPlan s;
  uint32_t o = 0;
  s.a = o; o += (uint32_t)Rpad * 64;          // [4 chunks][Rpad rows][16 B]
  s.b = o; o += (uint32_t)Rpad * 64;
  s.bt = o; o += (uint32_t)Rpad * 64;         // third operand, same layout, read MN-major by MMA 2
  s.lin = o; o += (uint32_t)Rpad * 4;
  s.reg = o; o += (uint32_t)((Rpad + 15) / 16 * 16);
  s.tab = o; o += (uint32_t)((tab * 4 + 15) / 16 * 16);
  s.dtab = o; o += bwd ? (uint32_t)((tab * 4 + 15) / 16 * 16) : 0;
  s.bars = o; o += 32;
  s.tmem_slot = o; o += 16;
  s.total = o;
  return s;
}

// operand tile: element (row r, dim d) in 16-B chunk d/8 of the 64-B row r, SWIZZLE_64B
__device__ __forceinline__ uint32_t plain_off(int /*Rpad*/, int r, int c16) { return (uint32_t)r * 64 + (uint32_t)((c16 ^ ((r >> 1) & 3)) << 4); }

// stage rows [0, N) of a u8 {0,1} [N, 32] block
template <int BF>
__device__ __forceinline__ void stage_plain_u8(uint8_t* smem, uint32_t base, int Rpad, const uint8_t* src, int N) {
  for (int i = threadIdx.x; i < N * 2; i += blockDim.x) {       // one 16-byte half-row per item
    const int r = i >> 1, hf = i & 1;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)r * 32 + hf * 16));
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t b = ws[j];
      o[2 * j] = ((b & 0xFF) ? one16<BF>() : 0) | (((b >> 8) & 0xFF) ? (uint32_t)one16<BF>() << 16 : 0);
      o[2 * j + 1] = (((b >> 16) & 0xFF) ? one16<BF>() : 0) | (((b >> 24) & 0xFF) ? (uint32_t)one16<BF>() << 16 : 0);
    }
    *reinterpret_cast<uint4*>(smem + base + plain_off(Rpad, r, hf * 2)) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(smem + base + plain_off(Rpad, r, hf * 2 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
  }
}
// fp32 rows gathered from the proj-input layout: token n of pair (m', h') lives at row (t*M + m')*P + pos, col h'*32
__device__ __forceinline__ const float* go_row(const QktvP& p, int64_t mwin, int64_t head, int n) {
  const int64_t t = n / p.P, pos = n - t * p.P;
  return p.grad_out + ((t * p.M + mwin) * p.P + pos) * (p.nH * 32) + head * 32;
}
template <int BF>
__device__ __forceinline__ void stage_plain_f32(uint8_t* smem, uint32_t base, int Rpad, const QktvP& p, int64_t mwin, int64_t head) {
  for (int i = threadIdx.x; i < p.N * 4; i += blockDim.x) {     // 8 floats -> one 16-byte chunk
    const int r = i >> 2, c = i & 3;
    const float* src = go_row(p, mwin, head, r) + c * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
    uint4 o;
    o.x = f2h<BF>(a.x) | ((uint32_t)f2h<BF>(a.y) << 16); o.y = f2h<BF>(a.z) | ((uint32_t)f2h<BF>(a.w) << 16);
    o.z = f2h<BF>(b.x) | ((uint32_t)f2h<BF>(b.y) << 16); o.w = f2h<BF>(b.z) | ((uint32_t)f2h<BF>(b.w) << 16);
    *reinterpret_cast<uint4*>(smem + base + plain_off(Rpad, r, c)) = o;
  }
}
// PHASE 0: forward O = T V           A=Q  B=K  Bt=V^T   out -> p.out (permuted rows)
// PHASE 1: dQ = scale*(dO V^T) K     A=dO B=V  Bt=K^T   out -> grad_q, side effect d(bias table)
// PHASE 2: dK = scale*(V dO^T) Q     A=V  B=dO Bt=Q^T   out -> grad_k
// PHASE 3: dV = T^T dO               A=K  B=Q  Bt=dO^T  out -> grad_v      (T^T[j][i]: roles of i, j swapped)
// NSLOT = 2: one 512-thread CTA per SM working on two M-tiles at a time (large windows, smem-bound occupancy);
// NSLOT = 1: 256-thread CTAs, two per SM, each with its own 256 TMEM columns — independent pairs overlap each
// other's staging / MMA / epilogue latencies.
template <int PHASE, bool MASK, bool DBG, int NSLOT>
__global__ void __launch_bounds__(kSlotThreads * NSLOT, NSLOT == 2 ? 1 : 4) qktv_kernel(const QktvP p) {
  constexpr int kThreads = kSlotThreads * NSLOT;
  constexpr int BF = PHASE == 0 ? 0 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemPlan sp = plan_smem(p.Rpad, p.tab, PHASE == 1);
  int* lin_s = reinterpret_cast<int*>(smem + sp.lin);
  uint8_t* reg_s = smem + sp.reg;
  float* tab_s = reinterpret_cast<float*>(smem + sp.tab);
  float* dtab_s = reinterpret_cast<float*>(smem + sp.dtab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sp.tmem_slot);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, Rpad = p.Rpad;
  const int64_t head = blockIdx.x % p.nH;      // fixed pseudo-head per CTA (gridDim.x is a multiple of nH)

  // ---- one-time setup ----
  for (uint32_t i = tid * 16; i < sp.lin; i += kThreads * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  const int A_ = (2 * p.wh - 1) * (2 * p.ww - 1), B_ = 2 * p.ww - 1;
  for (int n = tid; n < Rpad; n += kThreads) {
    const int d = n / (p.wh * p.ww), rem = n - d * (p.wh * p.ww), hh = rem / p.ww, w = rem - hh * p.ww;
    lin_s[n] = n < N ? d * A_ + hh * B_ + w : 0;
    reg_s[n] = 0;
  }
  for (int i = tid; i < p.tab; i += kThreads) {
    tab_s[i] = (PHASE == 0 || PHASE == 3) ? __ldg(p.bias_table + (int64_t)i * p.nH + head) : 0.f;
    if (PHASE == 1) dtab_s[i] = 0.f;
  }
  const int lin_off = (p.wd - 1) * A_ + (p.wh - 1) * B_ + (p.ww - 1);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform copy for the MMA issuer
  uint32_t ph_s = 0, ph_o = 0;

  const int slot = NSLOT == 2 ? (warp >> 2) & 1 : 0;            // which M-tile of the current group of M-tiles
  const int half = NSLOT == 2 ? warp >> 3 : warp >> 2;          // which half of the column chunks this warp takes
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16; // this warp's TMEM lanes
  const uint32_t a_base = smem_u32(smem + sp.a), b_base = smem_u32(smem + sp.b), bt_base = smem_u32(smem + sp.bt);


  const int64_t n_pairs = p.M * p.nH;
  for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x) {
    const int64_t mwin = pair / p.nH;                            // pair % nH == head by construction
    const uint8_t* qp = p.q + pair * N * 32;
    const uint8_t* kp = p.k + pair * N * 32;
    const uint8_t* vp = p.v + pair * N * 32;
    // ---- stage operands ----
    if (PHASE == 0) { stage_plain_u8<BF>(smem, sp.a, Rpad, qp, N); stage_plain_u8<BF>(smem, sp.b, Rpad, kp, N); stage_plain_u8<BF>(smem, sp.bt, Rpad, vp, N); }
    if (PHASE == 1) { stage_plain_f32<BF>(smem, sp.a, Rpad, p, mwin, head); stage_plain_u8<BF>(smem, sp.b, Rpad, vp, N); stage_plain_u8<BF>(smem, sp.bt, Rpad, kp, N); }
    if (PHASE == 2) { stage_plain_u8<BF>(smem, sp.a, Rpad, vp, N); stage_plain_f32<BF>(smem, sp.b, Rpad, p, mwin, head); stage_plain_u8<BF>(smem, sp.bt, Rpad, qp, N); }
    if (PHASE == 3) { stage_plain_u8<BF>(smem, sp.a, Rpad, kp, N); stage_plain_u8<BF>(smem, sp.b, Rpad, qp, N); stage_plain_f32<BF>(smem, sp.bt, Rpad, p, mwin, head); }
    if (MASK && (PHASE == 0 || PHASE == 3)) {
      const uint8_t* rp = p.region + (mwin % p.nW) * N;
      for (int n = tid; n < N; n += kThreads) reg_s[n] = __ldg(rp + n);
    }
    fence_async_smem();
    __syncthreads();

    for (int mt0 = 0; mt0 < p.n_mt; mt0 += NSLOT) {
      const int n_slots = (p.n_mt - mt0) >= NSLOT ? NSLOT : 1;
      const int mt = mt0 + slot;
      const int row = mt * 128 + (warp & 3) * 32 + lane;         // A-operand row handled by this thread
      const bool row_ok = slot < n_slots && row < N;
      const bool warp_ok = slot < n_slots && (mt * 128 + (warp & 3) * 32) < N;   // any valid row in this warp's 32 lanes
      const int lin_i = row_ok ? lin_s[row] : 0;
      const int reg_i = row_ok ? reg_s[row] : 0;
      for (int kt = 0; kt < p.n_kt; ++kt) {
        const int key0 = kt * p.kt;
        int nk = N - key0;
        nk = nk > p.kt ? p.kt : ((nk + 15) & ~15);
        // ---- MMA 1: S[128 x nk] = A[128 x 32] * B[nk x 32]^T, both slots ----
        // Issued by one elected lane of warp 0 while the whole warp runs the (warp-uniform) address arithmetic:
        // descriptors then live in uniform registers and the UTCHMMA stream is not throttled by R2UR round trips.
        if (warp == 0) {
          tc_fence_after();
          const uint32_t idesc = make_idesc(BF, 128, nk);
          const uint32_t a_lo = desc_lo(a_base + (uint32_t)mt0 * 128 * 64), b_lo = desc_lo(b_base + (uint32_t)key0 * 64);
          if (elect_one()) {
            for (int s = 0; s < n_slots; ++s) {
              const uint32_t d = tm_u + s * p.slot_cols;
              mma_ss_lh<false>(d, a_lo + (uint32_t)s * (128 * 64 >> 4), b_lo, kDescHiSw64, idesc);
              mma_ss_lh<true>(d, a_lo + (uint32_t)s * (128 * 64 >> 4) + 2, b_lo + 2, kDescHiSw64, idesc);
            }
            tc_commit(&bars[0]);
          }
          __syncwarp();
        }
        mbar_wait(&bars[0], ph_s);
        ph_s ^= 1;
        tc_fence_after();
        // ---- epilogue 1: S -> T = hi + lo (in place); the two warp halves take alternate 16-column chunks ----
        if (warp_ok) {
          const uint32_t tcol = tmem_base + lane_base + slot * p.slot_cols;
          // bias index is linear in the token coordinates: tab[lin_i - lin_j + off] (PHASE 3: roles swapped).
          // Padding rows/columns need no guard: lin = region = 0 there keeps every index in range, the values
          // stay finite and meet all-zero V^T columns (or are never written out).
          const float* tab_i = tab_s + (PHASE == 3 ? lin_off - lin_i : lin_off + lin_i);
          for (int c0 = half * 16; c0 < nk; c0 += 32) {
            uint32_t r[16], o[16];
            tmem_ld16(tcol + c0, r);
            const int j0 = key0 + c0;
            int lj[16];
            uint32_t rj[4];
            if (PHASE == 0 || PHASE == 3) {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const int4 v4 = *reinterpret_cast<const int4*>(lin_s + j0 + q4 * 4);
                lj[q4 * 4] = v4.x; lj[q4 * 4 + 1] = v4.y; lj[q4 * 4 + 2] = v4.z; lj[q4 * 4 + 3] = v4.w;
              }
              if (MASK) {
                const uint4 r4 = *reinterpret_cast<const uint4*>(reg_s + j0);
                rj[0] = r4.x; rj[1] = r4.y; rj[2] = r4.z; rj[3] = r4.w;
              }
            }
            float t[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float sv = __uint_as_float(r[e]);
              if (PHASE == 0 || PHASE == 3) {
                t[e] = fmaf(sv, p.scale, PHASE == 3 ? tab_i[lj[e]] : tab_i[-lj[e]]);
                if (MASK) {
                  const int rg = (int)((rj[e >> 2] >> ((e & 3) * 8)) & 0xFF);
                  if (rg != reg_i) t[e] -= 100.f;
                }
                if (DBG && PHASE == 0 && row_ok && j0 + e < N) {
                  const int64_t o2 = (pair * N + row) * N + j0 + e;
                  if (p.s_dbg) p.s_dbg[o2] = (int32_t)sv;
                  if (p.attn_dbg) p.attn_dbg[o2] = t[e];
                }
              } else {
                // sv = dA[i][j] (PHASE 1, thread row = query i) or dA^T[j][i] (PHASE 2)
                if (PHASE == 1 && row_ok && j0 + e < N) atomicAdd(&dtab_s[lin_i - lin_s[j0 + e] + lin_off], sv);
                t[e] = sv * p.scale;
              }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t hi = pack2<BF>(t[2 * c], t[2 * c + 1]);
              float ha, hb;
              unpack2<BF>(hi, ha, hb);
              o[c] = hi;
              o[8 + c] = pack2<BF>(t[2 * c] - ha, t[2 * c + 1] - hb);
            }
            tmem_st16(tcol + c0, o);
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncthreads();
        // ---- MMA 2: O[128 x 32] (+)= T_hi * Bt + T_lo * Bt over the 16-key chunks of this tile ----
        if (warp == 0) {
          tc_fence_after();
          // B = the third operand read MN-major (rows = keys): a 16-key step advances the start address by 1024 B
          const uint32_t idesc = make_idesc(BF, 128, 32, 1);
          const uint32_t bt_lo = desc_lo(bt_base + (uint32_t)key0 * 64);
          if (elect_one()) {
            for (int s = 0; s < n_slots; ++s) {
              const uint32_t d = tm_u + s * p.slot_cols + p.kt;
              uint32_t a_hi = tm_u + s * p.slot_cols, bd = bt_lo;
              if (kt == 0) mma_ts_lh<false>(d, a_hi, bd, kDescHiSw64, idesc);
              else mma_ts_lh<true>(d, a_hi, bd, kDescHiSw64, idesc);
              mma_ts_lh<true>(d, a_hi + 8, bd, kDescHiSw64, idesc);
              for (int c0 = 16; c0 < nk; c0 += 16) {
                a_hi += 16; bd += 1024 >> 4;
                mma_ts_lh<true>(d, a_hi, bd, kDescHiSw64, idesc);
                mma_ts_lh<true>(d, a_hi + 8, bd, kDescHiSw64, idesc);
              }
            }
            if (kt == p.n_kt - 1) tc_commit(&bars[1]);
          }
          __syncwarp();
        }
      }
      // ---- epilogue 2: O rows -> global ----
      mbar_wait(&bars[1], ph_o);
      ph_o ^= 1;
      tc_fence_after();
      if (warp_ok) {
        const uint32_t tcol = tmem_base + lane_base + slot * p.slot_cols + p.kt + half * 16;
        uint32_t r0[16];
        tmem_ld16(tcol, r0);
        if (row_ok) {
          float* dst;
          if (PHASE == 0) {
            const int64_t t = row / p.P, pos = row - t * p.P;
            dst = p.out + ((t * p.M + mwin) * p.P + pos) * (p.nH * 32) + head * 32;
          } else {
            float* base = PHASE == 1 ? p.grad_q : (PHASE == 2 ? p.grad_k : p.grad_v);
            dst = base + (pair * N + row) * 32;
          }
          dst += half * 16;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_stream4(dst + c * 4, make_float4(__uint_as_float(r0[4 * c]), __uint_as_float(r0[4 * c + 1]),
                                                __uint_as_float(r0[4 * c + 2]), __uint_as_float(r0[4 * c + 3])));
        }
      }
      tc_fence_before();
      __syncthreads();   // TMEM and smem free for the next M-tile pair / next pair
    }
  }
  if (PHASE == 1) {
    __syncthreads();
    for (int i = tid; i < p.tab; i += kThreads)
      if (dtab_s[i] != 0.f) atomicAdd(p.grad_table + (int64_t)i * p.nH + head, dtab_s[i]);
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}


// ================================================================================================
// K3 v2 — warp-specialised forward for windows whose bias matrix fits in tensor memory
// ================================================================================================
// O = scale * ( S@V + (Bias_h/scale)@V ) - 100 * ( sum_j V_j - sum_{j in region(i)} V_j )
//   * S = Q K^T are exact integers <= 32, so S converts EXACTLY to fp16: the per-element epilogue is one packed
//     cvt per two elements (in place over S in TMEM) instead of bias lookup + mask + hi/lo split;
//   * Bias_h is fixed per CTA (fixed pseudo-head): it is built once, split hi + lo in fp16 (22 significant
//     bits) and stays RESIDENT in TMEM as the A operand of two more MMAs per key chunk;
//   * the shift mask only ever adds -100 * (number of spiking keys of the other regions): 27 region sums of V.
// Operand staging (round 2): Q and K are NOT converted any more — one TMA box each ([npad rows x 32 B] of the raw 1-byte
// spikes, SWIZZLE_32B K-major) lands them in shared memory and the first contraction S = Q K^T runs in tcgen05.mma kind::i8
// (u8 x u8 -> s32, one MMA per tile since K = 32 bytes is exactly one i8 K step); the conversion warps turn the exact integer
// counts into fp16 in place.  Only V (the MN-major B operand of the fp16 second contraction) is still expanded by the
// producer warps — a third of the former staging work.
// Roles: warps 0-3 epilogue (one per TMEM lane quarter), warp 4 MMA issuer, warps 5-8 producers (global ->
// fp16 canonical shared-memory layouts, double buffered).  All hand-offs are mbarriers; S is double buffered in
// TMEM so MMA 1 of item i+1 and MMA 2 of item i overlap the conversion of item i.
// TMEM map (columns): [0,176) Bias hi (two M-tiles x 88), [176,352) Bias lo, [352,416) S0, [416,480) S1, [480,512) O.
constexpr int kV2Threads = 544;               // warps 0-3 S-conversion, 4 MMA, 5-8 + 13-16 producers, 9-12 output
constexpr int kV2Producers = 256;
constexpr int kV2KT = 64;
constexpr int kV2S0 = 352, kV2S1 = 416, kV2O = 480;
constexpr int kV2Regions = 27;
constexpr int kV2Loads = (2 * 176 + kV2Producers - 1) / kV2Producers;   // 16-byte loads of V in flight per producer thread

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

struct V2Plan {
  uint32_t op[2];       // operand buffers: Q (u8, Rpad*32 B) | K (u8, Rpad*32 B) | V (fp16, Rpad*64 B)
  uint32_t rsum;        // 4 output warps x (kV2Regions + 1) x 32 ints: per-warp region sums of V, last row = the warp's total
  uint32_t reg;         // Rpad bytes: region id of every token of the current window
  uint32_t rowoff;      // Rpad x int64: element offset of token row i inside a window's output rows
  uint32_t ostage;      // 4 output warps x 32 rows x 144 B (128 B + 16 B pad: conflict-free both ways)
  uint32_t lin, tab, bars, tmem_slot, total;
};
// op_row_bytes: 128 for the forward (Q, K as raw bytes + fp16 V), 192 for the backward (three 16-bit operand slots)
__host__ __device__ inline V2Plan plan_v2(int Rpad, int tab, int op_row_bytes = 192) {
  V2Plan s;
  uint32_t o = 0;
  for (int b = 0; b < 2; ++b) { s.op[b] = o; o += (uint32_t)Rpad * op_row_bytes; }
  s.rsum = o; o += (4 * (kV2Regions + 1) + kV2Regions) * 32 * 4;   // + combined table: keys outside region r, per dim
  s.reg = o; o += (uint32_t)((Rpad + 15) / 16 * 16);
  s.rowoff = o; o += (uint32_t)Rpad * 8;
  s.ostage = o; o += 4 * 32 * 144;
  s.lin = o; o += (uint32_t)Rpad * 4;
  s.tab = o; o += (uint32_t)((tab * 4 + 15) / 16 * 16);
  s.bars = o; o += 16 * 8;
  s.tmem_slot = o; o += 16;
  s.total = o;
  return s;
}

#ifdef SDF_V2_TRACE
__device__ long long g_v2_trace[16 * 64];
#define V2_TR(slot, it) do { if (blockIdx.x == 0 && (it) < 64) g_v2_trace[(slot) * 64 + (it)] = clock64(); } while (0)
#else
#define V2_TR(slot, it) do { } while (0)
#endif

template <int NPAD, bool MASK>
__global__ void __launch_bounds__(kV2Threads, 1) qktv2_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                             const __grid_constant__ CUtensorMap tmK, const QktvP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const V2Plan sp = plan_v2(p.Rpad, p.tab, 128);
  int* lin_s = reinterpret_cast<int*>(smem + sp.lin);
  float* tab_s = reinterpret_cast<float*>(smem + sp.tab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* full = bars;          // [2] producers -> MMA/epilogue: operands of buffer b are in shared memory
  uint64_t* empty = bars + 2;     // [2] MMA commit + epilogue warps -> producers: buffer b may be overwritten
  uint64_t* s_full = bars + 4;    // [2] MMA commit -> epilogue: S buffer holds fresh counts
  uint64_t* s16_full = bars + 6;  // [2] epilogue -> MMA: S converted to fp16 in place
  uint64_t* o_full = bars + 8;    //     MMA commit -> epilogue: O of the current M-tile is complete
  uint64_t* o_free = bars + 9;    //     epilogue -> MMA: O has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sp.tmem_slot);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, Rpad = p.Rpad;
  constexpr int npad = NPAD;                  // keys padded to the MMA K granularity (compile time: the MMA issue
                                              // sequence is fully unrolled with immediate operands)
  constexpr int w2 = npad >> 1;               // TMEM columns of one fp16 bias tile
  constexpr int n_kt = (npad + kV2KT - 1) / kV2KT;
  constexpr int n_mt = npad > 128 ? 2 : 1;
  constexpr int n_items = n_mt * n_kt;
  const int64_t head = blockIdx.x % p.nH;
  const int A_ = (2 * p.wh - 1) * (2 * p.ww - 1), B_ = 2 * p.ww - 1;
  const int lin_off = (p.wd - 1) * A_ + (p.wh - 1) * B_ + (p.ww - 1);

  // ---- one-time setup (all threads) ----
  for (uint32_t i = tid * 16; i < sp.rsum; i += kV2Threads * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  for (int n = tid; n < Rpad; n += kV2Threads) {
    const int d = n / (p.wh * p.ww), rem = n - d * (p.wh * p.ww), hh = rem / p.ww, w = rem - hh * p.ww;
    lin_s[n] = n < N ? d * A_ + hh * B_ + w : 0;
    smem[sp.reg + n] = 0;
    // token (t = d, pos) of window mwin lives at row (t * M + mwin) * P + pos of the [wd * M * P, C] output
    reinterpret_cast<int64_t*>(smem + sp.rowoff)[n] = ((int64_t)d * p.M * p.P + rem) * (p.nH * 32);
  }
  const float inv_scale = 1.f / p.scale;
  for (int i = tid; i < p.tab; i += kV2Threads) tab_s[i] = __ldg(p.bias_table + (int64_t)i * p.nH + head) * inv_scale;
  if (tid == 0) {
    mbar_init(&full[0], kV2Producers / 32 + 1); mbar_init(&full[1], kV2Producers / 32 + 1);   // + the TMA transaction
    mbar_init(&empty[0], 5); mbar_init(&empty[1], 5);
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    mbar_init(&s16_full[0], 4); mbar_init(&s16_full[1], 4);
    mbar_init(o_full, 1); mbar_init(o_free, 4); mbar_init(&bars[10], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // resident bias: rows of both M-tiles, hi and lo halves, written by the epilogue warps (they own the lanes)
  if (warp < 4) {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int mt = 0; mt < n_mt; ++mt) {
      const int row = mt * 128 + warp * 32 + lane;
      const float* tab_i = tab_s + lin_off + (row < N ? lin_s[row] : 0);
      for (int c0 = 0; c0 < npad; c0 += 16) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int j0 = c0 + 2 * c;
          const float t0 = (row < N && j0 < N) ? tab_i[-lin_s[j0]] : 0.f;
          const float t1 = (row < N && j0 + 1 < N) ? tab_i[-lin_s[j0 + 1]] : 0.f;
          hi[c] = pack2<0>(t0, t1);
          float ha, hb;
          unpack2<0>(hi[c], ha, hb);
          lo[c] = pack2<0>(t0 - ha, t1 - hb);
        }
        // 16 keys = 8 TMEM columns: hi -> [mt*w2 + c0/2, +8), lo -> [(n_mt + mt)*w2 + c0/2, +8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_base + lane_base + mt * w2 + (c0 >> 1)),
                     "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_base + lane_base + (n_mt + mt) * w2 + (c0 >> 1)),
                     "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int64_t n_pairs = p.M * p.nH;


  if ((warp >= 5 && warp < 9) || warp >= 13) {
    // =========================== producers ===========================
    const int pt = (warp < 9 ? warp - 5 : warp - 9) * 32 + lane;  // 0..255
    int pi = 0;
    for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pi) {
      const int b = pi & 1;
      if (pt == 0) V2_TR(10, pi);
      mbar_wait(&empty[b], ((pi >> 1) & 1) ^ 1);
      if (pt == 0) V2_TR(11, pi);
      // Q and K: one TMA box each of the raw spike bytes (rows [pair*N, pair*N + npad) of the [M*nH*N, 32] views; rows past
      // N belong to the next pair and only ever meet zero rows of V / zero bias columns, rows past the tensor are zero-filled)
      if (pt == 0) {
        tc::mbar_expect_tx(&full[b], 2u * npad * 32u);
        tc::tma_load_2d(&tmQ, &full[b], smem_u32(smem + sp.op[b]), 0, (int)(pair * N));
        tc::tma_load_2d(&tmK, &full[b], smem_u32(smem + sp.op[b]) + (uint32_t)Rpad * 32u, 0, (int)(pair * N));
      }
      // V: all 16-byte half-rows of this pair, every load issued first, then expanded to fp16 and stored
      const uint8_t* src = p.v + pair * N * 32;
      const int n_it = 2 * N;
      uint4 w[kV2Loads];
#pragma unroll
      for (int u = 0; u < kV2Loads; ++u) {
        const int g = pt + u * kV2Producers;
        if (g < n_it) w[u] = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)g * 16));
      }
#pragma unroll
      for (int u = 0; u < kV2Loads; ++u) {
        const int g = pt + u * kV2Producers;
        if (g < n_it) {
          const int r = g >> 1, hf = g & 1;
          const uint32_t ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // bytes are 0/1: spread them to 16-bit lanes and multiply by the fp16 pattern of 1.0
            o[2 * j] = __byte_perm(ws[j], 0, 0x4140) * 0x3C00u;
            o[2 * j + 1] = __byte_perm(ws[j], 0, 0x4342) * 0x3C00u;
          }
          const uint32_t base = sp.op[b] + (uint32_t)Rpad * 64;
          *reinterpret_cast<uint4*>(smem + base + plain_off(Rpad, r, hf * 2)) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(smem + base + plain_off(Rpad, r, hf * 2 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[b]);
      if (pt == 0) V2_TR(12, pi);
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the warp-uniform control flow; one elected lane issues.  The tile structure is compile
    // time, so every descriptor / TMEM address below is (per-pair base + immediate): the issue stream is
    // back-to-back UTCHMMA, which is what lets N=32 MMAs run at their 16-cycle floor (tools/ubench/mma_chain.cu).
    {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t kDescHi = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512 B, version 1, SWIZZLE_64B
      constexpr uint32_t idesc2 = (1u << 4) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      int pi = 0, item = 0, tile = 0;
      for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pi) {
        const int b = pi & 1;
        mbar_wait(&full[b], (pi >> 1) & 1);
        tc_fence_after();
        // Q, K: u8 rows of 32 B, SWIZZLE_32B K-major (8-row groups 256 B apart); V: fp16 rows of 64 B, SWIZZLE_64B
        const uint32_t q_lo = ((smem_u32(smem + sp.op[b]) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t k_lo = q_lo + (uint32_t)((Rpad * 32) >> 4), v_lo = k_lo + (uint32_t)((Rpad * 32) >> 4);
        constexpr uint32_t kDescHiSw32 = (256u >> 4) | (1u << 14) | (6u << 29);
#pragma unroll
        for (int li = 0; li < n_items; ++li, ++item) {
          const int sb = item & 1;
          const uint32_t s_cur = tm + kV2S0 + (uint32_t)sb * 64u, s_nxt = tm + kV2S0 + (uint32_t)(sb ^ 1) * 64u;
          V2_TR(0, item);
          // MMA 1 of this item (first item of the pair only) and of the next one: kind::i8, u8 x u8 -> s32, K = 32
#pragma unroll
          for (int lj = (li == 0 ? 0 : li + 1); lj <= li + 1 && lj < n_items; ++lj) {
            const int mt1 = lj / n_kt, kt1 = lj % n_kt;
            const int key1 = kt1 * kV2KT;
            const int nk1 = (npad - key1) < kV2KT ? (npad - key1) : kV2KT;
            const uint32_t idesc1 = (2u << 4) | ((uint32_t)(nk1 >> 3) << 17) | ((128u >> 4) << 24);   // S32 accum, u8 x u8
            const uint32_t d1 = lj == li ? s_cur : s_nxt;
            if (elect_one()) {
              mma_ss_i8_lh(d1, q_lo + (uint32_t)(mt1 * 128 * 32 >> 4), k_lo + (uint32_t)(key1 * 32 >> 4), kDescHiSw32, idesc1);
              tc_commit(&s_full[lj == li ? sb : sb ^ 1]);
            }
            __syncwarp();
          }
          const int mt = li / n_kt, kt = li % n_kt;
          const int key0 = kt * kV2KT;
          const int nk = (npad - key0) < kV2KT ? (npad - key0) : kV2KT;
          V2_TR(3, item);
          mbar_wait(&s16_full[sb], (item >> 1) & 1);
          tc_fence_after();
          V2_TR(1, item);
          if (kt == 0) {
            mbar_wait(o_free, (tile & 1) ^ 1);
            tc_fence_after();
          }
          if (elect_one()) {
            const uint32_t d = tm + kV2O;
#pragma unroll
            for (int c0 = 0; c0 < nk; c0 += 16) {
              // B = V read MN-major: a 16-key step advances the start address by 16 rows * 64 B
              const uint32_t bv = v_lo + (uint32_t)((key0 + c0) * 64 >> 4);
              const uint32_t kc = (uint32_t)((key0 + c0) >> 1);
              if (kt == 0 && c0 == 0) mma_ts_lh<false>(d, tm + mt * w2 + kc, bv, kDescHi, idesc2);
              else mma_ts_lh<true>(d, tm + mt * w2 + kc, bv, kDescHi, idesc2);
              mma_ts_lh<true>(d, tm + (n_mt + mt) * w2 + kc, bv, kDescHi, idesc2);
              mma_ts_lh<true>(d, s_cur + (uint32_t)(c0 >> 1), bv, kDescHi, idesc2);
            }
            if (kt == n_kt - 1) tc_commit(o_full);
          }
          __syncwarp();
          if (kt == n_kt - 1) ++tile;
          V2_TR(2, item);
        }
        if (elect_one()) tc_commit(&empty[b]);
        __syncwarp();
      }
      // drain: every commit above has been delivered once this one has
      if (elect_one()) tc_commit(&bars[10]);
      __syncwarp();
      mbar_wait(&bars[10], 0);
    }
  } else if (warp < 4) {
    // =========================== S-conversion warps (one per TMEM lane quarter) ===========================
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int item = 0;
    for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll
      for (int li = 0; li < n_items; ++li, ++item) {
        const int sb = item & 1;
        const int kt = li % n_kt;
        const int key0 = kt * kV2KT;
        const int nk = (npad - key0) < kV2KT ? (npad - key0) : kV2KT;
        const uint32_t scol = tmem_base + lane_base + (sb ? kV2S1 : kV2S0);
        if (tid == 0) V2_TR(4, item);
        mbar_wait(&s_full[sb], (item >> 1) & 1);
        tc_fence_after();
        if (tid == 0) V2_TR(5, item);
        // S (s32 exact counts 0..32) -> fp16 pairs, in place: this warp is the only one touching these lanes.
        // int -> float without I2F: as_float(n + 0x4B400000) - 1.5 * 2^23 (exact for |n| < 2^22)
        uint32_t r0[32], o[32];
        tmem_ld32(scol, r0);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          o[c] = pack2<0>(__uint_as_float(r0[2 * c] + 0x4B400000u) - 12582912.f, __uint_as_float(r0[2 * c + 1] + 0x4B400000u) - 12582912.f);
        if (nk > 32) {
          tmem_ld32(scol + 32, r0);
#pragma unroll
          for (int c = 0; c < 16; ++c)
            o[16 + c] = pack2<0>(__uint_as_float(r0[2 * c] + 0x4B400000u) - 12582912.f, __uint_as_float(r0[2 * c + 1] + 0x4B400000u) - 12582912.f);
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) o[16 + c] = 0;
        }
        tmem_st32(scol, o);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s16_full[sb]);
        if (tid == 0) V2_TR(6, item);
      }
    }
  } else {
    // =========================== output warps (9..12, lane quarter = warp % 4) ===========================
    const int qw = warp & 3;
    const uint32_t lane_base = (uint32_t)(qw * 32) << 16;
    const int P = (int)p.P;
    const int64_t* rowoff_s = reinterpret_cast<const int64_t*>(smem + sp.rowoff);
    int pi = 0, tile = 0;
    for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pi) {
      const int b = pi & 1;
      const int64_t mwin = pair / p.nH;
      uint8_t rg_next[2] = {0, 0};
      if (MASK) {
        const uint8_t* rp = p.region + (int64_t)((int)mwin % (int)p.nW) * N;
        const int otid = qw * 32 + lane;
        rg_next[0] = __ldg(rp + min(otid, N - 1));
        rg_next[1] = __ldg(rp + min(otid + 128, N - 1));
        mbar_wait(&full[b], (pi >> 1) & 1);
      }
      int* rs_all = reinterpret_cast<int*>(smem + sp.rsum);
      int* rs_w = rs_all + qw * ((kV2Regions + 1) * 32);
      uint8_t* rg = smem + sp.reg;
      if (MASK) {
        // Region sums R[r][d] = sum over the keys j of region r of V[j][d], without atomics: lane = dim d, each
        // output warp walks a quarter of the keys; the region id is warp-uniform per key and keys of one region
        // come in runs, so a run accumulates in a register and is flushed into this warp's private copy.
        const int otid = qw * 32 + lane;
        if (otid == 0) V2_TR(13, pi);
        // region ids of this window's tokens; the tail up to the next multiple of 8 repeats the last id (those
        // rows of V are zero), so the pass below needs no bounds checks
        for (int n = otid; n < ((N + 7) & ~7); n += 128) rg[n] = rg_next[n >= 128];
        for (int r = 0; r <= kV2Regions; ++r) rs_w[r * 32 + lane] = 0;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (otid == 0) V2_TR(14, pi);
        // rows in groups of 8 (one swizzle period): the four chunk offsets are lane constants
        const int per_w = (((N + 3) >> 2) + 7) & ~7, j0 = qw * per_w, j1 = min((N + 7) & ~7, j0 + per_w);
        const uint8_t* vb = smem + sp.op[b] + (uint32_t)Rpad * 64 + (lane & 7) * 2;
        const int c = lane >> 3;
        const uint32_t o0 = (uint32_t)(c ^ 0) << 4, o1 = (uint32_t)(c ^ 1) << 4, o2 = (uint32_t)(c ^ 2) << 4, o3 = (uint32_t)(c ^ 3) << 4;
        int cur = j0 < j1 ? (int)rg[j0] : 0, acc = 0, tot = 0;
        for (int jb = j0; jb < j1; jb += 8) {
          const uint8_t* vr = vb + jb * 64;
          const uint2 r8 = *reinterpret_cast<const uint2*>(rg + jb);
          uint32_t vv[8];
          vv[0] = *reinterpret_cast<const uint16_t*>(vr + 0 * 64 + o0);
          vv[1] = *reinterpret_cast<const uint16_t*>(vr + 1 * 64 + o0);
          vv[2] = *reinterpret_cast<const uint16_t*>(vr + 2 * 64 + o1);
          vv[3] = *reinterpret_cast<const uint16_t*>(vr + 3 * 64 + o1);
          vv[4] = *reinterpret_cast<const uint16_t*>(vr + 4 * 64 + o2);
          vv[5] = *reinterpret_cast<const uint16_t*>(vr + 5 * 64 + o2);
          vv[6] = *reinterpret_cast<const uint16_t*>(vr + 6 * 64 + o3);
          vv[7] = *reinterpret_cast<const uint16_t*>(vr + 7 * 64 + o3);
          const uint32_t cur4 = (uint32_t)cur * 0x01010101u;
          if (r8.x == cur4 && r8.y == cur4) {      // the whole group continues the current run (the common case)
            acc += (int)(((vv[0] + vv[1]) + (vv[2] + vv[3]) + (vv[4] + vv[5]) + (vv[6] + vv[7])) / 0x3C00u);   // fp16 1.0 = 0x3C00
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int rj = (int)(((u < 4 ? r8.x : r8.y) >> ((u & 3) * 8)) & 0xFF);
              if (rj != cur) {                     // warp-uniform branch, once per run of equal region ids
                rs_w[cur * 32 + lane] += acc;
                tot += acc; acc = 0; cur = rj;
              }
              acc += (int)(vv[u] >> 13);
            }
          }
        }
        rs_w[cur * 32 + lane] += acc;
        rs_w[kV2Regions * 32 + lane] = tot + acc;
        if (otid == 0) V2_TR(15, pi);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        // combine the four copies: other[r][d] = (number of spiking keys, dim d) - (those inside region r)
        int* oth = rs_all + 4 * (kV2Regions + 1) * 32;
        for (int idx = otid; idx < kV2Regions * 32; idx += 128) {
          const int d = idx & 31;
          int o = 0;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int* cw = rs_all + w * ((kV2Regions + 1) * 32);
            o += cw[kV2Regions * 32 + d] - cw[idx];
          }
          oth[idx] = o;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (otid == 0) V2_TR(9, pi);
      }
      for (int mt = 0; mt < n_mt; ++mt, ++tile) {
        // O of this M-tile is complete: out = scale * O - 100 * (Vsum - R[region(i)])
        mbar_wait(o_full, tile & 1);
        tc_fence_after();
        if (lane == 0 && qw == 0) V2_TR(7, tile);
        uint32_t ov[32];
        tmem_ld32(tmem_base + lane_base + kV2O, ov);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
        if (lane == 0 && qw == 0) V2_TR(13, tile);
        // scale + mask term, then transpose through shared memory so that each store instruction writes four
        // complete 128-B rows (thread-per-row stores would touch 32 lines per instruction and clog the LSU)
        const int row0 = mt * 128 + qw * 32;
        uint8_t* stg = smem + sp.ostage + (uint32_t)qw * (32 * 144);
        {
          const int row = row0 + lane;
          const int ri = ((MASK && row < N) ? (int)rg[row] : 0) * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 v;
            int4 other = make_int4(0, 0, 0, 0);     // spiking keys outside this row's region, per dim
            if (MASK) other = *reinterpret_cast<const int4*>(rs_all + 4 * (kV2Regions + 1) * 32 + ri + 4 * c);
            v.x = fmaf(__uint_as_float(ov[4 * c]), p.scale, -100.f * (float)other.x);
            v.y = fmaf(__uint_as_float(ov[4 * c + 1]), p.scale, -100.f * (float)other.y);
            v.z = fmaf(__uint_as_float(ov[4 * c + 2]), p.scale, -100.f * (float)other.z);
            v.w = fmaf(__uint_as_float(ov[4 * c + 3]), p.scale, -100.f * (float)other.w);
            *reinterpret_cast<float4*>(stg + lane * 144 + c * 16) = v;
          }
        }
        __syncwarp();
        if (lane == 0 && qw == 0) V2_TR(14, tile);
        {
          const int sub = lane >> 3, ch = lane & 7;
          float* obase = p.out + (mwin * P) * (p.nH * 32) + head * 32 + ch * 4;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rl = sub + 4 * k, row = row0 + rl;
            if (row < N) {
              const float4 v = *reinterpret_cast<const float4*>(stg + rl * 144 + ch * 16);
              st_stream4(obase + rowoff_s[row], v);
            }
          }
        }
        __syncwarp();
        if (lane == 0 && qw == 0) V2_TR(8, tile);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
      if (MASK) asm volatile("bar.sync 2, 128;" ::: "memory");   // region tables are rewritten for the next pair
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ================================================================================================
// K4 v2 — the same warp-specialised pipeline for the three backward contractions (small windows)
// ================================================================================================
// attn = scale S + Bias + Mask is linear in Q, K, V and the bias table, so with dA = dO V^T (per window and head):
//   PHASE 1  dQ = scale (dA K),  dTable[m] = sum over (i, j) with rel(i, j) = m of dA[i][j]
//   PHASE 2  dK = scale (dA^T Q)
//   PHASE 3  dV = attn^T dO = scale (S^T dO + (Bias^T / scale) dO) - 100 (sum_i dO_i - sum_{i in region(j)} dO_i)
// Operand slots (bf16, SWIZZLE_64B): X1 = A of MMA 1 (M-tile rows), X2 = B of MMA 1 (64-row tiles), X3 = B of MMA 2
// (read MN-major):   PHASE 1: dO, V, K     PHASE 2: V, dO, Q     PHASE 3: K, Q, dO.     dO enters in bf16 (as in v1).
// PHASE 1/2: the S-conversion warps turn scale * dA into bf16 hi + lo in place (two MMAs per 16-row chunk).
// PHASE 3: S^T is exact in bf16; Bias^T / scale sits in TMEM as bf16 hi + lo like the forward's bias.
// dTable: instead of one shared-memory atomic per (i, j) per window, dA is ALSO accumulated over all the windows of
// this CTA (fixed head) into a persistent TMEM accumulator (the 352 columns the forward uses for the bias); the
// scatter over relative positions then runs once per CTA.
template <bool ACC_RT>
__device__ __forceinline__ void mma_ss_lh_rt(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}

template <int NPAD, int PHASE, bool MASK>
__global__ void __launch_bounds__(kV2Threads, 1) qktv2_bwd_kernel(const QktvP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const V2Plan sp = plan_v2(p.Rpad, p.tab);
  int* lin_s = reinterpret_cast<int*>(smem + sp.lin);
  float* tab_s = reinterpret_cast<float*>(smem + sp.tab);      // PHASE 3: bias / scale;  PHASE 1: dTable accumulator
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + 2;
  uint64_t* s_full = bars + 4;
  uint64_t* s16_full = bars + 6;
  uint64_t* o_full = bars + 8;
  uint64_t* o_free = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sp.tmem_slot);
  const int64_t* rowoff_s = reinterpret_cast<const int64_t*>(smem + sp.rowoff);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, Rpad = p.Rpad;
  constexpr int npad = NPAD;
  constexpr int w2 = npad >> 1;
  constexpr int n_kt = (npad + kV2KT - 1) / kV2KT;
  constexpr int n_mt = npad > 128 ? 2 : 1;
  constexpr int n_items = n_mt * n_kt;
  // which slot holds dO, and which global spike arrays feed the other two
  constexpr int kGoSlot = PHASE == 1 ? 0 : (PHASE == 2 ? 1 : 2);
  const int64_t head = blockIdx.x % p.nH;
  const int A_ = (2 * p.wh - 1) * (2 * p.ww - 1), B_ = 2 * p.ww - 1;
  const int lin_off = (p.wd - 1) * A_ + (p.wh - 1) * B_ + (p.ww - 1);

  // ---- one-time setup ----
  for (uint32_t i = tid * 16; i < sp.rsum; i += kV2Threads * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  for (int n = tid; n < Rpad; n += kV2Threads) {
    const int d = n / (p.wh * p.ww), rem = n - d * (p.wh * p.ww), hh = rem / p.ww, w = rem - hh * p.ww;
    lin_s[n] = n < N ? d * A_ + hh * B_ + w : 0;
    smem[sp.reg + n] = 0;
    reinterpret_cast<int64_t*>(smem + sp.rowoff)[n] = ((int64_t)d * p.M * p.P + rem) * (p.nH * 32);
  }
  if (PHASE == 3) {
    const float inv_scale = 1.f / p.scale;
    for (int i = tid; i < p.tab; i += kV2Threads) tab_s[i] = __ldg(p.bias_table + (int64_t)i * p.nH + head) * inv_scale;
  } else if (PHASE == 1) {
    for (int i = tid; i < p.tab; i += kV2Threads) tab_s[i] = 0.f;
  }
  if (tid == 0) {
    mbar_init(&full[0], kV2Producers / 32); mbar_init(&full[1], kV2Producers / 32);
    mbar_init(&empty[0], 5); mbar_init(&empty[1], 5);
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    mbar_init(&s16_full[0], 4); mbar_init(&s16_full[1], 4);
    mbar_init(o_full, 1); mbar_init(o_free, 4); mbar_init(&bars[10], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (PHASE == 3 && warp < 4) {
    // resident Bias^T / scale: A-operand row = key j, column = query i, value tab[rel(i, j)]
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int mt = 0; mt < n_mt; ++mt) {
      const int row = mt * 128 + warp * 32 + lane;
      const float* tab_j = tab_s + lin_off - (row < N ? lin_s[row] : 0);
      for (int c0 = 0; c0 < npad; c0 += 16) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int i0 = c0 + 2 * c;
          const float t0 = (row < N && i0 < N) ? tab_j[lin_s[i0]] : 0.f;
          const float t1 = (row < N && i0 + 1 < N) ? tab_j[lin_s[i0 + 1]] : 0.f;
          hi[c] = pack2<1>(t0, t1);
          float ha, hb;
          unpack2<1>(hi[c], ha, hb);
          lo[c] = pack2<1>(t0 - ha, t1 - hb);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_base + lane_base + mt * w2 + (c0 >> 1)),
                     "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_base + lane_base + (n_mt + mt) * w2 + (c0 >> 1)),
                     "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int64_t n_pairs = p.M * p.nH;

  if ((warp >= 5 && warp < 9) || warp >= 13) {
    // =========================== producers ===========================
    const int pt = (warp < 9 ? warp - 5 : warp - 9) * 32 + lane;
    const uint8_t* sa = PHASE == 1 ? p.v : (PHASE == 2 ? p.v : p.k);     // first spike array (slot order)
    const uint8_t* sb_ = PHASE == 1 ? p.k : p.q;                           // second spike array
    constexpr int slot_a = PHASE == 1 ? 1 : 0;                             // PHASE 1: V -> X2; PHASE 2: V -> X1; PHASE 3: K -> X1
    constexpr int slot_b = PHASE == 1 ? 2 : (PHASE == 2 ? 2 : 1);          // PHASE 1: K -> X3; PHASE 2: Q -> X3; PHASE 3: Q -> X2
    constexpr int kGoLoads = (4 * 176 + kV2Producers - 1) / kV2Producers, kSpLoads = (4 * 176 + kV2Producers - 1) / kV2Producers;
    int pi = 0;
    for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pi) {
      const int b = pi & 1;
      mbar_wait(&empty[b], ((pi >> 1) & 1) ^ 1);
      const int64_t mwin = pair / p.nH;
      const float* go = p.grad_out + (mwin * p.P) * (p.nH * 32) + head * 32;
      // dO: one item = 8 floats of a row -> one 16-B bf16 chunk;  spikes: one item = 16 bytes -> two chunks
      float4 g0[kGoLoads], g1[kGoLoads];
      uint4 w[kSpLoads];
      const int n_go = 4 * N, per = 2 * N, n_sp = 2 * per;
#pragma unroll
      for (int u = 0; u < kGoLoads; ++u) {
        const int g = pt + u * kV2Producers;
        if (g < n_go) {
          const float* src = go + rowoff_s[g >> 2] + (g & 3) * 8;
          g0[u] = __ldg(reinterpret_cast<const float4*>(src));
          g1[u] = __ldg(reinterpret_cast<const float4*>(src + 4));
        }
      }
#pragma unroll
      for (int u = 0; u < kSpLoads; ++u) {
        const int g = pt + u * kV2Producers;
        if (g < n_sp) {
          const int a = g >= per ? 1 : 0;
          const int i = g - a * per;
          w[u] = __ldg(reinterpret_cast<const uint4*>((a ? sb_ : sa) + pair * N * 32 + (int64_t)i * 16));
        }
      }
#pragma unroll
      for (int u = 0; u < kGoLoads; ++u) {
        const int g = pt + u * kV2Producers;
        if (g < n_go) {
          const uint4 o = make_uint4(pack2<1>(g0[u].x, g0[u].y), pack2<1>(g0[u].z, g0[u].w), pack2<1>(g1[u].x, g1[u].y), pack2<1>(g1[u].z, g1[u].w));
          *reinterpret_cast<uint4*>(smem + sp.op[b] + (uint32_t)kGoSlot * Rpad * 64 + plain_off(Rpad, g >> 2, g & 3)) = o;
        }
      }
#pragma unroll
      for (int u = 0; u < kSpLoads; ++u) {
        const int g = pt + u * kV2Producers;
        if (g < n_sp) {
          const int a = g >= per ? 1 : 0;
          const int i = g - a * per;
          const int r = i >> 1, hf = i & 1;
          const uint32_t ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[2 * j] = __byte_perm(ws[j], 0, 0x4140) * 0x3F80u;       // bf16 1.0
            o[2 * j + 1] = __byte_perm(ws[j], 0, 0x4342) * 0x3F80u;
          }
          const uint32_t base = sp.op[b] + (uint32_t)(a ? slot_b : slot_a) * Rpad * 64;
          *reinterpret_cast<uint4*>(smem + base + plain_off(Rpad, r, hf * 2)) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(smem + base + plain_off(Rpad, r, hf * 2 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[b]);
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t kDescHi = kDescHiSw64;
    constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);   // bf16, B MN-major
    int pi = 0, item = 0, tile = 0;
    for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pi) {
      const int b = pi & 1;
      mbar_wait(&full[b], (pi >> 1) & 1);
      tc_fence_after();
      const uint32_t x1_lo = desc_lo(smem_u32(smem + sp.op[b]));
      const uint32_t x2_lo = x1_lo + (uint32_t)((Rpad * 64) >> 4), x3_lo = x2_lo + (uint32_t)((Rpad * 64) >> 4);
      const uint32_t acc_p = pi > 0 ? 1u : 0u;
#pragma unroll
      for (int li = 0; li < n_items; ++li, ++item) {
        const int sb = item & 1;
        const uint32_t s_cur = tm + kV2S0 + (uint32_t)sb * 64u, s_nxt = tm + kV2S0 + (uint32_t)(sb ^ 1) * 64u;
#pragma unroll
        for (int lj = (li == 0 ? 0 : li + 1); lj <= li + 1 && lj < n_items; ++lj) {
          const int mt1 = lj / n_kt, kt1 = lj % n_kt;
          const int key1 = kt1 * kV2KT;
          const int nk1 = (npad - key1) < kV2KT ? (npad - key1) : kV2KT;
          const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nk1 >> 3) << 17) | ((128u >> 4) << 24);
          const uint32_t d1 = lj == li ? s_cur : s_nxt;
          const uint32_t a1 = x1_lo + (uint32_t)(mt1 * 128 * 64 >> 4), b1 = x2_lo + (uint32_t)(key1 * 64 >> 4);
          if (elect_one()) {
            mma_ss_lh<false>(d1, a1, b1, kDescHi, idesc1);
            mma_ss_lh<true>(d1, a1 + 2, b1 + 2, kDescHi, idesc1);
            if (PHASE == 1) {     // the same product once more, summed over this CTA's windows (dTable)
              const uint32_t dp = tm + (uint32_t)(mt1 * npad + key1);
              mma_ss_lh_rt<true>(dp, a1, b1, kDescHi, idesc1, acc_p);
              mma_ss_lh<true>(dp, a1 + 2, b1 + 2, kDescHi, idesc1);
            }
            tc_commit(&s_full[lj == li ? sb : sb ^ 1]);
          }
          __syncwarp();
        }
        const int mt = li / n_kt, kt = li % n_kt;
        const int key0 = kt * kV2KT;
        const int nk = (npad - key0) < kV2KT ? (npad - key0) : kV2KT;
        mbar_wait(&s16_full[sb], (item >> 1) & 1);
        tc_fence_after();
        if (kt == 0) {
          mbar_wait(o_free, (tile & 1) ^ 1);
          tc_fence_after();
        }
        if (elect_one()) {
          const uint32_t d = tm + kV2O;
#pragma unroll
          for (int c0 = 0; c0 < nk; c0 += 16) {
            const uint32_t bv = x3_lo + (uint32_t)((key0 + c0) * 64 >> 4);
            if (PHASE == 3) {
              const uint32_t kc = (uint32_t)((key0 + c0) >> 1);
              if (kt == 0 && c0 == 0) mma_ts_lh<false>(d, tm + mt * w2 + kc, bv, kDescHi, idesc2);
              else mma_ts_lh<true>(d, tm + mt * w2 + kc, bv, kDescHi, idesc2);
              mma_ts_lh<true>(d, tm + (n_mt + mt) * w2 + kc, bv, kDescHi, idesc2);
              mma_ts_lh<true>(d, s_cur + (uint32_t)(c0 >> 1), bv, kDescHi, idesc2);
            } else {
              if (kt == 0 && c0 == 0) mma_ts_lh<false>(d, s_cur + (uint32_t)(c0 >> 1), bv, kDescHi, idesc2);
              else mma_ts_lh<true>(d, s_cur + (uint32_t)(c0 >> 1), bv, kDescHi, idesc2);
              mma_ts_lh<true>(d, s_cur + 32u + (uint32_t)(c0 >> 1), bv, kDescHi, idesc2);
            }
          }
          if (kt == n_kt - 1) tc_commit(o_full);
        }
        __syncwarp();
        if (kt == n_kt - 1) ++tile;
      }
      if (elect_one()) tc_commit(&empty[b]);
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bars[10]);
    __syncwarp();
    mbar_wait(&bars[10], 0);
  } else if (warp < 4) {
    // =========================== S-conversion warps ===========================
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int item = 0;
    for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll
      for (int li = 0; li < n_items; ++li, ++item) {
        const int sb = item & 1;
        const int kt = li % n_kt;
        const int key0 = kt * kV2KT;
        const int nk = (npad - key0) < kV2KT ? (npad - key0) : kV2KT;
        const uint32_t scol = tmem_base + lane_base + (sb ? kV2S1 : kV2S0);
        mbar_wait(&s_full[sb], (item >> 1) & 1);
        tc_fence_after();
        uint32_t r0[32], hi[32], lo[32];
        tmem_ld32(scol, r0);
        if (PHASE == 3) {
#pragma unroll
          for (int c = 0; c < 16; ++c) hi[c] = pack2<1>(__uint_as_float(r0[2 * c]), __uint_as_float(r0[2 * c + 1]));   // exact
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float t0 = __uint_as_float(r0[2 * c]) * p.scale, t1 = __uint_as_float(r0[2 * c + 1]) * p.scale;
            hi[c] = pack2<1>(t0, t1);
            float ha, hb;
            unpack2<1>(hi[c], ha, hb);
            lo[c] = pack2<1>(t0 - ha, t1 - hb);
          }
        }
        if (nk > 32) {
          tmem_ld32(scol + 32, r0);
          if (PHASE == 3) {
#pragma unroll
            for (int c = 0; c < 16; ++c) hi[16 + c] = pack2<1>(__uint_as_float(r0[2 * c]), __uint_as_float(r0[2 * c + 1]));
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float t0 = __uint_as_float(r0[2 * c]) * p.scale, t1 = __uint_as_float(r0[2 * c + 1]) * p.scale;
              hi[16 + c] = pack2<1>(t0, t1);
              float ha, hb;
              unpack2<1>(hi[16 + c], ha, hb);
              lo[16 + c] = pack2<1>(t0 - ha, t1 - hb);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) { hi[16 + c] = 0; lo[16 + c] = 0; }
        }
        tmem_st32(scol, hi);                       // both loads are complete: safe to overwrite in place
        if (PHASE != 3) tmem_st32(scol + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s16_full[sb]);
      }
    }
    if (PHASE == 1) {
      // every window of this CTA is in the persistent accumulator: scatter it over the relative positions once.
      // Lanes are consecutive query rows at one key column -> distinct table entries within an instruction.
      tc_fence_after();
      for (int mt = 0; mt < n_mt; ++mt) {
        const int row = mt * 128 + warp * 32 + lane;
        const int lin_i = row < N ? lin_s[row] : 0;
        for (int c0 = 0; c0 < npad; c0 += 32) {
          uint32_t r0[32];
          tmem_ld32(tmem_base + lane_base + (uint32_t)(mt * npad + c0), r0);
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (row < N && c0 + e < N) atomicAdd(&tab_s[lin_off + lin_i - lin_s[c0 + e]], __uint_as_float(r0[e]));
        }
      }
      asm volatile("bar.sync 3, 128;" ::: "memory");
      for (int i = tid; i < p.tab; i += 128)
        if (tab_s[i] != 0.f) atomicAdd(p.grad_table + (int64_t)i * p.nH + head, tab_s[i]);
    }
  } else {
    // =========================== output warps ===========================
    const int qw = warp & 3;
    const uint32_t lane_base = (uint32_t)(qw * 32) << 16;
    float* gout = PHASE == 1 ? p.grad_q : (PHASE == 2 ? p.grad_k : p.grad_v);
    int pi = 0, tile = 0;
    for (int64_t pair = (int64_t)blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pi) {
      const int b = pi & 1;
      const int64_t mwin = pair / p.nH;
      float* rs_all = reinterpret_cast<float*>(smem + sp.rsum);
      float* rs_w = rs_all + qw * ((kV2Regions + 1) * 32);
      uint8_t* rg = smem + sp.reg;
      if (PHASE == 3 && MASK) {
        // float region sums of dO (bf16-rounded, as the MMAs see it): same run-based walk as the forward's
        const int otid = qw * 32 + lane;
        const uint8_t* rp = p.region + (int64_t)((int)mwin % (int)p.nW) * N;
        const uint8_t rg0 = __ldg(rp + min(otid, N - 1)), rg1 = __ldg(rp + min(otid + 128, N - 1));
        mbar_wait(&full[b], (pi >> 1) & 1);
        for (int n = otid; n < ((N + 7) & ~7); n += 128) rg[n] = n >= 128 ? rg1 : rg0;
        for (int r = 0; r <= kV2Regions; ++r) rs_w[r * 32 + lane] = 0.f;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        const int per_w = (((N + 3) >> 2) + 7) & ~7, j0 = qw * per_w, j1 = min((N + 7) & ~7, j0 + per_w);
        const uint8_t* vb = smem + sp.op[b] + 2u * Rpad * 64 + (lane & 7) * 2;
        const int c = lane >> 3;
        const uint32_t o0 = (uint32_t)(c ^ 0) << 4, o1 = (uint32_t)(c ^ 1) << 4, o2 = (uint32_t)(c ^ 2) << 4, o3 = (uint32_t)(c ^ 3) << 4;
        int cur = j0 < j1 ? (int)rg[j0] : 0;
        float acc = 0.f, tot = 0.f;
        for (int jb = j0; jb < j1; jb += 8) {
          const uint8_t* vr = vb + jb * 64;
          const uint2 r8 = *reinterpret_cast<const uint2*>(rg + jb);
          float vv[8];
          vv[0] = __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 0 * 64 + o0) << 16);
          vv[1] = __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 1 * 64 + o0) << 16);
          vv[2] = __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 2 * 64 + o1) << 16);
          vv[3] = __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 3 * 64 + o1) << 16);
          vv[4] = __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 4 * 64 + o2) << 16);
          vv[5] = __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 5 * 64 + o2) << 16);
          vv[6] = __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 6 * 64 + o3) << 16);
          vv[7] = __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 7 * 64 + o3) << 16);
          const uint32_t cur4 = (uint32_t)cur * 0x01010101u;
          if (r8.x == cur4 && r8.y == cur4) {
            acc += ((vv[0] + vv[1]) + (vv[2] + vv[3])) + ((vv[4] + vv[5]) + (vv[6] + vv[7]));
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int rj = (int)(((u < 4 ? r8.x : r8.y) >> ((u & 3) * 8)) & 0xFF);
              if (rj != cur) {
                rs_w[cur * 32 + lane] += acc;
                tot += acc; acc = 0.f; cur = rj;
              }
              acc += vv[u];
            }
          }
        }
        rs_w[cur * 32 + lane] += acc;
        rs_w[kV2Regions * 32 + lane] = tot + acc;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        float* oth = rs_all + 4 * (kV2Regions + 1) * 32;
        for (int idx = otid; idx < kV2Regions * 32; idx += 128) {
          const int d = idx & 31;
          float o = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float* cw = rs_all + w * ((kV2Regions + 1) * 32);
            o += cw[kV2Regions * 32 + d] - cw[idx];
          }
          oth[idx] = o;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      for (int mt = 0; mt < n_mt; ++mt, ++tile) {
        mbar_wait(o_full, tile & 1);
        tc_fence_after();
        uint32_t ov[32];
        tmem_ld32(tmem_base + lane_base + kV2O, ov);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
        const int row0 = mt * 128 + qw * 32;
        uint8_t* stg = smem + sp.ostage + (uint32_t)qw * (32 * 144);
        {
          const int row = row0 + lane;
          const int ri = ((PHASE == 3 && MASK && row < N) ? (int)rg[row] : 0) * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 v;
            if (PHASE == 3) {
              float4 other = make_float4(0.f, 0.f, 0.f, 0.f);
              if (MASK) other = *reinterpret_cast<const float4*>(rs_all + 4 * (kV2Regions + 1) * 32 + ri + 4 * c);
              v.x = fmaf(__uint_as_float(ov[4 * c]), p.scale, -100.f * other.x);
              v.y = fmaf(__uint_as_float(ov[4 * c + 1]), p.scale, -100.f * other.y);
              v.z = fmaf(__uint_as_float(ov[4 * c + 2]), p.scale, -100.f * other.z);
              v.w = fmaf(__uint_as_float(ov[4 * c + 3]), p.scale, -100.f * other.w);
            } else {
              v = make_float4(__uint_as_float(ov[4 * c]), __uint_as_float(ov[4 * c + 1]), __uint_as_float(ov[4 * c + 2]), __uint_as_float(ov[4 * c + 3]));
            }
            *reinterpret_cast<float4*>(stg + lane * 144 + c * 16) = v;
          }
        }
        __syncwarp();
        {
          // gradient rows of one (window, head) are contiguous: [pair * N + row][32]
          const int sub = lane >> 3, ch = lane & 7;
          float* obase = gout + pair * N * 32 + ch * 4;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rl = sub + 4 * k, row = row0 + rl;
            if (row < N) st_stream4(obase + (int64_t)row * 32, *reinterpret_cast<const float4*>(stg + rl * 144 + ch * 16));
          }
        }
        __syncwarp();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
      if (PHASE == 3 && MASK) asm volatile("bar.sync 2, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int NPAD>
static void launch_v2_bwd_npad(bool mask, int grid, uint32_t smem_bytes, cudaStream_t stream, const QktvP& p) {
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    kern<<<grid, kV2Threads, smem_bytes, stream>>>(p);
  };
  if (mask) go(qktv2_bwd_kernel<NPAD, 3, true>); else go(qktv2_bwd_kernel<NPAD, 3, false>);
  go(qktv2_bwd_kernel<NPAD, 1, false>);
  go(qktv2_bwd_kernel<NPAD, 2, false>);
}
static void launch_v2_bwd(int npad, bool mask, int grid, uint32_t smem_bytes, cudaStream_t stream, const QktvP& p) {
  switch (npad) {
    case 16: launch_v2_bwd_npad<16>(mask, grid, smem_bytes, stream, p); break;
    case 32: launch_v2_bwd_npad<32>(mask, grid, smem_bytes, stream, p); break;
    case 48: launch_v2_bwd_npad<48>(mask, grid, smem_bytes, stream, p); break;
    case 64: launch_v2_bwd_npad<64>(mask, grid, smem_bytes, stream, p); break;
    case 80: launch_v2_bwd_npad<80>(mask, grid, smem_bytes, stream, p); break;
    case 96: launch_v2_bwd_npad<96>(mask, grid, smem_bytes, stream, p); break;
    case 112: launch_v2_bwd_npad<112>(mask, grid, smem_bytes, stream, p); break;
    case 128: launch_v2_bwd_npad<128>(mask, grid, smem_bytes, stream, p); break;
    case 144: launch_v2_bwd_npad<144>(mask, grid, smem_bytes, stream, p); break;
    case 160: launch_v2_bwd_npad<160>(mask, grid, smem_bytes, stream, p); break;
    default: launch_v2_bwd_npad<176>(mask, grid, smem_bytes, stream, p); break;
  }
}

template <int NPAD>
static int launch_v2_npad(bool mask, int grid, uint32_t smem_bytes, cudaStream_t stream, const QktvP& p) {
  // Q / K as [M*nH*N rows, 32 bytes] tensors; one box = the npad rows of a (window, pseudo-head) pair
  CUtensorMap tmQ, tmK;
  const uint64_t dims[2] = {32, (uint64_t)(p.M * p.nH * p.N)};
  const uint64_t str[1] = {32};
  const uint32_t box[2] = {32, (uint32_t)NPAD};
  int st = make_tmap(&tmQ, 0, 2, p.q, dims, str, box, nullptr, 32);
  if (st) return st;
  st = make_tmap(&tmK, 0, 2, p.k, dims, str, box, nullptr, 32);
  if (st) return st;
  if (mask) {
    cudaFuncSetAttribute(qktv2_kernel<NPAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    qktv2_kernel<NPAD, true><<<grid, kV2Threads, smem_bytes, stream>>>(tmQ, tmK, p);
  } else {
    cudaFuncSetAttribute(qktv2_kernel<NPAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    qktv2_kernel<NPAD, false><<<grid, kV2Threads, smem_bytes, stream>>>(tmQ, tmK, p);
  }
  return SDF_OK;
}
static int launch_v2(int npad, bool mask, int grid, uint32_t smem_bytes, cudaStream_t stream, const QktvP& p) {
  switch (npad) {
    case 16: return launch_v2_npad<16>(mask, grid, smem_bytes, stream, p);
    case 32: return launch_v2_npad<32>(mask, grid, smem_bytes, stream, p);
    case 48: return launch_v2_npad<48>(mask, grid, smem_bytes, stream, p);
    case 64: return launch_v2_npad<64>(mask, grid, smem_bytes, stream, p);
    case 80: return launch_v2_npad<80>(mask, grid, smem_bytes, stream, p);
    case 96: return launch_v2_npad<96>(mask, grid, smem_bytes, stream, p);
    case 112: return launch_v2_npad<112>(mask, grid, smem_bytes, stream, p);
    case 128: return launch_v2_npad<128>(mask, grid, smem_bytes, stream, p);
    case 144: return launch_v2_npad<144>(mask, grid, smem_bytes, stream, p);
    case 160: return launch_v2_npad<160>(mask, grid, smem_bytes, stream, p);
    default: return launch_v2_npad<176>(mask, grid, smem_bytes, stream, p);
  }
}
#ifdef SDF_V2_TRACE
extern "C" int sdf_debug_v2_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, sdf::g_v2_trace, sizeof(long long) * 16 * 64);
}
#endif
static int qktv_setup(int64_t M, int64_t nH, int64_t nW, int64_t wd, int64_t wh, int64_t ww, double scale, bool bwd,
                      QktvP* p, SmemPlan* sp, int* grid, int* nslot) {
  SDF_REQUIRE(M > 0 && nH > 0 && wd > 0 && wh > 0 && ww > 0, "qktv: bad dims");
  const int64_t N = wd * wh * ww;
  SDF_REQUIRE(N >= 8 && N <= 1024, "qktv: window tokens N=%lld must be in [8, 1024]", (long long)N);
  SDF_REQUIRE(nW > 0 && M % nW == 0, "qktv: M must be a multiple of windows-per-sample nW");
  p->M = M; p->nH = nH; p->nW = nW; p->P = wh * ww; p->N = (int)N; p->wd = (int)wd; p->wh = (int)wh; p->ww = (int)ww;
  p->Rpad = (int)((N + 127) / 128 * 128);
  p->n_mt = p->Rpad / 128;
  p->tab = (int)((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1));
  p->scale = (float)scale;
  *sp = plan_smem(p->Rpad, p->tab, bwd);
  SDF_REQUIRE(sp->total <= 227 * 1024, "qktv: window too large for shared memory (%u B)", sp->total);
  // small windows: 256-thread CTAs with 128 TMEM columns each (key tile 96 + 32 O columns), 3-4 per SM, so that
  // independent pairs overlap each other's staging / MMA / epilogue latencies; large windows: one 512-thread CTA
  // per SM working on two M-tiles (2 x (192 + 32) columns).
  int ctas = 1;
  if (sp->total <= 55 * 1024) { *nslot = 1; ctas = 4; p->kt = 96; p->slot_cols = 128; p->tmem_cols = 128; }
  else if (sp->total <= 72 * 1024) { *nslot = 1; ctas = 3; p->kt = 96; p->slot_cols = 128; p->tmem_cols = 128; }
  else if (sp->total <= 110 * 1024) { *nslot = 1; ctas = 2; p->kt = kKTmax; p->slot_cols = 256; p->tmem_cols = 256; }
  else { *nslot = 2; p->kt = kKTmax; p->slot_cols = 256; p->tmem_cols = 512; }
  p->n_kt = (int)((N + p->kt - 1) / p->kt);
  int g = (kNumSMs * ctas) / (int)nH * (int)nH;
  if (g < nH) g = (int)nH;
  if ((int64_t)g > M * nH) g = (int)(M * nH);
  *grid = g;
  return SDF_OK;
}

}  // namespace sdf

using namespace sdf;

extern "C" int sdf_attn_qktv_fwd(const sdf_attn_qktv_fwd_args* a) {
  SDF_REQUIRE(a && a->q && a->k && a->v && a->bias_table && a->out, "sdf_attn_qktv_fwd: null argument");
  SDF_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->out), "sdf_attn_qktv_fwd: 16-byte alignment");
  QktvP p = {};
  SmemPlan sp;
  int grid, nslot;
  int st = qktv_setup(a->M, a->nH, a->nW, a->wd, a->wh, a->ww, a->scale, false, &p, &sp, &grid, &nslot);
  if (st) return st;
  p.q = a->q; p.k = a->k; p.v = a->v; p.bias_table = a->bias_table; p.region = a->region; p.has_mask = a->region != nullptr;
  p.out = a->out; p.s_dbg = a->s_dbg; p.attn_dbg = a->attn_dbg;
  const bool dbg = a->s_dbg || a->attn_dbg;
  cudaStream_t stream = (cudaStream_t)a->stream;
  {
    // v2 (warp-specialised, bias resident in TMEM) when both M-tiles' fp16 hi/lo bias fit in 352 columns
    static const int use_v2 = [] { const char* e = getenv("SDF_QKTV_V2"); return e ? atoi(e) : 1; }();
    const int npad = (p.N + 15) & ~15;
    const V2Plan vp = plan_v2(p.Rpad, p.tab, 128);
    if (use_v2 && !dbg && p.n_mt * npad <= 352 && vp.total <= 200 * 1024 && a->scale != 0.0) {
      int g = kNumSMs / (int)a->nH * (int)a->nH;
      if (g < a->nH) g = (int)a->nH;
      if ((int64_t)g > a->M * a->nH) g = (int)(a->M * a->nH);
      int st2 = launch_v2(npad, p.has_mask != 0, g, vp.total, stream, p);
      if (st2) return st2;
      return finish_launch("sdf_attn_qktv_fwd(v2)");
    }
  }
#define QKTV_LAUNCH(PH, MK, DB)                                                                                       \
  do {                                                                                                                \
    if (nslot == 1) {                                                                                                 \
      cudaFuncSetAttribute(qktv_kernel<PH, MK, DB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.total);   \
      qktv_kernel<PH, MK, DB, 1><<<grid, kSlotThreads, sp.total, stream>>>(p);                                        \
    } else {                                                                                                          \
      cudaFuncSetAttribute(qktv_kernel<PH, MK, DB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.total);   \
      qktv_kernel<PH, MK, DB, 2><<<grid, kSlotThreads * 2, sp.total, stream>>>(p);                                    \
    }                                                                                                                 \
  } while (0)
  if (dbg) { if (p.has_mask) QKTV_LAUNCH(0, true, true); else QKTV_LAUNCH(0, false, true); }
  else { if (p.has_mask) QKTV_LAUNCH(0, true, false); else QKTV_LAUNCH(0, false, false); }
  return finish_launch("sdf_attn_qktv_fwd");
}

extern "C" int sdf_attn_qktv_bwd(const sdf_attn_qktv_bwd_args* a) {
  SDF_REQUIRE(a && a->q && a->k && a->v && a->bias_table && a->grad_out && a->grad_q && a->grad_k && a->grad_v && a->grad_bias_table,
              "sdf_attn_qktv_bwd: null argument");
  SDF_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->grad_out) && aligned16(a->grad_q) &&
                  aligned16(a->grad_k) && aligned16(a->grad_v), "sdf_attn_qktv_bwd: 16-byte alignment");
  QktvP p = {};
  SmemPlan sp;
  int grid, nslot;
  int st = qktv_setup(a->M, a->nH, a->nW, a->wd, a->wh, a->ww, a->scale, true, &p, &sp, &grid, &nslot);
  if (st) return st;
  p.q = a->q; p.k = a->k; p.v = a->v; p.bias_table = a->bias_table; p.region = a->region; p.has_mask = a->region != nullptr;
  p.grad_out = a->grad_out; p.grad_q = a->grad_q; p.grad_k = a->grad_k; p.grad_v = a->grad_v; p.grad_table = a->grad_bias_table;
  cudaStream_t stream = (cudaStream_t)a->stream;
  {
    // small windows: the warp-specialised kernels (dV, dQ + dTable, dK), same eligibility as the forward's
    static const int use_v2 = [] { const char* e = getenv("SDF_QKTV_V2"); return e ? atoi(e) : 1; }();
    const int npad = (p.N + 15) & ~15;
    const V2Plan vp = plan_v2(p.Rpad, p.tab);
    if (use_v2 && p.n_mt * npad <= 352 && vp.total <= 200 * 1024 && a->scale != 0.0) {
      int g = kNumSMs / (int)a->nH * (int)a->nH;
      if (g < a->nH) g = (int)a->nH;
      if ((int64_t)g > a->M * a->nH) g = (int)(a->M * a->nH);
      launch_v2_bwd(npad, p.has_mask != 0, g, vp.total, stream, p);
      count_launch(); count_launch();
      return finish_launch("sdf_attn_qktv_bwd(v2)");
    }
  }
  QKTV_LAUNCH(1, false, false);
  st = finish_launch("sdf_attn_qktv_bwd(dQ)");
  if (st) return st;
  QKTV_LAUNCH(2, false, false);
  st = finish_launch("sdf_attn_qktv_bwd(dK)");
  if (st) return st;
  if (p.has_mask) QKTV_LAUNCH(3, true, false); else QKTV_LAUNCH(3, false, false);
  return finish_launch("sdf_attn_qktv_bwd(dV)");
}
