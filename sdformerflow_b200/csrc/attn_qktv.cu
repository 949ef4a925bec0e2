// attn_qktv.cu — K3/K4: Q K^T V spiking window attention on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Replaces reference models/STSwinNet_SNN/Spiking_swin_transformer3D.py:320-363 / :438-485
// (Spiking_BN_WindowAttention3D / SDSA_WindowAttention3D):
//     attn = (q*scale) @ k^T + relative_position_bias[h'] + mask[w]        (no softmax, :356-358)
//     x    = (attn @ v).reshape(B_,nH,T,H,W,hd).permute(2,0,3,4,1,5)        (:362-363)
// on the raw [M*nH, N, 32] reinterpretation of the (wd,M,wh,ww,C) spike buffers (Appendix B.4):
// the rows of one (window m', pseudo-head h') pair are 32 contiguous bytes each, N*32 B in all.
//
// One persistent CTA per SM walks pairs with a fixed pseudo-head.  Per pair:
//   stage  Q, K, V (row-major bytes -> UMMA canonical no-swizzle layout [16-B chunk][row][16 B]) into
//          shared memory, u8 {0,1} -> fp16 (exact); Q, K are read K-major, V MN-major (no transpose);
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
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "sdf_common.cuh"

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

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
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

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): core matrix = 8 rows x 16 B
// contiguous; SBO = byte distance between 8-row groups, LBO = byte distance between 16-B K chunks.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
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

// plain operand: element (row r, dim d) at chunk (d/8), row r  -> [c][r][16 B]
__device__ __forceinline__ uint32_t plain_off(int Rpad, int r, int c16) { return (uint32_t)(c16 * Rpad + r) * 16; }

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
  uint32_t ph_s = 0, ph_o = 0;

  const int slot = NSLOT == 2 ? (warp >> 2) & 1 : 0;            // which M-tile of the current group of M-tiles
  const int half = NSLOT == 2 ? warp >> 3 : warp >> 2;          // which half of the column chunks this warp takes
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16; // this warp's TMEM lanes
  const uint32_t a_base = smem_u32(smem + sp.a), b_base = smem_u32(smem + sp.b), bt_base = smem_u32(smem + sp.bt);
  const uint32_t lbo_plain = (uint32_t)Rpad * 16, sbo = 128;

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
        if (tid == 0) {
          tc_fence_after();
          const uint32_t idesc = make_idesc(BF, 128, nk);
          for (int s = 0; s < n_slots; ++s) {
            const uint32_t d = tmem_base + s * p.slot_cols;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t ad = make_desc(a_base + ks * 2 * lbo_plain + (uint32_t)(mt0 + s) * 128 * 16, lbo_plain, sbo);
              const uint64_t bd = make_desc(b_base + ks * 2 * lbo_plain + (uint32_t)key0 * 16, lbo_plain, sbo);
              mma_ss(d, ad, bd, idesc, ks);
            }
          }
          tc_commit(&bars[0]);
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
        if (tid == 0) {
          tc_fence_after();
          // B = the third operand in its plain [dim chunk][token][16 B] layout read MN-major: N (=dim) chunks of 8
          // are SBO = Rpad*16 B apart, K (=token) groups of 8 are LBO = 128 B apart.
          const uint32_t idesc = make_idesc(BF, 128, 32, 1);
          for (int s = 0; s < n_slots; ++s) {
            const uint32_t d = tmem_base + s * p.slot_cols + p.kt;
            for (int c0 = 0; c0 < nk; c0 += 16) {
              const uint64_t bd = make_desc(bt_base + (uint32_t)(key0 + c0) * 16, 128, lbo_plain);
              const uint32_t a_hi = tmem_base + s * p.slot_cols + c0;
              mma_ts(d, a_hi, bd, idesc, (kt > 0 || c0 > 0) ? 1u : 0u);
              mma_ts(d, a_hi + 8, bd, idesc, 1u);
            }
          }
          if (kt == p.n_kt - 1) tc_commit(&bars[1]);
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
  QKTV_LAUNCH(1, false, false);
  st = finish_launch("sdf_attn_qktv_bwd(dQ)");
  if (st) return st;
  QKTV_LAUNCH(2, false, false);
  st = finish_launch("sdf_attn_qktv_bwd(dK)");
  if (st) return st;
  if (p.has_mask) QKTV_LAUNCH(3, true, false); else QKTV_LAUNCH(3, false, false);
  return finish_launch("sdf_attn_qktv_bwd(dV)");
}
