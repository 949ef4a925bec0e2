// spike_wgrad.cu — G3: weight gradient of a Linear / convolution on a spike operand, dW = G^T S, on tcgen05 + TMA.
//
// Replaces the autograd weight-gradient GEMMs of the reference's sj_layer.Linear / sj_layer.Conv2d on spike tensors
// (Spiking_swin_transformer3D.py:126-131,267-290,632-652,909; Spiking_modules.py:268,318,803,845-846), which cuBLAS / cuDNN
// run as [Cout, rows] x [rows, K] contractions over the (huge) row / pixel axis with fp32 spikes.
//
//   dW[co, tap, ci] = sum_p G[p, co] * S[p*stride + off(tap), ci]          p = output pixel (Linear: row, one tap)
//
// D[M = co, N = (tap, ci)] accumulates in TMEM over a slab of pixels; both operands are "MN-major" for the MMA (the
// contraction index is the slow one in memory), which tcgen05 takes directly through the descriptor major bits — no transposes.
// The MMA kind is f16 with bf16 operands: MN-major 16-bit operands use the ordinary SWIZZLE_128B layout and run at the full
// bf16 rate (MN-major TF32, the first version of this kernel, exists only in the 32-byte-atom layout and measured 45 % of the
// TF32 rate: the whole kernel was bound by it).
//   A = G, fp32 in HBM: ONE TMA box of [32 pixels][128 channels] per stage lands raw; four converter warps (thread = output
//       channel = TMEM lane) split every value into bf16 hi + bf16 lo (hi = rn(g), lo = rn(g - hi): 16 significant bits,
//       more than TF32's 11) and write both planes to TENSOR MEMORY (tcgen05.st): the MMA takes A from TMEM, so G costs no
//       shared-memory store and no shared-memory operand read (the kernel is bound by shared-memory bandwidth otherwise);
//       both planes are multiplied into the same accumulator.
//   B = S, 1-byte spikes: TMA brings the raw bytes (4-D box shifted by the tap offset, zero fill outside the image = the
//       convolution padding); the converter warps expand them to bf16 (exact) in the MN-major layout.
// One CTA = one (128-channel M tile, <= 384-column N tile, pixel slab); partial tiles go to a workspace and a second kernel
// reduces the slabs in a fixed order (deterministic) into the parameter's own layout (Linear [Cout,K], Conv OIHW).
#include <cstdlib>
#include "sdf_common.cuh"
#include "tc_ptx.cuh"

namespace sdf {
using namespace tc;

constexpr int kWgConv = 512;         // converter threads (16 warps: the conversion is latency-bound, 8 warps left it exposed)
constexpr int kWgThreads = kWgConv + 64;   // warps 0-15: converters, then epilogue (warps 0-3); then the TMA producer warp and the MMA issuer warp
constexpr int kWgProdWarp = kWgConv / 32, kWgMmaWarp = kWgConv / 32 + 1;
constexpr int kWgRB = 32;            // pixels (GEMM K) per pipeline stage
constexpr int kWgM = 128;            // output channels per tile
constexpr int kWgMaxN = 384;         // columns per tile (TMEM columns of the accumulator; issued as <= 2 MMAs of N <= 256)
constexpr int kWgMaxStages = 8;      // TMA ring (raw G tile + raw spike bytes): deep, the L2/HBM round trip is ~2 us under load
constexpr int kWgBSlots = 2;         // converted-operand ring (A planes in TMEM + B in shared memory)
constexpr int kWgGConv = 128;        // threads 0..127 convert G (one output channel = one TMEM lane each), the rest convert spikes
                                     // (8 G warps, 16 pixels each, measured 5-10 % slower: the spike side is the critical one)
constexpr int kWgACols = 32;         // TMEM columns of one A slot: [hi px 0-15 | hi px 16-31 | lo px 0-15 | lo px 16-31], 8 columns each
constexpr int kWgPatchW = 16, kWgPatchH = 2;
constexpr uint32_t kWgBlk = kWgRB * 128;   // one 64-column MN block of bf16: 32 pixel rows x 128 B

struct WgradP {
  int n_chunks;            // pixel chunks (of 32) in the whole problem
  int chunks_per_slab, n_slabs;
  int n_mtiles, n_ntiles;
  int Cout, Cin;
  int taps_per_tile;       // Cin <= kWgMaxN: taps grouped per tile (all Cin channels each)
  int ci_tiles;            // Cin > kWgMaxN: channel slices per tap (taps_per_tile = 1)
  int ci_width;            // width of a channel slice (multiple of 32)
  int taps;
  int ncols_total;         // taps * Cin: row length of the partial tiles
  int box_w, nbox;         // spike TMA boxes: nbox boxes of box_w (<= 256) channels per tap slice; box_w = staging row pitch
  int max_cols;            // widest tile of this launch (sizes the shared-memory plan)
  int stages;              // TMA ring depth
  int conv, tiles_h, tiles_w, stride;
  int dh[9], dw[9];
  float* partial;          // [n_slabs][Cout][ncols_total]
  float* db_partial;       // [n_slabs][Cout] bias-gradient partial sums (column sums of G), or nullptr
  int halo;                // stride-1 conv, one kernel row per tile: ONE spike box [2][16 + kw - 1][Cin] serves the kw taps
  int binary;              // spikes are 0/1 (cheaper byte -> bf16 expansion)
  int stg_bytes;           // raw spike bytes per stage
  int debug;               // timing experiments (SDF_WGRAD_DEBUG): 1 skip conversion, 2 also skip MMA, 3 MMA only (no TMA)
  // tap tiles: N tile n = k * ci_tiles + (channel slice), tap tile k covers taps [tt_tap0[k], tt_tap0[k] + tt_ntap[k]) and reads G
  // through tensor map tt_cls[k] (one map per output parity class of a transposed convolution; 0 otherwise)
  int n_cls;
  unsigned char tt_tap0[9], tt_ntap[9], tt_cls[9], tt_first[9];   // tt_first: this tap tile writes its class's bias-gradient partial
};

struct WgSmem { uint32_t b, g, stg, bars, tmem_slot, b_slot, g_stage, stg_stage, total; };
__host__ __device__ inline WgSmem wg_smem_plan(int max_cols, int stages, int stg_bytes) {
  WgSmem s;
  uint32_t o = 0;
  s.b_slot = (uint32_t)((max_cols + 63) / 64) * kWgBlk;             // bf16 spikes in whole 64-column MN blocks, <= 24 KB
  s.g_stage = kWgM * kWgRB * 4;                                     // raw fp32 G box: 16 KB
  s.stg_stage = ((uint32_t)stg_bytes + 1023) / 1024 * 1024;         // raw spike bytes
  s.b = o; o += kWgBSlots * s.b_slot;
  s.g = o; o += stages * s.g_stage;
  s.stg = o; o += stages * s.stg_stage;
  s.bars = o; o += (2 * kWgMaxStages + 2 * kWgBSlots + 1) * 8;
  s.tmem_slot = o; o += 16;
  s.total = o;
  return s;
}

// two fp32 -> packed bf16x2 (x in the low half), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16(float x, float y) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
  return d;
}
// hi / lo bf16 split of two values: hi = rn(v), lo = rn(v - hi)
__device__ __forceinline__ void split_bf16(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(x, y);
  lo = pack_bf16(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xFFFF0000u));
}
// four spike bytes -> four bf16 (exact for 0..255): 0x4B000000 | b is 2^23 + b in fp32; the top half of (that - 2^23) is bf16(b)
__device__ __forceinline__ uint2 bytes_to_bf16(uint32_t w) {
  const uint32_t f0 = __float_as_uint(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)) - 8388608.f);
  const uint32_t f1 = __float_as_uint(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)) - 8388608.f);
  const uint32_t f2 = __float_as_uint(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652)) - 8388608.f);
  const uint32_t f3 = __float_as_uint(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653)) - 8388608.f);
  return make_uint2(__byte_perm(f0, f1, 0x7632), __byte_perm(f2, f3, 0x7632));
}

// the same for 0/1 bytes: bf16(1) = 0x3F80, so the high bytes are w * 0x3F and the low bytes w << 7 (no carries between bytes)
__device__ __forceinline__ uint2 bits_to_bf16(uint32_t w) {
  const uint32_t h = w * 0x3Fu, l = w << 7;
  return make_uint2(__byte_perm(l, h, 0x5140), __byte_perm(l, h, 0x7362));
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmG1, const __grid_constant__ CUtensorMap tmG2,
             const __grid_constant__ CUtensorMap tmG3, const __grid_constant__ CUtensorMap tmS, const WgradP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const WgSmem sp = wg_smem_plan(p.max_cols, p.stages, p.stg_bytes);
  const int kWgStages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* full_tma = bars;                                   // TMA landed (raw G + raw spike bytes)         [stages]
  uint64_t* empty = bars + kWgMaxStages;                       // converters finished reading the raw stage   [stages]
  uint64_t* full_b = bars + 2 * kWgMaxStages;                  // converters finished an operand slot         [kWgBSlots]
  uint64_t* empty_b = bars + 2 * kWgMaxStages + kWgBSlots;     // MMAs reading the operand slot retired       [kWgBSlots]
  uint64_t* done = bars + 2 * kWgMaxStages + 2 * kWgBSlots;    // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sp.tmem_slot);
  const int tid = threadIdx.x, warp = tid >> 5;

  // tile decode: blockIdx.x = (slab * n_mtiles + m_tile) * n_ntiles + n_tile  (N tiles of one slab adjacent: they share G in L2)
  const int n_tile = blockIdx.x % p.n_ntiles;
  const int m_tile = (blockIdx.x / p.n_ntiles) % p.n_mtiles;
  const int slab = blockIdx.x / (p.n_ntiles * p.n_mtiles);
  // this tile's columns: taps [tap0, tap0+ntap) x channels [ci0, ci0+width)
  const int tk = n_tile / p.ci_tiles, cit = n_tile - tk * p.ci_tiles;
  const int tap0 = p.tt_tap0[tk], ntap = p.tt_ntap[tk], cls = p.tt_cls[tk];
  const int ci0 = cit * p.ci_width, width = min(p.ci_width, p.Cin - ci0);
  const CUtensorMap* tmGc = cls == 0 ? &tmG : cls == 1 ? &tmG1 : cls == 2 ? &tmG2 : &tmG3;
  const int ncols = ntap * width;               // multiple of 16 (host-checked), <= kWgMaxN
  // <= 2 MMAs per K step: N0 + N1 = ncols, both multiples of 16 and <= 256, N0 a multiple of 64 (whole MN blocks)
  const int N0 = ncols <= 256 ? ncols : ((ncols / 2 + 63) / 64) * 64;
  const int N1 = ncols - N0;
  const int c_begin = slab * p.chunks_per_slab;
  const int c_end = min(p.n_chunks, c_begin + p.chunks_per_slab);
  const int n_iter = c_end - c_begin;

  if (tid == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(&full_tma[i], 1); mbar_init(&empty[i], kWgConv / 32); }
    for (int i = 0; i < kWgBSlots; ++i) { mbar_init(&full_b[i], kWgConv / 32); mbar_init(&empty_b[i], 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == kWgProdWarp && elect_one()) { tma_prefetch_desc(tmGc); tma_prefetch_desc(&tmS); }
  if (warp == kWgMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t b_base = smem_u32(smem + sp.b);
  const uint32_t g_base = smem_u32(smem + sp.g), stg_base = smem_u32(smem + sp.stg);
  const uint32_t kBSlot = sp.b_slot, kGStage = sp.g_stage, kStgStage = sp.stg_stage;

  if (warp == kWgProdWarp) {
    // ===== TMA producer: lane 0 waits for the stage, posts the byte count and issues the G box; lanes 4.. issue ONE spike box
    // each (a bulk-tensor copy costs its issuing thread a few hundred cycles) =====
    const int lane = tid & 31;
    const int per_img = p.tiles_h * p.tiles_w;
    const int n_sbox = p.halo ? 1 : ntap * p.nbox;
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kWgStages;
      const uint32_t ph = (it / kWgStages) & 1;
      if (lane == 0) {
        mbar_wait(&empty[s], ph ^ 1);
        if (p.debug == 3) mbar_arrive(&full_tma[s]);
        else mbar_expect_tx(&full_tma[s], kGStage + (uint32_t)(p.halo ? p.stg_bytes : n_sbox * p.box_w * kWgRB));
      }
      __syncwarp();
      if (p.debug == 3) continue;
      const int chunk = c_begin + it;
      if (!p.conv) {
        if (lane == 0) tma_load_2d(tmGc, &full_tma[s], g_base + s * kGStage, m_tile * kWgM, chunk * kWgRB);
        else if (lane >= 4 && lane - 4 < p.nbox) {
          const int j = lane - 4;
          tma_load_2d(&tmS, &full_tma[s], stg_base + s * kStgStage + j * (p.box_w * kWgRB), ci0 + j * p.box_w, chunk * kWgRB);
        }
      } else {
        const int img = chunk / per_img, rem = chunk - img * per_img;
        const int py = rem / p.tiles_w, px = rem - py * p.tiles_w;
        const int w0 = px * kWgPatchW, h0 = py * kWgPatchH;
        if (lane == 0) tma_load_4d(tmGc, &full_tma[s], g_base + s * kGStage, m_tile * kWgM, w0, h0, img);
        else if (p.halo) {
          if (lane == 4) tma_load_4d(&tmS, &full_tma[s], stg_base + s * kStgStage, ci0, w0 + p.dw[tap0], h0 + p.dh[tap0], img);
        } else if (lane >= 4 && lane - 4 < n_sbox) {
          const int b = lane - 4, t = b / p.nbox, j = b - t * p.nbox;
          tma_load_4d(&tmS, &full_tma[s], stg_base + s * kStgStage + b * (p.box_w * kWgRB), ci0 + j * p.box_w,
                      w0 * p.stride + p.dw[tap0 + t], h0 * p.stride + p.dh[tap0 + t], img);
        }
      }
      __syncwarp();
    }
  } else if (warp == kWgMmaWarp) {
    if (elect_one()) {
      const uint32_t idesc0 = idesc_bf16(kWgM, N0, 0, 1);      // A: TMEM (K-major by construction), B: MN-major
      const uint32_t idesc1 = idesc_bf16(kWgM, N1 > 0 ? N1 : 16, 0, 1);
      constexpr uint32_t hi = desc_hi_sw128(1024);     // K groups of 8 pixels are 8 x 128 B apart
      for (int it = 0; it < n_iter; ++it) {
        const int bs = it % kWgBSlots;
        const uint32_t bph = (it / kWgBSlots) & 1;
        mbar_wait(&full_b[bs], bph);
        tc_fence_after();
        if (p.debug != 2) {
#pragma unroll
          for (int g = 0; g < kWgRB / 16; ++g) {       // bf16 MMA: K = 16 pixels = two 8-row groups, 2048 B per step
            const uint32_t b_lo = desc_lo(b_base + bs * kBSlot + g * 2048, kWgBlk);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {           // G hi plane, then G lo plane, into the same accumulator
              const uint32_t a_t = tmem_base + kWgMaxN + bs * kWgACols + pl * 16 + g * 8;
              const uint32_t acc = (it | g | pl) != 0 ? 1u : 0u;
              mma_ts_f16(tmem_base, a_t, b_lo, hi, idesc0, acc);
              if (N1 > 0) mma_ts_f16(tmem_base + N0, a_t, b_lo + (uint32_t)(N0 / 64) * (kWgBlk >> 4), hi, idesc1, acc);
            }
          }
        }
        tc_commit(&empty_b[bs]);
      }
      tc_commit(done);
    }
  } else {
    // ===== converters =====
    // threads 0..127 (warps 0-3): G.  Thread = output channel = TMEM lane: 32 pixel values of the raw box (a warp reads 128
    //   contiguous bytes per pixel row) -> bf16 hi / lo pairs -> 2 x 16 TMEM columns of the A slot.
    // threads 128..511: spikes.  Raw bytes -> bf16 in the MN-major SWIZZLE_128B layout (64-column blocks of [32 pixels][128 B],
    //   16-byte chunk index XOR (pixel & 7)); thread -> (fixed 8-spike word q of the tile row, row sub-phase).
    const bool g_thread = tid < kWgGConv;
    const int tid2 = tid - kWgGConv;
    const int q_per_row = ncols >> 3;           // 8-spike words per pixel row, <= 48
    const int rows_par = (kWgConv - kWgGConv) / q_per_row;   // pixel rows converted in parallel (>= 8)
    const int q = g_thread ? 0 : tid2 % q_per_row, rsub = g_thread ? kWgRB : tid2 / q_per_row;
    const int wq8 = width >> 3;
    const int tq = q / wq8, qc = q - tq * wq8;  // tap-local index, word within the tap's channels
    const int ch = qc * 8, bj = ch / p.box_w;   // channel within the tap slice -> (spike box, offset in the box row)
    // halo box: pixel (h, w) of tap tq is row h * (16 + ntap - 1) + w + tq of the one staged box
    const uint32_t src_off = p.halo ? (uint32_t)(tq * p.box_w + ch)
                                    : (uint32_t)((tq * p.nbox + bj) * (p.box_w * kWgRB) + (ch - bj * p.box_w));
    const int halo_extra = p.halo ? ntap - 1 : 0;
    const int c = q & 7;
    const uint32_t dst_blk = (uint32_t)(q >> 3) * kWgBlk;
    const bool conv_thread = !g_thread && rsub < rows_par;
    const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kWgMaxN;
    float bsum = 0.f;                           // G threads: column sum of this slab's G (bias gradient), fixed order
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kWgStages, bs = it % kWgBSlots;
      const uint32_t ph = (it / kWgStages) & 1, bph = (it / kWgBSlots) & 1;
      if ((tid & 31) == 0) {                    // one polling lane per warp
        mbar_wait(&empty_b[bs], bph ^ 1);       // the MMAs that read this operand slot two iterations ago have retired
        mbar_wait(&full_tma[s], ph);
      }
      __syncwarp();
      if (p.debug == 0) {
        if (g_thread) {
          tc_fence_after();
          const uint8_t* gsrc = smem + sp.g + s * kGStage + tid * 4;
          uint32_t h[16], l[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x0 = *reinterpret_cast<const float*>(gsrc + (2 * j) * 512);
            const float x1 = *reinterpret_cast<const float*>(gsrc + (2 * j + 1) * 512);
            bsum += x0 + x1;
            split_bf16(x0, x1, h[j], l[j]);
          }
          tmem_st16(a_lane + bs * kWgACols, h);
          tmem_st16(a_lane + bs * kWgACols + 16, l);
          tmem_st_wait_all();
          tc_fence_before();
        } else if (conv_thread) {
          const uint8_t* stg = smem + sp.stg + s * kStgStage + src_off;
          uint8_t* bdst = smem + sp.b + bs * kBSlot + dst_blk;
#pragma unroll 4
          for (int r = rsub; r < kWgRB; r += rows_par) {
            const uint2 w = *reinterpret_cast<const uint2*>(stg + (r + (r >> 4) * halo_extra) * p.box_w);
            const uint2 lo = p.binary ? bits_to_bf16(w.x) : bytes_to_bf16(w.x), hi2 = p.binary ? bits_to_bf16(w.y) : bytes_to_bf16(w.y);
            *reinterpret_cast<uint4*>(bdst + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(lo.x, lo.y, hi2.x, hi2.y);
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (elect_one()) { mbar_arrive(&full_b[bs]); mbar_arrive(&empty[s]); }
    }
    // ===== epilogue (warps 0-3): TMEM -> this slab's partial tile =====
    if (warp < 4) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int co = m_tile * kWgM + tid;
      if (p.db_partial != nullptr && cit == 0 && p.tt_first[tk] && co < p.Cout)
        p.db_partial[((int64_t)slab * p.n_cls + cls) * p.Cout + co] = bsum;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      float* dst = p.partial + ((int64_t)slab * p.Cout + co) * p.ncols_total + (int64_t)tap0 * p.Cin + ci0;
      for (int cc = 0; cc < ncols; cc += 16) {
        uint32_t v[16];
        tmem_ld16_nowait(trow + cc, v);
        tmem_ld_wait();
        if (co < p.Cout) {
          // columns of a tile are (tap-local, channel) pairs: tap t's channels sit Cin apart in the partial row
          const int t = cc / width, cch = cc - t * width;
          float* d = dst + (int64_t)t * p.Cin + cch;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(d + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWgMmaWarp) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// dW[co*s_co + ci*s_ci + tap*s_tap] (+)= sum_s partial[s][co][tap*Cin + ci], slabs added in index order
// tap_map: destination tap of local tap t (the launch's tap order need not be the parameter's); channels ci >= cin_store are
// dropped (operand channels appended as padding); n_cls: bias-gradient partial rows per slab (one per parity class)
struct WgTapMap { int t[9]; };
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int n_slabs, int Cout, int Cin,
                                    int taps, int64_t s_co, int64_t s_ci, int64_t s_tap, int accumulate,
                                    const float* __restrict__ db_partial, float* __restrict__ db, const WgTapMap tap_map,
                                    int cin_store, int n_cls) {
  const int64_t total = (int64_t)Cout * taps * Cin;
  const int64_t total_all = total + (db != nullptr ? Cout : 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_all; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= total) {                           // bias gradient: slabs added in index order
      const int64_t co = i - total;
      float acc = 0.f;
      for (int s = 0; s < n_slabs * n_cls; ++s) acc += __ldg(db_partial + (int64_t)s * Cout + co);
      db[co] = accumulate ? db[co] + acc : acc;
      continue;
    }
    float acc = 0.f;
    for (int s = 0; s < n_slabs; ++s) acc += __ldg(partial + (int64_t)s * total + i);
    const int64_t co = i / ((int64_t)taps * Cin);
    const int64_t rem = i - co * taps * Cin;
    const int64_t tap = rem / Cin, ci = rem - tap * Cin;
    if (ci >= cin_store) continue;
    float* d = dw + co * s_co + ci * s_ci + tap_map.t[tap] * s_tap;
    *d = accumulate ? *d + acc : acc;
  }
}

}  // namespace sdf

using namespace sdf;

// N tiling: whole taps grouped up to kWgMaxN columns (balanced), or — for wide layers — channel slices of a tap
static int wgrad_tiles(int Cin, int taps, int* taps_per_tile, int* ci_tiles, int* ci_width, int* n_ntiles) {
  if (Cin <= kWgMaxN) {
    const int max_tpt = kWgMaxN / Cin;
    const int tiles = (taps + max_tpt - 1) / max_tpt;
    *taps_per_tile = (taps + tiles - 1) / tiles;
    *ci_tiles = 1;
    *ci_width = Cin;
    *n_ntiles = (taps + *taps_per_tile - 1) / *taps_per_tile;
  } else {
    *taps_per_tile = 1;
    *ci_tiles = (Cin + kWgMaxN - 1) / kWgMaxN;
    *ci_width = ((Cin + *ci_tiles - 1) / *ci_tiles + 31) / 32 * 32;    // balanced slices, whole 32-column blocks
    *ci_tiles = (Cin + *ci_width - 1) / *ci_width;
    *n_ntiles = taps * *ci_tiles;
  }
  return 0;
}

// pixel slabs so that the grid is ONE wave of CTAs (tiles * slabs <= SMs; one CTA per SM), at least 4 chunks per slab
static void wgrad_slabs(WgradP& p) {
  const int tiles = p.n_mtiles * p.n_ntiles;
  int slabs = num_sms() / tiles;
  if (slabs < 1) slabs = 1;
  int max_slabs = (p.n_chunks + 3) / 4;
  if (max_slabs < 1) max_slabs = 1;
  if (slabs > max_slabs) slabs = max_slabs;
  p.chunks_per_slab = (p.n_chunks + slabs - 1) / slabs;
  p.n_slabs = (p.n_chunks + p.chunks_per_slab - 1) / p.chunks_per_slab;
}

extern "C" int64_t sdf_spike_wgrad_workspace_bytes(int64_t rows_or_pixels, int64_t Cout, int64_t Cin, int64_t taps) {
  WgradP p{};
  p.n_mtiles = (int)((Cout + kWgM - 1) / kWgM);
  wgrad_tiles((int)Cin, (int)taps, &p.taps_per_tile, &p.ci_tiles, &p.ci_width, &p.n_ntiles);
  p.n_chunks = (int)((rows_or_pixels + kWgRB - 1) / kWgRB);   // conv callers pass whole 2 x 16 patches
  wgrad_slabs(p);
  const int64_t slabs = p.n_slabs;
  return slabs * Cout * (taps * Cin + 1) * 4;     // partial tiles + the bias-gradient partial sums
}

static int wgrad_debug_mode() {
  static int m = [] { const char* e = getenv("SDF_WGRAD_DEBUG"); return e ? atoi(e) : 0; }();
  return m;
}

// uniform tap tiles of an ordinary launch: taps_per_tile taps each, one G map
static void wgrad_default_tap_tiles(WgradP& p) {
  if (p.n_cls > 0) return;                       // the caller filled the table (transposed convolution)
  p.n_cls = 1;
  const int n = p.ci_tiles > 1 ? p.taps : (p.taps + p.taps_per_tile - 1) / p.taps_per_tile;
  for (int k = 0; k < n; ++k) {
    const int tpt = p.ci_tiles > 1 ? 1 : p.taps_per_tile;
    p.tt_tap0[k] = (unsigned char)(k * tpt);
    p.tt_ntap[k] = (unsigned char)(p.taps - k * tpt < tpt ? p.taps - k * tpt : tpt);
    p.tt_cls[k] = 0;
    p.tt_first[k] = k == 0;
  }
}

static int wgrad_launch(WgradP& p, const CUtensorMap* tmG, const CUtensorMap& tmS, float* dw, int64_t s_co, int64_t s_ci,
                        int64_t s_tap, int accumulate, int64_t ws_bytes, cudaStream_t st, const char* what, float* db,
                        const int* tap_map = nullptr, int cin_store = -1) {
  wgrad_default_tap_tiles(p);
  wgrad_slabs(p);
  p.debug = wgrad_debug_mode();
  SDF_REQUIRE((int64_t)p.n_slabs * p.Cout * (p.ncols_total + p.n_cls) * 4 <= ws_bytes, "%s: workspace too small (%lld needed)", what,
              (long long)p.n_slabs * p.Cout * (p.ncols_total + p.n_cls) * 4);
  p.db_partial = db != nullptr ? p.partial + (int64_t)p.n_slabs * p.Cout * p.ncols_total : nullptr;
  if (p.max_cols == 0) p.max_cols = p.ci_tiles > 1 ? p.ci_width : p.taps_per_tile * p.Cin;
  SDF_REQUIRE(p.max_cols % 16 == 0 && p.max_cols <= kWgMaxN && p.box_w % 16 == 0, "%s: unsupported column tiling (%d columns, box %d)", what, p.max_cols, p.box_w);
  if (!p.halo) p.stg_bytes = p.max_cols * kWgRB;
  p.stages = 2;
  for (int st_ = kWgMaxStages; st_ >= 2; --st_)
    if (wg_smem_plan(p.max_cols, st_, p.stg_bytes).total <= 220 * 1024) { p.stages = st_; break; }
  static bool attr_done = false;
  const WgSmem sp = wg_smem_plan(p.max_cols, p.stages, p.stg_bytes);
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return SDF_ERR_CUDA; }
    attr_done = true;
  }
  wgrad_kernel<<<p.n_slabs * p.n_mtiles * p.n_ntiles, kWgThreads, sp.total, st>>>(tmG[0], tmG[p.n_cls > 1 ? 1 : 0], tmG[p.n_cls > 2 ? 2 : 0],
                                                                                  tmG[p.n_cls > 3 ? 3 : 0], tmS, p);
  int r = finish_launch(what);
  if (r) return r;
  const int64_t total = (int64_t)p.Cout * (p.ncols_total + 1);
  const int blocks = (int)((total + 255) / 256 < 4 * num_sms() ? (total + 255) / 256 : 4 * num_sms());
  WgTapMap tm;
  for (int i = 0; i < 9; ++i) tm.t[i] = tap_map != nullptr && i < p.taps ? tap_map[i] : i;
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(p.partial, dw, p.n_slabs, p.Cout, p.Cin, p.taps, s_co, s_ci, s_tap, accumulate,
                                              p.db_partial, db, tm, cin_store < 0 ? p.Cin : cin_store, p.n_cls);
  return finish_launch(what);
}

extern "C" int sdf_spike_wgrad(const sdf_spike_wgrad_args* a) {
  SDF_REQUIRE(a->g && a->s && a->dw && a->workspace, "spike_wgrad: null pointer");
  SDF_REQUIRE(a->rows > 0 && a->Cout > 0 && a->K > 0, "spike_wgrad: empty problem");
  SDF_REQUIRE(a->Cout % 4 == 0 && a->K % 16 == 0, "spike_wgrad: Cout %% 4 and K %% 16 must be 0");
  SDF_REQUIRE(aligned16(a->g) && aligned16(a->s) && aligned16(a->workspace), "spike_wgrad: pointers must be 16-byte aligned");
  WgradP p{};
  p.Cout = (int)a->Cout; p.Cin = (int)a->K; p.taps = 1; p.ncols_total = (int)a->K;
  p.n_mtiles = (p.Cout + kWgM - 1) / kWgM;
  wgrad_tiles(p.Cin, 1, &p.taps_per_tile, &p.ci_tiles, &p.ci_width, &p.n_ntiles);
  p.n_chunks = (int)((a->rows + kWgRB - 1) / kWgRB);
  p.partial = a->workspace;
  p.binary = a->s_max == 1;
  p.nbox = p.ci_width > 256 ? 2 : 1;
  p.box_w = p.ci_width / p.nbox;
  CUtensorMap tmG, tmS;
  {
    const uint64_t dims[2] = {(uint64_t)a->Cout, (uint64_t)a->rows};
    const uint64_t str[1] = {(uint64_t)a->ldg * 4};
    const uint32_t box[2] = {(uint32_t)kWgM, (uint32_t)kWgRB};
    int st = make_tmap(&tmG, 1, 2, a->g, dims, str, box, nullptr, 0);
    if (st) return st;
  }
  {
    const int width = p.box_w;
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->rows};
    const uint64_t str[1] = {(uint64_t)a->K};
    const uint32_t box[2] = {(uint32_t)width, (uint32_t)kWgRB};
    int st = make_tmap(&tmS, 0, 2, a->s, dims, str, box, nullptr, 0);
    if (st) return st;
  }
  return wgrad_launch(p, &tmG, tmS, a->dw, a->K, 1, 0, a->accumulate, a->workspace_bytes, (cudaStream_t)a->stream, "sdf_spike_wgrad",
                      a->db);
}

extern "C" int sdf_spike_conv_wgrad(const sdf_spike_conv_wgrad_args* a) {
  SDF_REQUIRE(a->g && a->x && a->dw && a->workspace, "spike_conv_wgrad: null pointer");
  SDF_REQUIRE(a->Cin % 16 == 0 && a->Cout % 4 == 0, "spike_conv_wgrad: Cin %% 16 and Cout %% 4 must be 0");
  SDF_REQUIRE(a->stride == 1 || a->stride == 2, "spike_conv_wgrad: stride must be 1 or 2");
  SDF_REQUIRE(a->kh * a->kw >= 1 && a->kh * a->kw <= 9, "spike_conv_wgrad: kernel size unsupported");
  SDF_REQUIRE(aligned16(a->g) && aligned16(a->x) && aligned16(a->workspace), "spike_conv_wgrad: pointers must be 16-byte aligned");
  WgradP p{};
  p.conv = 1;
  p.Cout = (int)a->Cout; p.Cin = (int)a->Cin; p.taps = (int)(a->kh * a->kw); p.ncols_total = p.taps * p.Cin;
  p.n_mtiles = (p.Cout + kWgM - 1) / kWgM;
  wgrad_tiles(p.Cin, p.taps, &p.taps_per_tile, &p.ci_tiles, &p.ci_width, &p.n_ntiles);
  p.tiles_h = (int)((a->Ho + kWgPatchH - 1) / kWgPatchH);
  p.tiles_w = (int)((a->Wo + kWgPatchW - 1) / kWgPatchW);
  p.n_chunks = (int)a->Nimg * p.tiles_h * p.tiles_w;
  p.stride = (int)a->stride;
  for (int i = 0; i < p.taps; ++i) { p.dh[i] = i / (int)a->kw - (int)a->pad; p.dw[i] = i % (int)a->kw - (int)a->pad; }
  p.partial = a->workspace;
  p.binary = a->s_max == 1;
  p.nbox = p.ci_width > 256 ? 2 : 1;
  p.box_w = p.ci_width / p.nbox;
  // one kernel row per N tile at stride 1: the kw taps of the row read overlapping pixels, one box with a (kw - 1)-pixel halo
  p.halo = a->stride == 1 && a->kw > 1 && p.ci_tiles == 1 && p.nbox == 1 && p.taps_per_tile == (int)a->kw;
  if (p.halo) p.stg_bytes = kWgPatchH * (kWgPatchW + (int)a->kw - 1) * p.box_w;
  CUtensorMap tmG, tmS;
  {
    const uint64_t dims[4] = {(uint64_t)a->Cout, (uint64_t)a->Wo, (uint64_t)a->Ho, (uint64_t)a->Nimg};
    const uint64_t str[3] = {(uint64_t)a->Cout * 4, (uint64_t)a->Wo * a->Cout * 4, (uint64_t)a->Ho * a->Wo * a->Cout * 4};
    const uint32_t box[4] = {(uint32_t)kWgM, (uint32_t)kWgPatchW, (uint32_t)kWgPatchH, 1};
    int st = make_tmap(&tmG, 1, 4, a->g, dims, str, box, nullptr, 0);
    if (st) return st;
  }
  {
    const int width = p.box_w;
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->Nimg};
    const uint64_t str[3] = {(uint64_t)a->Cin, (uint64_t)a->W * a->Cin, (uint64_t)a->H * a->W * a->Cin};
    const uint32_t box[4] = {(uint32_t)width, (uint32_t)(p.halo ? kWgPatchW + a->kw - 1 : kWgPatchW * a->stride),
                             (uint32_t)(kWgPatchH * a->stride), 1};
    const uint32_t es[4] = {1, (uint32_t)a->stride, (uint32_t)a->stride, 1};
    int st = make_tmap(&tmS, 0, 4, a->x, dims, str, box, es, 0);
    if (st) return st;
  }
  // parameter layout OIHW: (co, ci, tap) at co*Cin*taps + ci*taps + tap
  return wgrad_launch(p, &tmG, tmS, a->dw, (int64_t)p.Cin * p.taps, p.taps, 1, a->accumulate, a->workspace_bytes,
                      (cudaStream_t)a->stream, "sdf_spike_conv_wgrad", a->db);
}

// Weight (+ bias) gradient of ConvTranspose2d(k = 3, stride 2, padding 1, output_padding 1) on a spike operand:
//   dW[ci, co, kh, kw] = sum_{n,i,j} S[n, i, j, ci] * G[n, 2i - 1 + kh, 2j - 1 + kw, co]
// Per output parity class (a, b) the pixels G[2i + a, 2j + b] are a strided VIEW of G (a TMA tensor map with doubled pixel /
// row strides and a parity base offset) on the input grid, and the class's taps read the spikes at (i + dh, j + dw),
// dh, dw in {0, 1} (sdf_spike_deconv_class_taps).  ONE launch of the G3 kernel: its N tiles are the 1 + 2 + 2 + 4 taps grouped
// per class, each tile reading G through its class's tensor map — the tap shift stays on the 1-byte operand, G is never
// gathered or copied.  Replaces cuDNN's wgrad on fp32-expanded spikes (reference: SpikingTransposeDecoderLayer.deconv
// backward, Spiking_modules.py:398-474).
extern "C" int64_t sdf_spike_deconv_class_taps(int64_t cls, int64_t* src_tap, int64_t* dh, int64_t* dw);

static void deconv_wgrad_plan(WgradP& p, const sdf_spike_deconv_wgrad_args* a, int* tap_map) {
  p = WgradP{};
  p.conv = 1;
  p.Cout = (int)a->Cout; p.Cin = (int)a->Cin; p.taps = 9; p.ncols_total = 9 * p.Cin;
  p.n_mtiles = (p.Cout + kWgM - 1) / kWgM;
  int dummy_tpt, dummy_nt;
  wgrad_tiles(p.Cin, 1, &dummy_tpt, &p.ci_tiles, &p.ci_width, &dummy_nt);      // channel slicing only
  const int max_tpt = p.ci_tiles > 1 ? 1 : (kWgMaxN / p.Cin < 1 ? 1 : kWgMaxN / p.Cin);
  p.taps_per_tile = 1;
  int t = 0, k = 0, widest = 1;
  for (int cls = 0; cls < 4; ++cls) {
    int64_t src[4], dh[4], dw[4];
    const int n = (int)sdf_spike_deconv_class_taps(cls, src, dh, dw);
    for (int i = 0; i < n; ++i) { p.dh[t + i] = (int)dh[i]; p.dw[t + i] = (int)dw[i]; tap_map[t + i] = (int)src[i]; }
    const int groups = (n + max_tpt - 1) / max_tpt, per = (n + groups - 1) / groups;      // balanced tap groups of the class
    for (int g = 0; g * per < n; ++g, ++k) {
      p.tt_tap0[k] = (unsigned char)(t + g * per);
      p.tt_ntap[k] = (unsigned char)(n - g * per < per ? n - g * per : per);
      p.tt_cls[k] = (unsigned char)cls;
      p.tt_first[k] = g == 0;
      if (p.tt_ntap[k] > widest) widest = p.tt_ntap[k];
    }
    t += n;
  }
  p.n_cls = 4;
  p.n_ntiles = k * p.ci_tiles;
  p.max_cols = p.ci_tiles > 1 ? p.ci_width : widest * p.Cin;
  p.tiles_h = (int)((a->H + kWgPatchH - 1) / kWgPatchH);
  p.tiles_w = (int)((a->W + kWgPatchW - 1) / kWgPatchW);
  p.n_chunks = (int)a->Nimg * p.tiles_h * p.tiles_w;
  p.stride = 1;
  p.binary = a->s_max == 1;
  p.nbox = p.ci_width > 256 ? 2 : 1;
  p.box_w = p.ci_width / p.nbox;
}

extern "C" int64_t sdf_spike_deconv_wgrad_workspace_bytes(int64_t Nimg, int64_t H, int64_t W, int64_t Cout, int64_t Cin) {
  sdf_spike_deconv_wgrad_args a{};
  a.Nimg = Nimg; a.H = H; a.W = W; a.Cout = Cout; a.Cin = Cin;
  WgradP p; int tm[9];
  deconv_wgrad_plan(p, &a, tm);
  wgrad_slabs(p);
  return (int64_t)p.n_slabs * Cout * (p.ncols_total + p.n_cls) * 4;
}

extern "C" int sdf_spike_deconv_wgrad(const sdf_spike_deconv_wgrad_args* a) {
  SDF_REQUIRE(a->g && a->x && a->dw && a->workspace, "spike_deconv_wgrad: null pointer");
  SDF_REQUIRE(a->Nimg > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "spike_deconv_wgrad: empty problem");
  SDF_REQUIRE(a->Cin % 16 == 0 && a->Cout % 4 == 0, "spike_deconv_wgrad: Cin %% 16 and Cout %% 4 must be 0");
  SDF_REQUIRE(a->Cin_w >= 1 && a->Cin_w <= a->Cin, "spike_deconv_wgrad: Cin_w=%lld outside [1, Cin]", (long long)a->Cin_w);
  SDF_REQUIRE(aligned16(a->g) && aligned16(a->x) && aligned16(a->workspace), "spike_deconv_wgrad: pointers must be 16-byte aligned");
  const int64_t Ho = 2 * a->H, Wo = 2 * a->W;
  WgradP p; int tap_map[9];
  deconv_wgrad_plan(p, a, tap_map);
  p.partial = a->workspace;
  CUtensorMap tmG[4], tmS;
  for (int cls = 0; cls < 4; ++cls) {
    // the class's pixels of G on the input grid: (n, i, j) -> G[n, 2i + pa, 2j + pb, :]
    const int pa = cls >> 1, pb = cls & 1;
    const uint64_t dims[4] = {(uint64_t)a->Cout, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->Nimg};
    const uint64_t str[3] = {(uint64_t)2 * a->Cout * 4, (uint64_t)2 * Wo * a->Cout * 4, (uint64_t)Ho * Wo * a->Cout * 4};
    const uint32_t box[4] = {(uint32_t)kWgM, (uint32_t)kWgPatchW, (uint32_t)kWgPatchH, 1};
    int st = make_tmap(&tmG[cls], 1, 4, a->g + ((int64_t)pa * Wo + pb) * a->Cout, dims, str, box, nullptr, 0);
    if (st) return st;
  }
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->Nimg};
    const uint64_t str[3] = {(uint64_t)a->Cin, (uint64_t)a->W * a->Cin, (uint64_t)a->H * a->W * a->Cin};
    const uint32_t box[4] = {(uint32_t)p.box_w, (uint32_t)kWgPatchW, (uint32_t)kWgPatchH, 1};
    int st = make_tmap(&tmS, 0, 4, a->x, dims, str, box, nullptr, 0);
    if (st) return st;
  }
  // parameter layout IOHW: (ci, co, tap) at ci*Cout*9 + co*9 + tap
  return wgrad_launch(p, tmG, tmS, a->dw, 9, (int64_t)a->Cout * 9, 1, 0, a->workspace_bytes, (cudaStream_t)a->stream,
                      "sdf_spike_deconv_wgrad", a->db, tap_map, (int)a->Cin_w);
}
