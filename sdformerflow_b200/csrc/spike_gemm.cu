// spike_gemm.cu — G1/G2: Linear / convolution as (implicit) GEMM on the 5th-gen tensor cores, operands moved by TMA.
//
// Replaces the sj_layer.Linear / sj_layer.Conv2d calls on spike operands of the reference
// (models/STSwinNet_SNN/Spiking_swin_transformer3D.py:126-131 fc1/fc2, :267-290 / :632-652 linear_q/k/v + proj,
//  :909 reduction; models/STSwinNet_SNN/Spiking_modules.py:268,318,803,845-846 3x3 convolutions), which the reference runs
// as cuBLAS / cuDNN fp32 (fp16 under autocast) GEMMs on fp32 {0,1} tensors.
//
// G1 forward (KIND_I8): the A operand is the spike tensor itself — 1 byte per spike, loaded by TMA straight into the
// UMMA SWIZZLE_128B layout, no conversion warps.  The fp32 weights are pre-quantised once per optimizer step
// (sdf_spike_gemm_pack) to 23-bit fixed point per output channel (power-of-two scale) and split into three signed 8-bit
// digit planes stacked along N, so ONE tcgen05.mma kind::i8 (u8 x s8 -> s32, exact) per 32-deep K step produces the three
// partial accumulators side by side in TMEM; the epilogue recombines them in 64-bit integers, converts once to fp32
// (single rounding: the result is the correctly rounded dot product of the quantised weights, independent of tiling and
// summation order), adds the bias, writes fp32 rows through shared memory with a TMA store, and accumulates the per-channel
// sum / sum of squares that BatchNorm needs (the separate sdf_bn_stats pass over the output disappears).
// G2 (KIND_TF32): the same pipeline with fp32 operands read as TF32 (dgrad: dS = G W, both real-valued).
//
// Convolutions are the same GEMM with K = taps x Cin: an M tile is an 8 x 16 patch of output pixels, and each tap's A tile is
// a 4-D TMA box of the NHWC input shifted by the tap offset — zero padding is the TMA out-of-bounds fill, stride 2 is the
// tensor map's element stride; no im2col buffer and no padded copy (cuDNN's nhwcAddPaddingKernel) exist.
//
// Kernel anatomy (320 threads, one persistent CTA per SM, fixed N tile per CTA):
//   warp 8      TMA producer: A (and B unless the weights stay resident in shared memory) chunks into a ring of stages
//   warp 9      MMA issuer (one elected lane): 4 x tcgen05.mma per 128-byte K chunk, tcgen05.commit frees the stage;
//               accumulators are double buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 0-7   epilogue: tcgen05.ld (lane quarter = warp % 4, 16-column chunks interleaved between the two warps of a
//               quarter) -> combine/scale/bias -> swizzled smem -> TMA store.  BatchNorm sums stay in registers per
//               (row, column) across all the tiles of the CTA and are reduced over the 128 rows once at the end.
#include "sdf_common.cuh"
#include "tc_ptx.cuh"

namespace sdf {
using namespace tc;

constexpr int kGemmThreads = 320;               // warps 0-7 epilogue, warp 8 TMA producer, warp 9 MMA issuer
constexpr int kEpiThreads = 256;
constexpr int kTileM = 128;
constexpr int kChunkBytes = 128;                 // K bytes per operand row per chunk: one SWIZZLE_128B span
constexpr int kAChunk = kTileM * kChunkBytes;    // 16 KB
constexpr int kMaxStages = 8;
constexpr int kPatchH = 8, kPatchW = 16;         // conv M tile = 8 x 16 output pixels
constexpr int kMaxTaps = 9;

struct GemmP {
  int64_t rows;                 // linear: valid rows; conv: unused
  int n_mtiles, n_ntiles, n_kchunks, n_groups;
  int nt, ncols, Cout, stages, b_resident, kc_elems;
  int out_bufs;                 // staging buffers of the epilogue (1 or 2)
  int fast_cvt;                 // i8: accumulators < 2^22 in magnitude (spike operands): int -> float by magic-number add
  const float* wscale;
  const float* bias;
  float* bn_partials;
  int n_partial_cap;
  // convolution geometry (conv = 1): output (Nimg, Ho, Wo), tiles_h x tiles_w patches per image
  int conv, tiles_h, tiles_w, Ho, Wo, stride, taps, cpt;   // cpt = K chunks per tap
  int dh[kMaxTaps], dw[kMaxTaps];                          // input offset of tap t (already includes -padding)
  int btap[kMaxTaps];                                      // conv: tap t of this launch reads the B rows' K range of tap btap[t]
                                                           // (identity unless the launch uses a subset of the packed taps)
  int out_sh, out_sw, out_oh, out_ow;                      // transposed conv: output pixel = patch pixel * out_s + out_o
  int patch_h, patch_w;                                    // conv M tile = patch_h x patch_w output pixels (<= 128 of them: rows
                                                           // beyond patch_h*patch_w of a tile are never loaded, counted or stored)
  uint32_t a_tx;                                           // bytes one A chunk brings (patch rows * 128)
  int last_ks;                                             // 32-byte K steps of the LAST chunk of a tap / row that hold data (1..4):
                                                           // the MMAs over the zero-filled tail of the chunk are not issued
};

struct GemmSmem {
  uint32_t a, b, out, sc, bars, tmem_slot, total;
};
__host__ __device__ inline GemmSmem gemm_smem_plan(const GemmP& p) {
  GemmSmem s;
  uint32_t o = 0;
  s.a = o; o += (uint32_t)p.stages * kAChunk;
  s.b = o; o += (uint32_t)(p.b_resident ? p.n_kchunks : p.stages) * (uint32_t)p.ncols * kChunkBytes;
  s.out = o; o += (uint32_t)p.out_bufs * kTileM * p.nt * 4;   // epilogue staging (double buffered when it fits)
  s.sc = o; o += (uint32_t)p.nt * 8;
  s.bars = o; o += (2 * kMaxStages + 5) * 8;
  s.tmem_slot = o; o += 16;
  s.total = o;
  return s;
}

// tile -> coordinates
struct TileCoord {
  int c1, c2, c3;     // A / out coordinates beyond the innermost (row | w, h, img)
};
__device__ __forceinline__ TileCoord tile_coord(const GemmP& p, int m_tile) {
  TileCoord t;
  if (!p.conv) { t.c1 = m_tile * kTileM; t.c2 = 0; t.c3 = 0; return t; }
  const int per_img = p.tiles_h * p.tiles_w;
  const int img = m_tile / per_img, rem = m_tile - img * per_img;
  const int ph = rem / p.tiles_w, pw = rem - ph * p.tiles_w;
  t.c1 = pw * p.patch_w; t.c2 = ph * p.patch_h; t.c3 = img;
  return t;
}
// K coordinate (elements) of chunk kc in the B operand
__device__ __forceinline__ int b_kcoord(const GemmP& p, int kc) {
  if (!p.conv) return kc * p.kc_elems;
  const int tap = kc / p.cpt, cc = kc - tap * p.cpt;
  return (p.btap[tap] * p.cpt + cc) * p.kc_elems;
}
// number of valid rows mask: is tile row r inside the output?
__device__ __forceinline__ bool row_valid(const GemmP& p, const TileCoord& t, int r) {
  if (!p.conv) return (int64_t)t.c1 + r < p.rows;
  const int rh = r / p.patch_w;
  return rh < p.patch_h && (t.c2 + rh) < p.Ho && (t.c1 + r - rh * p.patch_w) < p.Wo;
}

template <int KIND>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmO, const GemmP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const GemmSmem sp = gemm_smem_plan(p);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;       // [2]
  uint64_t* tempty = bars + 2 * kMaxStages + 2;  // [2]
  uint64_t* bres = bars + 2 * kMaxStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sp.tmem_slot);
  float* sc_s = reinterpret_cast<float*>(smem + sp.sc);   // [nt] scale, [nt] bias
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tile = blockIdx.x % p.n_ntiles;             // fixed per CTA (gridDim.x = n_groups * n_ntiles)
  const int group = blockIdx.x / p.n_ntiles;
  const int n0 = n_tile * p.nt;
  const uint32_t bchunk = (uint32_t)p.ncols * kChunkBytes;

  if (tid == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiThreads / 32); }
    mbar_init(bres, 1);
    mbar_fence_init();
  }
  if (warp == 8 && elect_one()) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmO); }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < p.nt; i += kGemmThreads) {
    const int c = n0 + i;
    sc_s[i] = (KIND == KIND_I8 && c < p.Cout) ? __ldg(p.wscale + c) : 1.f;
    sc_s[p.nt + i] = (p.bias != nullptr && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_base = smem_u32(smem + sp.a), b_base = smem_u32(smem + sp.b);

  if (warp == 8) {
    // ===== TMA producer (one elected lane; splitting the A / B issue over two lanes measured slower) =====
    if (elect_one()) {
      if (p.b_resident) {
        mbar_expect_tx(bres, (uint32_t)p.n_kchunks * bchunk);
        for (int kc = 0; kc < p.n_kchunks; ++kc) tma_load_2d(&tmB, bres, b_base + kc * bchunk, b_kcoord(p, kc), n_tile * p.ncols);
      }
      int s = 0;                                 // ring slot and its phase, advanced incrementally (no division per chunk)
      uint32_t ph = 0;
      const int n_taps = p.conv ? p.taps : 1, cpt = p.conv ? p.cpt : p.n_kchunks;
      for (int m = group; m < p.n_mtiles; m += p.n_groups) {
        const TileCoord t = tile_coord(p, m);
        for (int tap = 0; tap < n_taps; ++tap) {
          // per-tap coordinates hoisted out of the chunk loop (no division on the issue path)
          const int aw = p.conv ? t.c1 * p.stride + p.dw[tap] : 0, ah = p.conv ? t.c2 * p.stride + p.dh[tap] : 0;
          const int bk0 = p.conv ? p.btap[tap] * cpt * p.kc_elems : 0;
          for (int cc = 0; cc < cpt; ++cc) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], p.a_tx + (p.b_resident ? 0u : bchunk));
            if (!p.conv) tma_load_2d(&tmA, &full[s], a_base + s * kAChunk, cc * p.kc_elems, t.c1);
            else tma_load_4d(&tmA, &full[s], a_base + s * kAChunk, cc * p.kc_elems, aw, ah, t.c3);
            if (!p.b_resident) tma_load_2d(&tmB, &full[s], b_base + s * bchunk, bk0 + cc * p.kc_elems, n_tile * p.ncols);
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = KIND == KIND_I8 ? idesc_i8_u8s8(kTileM, p.ncols) : idesc_tf32(kTileM, p.ncols);
      constexpr uint32_t hi = desc_hi_sw128(1024);
      if (p.b_resident) { mbar_wait(bres, 0); }
      uint32_t tile_i = 0, ph = 0;
      int s = 0;
      const int cpt_eff = p.conv ? p.cpt : p.n_kchunks;
      const int last_ks = p.last_ks > 0 ? p.last_ks : kChunkBytes / 32;
      for (int m = group; m < p.n_mtiles; m += p.n_groups, ++tile_i) {
        const uint32_t as = tile_i & 1, aph = (tile_i >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + as * (uint32_t)p.ncols;
        for (int kc = 0, cc = 0; kc < p.n_kchunks; ++kc) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_lo = desc_lo(a_base + s * kAChunk);
          const uint32_t b_lo = desc_lo(b_base + (p.b_resident ? kc : s) * bchunk);
          const int nks = cc == cpt_eff - 1 ? last_ks : kChunkBytes / 32;
#pragma unroll
          for (int ks = 0; ks < kChunkBytes / 32; ++ks)
            if (ks < nks) mma_ss<KIND>(d, a_lo + ks * 2, b_lo + ks * 2, hi, idesc, (kc | ks) != 0 ? 1u : 0u);
          tc_commit(&empty[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
          if (++cc == cpt_eff) cc = 0;
        }
        tc_commit(&tfull[as]);
      }
    }
  } else {
    // ===== epilogue (warps 0-7: TMEM lanes 32*(warp%4) .., chunks cc with cc % 2 == warp / 4) =====
    const int r = tid & 127;                             // tile row = TMEM lane
    const int half = tid >> 7;                           // which of the two warps of this lane quarter
    const int n_sub = p.nt / 16;
    const uint32_t out_bytes = (uint32_t)kTileM * p.nt * 4;
    constexpr int kStatChunks = 2;                       // i8: nt <= 64 -> at most 2 chunks per thread
    float ssum[kStatChunks][16], ssq[kStatChunks][16];   // BN sums of this thread's (row, column)s over all tiles
#pragma unroll
    for (int a = 0; a < kStatChunks; ++a)
#pragma unroll
      for (int i = 0; i < 16; ++i) { ssum[a][i] = 0.f; ssq[a][i] = 0.f; }
    const bool want_stats = KIND == KIND_I8 && p.bn_partials != nullptr;
    uint32_t tile_i = 0;
    for (int m = group; m < p.n_mtiles; m += p.n_groups, ++tile_i) {
      const TileCoord t = tile_coord(p, m);
      const uint32_t as = tile_i & 1, aph = (tile_i >> 1) & 1;
      uint8_t* out_s = smem + sp.out + (p.out_bufs == 2 ? as : 0u) * out_bytes;
      const uint32_t out_base = smem_u32(out_s);
      if ((tid & 31) == 0) mbar_wait(&tfull[as], aph);       // one polling lane per warp
      // the TMA stores issued from this staging buffer (two tiles ago when double buffered) must have finished reading it
      if (tid == 0) { if (p.out_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
      named_bar_sync(1, kEpiThreads);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + as * (uint32_t)p.ncols;
      const bool valid = want_stats && row_valid(p, t, r);
      const int x = (r >> 1) & 3;
      auto store_chunk = [&](int cc, const float (&y)[16]) {
        uint8_t* row = out_s + cc * (kTileM * 64) + r * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(row + ((j ^ x) << 4)) = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
      };
      if (KIND == KIND_I8) {
#pragma unroll
        for (int a = 0; a < kStatChunks; ++a) {
          const int cc = half + 2 * a;
          if (cc < n_sub) {
            uint32_t lo[16], mid[16], hi3[16];
            float y[16];
            tmem_ld16_nowait(trow + cc * 16, lo);
            tmem_ld16_nowait(trow + p.nt + cc * 16, mid);
            tmem_ld16_nowait(trow + 2 * p.nt + cc * 16, hi3);
            tmem_ld_wait();
            if (p.fast_cvt) {
              // |accumulator| < 2^22: float(x) = as_float(x + 0x4B400000) - 1.5*2^23, exact; then the three digit sums are
              // combined hi*65536 + (mid*256 + lo): the inner sum is < 2^31 and rounds once, the outer fma once more
              // (<= 1 ulp from the exact quantised dot product, a pure function of the integer triple: order independent)
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float fl = __int_as_float((int)lo[i] + 0x4B400000) - 12582912.f;
                const float fm = __int_as_float((int)mid[i] + 0x4B400000) - 12582912.f;
                const float fh = __int_as_float((int)hi3[i] + 0x4B400000) - 12582912.f;
                const float q = fmaf(fh, 65536.f, fmaf(fm, 256.f, fl));
                y[i] = __fadd_rn(__fmul_rn(q, sc_s[cc * 16 + i]), sc_s[p.nt + cc * 16 + i]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const long long q = ((long long)(int)hi3[i] << 16) + ((long long)(int)mid[i] << 8) + (long long)(int)lo[i];
                y[i] = __fadd_rn(__fmul_rn(__ll2float_rn(q), sc_s[cc * 16 + i]), sc_s[p.nt + cc * 16 + i]);
              }
            }
            store_chunk(cc, y);
            if (valid) {
#pragma unroll
              for (int i = 0; i < 16; ++i) { ssum[a][i] += y[i]; ssq[a][i] = fmaf(y[i], y[i], ssq[a][i]); }
            }
          }
        }
      } else {
        for (int cc = half; cc < n_sub; cc += 2) {
          uint32_t v[16];
          float y[16];
          tmem_ld16_nowait(trow + cc * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(v[i]) + sc_s[p.nt + cc * 16 + i];
          store_chunk(cc, y);
        }
      }
      // accumulator stage free for the MMA warp
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(&tempty[as]);
      fence_async_smem();
      named_bar_sync(1, kEpiThreads);
      if (tid == 0) {
        for (int cc = 0; cc < n_sub; ++cc) {
          if (n0 + cc * 16 >= p.Cout) break;
          if (!p.conv) tma_store_2d(&tmO, out_base + cc * (kTileM * 64), n0 + cc * 16, t.c1);
          else tma_store_4d(&tmO, out_base + cc * (kTileM * 64), n0 + cc * 16, t.c1, t.c2, t.c3);
        }
        tma_store_commit();
      }
    }
    if (tid == 0) tma_store_wait_all();
    if (want_stats) {
      // reduce the per-(row, column) sums over the 128 rows: [column][row ^ (column & 31)] in the staging buffer, one
      // statistic at a time (nt * 128 floats = one staging buffer)
      float* st = reinterpret_cast<float*>(smem + sp.out);
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        named_bar_sync(1, kEpiThreads);
#pragma unroll
        for (int a = 0; a < kStatChunks; ++a) {
          const int cc = half + 2 * a;
          if (cc < n_sub) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = cc * 16 + i;
              st[col * kTileM + (r ^ (col & 31))] = which == 0 ? ssum[a][i] : ssq[a][i];
            }
          }
        }
        named_bar_sync(1, kEpiThreads);
        if (tid < p.nt && n0 + tid < p.Cout) {
          float acc = 0.f;
          for (int rr = 0; rr < kTileM; ++rr) acc += st[tid * kTileM + (rr ^ (tid & 31))];
          p.bn_partials[((int64_t)group * 2 + which) * p.Cout + n0 + tid] = acc;
          // zero-fill the unused partial rows (the caller's finalize reduces all n_partial_cap rows)
          if (group == 0)
            for (int g = p.n_groups; g < p.n_partial_cap; ++g) p.bn_partials[((int64_t)g * 2 + which) * p.Cout + n0 + tid] = 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// ---- weight packing --------------------------------------------------------------------------------------------
// w element (co, ci, tap) at co*s_co + ci*s_ci + tap*s_tap (Linear: taps = 1, s_co = K, s_ci = 1; Conv2d OIHW: s_co = Cin*taps,
// s_ci = taps, s_tap = 1; ConvTranspose2d IOHW: s_ci = Cout*taps, s_co = taps).  tap_map[t] = source tap of GEMM tap t.
struct PackP {
  const float* w;
  int8_t* wq;
  float* wscale;
  float* wt;      // optional: fp32 transposed copy wt[ci][tap*Cout + co] (B operand of the data-gradient GEMMs)
  int Cout, Cin, taps, nt, n_ntiles, Kpad, cpt;
  int64_t s_co, s_ci, s_tap;
  int tap_map[kMaxTaps];
};
__global__ void pack_kernel(const PackP p) {
  const int co = blockIdx.x;                      // one block per (padded) output channel
  __shared__ float red[32];
  float mx = 0.f;
  if (co < p.Cout)
    for (int i = threadIdx.x; i < p.Cin * p.taps; i += blockDim.x) {
      const int ci = i / p.taps, tp = i - ci * p.taps;
      mx = fmaxf(mx, fabsf(__ldg(p.w + co * p.s_co + ci * p.s_ci + p.tap_map[tp] * p.s_tap)));
    }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  int e = 0;
  if (mx > 0.f) frexpf(mx, &e);                   // mx = m * 2^e, m in [0.5, 1)  ->  |w| * 2^(22-e) < 2^22
  const float qs = ldexpf(1.f, 22 - e);
  if (threadIdx.x == 0 && co < p.Cout) p.wscale[co] = ldexpf(1.f, e - 22);
  const int n_tile = co / p.nt, j = co - n_tile * p.nt;
  int8_t* base = p.wq + ((int64_t)n_tile * 3 * p.nt + j) * p.Kpad;   // plane s at + s*nt*Kpad
  const int kpt = p.cpt * kChunkBytes;            // padded K per tap
  for (int k = threadIdx.x; k < p.Kpad; k += blockDim.x) {
    const int tp = k / kpt, ci = k - tp * kpt;
    int q = 0;
    if (co < p.Cout && ci < p.Cin && tp < p.taps) {
      const float wv = __ldg(p.w + co * p.s_co + ci * p.s_ci + p.tap_map[tp] * p.s_tap);
      q = __float2int_rn(wv * qs);
      if (p.wt) p.wt[((int64_t)ci * p.taps + tp) * p.Cout + co] = wv;
    }
    const int lo = ((q + 128) & 255) - 128;
    const int q1 = (q - lo) >> 8;
    const int mid = ((q1 + 128) & 255) - 128;
    const int hi = (q1 - mid) >> 8;
    base[k] = (int8_t)lo;
    base[(int64_t)p.nt * p.Kpad + k] = (int8_t)mid;
    base[(int64_t)2 * p.nt * p.Kpad + k] = (int8_t)hi;
  }
}

// conv M tile: 8 x 16 output pixels unless that wastes more than 10 % of the tile rows on this image size (small feature maps:
// 9 x 12 fills 42 % of two 8 x 16 patches but 84 % of one 9 x 12 patch) — then the patch_h x patch_w <= 128 with the fewest tiles
static void pick_patch(GemmP& p) {
  auto tiles = [&](int h, int w) { return (int64_t)((p.Ho + h - 1) / h) * ((p.Wo + w - 1) / w); };
  int bh = kPatchH, bw = kPatchW;
  int64_t best = tiles(bh, bw);
  if ((double)p.Ho * p.Wo < 0.9 * (double)best * kTileM) {
    for (int w = 4; w <= kTileM && w <= p.Wo; ++w) {
      int h = kTileM / w;
      if (h > p.Ho) h = p.Ho;
      const int64_t t = tiles(h, w);
      if (t < best || (t == best && h * w > bh * bw)) { best = t; bh = h; bw = w; }
    }
  }
  p.patch_h = bh; p.patch_w = bw;
  p.tiles_h = (p.Ho + bh - 1) / bh;
  p.tiles_w = (p.Wo + bw - 1) / bw;
  p.a_tx = (uint32_t)(bh * bw) * kChunkBytes;
}

// N tile width: multiple of 16, minimal padding then as wide as possible
static int pick_nt(int C, int max_nt) {
  int best = 16, best_pad = 1 << 30;
  for (int nt = max_nt; nt >= 16; nt -= 16) {
    const int pad = (C + nt - 1) / nt * nt - C;
    if (pad < best_pad) { best_pad = pad; best = nt; }
  }
  return best;
}

}  // namespace sdf

using namespace sdf;

extern "C" int64_t sdf_spike_gemm_nt(int64_t Cout) { return pick_nt((int)Cout, 64); }

extern "C" int sdf_spike_gemm_pack(const sdf_spike_gemm_pack_args* a) {
  SDF_REQUIRE(a->w && a->wq && a->wscale, "spike_gemm_pack: null pointer");
  SDF_REQUIRE(a->taps >= 1 && a->taps <= kMaxTaps, "spike_gemm_pack: taps=%lld", (long long)a->taps);
  PackP p;
  p.w = a->w; p.wq = a->wq; p.wscale = a->wscale; p.wt = a->wt;
  p.Cout = (int)a->Cout; p.Cin = (int)a->Cin; p.taps = (int)a->taps;
  p.nt = pick_nt(p.Cout, 64);
  p.n_ntiles = (p.Cout + p.nt - 1) / p.nt;
  p.cpt = (p.Cin + kChunkBytes - 1) / kChunkBytes;
  p.Kpad = p.taps * p.cpt * kChunkBytes;
  SDF_REQUIRE(a->wq_bytes >= (int64_t)p.n_ntiles * 3 * p.nt * p.Kpad, "spike_gemm_pack: wq buffer too small (%lld < %lld)",
              (long long)a->wq_bytes, (long long)p.n_ntiles * 3 * p.nt * p.Kpad);
  p.s_co = a->s_co; p.s_ci = a->s_ci; p.s_tap = a->s_tap;
  for (int t = 0; t < kMaxTaps; ++t) p.tap_map[t] = t < p.taps ? (int)a->tap_map[t] : 0;
  pack_kernel<<<p.n_ntiles * p.nt, 128, 0, (cudaStream_t)a->stream>>>(p);
  return finish_launch("sdf_spike_gemm_pack");
}

extern "C" int64_t sdf_spike_gemm_wq_bytes(int64_t Cout, int64_t Cin, int64_t taps) {
  const int nt = pick_nt((int)Cout, 64);
  const int64_t n_ntiles = (Cout + nt - 1) / nt;
  const int64_t cpt = (Cin + kChunkBytes - 1) / kChunkBytes;
  return n_ntiles * 3 * nt * taps * cpt * kChunkBytes;
}

namespace sdf {
// common launch: fills the stage count / residency policy and launches
template <int KIND>
static int launch_gemm(GemmP& p, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, cudaStream_t st,
                       const char* what) {
  const int sms = num_sms();
  SDF_REQUIRE(p.n_ntiles <= sms, "%s: Cout too large (%d N tiles)", what, p.n_ntiles);
  p.n_groups = sms / p.n_ntiles;
  if (p.n_groups > p.n_mtiles) p.n_groups = p.n_mtiles;
  if (p.n_groups < 1) p.n_groups = 1;
  if (p.bn_partials) {
    SDF_REQUIRE(p.n_partial_cap >= 1, "%s: n_partial_blocks must be >= 1", what);
    if (p.n_groups > p.n_partial_cap) p.n_groups = p.n_partial_cap;
  }
  const uint32_t budget = 220 * 1024;
  const uint32_t bchunk = (uint32_t)p.ncols * kChunkBytes;
  // policy: double-buffer the staging tile and keep the weights resident whenever >= 4 (resp. 3) ring stages still fit
  p.b_resident = 0;
  p.out_bufs = 1;
  p.stages = 2;
  {
    GemmP q = p; q.out_bufs = 2; q.stages = 4;
    if (gemm_smem_plan(q).total <= budget) p.out_bufs = 2;
    q = p; q.b_resident = 1; q.stages = 3;
    if (gemm_smem_plan(q).total <= budget && p.n_mtiles / p.n_groups >= 2) p.b_resident = 1;
  }
  for (int s = kMaxStages; s >= 2; --s) {
    GemmP q = p; q.stages = s;
    if (gemm_smem_plan(q).total <= budget) { p.stages = s; break; }
    if (s == 2) { set_error("%s: tile does not fit shared memory (ncols %d, K chunks %d)", what, p.ncols, p.n_kchunks); return SDF_ERR_UNSUPPORTED; }
  }
  (void)bchunk;
  const GemmSmem sp = gemm_smem_plan(p);
  static bool attr_done[2] = {false, false};
  if (!attr_done[KIND]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return SDF_ERR_CUDA; }
    attr_done[KIND] = true;
  }
  gemm_kernel<KIND><<<p.n_groups * p.n_ntiles, kGemmThreads, sp.total, st>>>(tmA, tmB, tmO, p);
  return finish_launch(what);
}
}  // namespace sdf

extern "C" int sdf_spike_gemm_fwd(const sdf_spike_gemm_fwd_args* a) {
  SDF_REQUIRE(a->a && a->wq && a->wscale && a->out, "spike_gemm_fwd: null pointer");
  SDF_REQUIRE(a->rows > 0 && a->K > 0 && a->Cout > 0, "spike_gemm_fwd: empty problem");
  SDF_REQUIRE(a->K % 16 == 0, "spike_gemm_fwd: K=%lld must be a multiple of 16 (TMA row pitch)", (long long)a->K);
  SDF_REQUIRE(a->Cout % 4 == 0 && a->ld_out % 4 == 0, "spike_gemm_fwd: Cout / ld_out must be multiples of 4");
  SDF_REQUIRE(aligned16(a->a) && aligned16(a->out) && aligned16(a->wq), "spike_gemm_fwd: pointers must be 16-byte aligned");
  GemmP p{};
  p.rows = a->rows;
  p.nt = pick_nt((int)a->Cout, 64);
  p.n_ntiles = ((int)a->Cout + p.nt - 1) / p.nt;
  p.ncols = 3 * p.nt;
  p.Cout = (int)a->Cout;
  p.n_mtiles = (int)((a->rows + kTileM - 1) / kTileM);
  p.n_kchunks = (int)((a->K + kChunkBytes - 1) / kChunkBytes);
  p.kc_elems = kChunkBytes;
  p.wscale = a->wscale; p.bias = a->bias; p.bn_partials = a->bn_partials; p.n_partial_cap = (int)a->n_partial_blocks;
  p.fast_cvt = (a->a_max > 0 && a->K * a->a_max < 32768) ? 1 : 0;
  p.a_tx = kAChunk;
  p.last_ks = (int)((a->K - (int64_t)(p.n_kchunks - 1) * kChunkBytes + 31) / 32);
  const int64_t Kpad = (int64_t)p.n_kchunks * kChunkBytes;
  CUtensorMap tmA, tmB, tmO;
  {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->rows};
    const uint64_t str[1] = {(uint64_t)a->K};
    const uint32_t box[2] = {(uint32_t)kChunkBytes, (uint32_t)kTileM};
    int st = make_tmap(&tmA, 0, 2, a->a, dims, str, box, nullptr, 128);
    if (st) return st;
  }
  {
    const uint64_t dims[2] = {(uint64_t)Kpad, (uint64_t)p.n_ntiles * p.ncols};
    const uint64_t str[1] = {(uint64_t)Kpad};
    const uint32_t box[2] = {(uint32_t)kChunkBytes, (uint32_t)p.ncols};
    int st = make_tmap(&tmB, 0, 2, a->wq, dims, str, box, nullptr, 128);
    if (st) return st;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->Cout, (uint64_t)a->rows};
    const uint64_t str[1] = {(uint64_t)a->ld_out * 4};
    const uint32_t box[2] = {16, (uint32_t)kTileM};
    int st = make_tmap(&tmO, 1, 2, a->out, dims, str, box, nullptr, 64);
    if (st) return st;
  }
  return launch_gemm<KIND_I8>(p, tmA, tmB, tmO, (cudaStream_t)a->stream, "sdf_spike_gemm_fwd");
}

extern "C" int sdf_gemm_tf32(const sdf_gemm_tf32_args* a) {
  SDF_REQUIRE(a->a && a->b && a->out, "gemm_tf32: null pointer");
  SDF_REQUIRE(a->rows > 0 && a->K > 0 && a->N > 0, "gemm_tf32: empty problem");
  SDF_REQUIRE(a->K % 4 == 0 && a->N % 4 == 0 && a->ld_out % 4 == 0, "gemm_tf32: K, N, ld_out must be multiples of 4");
  SDF_REQUIRE(aligned16(a->a) && aligned16(a->b) && aligned16(a->out), "gemm_tf32: pointers must be 16-byte aligned");
  GemmP p{};
  p.rows = a->rows;
  p.nt = pick_nt((int)a->N, 128);
  p.n_ntiles = ((int)a->N + p.nt - 1) / p.nt;
  p.ncols = p.nt;
  p.Cout = (int)a->N;
  p.n_mtiles = (int)((a->rows + kTileM - 1) / kTileM);
  p.n_kchunks = (int)((a->K + 31) / 32);
  p.kc_elems = 32;
  p.a_tx = kAChunk;
  p.last_ks = (int)((a->K - (int64_t)(p.n_kchunks - 1) * 32 + 7) / 8);
  p.bias = a->bias;
  CUtensorMap tmA, tmB, tmO;
  {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->rows};
    const uint64_t str[1] = {(uint64_t)a->lda * 4};
    const uint32_t box[2] = {32, (uint32_t)kTileM};
    int st = make_tmap(&tmA, 1, 2, a->a, dims, str, box, nullptr, 128);
    if (st) return st;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    const uint64_t str[1] = {(uint64_t)a->ldb * 4};
    const uint32_t box[2] = {32, (uint32_t)p.nt};
    int st = make_tmap(&tmB, 1, 2, a->b, dims, str, box, nullptr, 128);
    if (st) return st;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->rows};
    const uint64_t str[1] = {(uint64_t)a->ld_out * 4};
    const uint32_t box[2] = {16, (uint32_t)kTileM};
    int st = make_tmap(&tmO, 1, 2, a->out, dims, str, box, nullptr, 64);
    if (st) return st;
  }
  return launch_gemm<KIND_TF32>(p, tmA, tmB, tmO, (cudaStream_t)a->stream, "sdf_gemm_tf32");
}

namespace sdf {
// One implicit-GEMM launch of the spike convolution: `taps` taps with input offsets (dh, dw) and the packed weights `wq`
// (tap order of the pack); output pixel (h, w) of the (Ho, Wo) tile grid is written at out_base + ((h*osh)*ld_row + w*osw)*Cout
// (osh = osw = 1, ld_row = Wo: an ordinary convolution; 2, 2 and a parity offset in out_base: one phase of a transposed one).
struct ConvLaunch {
  const uint8_t* x; int64_t Nimg, H, W, Cin;
  const int8_t* wq; const float* wscale; const float* bias;
  float* out_base; int64_t Ho, Wo, Cout, osh, osw, out_ld_row, out_img_elems;
  int taps, stride; int dh[kMaxTaps], dw[kMaxTaps];
  float* bn_partials; int64_t n_partial_blocks; int64_t a_max;
  cudaStream_t stream;
};
static int conv_fwd_launch(const ConvLaunch& c, const char* what) {
  GemmP p{};
  p.conv = 1;
  p.nt = pick_nt((int)c.Cout, 64);
  p.n_ntiles = ((int)c.Cout + p.nt - 1) / p.nt;
  p.ncols = 3 * p.nt;
  p.Cout = (int)c.Cout;
  p.Ho = (int)c.Ho; p.Wo = (int)c.Wo;
  pick_patch(p);
  p.n_mtiles = (int)c.Nimg * p.tiles_h * p.tiles_w;
  p.stride = c.stride;
  p.taps = c.taps;
  p.cpt = (int)((c.Cin + kChunkBytes - 1) / kChunkBytes);
  p.n_kchunks = p.taps * p.cpt;
  p.kc_elems = kChunkBytes;
  p.last_ks = (int)((c.Cin - (int64_t)(p.cpt - 1) * kChunkBytes + 31) / 32);
  for (int i = 0; i < p.taps; ++i) { p.dh[i] = c.dh[i]; p.dw[i] = c.dw[i]; p.btap[i] = i; }
  p.wscale = c.wscale; p.bias = c.bias; p.bn_partials = c.bn_partials; p.n_partial_cap = (int)c.n_partial_blocks;
  p.fast_cvt = (c.a_max > 0 && c.Cin * c.taps * c.a_max < 32768) ? 1 : 0;
  const int64_t Kpad = (int64_t)p.n_kchunks * kChunkBytes;
  CUtensorMap tmA, tmB, tmO;
  {
    const uint64_t dims[4] = {(uint64_t)c.Cin, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.Nimg};
    const uint64_t str[3] = {(uint64_t)c.Cin, (uint64_t)c.W * c.Cin, (uint64_t)c.H * c.W * c.Cin};
    const uint32_t box[4] = {(uint32_t)kChunkBytes, (uint32_t)(p.patch_w * c.stride), (uint32_t)(p.patch_h * c.stride), 1};
    const uint32_t es[4] = {1, (uint32_t)c.stride, (uint32_t)c.stride, 1};
    int st = make_tmap(&tmA, 0, 4, c.x, dims, str, box, es, 128);
    if (st) return st;
  }
  {
    const uint64_t dims[2] = {(uint64_t)Kpad, (uint64_t)p.n_ntiles * p.ncols};
    const uint64_t str[1] = {(uint64_t)Kpad};
    const uint32_t box[2] = {(uint32_t)kChunkBytes, (uint32_t)p.ncols};
    int st = make_tmap(&tmB, 0, 2, c.wq, dims, str, box, nullptr, 128);
    if (st) return st;
  }
  {
    const uint64_t dims[4] = {(uint64_t)c.Cout, (uint64_t)c.Wo, (uint64_t)c.Ho, (uint64_t)c.Nimg};
    const uint64_t str[3] = {(uint64_t)c.osw * c.Cout * 4, (uint64_t)c.osh * c.out_ld_row * c.Cout * 4, (uint64_t)c.out_img_elems * 4};
    const uint32_t box[4] = {16, (uint32_t)p.patch_w, (uint32_t)p.patch_h, 1};
    int st = make_tmap(&tmO, 1, 4, c.out_base, dims, str, box, nullptr, 64);
    if (st) return st;
  }
  return launch_gemm<KIND_I8>(p, tmA, tmB, tmO, c.stream, what);
}
}  // namespace sdf

extern "C" int sdf_spike_conv_fwd(const sdf_spike_conv_fwd_args* a) {
  SDF_REQUIRE(a->x && a->wq && a->wscale && a->out, "spike_conv_fwd: null pointer");
  SDF_REQUIRE(a->Nimg > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "spike_conv_fwd: empty problem");
  SDF_REQUIRE(a->Cin % 16 == 0, "spike_conv_fwd: Cin=%lld must be a multiple of 16 (TMA pixel pitch)", (long long)a->Cin);
  SDF_REQUIRE(a->Cout % 4 == 0, "spike_conv_fwd: Cout must be a multiple of 4");
  SDF_REQUIRE(a->stride == 1 || a->stride == 2, "spike_conv_fwd: stride must be 1 or 2");
  SDF_REQUIRE(a->kh * a->kw >= 1 && a->kh * a->kw <= kMaxTaps, "spike_conv_fwd: kernel %lldx%lld unsupported", (long long)a->kh, (long long)a->kw);
  SDF_REQUIRE(a->Ho == (a->H + 2 * a->pad - a->kh) / a->stride + 1 && a->Wo == (a->W + 2 * a->pad - a->kw) / a->stride + 1,
              "spike_conv_fwd: output size does not match the geometry");
  SDF_REQUIRE(aligned16(a->x) && aligned16(a->out) && aligned16(a->wq), "spike_conv_fwd: pointers must be 16-byte aligned");
  ConvLaunch c{};
  c.x = a->x; c.Nimg = a->Nimg; c.H = a->H; c.W = a->W; c.Cin = a->Cin;
  c.wq = a->wq; c.wscale = a->wscale; c.bias = a->bias;
  c.out_base = a->out; c.Ho = a->Ho; c.Wo = a->Wo; c.Cout = a->Cout; c.osh = 1; c.osw = 1; c.out_ld_row = a->Wo;
  c.out_img_elems = a->Ho * a->Wo * a->Cout;
  c.taps = (int)(a->kh * a->kw); c.stride = (int)a->stride;
  for (int i = 0; i < c.taps; ++i) { c.dh[i] = i / (int)a->kw - (int)a->pad; c.dw[i] = i % (int)a->kw - (int)a->pad; }
  c.bn_partials = a->bn_partials; c.n_partial_blocks = a->n_partial_blocks; c.a_max = a->a_max;
  c.stream = (cudaStream_t)a->stream;
  return conv_fwd_launch(c, "sdf_spike_conv_fwd");
}

// ConvTranspose2d(k = 3, stride 2, padding 1, output_padding 1) on spikes: out (Nimg, 2H, 2W, Cout).  The output pixels of
// parity (a, b) are an ordinary stride-1 convolution of the input with the kernel taps that land on that parity
//   a = 0: kh = 1 (input row i);          a = 1: kh = 2 (row i) and kh = 0 (row i + 1)          (same for columns)
// i.e. 1 + 2 + 2 + 4 = 9 taps over four launches of the same implicit-GEMM kernel, each writing its quarter of the output
// through a strided TMA tensor map (no zero-stuffed input, no scatter pass; rows past the image are the TMA zero fill).
extern "C" int64_t sdf_spike_deconv_class_taps(int64_t cls, int64_t* src_tap, int64_t* dh, int64_t* dw) {
  const int a = (int)(cls >> 1), b = (int)(cls & 1);
  const int kh_list[2][2] = {{1, -1}, {2, 0}}, d_list[2][2] = {{0, 0}, {0, 1}};
  int n = 0;
  for (int i = 0; i < (a ? 2 : 1); ++i)
    for (int j = 0; j < (b ? 2 : 1); ++j) {
      src_tap[n] = kh_list[a][i] * 3 + kh_list[b][j];
      dh[n] = d_list[a][i];
      dw[n] = d_list[b][j];
      ++n;
    }
  return n;
}

extern "C" int sdf_spike_deconv_fwd(const sdf_spike_deconv_fwd_args* a) {
  SDF_REQUIRE(a->x && a->out, "spike_deconv_fwd: null pointer");
  SDF_REQUIRE(a->Nimg > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "spike_deconv_fwd: empty problem");
  SDF_REQUIRE(a->Cin % 16 == 0 && a->Cout % 4 == 0, "spike_deconv_fwd: Cin %% 16 and Cout %% 4 must be 0");
  SDF_REQUIRE(aligned16(a->x) && aligned16(a->out), "spike_deconv_fwd: pointers must be 16-byte aligned");
  const int64_t Ho = 2 * a->H, Wo = 2 * a->W;
  for (int cls = 0; cls < 4; ++cls) {
    SDF_REQUIRE(a->wq[cls] && a->wscale[cls] && aligned16(a->wq[cls]), "spike_deconv_fwd: null / unaligned weight planes of class %d", cls);
    ConvLaunch c{};
    c.x = a->x; c.Nimg = a->Nimg; c.H = a->H; c.W = a->W; c.Cin = a->Cin;
    c.wq = a->wq[cls]; c.wscale = a->wscale[cls]; c.bias = a->bias;
    const int pa = cls >> 1, pb = cls & 1;
    c.out_base = a->out + ((int64_t)pa * Wo + pb) * a->Cout;
    c.Ho = a->H; c.Wo = a->W; c.Cout = a->Cout; c.osh = 2; c.osw = 2; c.out_ld_row = Wo; c.out_img_elems = Ho * Wo * a->Cout;
    int64_t src[4], dh[4], dw[4];
    c.taps = (int)sdf_spike_deconv_class_taps(cls, src, dh, dw);
    for (int i = 0; i < c.taps; ++i) { c.dh[i] = (int)dh[i]; c.dw[i] = (int)dw[i]; }
    c.stride = 1;
    c.bn_partials = a->bn_partials ? a->bn_partials + (int64_t)cls * a->n_partial_blocks * 2 * a->Cout : nullptr;
    c.n_partial_blocks = a->n_partial_blocks; c.a_max = a->a_max;
    c.stream = (cudaStream_t)a->stream;
    int st = conv_fwd_launch(c, "sdf_spike_deconv_fwd");
    if (st) return st;
  }
  return SDF_OK;
}

// Data gradient of a stride-1 NHWC convolution: dX[n,h,w,ci] = sum_{kh,kw,co} G[n, h+pad-kh, w+pad-kw, co] * W[co,ci,kh,kw],
// the same implicit GEMM with G (fp32, read as TF32) as the tapped operand and the weights re-laid as
// wd[ci][tap*Cout + co] (K-major rows; the caller permutes the parameter once per step).
extern "C" int sdf_conv_dgrad_tf32(const sdf_conv_dgrad_tf32_args* a) {
  SDF_REQUIRE(a->g && a->wd && a->out, "conv_dgrad_tf32: null pointer");
  SDF_REQUIRE(a->Nimg > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "conv_dgrad_tf32: empty problem");
  SDF_REQUIRE(a->Cout % 32 == 0, "conv_dgrad_tf32: Cout=%lld must be a multiple of 32 (one 128-byte K chunk)", (long long)a->Cout);
  SDF_REQUIRE(a->Cin % 4 == 0, "conv_dgrad_tf32: Cin must be a multiple of 4");
  SDF_REQUIRE(a->kh * a->kw >= 1 && a->kh * a->kw <= kMaxTaps, "conv_dgrad_tf32: kernel size unsupported");
  SDF_REQUIRE(a->Ho == a->H + 2 * a->pad - a->kh + 1 && a->Wo == a->W + 2 * a->pad - a->kw + 1, "conv_dgrad_tf32: stride-1 geometry expected");
  SDF_REQUIRE(aligned16(a->g) && aligned16(a->wd) && aligned16(a->out), "conv_dgrad_tf32: pointers must be 16-byte aligned");
  GemmP p{};
  p.conv = 1;
  p.nt = pick_nt((int)a->Cin, 128);
  p.n_ntiles = ((int)a->Cin + p.nt - 1) / p.nt;
  p.ncols = p.nt;
  p.Cout = (int)a->Cin;                         // GEMM N = input channels
  p.Ho = (int)a->H; p.Wo = (int)a->W;           // the output of this GEMM is the input image
  pick_patch(p);
  p.n_mtiles = (int)a->Nimg * p.tiles_h * p.tiles_w;
  p.stride = 1;
  p.taps = (int)(a->kh * a->kw);
  p.cpt = (int)(a->Cout / 32);
  p.n_kchunks = p.taps * p.cpt;
  p.kc_elems = 32;
  for (int i = 0; i < p.taps; ++i) { p.dh[i] = (int)a->pad - i / (int)a->kw; p.dw[i] = (int)a->pad - i % (int)a->kw; p.btap[i] = i; }
  CUtensorMap tmA, tmB, tmO;
  {
    const uint64_t dims[4] = {(uint64_t)a->Cout, (uint64_t)a->Wo, (uint64_t)a->Ho, (uint64_t)a->Nimg};
    const uint64_t str[3] = {(uint64_t)a->Cout * 4, (uint64_t)a->Wo * a->Cout * 4, (uint64_t)a->Ho * a->Wo * a->Cout * 4};
    const uint32_t box[4] = {32, (uint32_t)p.patch_w, (uint32_t)p.patch_h, 1};
    int st = make_tmap(&tmA, 1, 4, a->g, dims, str, box, nullptr, 128);
    if (st) return st;
  }
  {
    const uint64_t K = (uint64_t)p.taps * a->Cout;
    const uint64_t dims[2] = {K, (uint64_t)a->Cin};
    const uint64_t str[1] = {K * 4};
    const uint32_t box[2] = {32, (uint32_t)p.nt};
    int st = make_tmap(&tmB, 1, 2, a->wd, dims, str, box, nullptr, 128);
    if (st) return st;
  }
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->Nimg};
    const uint64_t str[3] = {(uint64_t)a->Cin * 4, (uint64_t)a->W * a->Cin * 4, (uint64_t)a->H * a->W * a->Cin * 4};
    const uint32_t box[4] = {16, (uint32_t)p.patch_w, (uint32_t)p.patch_h, 1};
    int st = make_tmap(&tmO, 1, 4, a->out, dims, str, box, nullptr, 64);
    if (st) return st;
  }
  return launch_gemm<KIND_TF32>(p, tmA, tmB, tmO, (cudaStream_t)a->stream, "sdf_conv_dgrad_tf32");
}

namespace sdf {
// One TF32 implicit-GEMM launch over a tapped fp32 NHWC operand `g` (Nimg, Hg, Wg, Cg), read with element stride `stride`:
//   out[n, h, w, c] = sum_{t < taps} sum_{k < Cg} g[n, h*stride + dh[t], w*stride + dw[t], k] * wd[c][(btap[t]*cpt + k/32)*32 + k%32]
// for the (Hc, Wc) grid of output pixels, written at out_base + (h*osh + w*osw + n*oimg) elements + c.  wd has n_brows valid
// rows (rows beyond are TMA zero fill: output channels of zero-padded operand slices come out as zeros) of total_taps*cpt*32
// floats, cpt = ceil(Cg / 32) (a tap's channels padded to whole 128-byte K chunks).
struct ConvTf32Launch {
  const float* g; int64_t Nimg, Hg, Wg, Cg; int stride;
  const float* wd; int64_t n_brows; int total_taps;
  float* out_base; int64_t Hc, Wc, N, osw, osh, oimg;
  int taps; int dh[kMaxTaps], dw[kMaxTaps], btap[kMaxTaps];
  cudaStream_t stream;
};
static int conv_tf32_launch(const ConvTf32Launch& c, const char* what) {
  GemmP p{};
  p.conv = 1;
  // as few N tiles as possible (every N tile re-reads the tapped operand), each as narrow as that allows
  const int tiles = (int)((c.N + 127) / 128);
  p.nt = (int)(((c.N + tiles - 1) / tiles + 15) / 16 * 16);
  p.n_ntiles = ((int)c.N + p.nt - 1) / p.nt;
  p.ncols = p.nt;
  p.Cout = (int)c.N;
  p.Ho = (int)c.Hc; p.Wo = (int)c.Wc;
  pick_patch(p);
  p.n_mtiles = (int)c.Nimg * p.tiles_h * p.tiles_w;
  p.stride = c.stride;
  p.taps = c.taps;
  p.cpt = (int)((c.Cg + 31) / 32);
  p.n_kchunks = p.taps * p.cpt;
  p.kc_elems = 32;
  p.last_ks = (int)((c.Cg - (int64_t)(p.cpt - 1) * 32 + 7) / 8);
  for (int i = 0; i < p.taps; ++i) { p.dh[i] = c.dh[i]; p.dw[i] = c.dw[i]; p.btap[i] = c.btap[i]; }
  CUtensorMap tmA, tmB, tmO;
  {
    const uint64_t dims[4] = {(uint64_t)c.Cg, (uint64_t)c.Wg, (uint64_t)c.Hg, (uint64_t)c.Nimg};
    const uint64_t str[3] = {(uint64_t)c.Cg * 4, (uint64_t)c.Wg * c.Cg * 4, (uint64_t)c.Hg * c.Wg * c.Cg * 4};
    const uint32_t box[4] = {32, (uint32_t)(p.patch_w * c.stride), (uint32_t)(p.patch_h * c.stride), 1};
    const uint32_t es[4] = {1, (uint32_t)c.stride, (uint32_t)c.stride, 1};
    int st = make_tmap(&tmA, 1, 4, c.g, dims, str, box, es, 128);
    if (st) return st;
  }
  {
    const uint64_t K = (uint64_t)c.total_taps * p.cpt * 32;
    const uint64_t dims[2] = {K, (uint64_t)c.n_brows};
    const uint64_t str[1] = {K * 4};
    const uint32_t box[2] = {32, (uint32_t)p.nt};
    int st = make_tmap(&tmB, 1, 2, c.wd, dims, str, box, nullptr, 128);
    if (st) return st;
  }
  {
    const uint64_t dims[4] = {(uint64_t)c.N, (uint64_t)c.Wc, (uint64_t)c.Hc, (uint64_t)c.Nimg};
    const uint64_t str[3] = {(uint64_t)c.osw * 4, (uint64_t)c.osh * 4, (uint64_t)c.oimg * 4};
    const uint32_t box[4] = {16, (uint32_t)p.patch_w, (uint32_t)p.patch_h, 1};
    int st = make_tmap(&tmO, 1, 4, c.out_base, dims, str, box, nullptr, 64);
    if (st) return st;
  }
  return launch_gemm<KIND_TF32>(p, tmA, tmB, tmO, c.stream, what);
}
}  // namespace sdf

// Data gradient of a 3x3, stride-2, padding-1 NHWC convolution (the strided convolutions of the patch embedding):
//   dX[n, h, w, ci] = sum_{kh, kw, co : h + 1 - kh = 2 ho, w + 1 - kw = 2 wo} G[n, ho, wo, co] * W[co, ci, kh, kw]
// The input pixels of parity (a, b) = (h & 1, w & 1) are an ordinary stride-1 convolution of G with the taps that land on that
// parity — the classes of sdf_spike_deconv_class_taps (a strided convolution's data gradient IS a transposed convolution):
// four launches of the TF32 implicit GEMM, each reading its taps' K range of wd and writing its quarter of dX through a
// strided TMA tensor map.  Replaces cuDNN's strided_dgrad engine (+ its NHWC repacking) on this path.
extern "C" int sdf_conv_dgrad_s2_tf32(const sdf_conv_dgrad_s2_tf32_args* a) {
  SDF_REQUIRE(a->g && a->wd && a->out, "conv_dgrad_s2_tf32: null pointer");
  SDF_REQUIRE(a->Nimg > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "conv_dgrad_s2_tf32: empty problem");
  SDF_REQUIRE(a->Cout % 32 == 0, "conv_dgrad_s2_tf32: Cout=%lld must be a multiple of 32 (one 128-byte K chunk)", (long long)a->Cout);
  SDF_REQUIRE(a->Cin % 4 == 0, "conv_dgrad_s2_tf32: Cin must be a multiple of 4");
  SDF_REQUIRE(a->Ho == (a->H - 1) / 2 + 1 && a->Wo == (a->W - 1) / 2 + 1, "conv_dgrad_s2_tf32: 3x3 / stride 2 / padding 1 geometry expected");
  SDF_REQUIRE(aligned16(a->g) && aligned16(a->wd) && aligned16(a->out), "conv_dgrad_s2_tf32: pointers must be 16-byte aligned");
  for (int cls = 0; cls < 4; ++cls) {
    const int pa = cls >> 1, pb = cls & 1;
    ConvTf32Launch c{};
    c.Hc = (a->H - pa + 1) / 2; c.Wc = (a->W - pb + 1) / 2;
    if (c.Hc <= 0 || c.Wc <= 0) continue;
    c.g = a->g; c.Nimg = a->Nimg; c.Hg = a->Ho; c.Wg = a->Wo; c.Cg = a->Cout; c.stride = 1;
    c.wd = a->wd; c.n_brows = a->Cin; c.total_taps = 9;
    c.out_base = a->out + ((int64_t)pa * a->W + pb) * a->Cin;
    c.N = a->Cin; c.osw = 2 * a->Cin; c.osh = 2 * a->W * a->Cin; c.oimg = a->H * a->W * a->Cin;
    int64_t src[4], dh[4], dw[4];
    c.taps = (int)sdf_spike_deconv_class_taps(cls, src, dh, dw);
    for (int i = 0; i < c.taps; ++i) { c.dh[i] = (int)dh[i]; c.dw[i] = (int)dw[i]; c.btap[i] = (int)src[i]; }
    c.stream = (cudaStream_t)a->stream;
    int st = conv_tf32_launch(c, "sdf_conv_dgrad_s2_tf32");
    if (st) return st;
  }
  return SDF_OK;
}

// Data gradient of ConvTranspose2d(k = 3, stride 2, padding 1, output_padding 1):
//   dX[n, i, j, ci] = sum_{kh, kw, co} G[n, 2i - 1 + kh, 2j - 1 + kw, co] * W[ci, co, kh, kw]
// i.e. a stride-2, padding-1 convolution of G: ONE launch of the TF32 implicit GEMM whose tapped operand is read through a
// TMA tensor map with element stride 2 (rows / columns -1 and 2H are the TMA zero fill).  Replaces cuDNN's fprop engine on
// this path (reference: SpikingTransposeDecoderLayer.deconv backward, Spiking_modules.py:398-474).
extern "C" int sdf_deconv_dgrad_tf32(const sdf_deconv_dgrad_tf32_args* a) {
  SDF_REQUIRE(a->g && a->wd && a->out, "deconv_dgrad_tf32: null pointer");
  SDF_REQUIRE(a->Nimg > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "deconv_dgrad_tf32: empty problem");
  SDF_REQUIRE(a->Cout % 4 == 0 && a->Cin % 4 == 0, "deconv_dgrad_tf32: Cin, Cout must be multiples of 4");
  SDF_REQUIRE(a->Cin_w >= 1 && a->Cin_w <= a->Cin, "deconv_dgrad_tf32: Cin_w=%lld outside [1, Cin]", (long long)a->Cin_w);
  SDF_REQUIRE(aligned16(a->g) && aligned16(a->wd) && aligned16(a->out), "deconv_dgrad_tf32: pointers must be 16-byte aligned");
  ConvTf32Launch c{};
  c.g = a->g; c.Nimg = a->Nimg; c.Hg = 2 * a->H; c.Wg = 2 * a->W; c.Cg = a->Cout; c.stride = 2;
  c.wd = a->wd; c.n_brows = a->Cin_w; c.total_taps = 9;
  c.out_base = a->out; c.Hc = a->H; c.Wc = a->W; c.N = a->Cin;
  c.osw = a->Cin; c.osh = a->W * a->Cin; c.oimg = a->H * a->W * a->Cin;
  c.taps = 9;
  for (int i = 0; i < 9; ++i) { c.dh[i] = i / 3 - 1; c.dw[i] = i % 3 - 1; c.btap[i] = i; }
  c.stream = (cudaStream_t)a->stream;
  return conv_tf32_launch(c, "sdf_deconv_dgrad_tf32");
}
