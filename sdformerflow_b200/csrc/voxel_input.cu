// voxel_input.cu — input pipeline of the spiking patch embedding in one pass over the voxel grid.
//
// Replaces, fused:
//   train_flow_parallel_supervised_SNN.py:261-265 / eval_DSEC_flow_SNN.py:196-200   pos/neg polarity split
//        neg = relu(-chunk); pos = relu(chunk); chunk = cat((pos.unsqueeze(2), neg.unsqueeze(2)), dim=2)      (B,bins,2,H,W)
//   :278-284 / eval :203-212   min-max normalisation over the NON-ZERO entries
//        chunk[chunk != 0] = (chunk[chunk != 0] - min) / (max - min)          (skipped when min == max)
//   models/STSwinNet_SNN/Spiking_modules.py:1772-1786   bins -> (steps, channels) regroup of the patch embedding
//        new[:, i, :, :, t] = x[:, (i // 2) * steps + t, i % 2]
// and writes the result straight in the channels-last layout (B, steps, H, W, num_ch) the conv stack consumes.
// The reference makes 6 full-tensor passes (2 relu, cat, 2 masked min/max, masked scatter) + a zero-fill and a per-channel copy
// loop; here: one reduction pass over the signed grid (min / max of |x| over x != 0) and one write pass.
#include <cfloat>
#include "sdf_common.cuh"

namespace sdf {

constexpr int kVoxBlocks = 444;

// pass 1: per-block min / max of |x| over x != 0  ->  partial[2 * block + {0,1}]
__global__ void __launch_bounds__(256) voxel_minmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partial) {
  float mn = FLT_MAX, mx = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld_stream4(x + 4 * i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = fabsf(f4(v, j));
      if (a != 0.f) { mn = fminf(mn, a); mx = fmaxf(mx, a); }
    }
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = fabsf(x[i]);
    if (a != 0.f) { mn = fminf(mn, a); mx = fmaxf(mx, a); }
  }
  __shared__ float s_mn[8], s_mx[8];
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { mn = fminf(mn, s_mn[i]); mx = fmaxf(mx, s_mx[i]); }
    partial[2 * blockIdx.x] = mn;
    partial[2 * blockIdx.x + 1] = mx;
  }
}

struct VoxP {
  const float* x; float* out; const float* partial; float* minmax_out;
  int n_partial;
  int64_t B, bins, H, W, steps;
  int split;       // 1: x is the signed grid (B,bins,H,W); 0: x is already (B,bins,2,H,W)
  int normalize;
};

// pass 2: one thread per (b, t, h, w) writes its num_ch = 2 * bins / steps channels
__global__ void __launch_bounds__(256) voxel_prepare_kernel(const VoxP p) {
  __shared__ float s_lo, s_scale;
  if (p.normalize) {
    float mn = FLT_MAX, mx = 0.f;
    for (int i = threadIdx.x; i < p.n_partial; i += blockDim.x) { mn = fminf(mn, p.partial[2 * i]); mx = fmaxf(mx, p.partial[2 * i + 1]); }
    __shared__ float r_mn[8], r_mx[8];
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { r_mn[threadIdx.x >> 5] = mn; r_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < 8; ++i) { mn = fminf(mn, r_mn[i]); mx = fmaxf(mx, r_mx[i]); }
      const bool on = mx > mn && mn != FLT_MAX;            // reference: `if not min == max`
      s_lo = on ? mn : 0.f;
      s_scale = on ? mx - mn : 1.f;
      if (blockIdx.x == 0 && p.minmax_out) { p.minmax_out[0] = mn; p.minmax_out[1] = mx; }
    }
    __syncthreads();
  }
  const float lo = p.normalize ? s_lo : 0.f, sc = p.normalize ? s_scale : 1.f;
  const int64_t HW = p.H * p.W, groups = p.bins / p.steps, nch = 2 * groups;
  const int64_t total = p.B * p.steps * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t hw = i % HW, t = (i / HW) % p.steps, b = i / (HW * p.steps);
    float* o = p.out + i * nch;
    for (int64_t g = 0; g < groups; ++g) {
      const int64_t bin = g * p.steps + t;
      float pos, neg;
      if (p.split) {
        const float v = __ldg(p.x + (b * p.bins + bin) * HW + hw);
        pos = fmaxf(v, 0.f); neg = fmaxf(-v, 0.f);
      } else {
        pos = __ldg(p.x + ((b * p.bins + bin) * 2 + 0) * HW + hw);
        neg = __ldg(p.x + ((b * p.bins + bin) * 2 + 1) * HW + hw);
      }
      if (p.normalize) {                                    // exact op order of the reference: (x - min) / (max - min)
        if (pos != 0.f) pos = __fdiv_rn(__fsub_rn(pos, lo), sc);
        if (neg != 0.f) neg = __fdiv_rn(__fsub_rn(neg, lo), sc);
      }
      o[2 * g] = pos;
      o[2 * g + 1] = neg;
    }
  }
}

}  // namespace sdf

using namespace sdf;

extern "C" int sdf_voxel_prepare(const sdf_voxel_prepare_args* a) {
  SDF_REQUIRE(a && a->x && a->out, "sdf_voxel_prepare: null argument");
  SDF_REQUIRE(a->B > 0 && a->bins > 0 && a->H > 0 && a->W > 0 && a->steps > 0 && a->bins % a->steps == 0,
              "sdf_voxel_prepare: bins=%lld must be a positive multiple of steps=%lld", (long long)a->bins, (long long)a->steps);
  SDF_REQUIRE(!a->normalize || (a->workspace && a->split), "sdf_voxel_prepare: normalisation runs on the signed grid and needs the workspace");
  cudaStream_t st = (cudaStream_t)a->stream;
  const int64_t n_in = a->B * a->bins * a->H * a->W * (a->split ? 1 : 2);
  VoxP p{};
  p.x = a->x; p.out = a->out; p.partial = a->workspace; p.minmax_out = a->minmax; p.n_partial = kVoxBlocks;
  p.B = a->B; p.bins = a->bins; p.H = a->H; p.W = a->W; p.steps = a->steps; p.split = a->split ? 1 : 0; p.normalize = a->normalize ? 1 : 0;
  if (a->normalize) {
    SDF_REQUIRE(aligned16(a->x), "sdf_voxel_prepare: x must be 16-byte aligned");
    voxel_minmax_kernel<<<kVoxBlocks, 256, 0, st>>>(a->x, n_in, a->workspace);
    int r = finish_launch("sdf_voxel_prepare(minmax)");
    if (r) return r;
  }
  const int64_t total = a->B * a->steps * a->H * a->W;
  const int64_t need = (total + 255) / 256;
  voxel_prepare_kernel<<<(unsigned)(need < kNumSMs * 8 ? need : kNumSMs * 8), 256, 0, st>>>(p);
  return finish_launch("sdf_voxel_prepare");
}
