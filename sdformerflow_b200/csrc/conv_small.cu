// conv_small.cu — direct 3x3 / stride 1 / pad 1 convolution for a handful of input channels, channels-last.
//
// Replaces the first convolution of the patch embedding — SpikingConvEncoderLayer.conv of
// MS_PED_Spiking_PatchEmbed_Conv_sfn.head (reference models/STSwinNet_SNN/Spiking_modules.py:1737-1745,
// :270-277): 2 event-polarity channels -> embed_dim/2 on the full-resolution T*B frames — and the weight / bias
// gradient of the same layer (autograd of that nn.Conv2d).  With K = 9*Cin = 18 the library falls back to a SIMT
// implicit-GEMM at ~2 TFLOP/s (forward) and a 0.8 ms wgrad engine; the op is really bandwidth bound on the
// (N, H, W, Cout) side (4*Cout B per pixel written forward, read backward), so both directions are direct kernels:
// thread = (pixel lane, group of 4 output channels), the 9*Cin x 4 weights (forward) or weight-gradient accumulators
// (backward) of the thread's channels live in REGISTERS for the whole grid-stride loop, the 9*Cin input values of a pixel
// come through L1, pixel coordinates are advanced incrementally (no integer division in the loop).
// True fp32 (the input is real-valued, so no TF32 here).
#include "sdf_common.cuh"

namespace sdf {

struct ConvP {
  const float* x; const float* w; const float* bias; float* y;
  const float* g; float* partial;   // backward: dL/dy, per-block partial sums [blocks][(9*Cin + 1)*Cout]
  int64_t npix;       // N*H*W (< 2^31)
  int H, W, Cin, Cout, px_per_block;
};

// (row, column) of a pixel inside its image, advanced by the grid stride without divisions
struct PixWalk {
  int pix, wx, hy, step, step_w, step_h, W, H;
  __device__ __forceinline__ PixWalk(const ConvP& p, int first, int stride) {
    pix = first; W = p.W; H = p.H; step = stride;
    wx = first % W; hy = (first / W) % H;
    step_w = stride % W; step_h = (stride / W) % H;
  }
  __device__ __forceinline__ void next() {
    pix += step; wx += step_w; hy += step_h;
    if (wx >= W) { wx -= W; ++hy; }
    if (hy >= H) hy -= H;
  }
};

// the 9*CIN input values around the pixel (zero outside the image), tap-major
template <int CIN>
__device__ __forceinline__ void load_patch(const ConvP& p, const PixWalk& q, float (&xv)[9 * CIN]) {
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = q.hy + ky - 1;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = q.wx + kx - 1;
      const bool in = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      const float* xp = p.x + ((int64_t)q.pix + (ky - 1) * p.W + (kx - 1)) * CIN;
      if (CIN == 2) {
        const float2 v = in ? __ldg(reinterpret_cast<const float2*>(xp)) : make_float2(0.f, 0.f);
        xv[(ky * 3 + kx) * CIN] = v.x;
        xv[(ky * 3 + kx) * CIN + (CIN > 1 ? 1 : 0)] = v.y;
      } else if (CIN == 4) {
        const float4 v = in ? __ldg(reinterpret_cast<const float4*>(xp)) : make_float4(0.f, 0.f, 0.f, 0.f);
        xv[(ky * 3 + kx) * CIN] = v.x; xv[(ky * 3 + kx) * CIN + (CIN > 1 ? 1 : 0)] = v.y;
        xv[(ky * 3 + kx) * CIN + (CIN > 2 ? 2 : 0)] = v.z; xv[(ky * 3 + kx) * CIN + (CIN > 3 ? 3 : 0)] = v.w;
      } else {
#pragma unroll
        for (int c = 0; c < CIN; ++c) xv[(ky * 3 + kx) * CIN + c] = in ? __ldg(xp + c) : 0.f;
      }
    }
  }
}

template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_cl_kernel(const ConvP p) {
  const int cg = p.Cout / 4;                    // channel groups (threads per pixel)
  const int g = threadIdx.x % cg, lp = threadIdx.x / cg;
  if (lp >= p.px_per_block) return;
  // torch layout (Cout, Cin, 3, 3): the weights of this thread's 4 output channels, [tap*CIN + c]
  float4 w[9 * CIN];
#pragma unroll
  for (int i = 0; i < 9 * CIN; ++i) {
    const int c = i % CIN, tap = i / CIN;
    const float* wp = p.w + ((int64_t)(g * 4) * CIN + c) * 9 + tap;
    w[i] = make_float4(__ldg(wp), __ldg(wp + CIN * 9), __ldg(wp + 2 * CIN * 9), __ldg(wp + 3 * CIN * 9));
  }
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) b4 = *reinterpret_cast<const float4*>(p.bias + g * 4);
  const int npix = (int)p.npix;
  for (PixWalk q(p, blockIdx.x * p.px_per_block + lp, gridDim.x * p.px_per_block); q.pix < npix; q.next()) {
    float xv[9 * CIN];
    load_patch<CIN>(p, q, xv);
    float4 acc = b4;
#pragma unroll
    for (int i = 0; i < 9 * CIN; ++i) {         // fixed order: tap-major, then channel
      acc.x = fmaf(xv[i], w[i].x, acc.x); acc.y = fmaf(xv[i], w[i].y, acc.y);
      acc.z = fmaf(xv[i], w[i].z, acc.z); acc.w = fmaf(xv[i], w[i].w, acc.w);
    }
    st_stream4(p.y + (int64_t)q.pix * p.Cout + g * 4, acc);
  }
}

// dW[o, c, tap] = sum_pix g[pix, o] * x[pix + off(tap), c],  db[o] = sum_pix g[pix, o]: per-thread register accumulators over
// the grid-stride loop, reduced over the block's pixel lanes in shared memory (fixed order), one partial row per block.
template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_cl_wgrad_kernel(const ConvP p) {
  extern __shared__ float red[];                // [px_per_block][cg][(9*CIN + 1) * 4 + 1]
  constexpr int NA = 9 * CIN + 1;               // + the bias gradient
  const int cg = p.Cout / 4;
  const int g = threadIdx.x % cg, lp = threadIdx.x / cg;
  float4 acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lp < p.px_per_block) {
    const int npix = (int)p.npix;
    for (PixWalk q(p, blockIdx.x * p.px_per_block + lp, gridDim.x * p.px_per_block); q.pix < npix; q.next()) {
      float xv[9 * CIN];
      load_patch<CIN>(p, q, xv);
      const float4 gv = ld_stream4(p.g + (int64_t)q.pix * p.Cout + g * 4);
#pragma unroll
      for (int i = 0; i < 9 * CIN; ++i) {
        acc[i].x = fmaf(xv[i], gv.x, acc[i].x); acc[i].y = fmaf(xv[i], gv.y, acc[i].y);
        acc[i].z = fmaf(xv[i], gv.z, acc[i].z); acc[i].w = fmaf(xv[i], gv.w, acc[i].w);
      }
      acc[NA - 1].x += gv.x; acc[NA - 1].y += gv.y; acc[NA - 1].z += gv.z; acc[NA - 1].w += gv.w;
    }
  }
  constexpr int ROW = NA * 4 + 1;               // +1: odd pitch, no bank conflicts in the column walk below
  if (lp < p.px_per_block) {
    float* r = red + ((int64_t)lp * cg + g) * ROW;
#pragma unroll
    for (int i = 0; i < NA; ++i) { r[4 * i] = acc[i].x; r[4 * i + 1] = acc[i].y; r[4 * i + 2] = acc[i].z; r[4 * i + 3] = acc[i].w; }
  }
  __syncthreads();
  // element e = (g, i, j): channel o = 4g + j, accumulator i;  partial row layout: [i][o] (i = tap*CIN + c, last = bias)
  const int n_el = cg * NA * 4;
  for (int e = threadIdx.x; e < n_el; e += blockDim.x) {
    const int gg = e / (NA * 4), ij = e % (NA * 4);
    float s = 0.f;
    for (int l = 0; l < p.px_per_block; ++l) s += red[((int64_t)l * cg + gg) * ROW + ij];
    const int i = ij >> 2, j = ij & 3;
    p.partial[(int64_t)blockIdx.x * (NA * p.Cout) + (int64_t)i * p.Cout + gg * 4 + j] = s;
  }
}

// blocks summed in index order -> dW in the torch layout (Cout, Cin, 3, 3) and db [Cout]
__global__ void conv3x3_cl_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, int Cin, int Cout, float* __restrict__ dw,
                                               float* __restrict__ db) {
  const int NA = 9 * Cin + 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NA * Cout) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += __ldg(partial + (int64_t)b * (NA * Cout) + e);
  const int i = e / Cout, o = e % Cout;
  if (i == NA - 1) { if (db) db[o] = s; }
  else { const int c = i % Cin, tap = i / Cin; dw[((int64_t)o * Cin + c) * 9 + tap] = s; }
}

}  // namespace sdf

using namespace sdf;

extern "C" int sdf_conv3x3_cl_fwd(const sdf_conv3x3_cl_args* a) {
  SDF_REQUIRE(a && a->x && a->w && a->y, "sdf_conv3x3_cl_fwd: null argument");
  SDF_REQUIRE(a->Cin >= 1 && a->Cin <= 4, "sdf_conv3x3_cl_fwd: Cin=%lld not in [1,4]", (long long)a->Cin);
  SDF_REQUIRE(a->Cout % 4 == 0 && a->Cout >= 4 && a->Cout <= 512, "sdf_conv3x3_cl_fwd: Cout must be a multiple of 4, <= 512");
  SDF_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && aligned16(a->y) && (!a->bias || aligned16(a->bias)), "sdf_conv3x3_cl_fwd: bad shape/alignment");
  SDF_REQUIRE(a->N * a->H * a->W < (1LL << 31) - (1 << 24) && a->Cout <= 128, "sdf_conv3x3_cl_fwd: N*H*W must be < 2^31, Cout <= 128");
  ConvP p;
  p.x = a->x; p.w = a->w; p.bias = a->bias; p.y = a->y;
  p.npix = a->N * a->H * a->W; p.H = (int)a->H; p.W = (int)a->W; p.Cin = (int)a->Cin; p.Cout = (int)a->Cout;
  const int cg = p.Cout / 4;
  p.px_per_block = 256 / cg;
  if (p.px_per_block < 1) { p.px_per_block = 1; }
  const int threads = cg * p.px_per_block > 256 ? cg : 256;
  SDF_REQUIRE(cg <= 256 || true, "unreachable");
  const size_t smem = 0;
  int64_t need = (p.npix + p.px_per_block - 1) / p.px_per_block;
  const int blocks = (int)(need < kNumSMs * 8 ? need : kNumSMs * 8);
  cudaStream_t stream = (cudaStream_t)a->stream;
  switch (p.Cin) {
    case 1: conv3x3_cl_kernel<1><<<blocks, threads, smem, stream>>>(p); break;
    case 2: conv3x3_cl_kernel<2><<<blocks, threads, smem, stream>>>(p); break;
    case 3: conv3x3_cl_kernel<3><<<blocks, threads, smem, stream>>>(p); break;
    default: conv3x3_cl_kernel<4><<<blocks, threads, smem, stream>>>(p); break;
  }
  return finish_launch("sdf_conv3x3_cl_fwd");
}

static int conv_small_setup(int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout, ConvP* p, int* threads) {
  p->npix = N * H * W; p->H = (int)H; p->W = (int)W; p->Cin = (int)Cin; p->Cout = (int)Cout;
  const int cg = p->Cout / 4;
  p->px_per_block = 256 / cg;
  if (p->px_per_block < 1) p->px_per_block = 1;
  *threads = cg * p->px_per_block > 256 ? cg : 256;
  return SDF_OK;
}

extern "C" int64_t sdf_conv3x3_cl_wgrad_workspace_bytes(int64_t Cin, int64_t Cout) {
  return (int64_t)kNumSMs * 2 * (9 * Cin + 1) * Cout * 4;
}

extern "C" int sdf_conv3x3_cl_wgrad(const sdf_conv3x3_cl_wgrad_args* a) {
  SDF_REQUIRE(a && a->x && a->g && a->dw && a->workspace, "sdf_conv3x3_cl_wgrad: null argument");
  SDF_REQUIRE(a->Cin >= 1 && a->Cin <= 4, "sdf_conv3x3_cl_wgrad: Cin=%lld not in [1,4]", (long long)a->Cin);
  SDF_REQUIRE(a->Cout % 4 == 0 && a->Cout >= 4 && a->Cout <= 128, "sdf_conv3x3_cl_wgrad: Cout must be a multiple of 4, <= 128");
  SDF_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && aligned16(a->g) && aligned16(a->x), "sdf_conv3x3_cl_wgrad: bad shape/alignment");
  SDF_REQUIRE(a->N * a->H * a->W < (1LL << 31) - (1 << 24), "sdf_conv3x3_cl_wgrad: N*H*W must be < 2^31");
  SDF_REQUIRE(a->workspace_bytes >= sdf_conv3x3_cl_wgrad_workspace_bytes(a->Cin, a->Cout), "sdf_conv3x3_cl_wgrad: workspace too small");
  ConvP p = {};
  int threads;
  conv_small_setup(a->N, a->H, a->W, a->Cin, a->Cout, &p, &threads);
  p.x = a->x; p.g = a->g; p.partial = a->workspace;
  const int cg = p.Cout / 4;
  const int NA = 9 * p.Cin + 1;
  int64_t need = (p.npix + p.px_per_block - 1) / p.px_per_block;
  const int blocks = (int)(need < kNumSMs * 2 ? need : kNumSMs * 2);
  const size_t smem = sizeof(float) * (size_t)p.px_per_block * cg * (NA * 4 + 1);
  SDF_REQUIRE(smem <= 200 * 1024, "sdf_conv3x3_cl_wgrad: shared-memory plan too large");
  cudaStream_t stream = (cudaStream_t)a->stream;
#define WG_CASE(C)                                                                                                        \
  {                                                                                                                       \
    static bool attr = false;                                                                                             \
    if (!attr) {                                                                                                          \
      cudaFuncSetAttribute(conv3x3_cl_wgrad_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);          \
      attr = true;                                                                                                        \
    }                                                                                                                     \
    conv3x3_cl_wgrad_kernel<C><<<blocks, threads, smem, stream>>>(p);                                                     \
  }
  switch (p.Cin) {
    case 1: WG_CASE(1) break;
    case 2: WG_CASE(2) break;
    case 3: WG_CASE(3) break;
    default: WG_CASE(4) break;
  }
#undef WG_CASE
  int st = finish_launch("sdf_conv3x3_cl_wgrad");
  if (st) return st;
  const int n = NA * p.Cout;
  conv3x3_cl_wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a->workspace, blocks, p.Cin, p.Cout, a->dw, a->db);
  return finish_launch("sdf_conv3x3_cl_wgrad(reduce)");
}
