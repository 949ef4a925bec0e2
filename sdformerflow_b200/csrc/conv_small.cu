// conv_small.cu — direct 3x3 / stride 1 / pad 1 convolution for a handful of input channels, channels-last.
//
// Replaces the first convolution of the patch embedding — SpikingConvEncoderLayer.conv of
// MS_PED_Spiking_PatchEmbed_Conv_sfn.head (reference models/STSwinNet_SNN/Spiking_modules.py:1737-1745,
// :270-277): 2 event-polarity channels -> embed_dim/2 on the full-resolution T*B frames.  With K = 9*Cin = 18
// the library falls back to a SIMT implicit-GEMM at ~2 TFLOP/s; the op is really output-bandwidth bound
// (4*Cout B written per pixel), so a direct kernel with the 9*Cin*Cout weights in shared memory runs it at
// HBM speed in true fp32 (the input is real-valued, so no TF32 here).
#include "sdf_common.cuh"

namespace sdf {

struct ConvP {
  const float* x; const float* w; const float* bias; float* y;
  int64_t npix;       // N*H*W
  int H, W, Cin, Cout, px_per_block;
};

template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_cl_kernel(const ConvP p) {
  extern __shared__ float ws[];                 // [9*CIN][Cout]
  const int cg = p.Cout / 4;                    // channel groups (threads per pixel)
  for (int i = threadIdx.x; i < 9 * CIN * p.Cout; i += blockDim.x) {
    // torch layout (Cout, Cin, 3, 3) -> [tap*CIN + c][o]
    const int o = i % p.Cout, tc = i / p.Cout, c = tc % CIN, tap = tc / CIN;
    ws[i] = p.w[((int64_t)o * CIN + c) * 9 + tap];
  }
  __syncthreads();
  const int g = threadIdx.x % cg, lp = threadIdx.x / cg;
  if (lp >= p.px_per_block) return;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) b4 = *reinterpret_cast<const float4*>(p.bias + g * 4);
  for (int64_t pix = (int64_t)blockIdx.x * p.px_per_block + lp; pix < p.npix; pix += (int64_t)gridDim.x * p.px_per_block) {
    const int wx = (int)(pix % p.W);
    const int hy = (int)((pix / p.W) % p.H);
    float4 acc = b4;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = hy + ky - 1;
      if (yy < 0 || yy >= p.H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = wx + kx - 1;
        if (xx < 0 || xx >= p.W) continue;
        const float* xp = p.x + (pix + (int64_t)(ky - 1) * p.W + (kx - 1)) * CIN;
        const float4* wp = reinterpret_cast<const float4*>(ws + (ky * 3 + kx) * CIN * p.Cout) + g;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float xv = __ldg(xp + c);
          const float4 wv = wp[c * cg];
          acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y);
          acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
        }
      }
    }
    st_stream4(p.y + pix * p.Cout + g * 4, acc);
  }
}

}  // namespace sdf

using namespace sdf;

extern "C" int sdf_conv3x3_cl_fwd(const sdf_conv3x3_cl_args* a) {
  SDF_REQUIRE(a && a->x && a->w && a->y, "sdf_conv3x3_cl_fwd: null argument");
  SDF_REQUIRE(a->Cin >= 1 && a->Cin <= 4, "sdf_conv3x3_cl_fwd: Cin=%lld not in [1,4]", (long long)a->Cin);
  SDF_REQUIRE(a->Cout % 4 == 0 && a->Cout >= 4 && a->Cout <= 512, "sdf_conv3x3_cl_fwd: Cout must be a multiple of 4, <= 512");
  SDF_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && aligned16(a->y) && (!a->bias || aligned16(a->bias)), "sdf_conv3x3_cl_fwd: bad shape/alignment");
  ConvP p;
  p.x = a->x; p.w = a->w; p.bias = a->bias; p.y = a->y;
  p.npix = a->N * a->H * a->W; p.H = (int)a->H; p.W = (int)a->W; p.Cin = (int)a->Cin; p.Cout = (int)a->Cout;
  const int cg = p.Cout / 4;
  p.px_per_block = 256 / cg;
  if (p.px_per_block < 1) { p.px_per_block = 1; }
  const int threads = cg * p.px_per_block > 256 ? cg : 256;
  SDF_REQUIRE(cg <= 256 || true, "unreachable");
  const size_t smem = sizeof(float) * 9 * p.Cin * p.Cout;
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
