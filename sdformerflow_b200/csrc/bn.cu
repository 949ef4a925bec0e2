// bn.cu — K6: BatchNorm over channels-last rows, split so that the apply half can live in the
// prologue of the consuming kernel (K1 LIF, window scatter):
//   sdf_bn_stats        per-block partial (sum, sum of squares)          4 B/element read
//   sdf_bn_finalize     partials -> mean/rstd -> scale/shift (+ running-stat update as torch)
//   sdf_bn_apply        out = res + (u*scale + shift)  (MLP tail + residual)
//   sdf_bn_bwd_reduce   partial (sum dy, sum dy*u) for BN sites not preceded by a neuron
//   sdf_bn_bwd_finalize partials -> dweight, dbias and the coefficients of du = a*dy + b*u + c
//   sdf_bn_bwd_apply    du = a*dy + b*u + c
// Replaces sj_layer.BatchNorm2d reached via SpikingNormLayer (reference
// models/STSwinNet_SNN/Spiking_modules.py:101-146) on the permuted, non-contiguous views of
// Spiking_swin_transformer3D.py:153,159,310,314,318,367,673,677,714,933,972 — those permutes
// only move the channel axis, so on a channels-last [rows, C] buffer they are no-ops.
#include "sdf_common.cuh"

namespace sdf {

struct RowsP {
  const float* a; int64_t ld_a;
  const float* b; int64_t ld_b;
  float* out; int64_t ld_out;
  const float* v0; const float* v1; const float* v2;
  float* partials;
  int64_t rows, C, tile_w;
  int R, k;
};

// MODE 0: stats (sum a, sum a^2); MODE 1: bwd reduce (sum a, sum a*b)
template <int MODE>
__global__ void __launch_bounds__(512) bn_reduce_kernel(const RowsP p) {
  extern __shared__ float smem[];
  const int rx = threadIdx.x % p.R, ry = threadIdx.x / p.R;
  const int64_t col = (int64_t)blockIdx.y * p.tile_w + (int64_t)rx * 4;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  // 4 independent row loads in flight per thread
  const int64_t stride = (int64_t)gridDim.x * p.k;
  int64_t row = (int64_t)blockIdx.x * p.k + ry;
  for (; row + 3 * stride < p.rows; row += 4 * stride) {
    float4 x[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = ld_stream4(p.a + (row + j * stride) * p.ld_a + col);
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = ld_stream4(p.b + (row + j * stride) * p.ld_b + col);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float xv = f4(x[j], i);
        acc[0][i] += xv;
        acc[1][i] += xv * (MODE == 1 ? f4(y[j], i) : xv);
      }
    }
  }
  for (; row < p.rows; row += stride) {
    float4 x = ld_stream4(p.a + row * p.ld_a + col);
    float4 y = x;
    if (MODE == 1) y = ld_stream4(p.b + row * p.ld_b + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[0][i] += f4(x, i);
      acc[1][i] += f4(x, i) * f4(y, i);
    }
  }
  block_reduce_rows_to_partials<2>(acc, smem, p.partials, p.R, p.k, p.C, (int64_t)blockIdx.y * p.tile_w);
}

// MODE 0: out = res + (a*scale + shift)   (v0 = scale, v1 = shift, b = res; any may be NULL)
// MODE 1: out = v0*a + v1*b + v2          (BN backward apply)
template <int MODE>
__global__ void __launch_bounds__(512) bn_rows_kernel(const RowsP p) {
  const int rx = threadIdx.x % p.R, ry = threadIdx.x / p.R;
  const int64_t col = (int64_t)blockIdx.y * p.tile_w + (int64_t)rx * 4;
  float4 c0 = make_float4(1.f, 1.f, 1.f, 1.f), c1 = make_float4(0.f, 0.f, 0.f, 0.f), c2 = c1;
  if (p.v0) c0 = *reinterpret_cast<const float4*>(p.v0 + col);
  if (p.v1) c1 = *reinterpret_cast<const float4*>(p.v1 + col);
  if (p.v2) c2 = *reinterpret_cast<const float4*>(p.v2 + col);
  const int64_t stride = (int64_t)gridDim.x * p.k;
  int64_t row = (int64_t)blockIdx.x * p.k + ry;
  // BN backward apply: two rows per iteration, four 16-byte loads in flight per thread (0.78 -> 0.86 of HBM in the step); the
  // forward apply + residual (MODE 0) measured slower that way (1.20 -> 1.35 ms per step) and keeps one row per iteration
  for (; MODE == 1 && row + stride < p.rows; row += 2 * stride) {
    const int64_t r1 = row + stride;
    float4 x0 = ld_stream4(p.a + row * p.ld_a + col), x1 = ld_stream4(p.a + r1 * p.ld_a + col);
    float4 y0 = make_float4(0.f, 0.f, 0.f, 0.f), y1 = y0;
    if (p.b) { y0 = ld_stream4(p.b + row * p.ld_b + col); y1 = ld_stream4(p.b + r1 * p.ld_b + col); }
    float4 o0, o1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (MODE == 0) {
        f4(o0, i) = f4(y0, i) + fmaf(f4(x0, i), f4(c0, i), f4(c1, i));
        f4(o1, i) = f4(y1, i) + fmaf(f4(x1, i), f4(c0, i), f4(c1, i));
      } else {
        f4(o0, i) = fmaf(f4(c0, i), f4(x0, i), fmaf(f4(c1, i), f4(y0, i), f4(c2, i)));
        f4(o1, i) = fmaf(f4(c0, i), f4(x1, i), fmaf(f4(c1, i), f4(y1, i), f4(c2, i)));
      }
    }
    st_stream4(p.out + row * p.ld_out + col, o0);
    st_stream4(p.out + r1 * p.ld_out + col, o1);
  }
  for (; row < p.rows; row += stride) {
    float4 x = ld_stream4(p.a + row * p.ld_a + col);
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.b) y = ld_stream4(p.b + row * p.ld_b + col);
    float4 o;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (MODE == 0) f4(o, i) = f4(y, i) + fmaf(f4(x, i), f4(c0, i), f4(c1, i));
      else f4(o, i) = fmaf(f4(c0, i), f4(x, i), fmaf(f4(c1, i), f4(y, i), f4(c2, i)));
    }
    st_stream4(p.out + row * p.ld_out + col, o);
  }
}

struct FinP {
  const float* partials; int64_t nblk; int64_t count; int64_t C;
  const float* weight; const float* bias; float* running_mean; float* running_var;
  float momentum, eps; int training;
  float* scale; float* shift; float* mean; float* rstd;
  // backward
  const float* mean_in; const float* rstd_in; float* gw; float* gb; float* coef;
};

// Block of 32 channels x 32 row groups: thread (rg, cl) sums the partial rows rg, rg + 32, ... of channel c0 + cl (a warp
// reads 32 consecutive channels of one row: coalesced; ~14 independent loads per thread instead of a serial chain of 14 per
// lane), the 32 row-group sums are then added in index order by the rg == 0 threads.  Everything in double, fixed order:
// deterministic, and the single rounding to fp32 at the end keeps mean/var within 1 ulp of exact.
constexpr int kFinThreads = 1024;
constexpr int kFinRows = 14;    // partial rows per thread and round: 32 row groups x 14 = 448 >= the 444-row workspaces
__device__ __forceinline__ void block_sum_partials(const float* __restrict__ partials, int64_t nblk, int64_t C, int64_t c, double& s0,
                                                   double& s1) {
  __shared__ double sh[2][32][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  double a = 0.0, b = 0.0;
  if (c < C) {
    // all the loads of a round are issued before the first add (one memory round trip per 448 rows, not one per 4 rows:
    // these kernels are pure latency — a few CTAs reading ~100 KB that the producer has just written)
    for (int64_t base = rg; base < nblk; base += 32 * kFinRows) {
      float va[kFinRows], vb[kFinRows];
#pragma unroll
      for (int i = 0; i < kFinRows; ++i) {
        const int64_t blk = base + 32 * i;
        const bool ok = blk < nblk;
        va[i] = ok ? __ldg(partials + (blk * 2 + 0) * C + c) : 0.f;
        vb[i] = ok ? __ldg(partials + (blk * 2 + 1) * C + c) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kFinRows; ++i) { a += (double)va[i]; b += (double)vb[i]; }
    }
  }
  sh[0][rg][cl] = a;
  sh[1][rg][cl] = b;
  __syncthreads();
  a = 0.0; b = 0.0;
  if (rg == 0) {
#pragma unroll 8
    for (int r = 0; r < 32; ++r) { a += sh[0][r][cl]; b += sh[1][r][cl]; }
  }
  s0 = a; s1 = b;
}

__global__ void __launch_bounds__(kFinThreads) bn_finalize_kernel(const FinP p) {
  const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const bool lead = (threadIdx.x >> 5) == 0 && c < p.C;
  // parameters and running statistics are read BEFORE the reduction, so that their (cold) round trips overlap with it
  // instead of forming a chain of dependent load -> store pairs behind it
  float w = 1.f, b = 0.f, rm = 0.f, rv = 1.f;
  if (lead) {
    if (p.weight) w = p.weight[c];
    if (p.bias) b = p.bias[c];
    if (p.running_mean) rm = p.running_mean[c];
    if (p.running_var) rv = p.running_var[c];
  }
  float mean = 0.f, var = 1.f;
  if (p.training) {
    double s, ss;
    block_sum_partials(p.partials, p.nblk, p.C, c, s, ss);
    if (!lead) return;
    const double n = (double)p.count;
    const double m = s / n;
    double v = ss / n - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (p.running_mean) p.running_mean[c] = (1.f - p.momentum) * rm + p.momentum * mean;
    if (p.running_var) {
      const float unbiased = (float)(v * (n / (n > 1.0 ? n - 1.0 : 1.0)));
      p.running_var[c] = (1.f - p.momentum) * rv + p.momentum * unbiased;
    }
  } else {
    if (!lead) return;
    mean = rm;
    var = rv;
  }
  const float rstd = 1.f / sqrtf(var + p.eps);
  const float sc = w * rstd;
  p.scale[c] = sc;
  p.shift[c] = b - mean * sc;
  if (p.mean) p.mean[c] = mean;
  if (p.rstd) p.rstd[c] = rstd;
}

__global__ void __launch_bounds__(kFinThreads) bn_bwd_finalize_kernel(const FinP p) {
  const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const bool lead = (threadIdx.x >> 5) == 0 && c < p.C;
  double mean = 0.0, rstd = 1.0, w = 1.0;
  if (lead) {                     // issued before the reduction (see bn_finalize_kernel)
    mean = p.mean_in[c]; rstd = p.rstd_in[c];
    if (p.weight) w = (double)p.weight[c];
  }
  double s, su;
  block_sum_partials(p.partials, p.nblk, p.C, c, s, su);
  if (!lead) return;
  const double dgamma = (su - mean * s) * rstd;  // sum dy * xhat
  const double dbeta = s;
  if (p.gw) p.gw[c] = (float)dgamma;
  if (p.gb) p.gb[c] = (float)dbeta;
  if (p.coef) {
    if (p.training) {
      // du = w*rstd * (dy - dbeta/n - xhat * dgamma/n),  xhat = (u - mean)*rstd
      const double n = (double)p.count;
      const double a = w * rstd;
      const double k2 = a * rstd * dgamma / n;
      p.coef[c] = (float)a;
      p.coef[p.C + c] = (float)(-k2);
      p.coef[2 * p.C + c] = (float)(-a * dbeta / n + k2 * mean);
    } else {
      p.coef[c] = (float)(w * rstd);
      p.coef[p.C + c] = 0.f;
      p.coef[2 * p.C + c] = 0.f;
    }
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = w[i];
    const float h = __uint_as_float((__float_as_uint(x) + 0x1000u) & ~0x1FFFu);   // round to nearest on the 13 dropped bits
    const float l = __fsub_rn(x, h);                                              // exact
    hi[i] = h;
    lo[i] = __uint_as_float((__float_as_uint(l) + 0x1000u) & ~0x1FFFu);
  }
}

static int rows_launch_setup(int64_t rows, int64_t C, int64_t max_blocks, RowsP* p, dim3* grid, int* threads) {
  RowTiling rt;
  SDF_REQUIRE(C % 4 == 0 && make_row_tiling(rows, C, 4, 512, (int)max_blocks, &rt), "bn: C=%lld must be a multiple of 4", (long long)C);
  p->rows = rows; p->C = C; p->tile_w = rt.tile_w; p->R = rt.R; p->k = rt.k;
  *grid = dim3(rt.blocks, rt.ncol, 1);
  *threads = rt.threads;
  return SDF_OK;
}

}  // namespace sdf

using namespace sdf;

extern "C" int sdf_bn_stats(const sdf_bn_stats_args* a) {
  SDF_REQUIRE(a && a->x && a->partials, "sdf_bn_stats: null argument");
  SDF_REQUIRE(a->rows > 0 && a->ld >= a->C && a->ld % 4 == 0 && aligned16(a->x), "sdf_bn_stats: bad shape/alignment");
  SDF_REQUIRE(a->n_partial_blocks >= 1, "sdf_bn_stats: n_partial_blocks < 1");
  RowsP p = {};
  p.a = a->x; p.ld_a = a->ld; p.partials = a->partials;
  dim3 grid; int threads;
  int64_t cap = a->n_partial_blocks < kNumSMs * 2 ? a->n_partial_blocks : kNumSMs * 2;
  int st = rows_launch_setup(a->rows, a->C, cap, &p, &grid, &threads);
  if (st) return st;
  cudaStream_t stream = (cudaStream_t)a->stream;
  if (a->n_partial_blocks > (int64_t)grid.x)
    cudaMemsetAsync(a->partials + (int64_t)grid.x * 2 * a->C, 0, sizeof(float) * (a->n_partial_blocks - grid.x) * 2 * a->C, stream);
  bn_reduce_kernel<0><<<grid, threads, sizeof(float) * 4 * threads, stream>>>(p);
  return finish_launch("sdf_bn_stats");
}

extern "C" int sdf_bn_bwd_reduce(const sdf_bn_bwd_reduce_args* a) {
  SDF_REQUIRE(a && a->dy && a->u && a->partials, "sdf_bn_bwd_reduce: null argument");
  SDF_REQUIRE(a->rows > 0 && a->ld_u >= a->C && a->ld_u % 4 == 0 && aligned16(a->dy) && aligned16(a->u), "sdf_bn_bwd_reduce: bad shape/alignment");
  SDF_REQUIRE(a->n_partial_blocks >= 1, "sdf_bn_bwd_reduce: n_partial_blocks < 1");
  RowsP p = {};
  p.a = a->dy; p.ld_a = a->C; p.b = a->u; p.ld_b = a->ld_u; p.partials = a->partials;
  dim3 grid; int threads;
  int64_t cap = a->n_partial_blocks < kNumSMs * 2 ? a->n_partial_blocks : kNumSMs * 2;
  int st = rows_launch_setup(a->rows, a->C, cap, &p, &grid, &threads);
  if (st) return st;
  cudaStream_t stream = (cudaStream_t)a->stream;
  if (a->n_partial_blocks > (int64_t)grid.x)
    cudaMemsetAsync(a->partials + (int64_t)grid.x * 2 * a->C, 0, sizeof(float) * (a->n_partial_blocks - grid.x) * 2 * a->C, stream);
  bn_reduce_kernel<1><<<grid, threads, sizeof(float) * 4 * threads, stream>>>(p);
  return finish_launch("sdf_bn_bwd_reduce");
}

extern "C" int sdf_bn_finalize(const sdf_bn_finalize_args* a) {
  SDF_REQUIRE(a && a->scale && a->shift && a->C > 0, "sdf_bn_finalize: null argument");
  if (a->training) SDF_REQUIRE(a->partials && a->count > 0 && a->n_partial_blocks > 0, "sdf_bn_finalize: training needs partials");
  else SDF_REQUIRE(a->running_mean && a->running_var, "sdf_bn_finalize: eval needs running statistics");
  FinP p = {};
  p.partials = a->partials; p.nblk = a->n_partial_blocks; p.count = a->count; p.C = a->C;
  p.weight = a->weight; p.bias = a->bias; p.running_mean = a->running_mean; p.running_var = a->running_var;
  p.momentum = (float)a->momentum; p.eps = (float)a->eps; p.training = a->training;
  p.scale = a->scale; p.shift = a->shift; p.mean = a->mean; p.rstd = a->rstd;
  bn_finalize_kernel<<<(unsigned)((a->C + 31) / 32), kFinThreads, 0, (cudaStream_t)a->stream>>>(p);
  return finish_launch("sdf_bn_finalize");
}

extern "C" int sdf_bn_bwd_finalize(const sdf_bn_bwd_finalize_args* a) {
  SDF_REQUIRE(a && a->partials && a->mean && a->rstd && a->C > 0 && a->count > 0, "sdf_bn_bwd_finalize: null argument");
  FinP p = {};
  p.partials = a->partials; p.nblk = a->n_partial_blocks; p.count = a->count; p.C = a->C;
  p.weight = a->weight; p.mean_in = a->mean; p.rstd_in = a->rstd; p.gw = a->grad_weight; p.gb = a->grad_bias;
  p.coef = a->coef; p.training = a->training;
  bn_bwd_finalize_kernel<<<(unsigned)((a->C + 31) / 32), kFinThreads, 0, (cudaStream_t)a->stream>>>(p);
  return finish_launch("sdf_bn_bwd_finalize");
}

extern "C" int sdf_bn_bwd_apply(const sdf_bn_bwd_apply_args* a) {
  SDF_REQUIRE(a && a->dy && a->u && a->du && a->coef, "sdf_bn_bwd_apply: null argument");
  SDF_REQUIRE(a->rows > 0 && a->ld_u % 4 == 0 && a->ld_du % 4 == 0 && aligned16(a->dy) && aligned16(a->u) && aligned16(a->du),
              "sdf_bn_bwd_apply: bad shape/alignment");
  RowsP p = {};
  p.a = a->dy; p.ld_a = a->C; p.b = a->u; p.ld_b = a->ld_u; p.out = a->du; p.ld_out = a->ld_du;
  p.v0 = a->coef; p.v1 = a->coef + a->C; p.v2 = a->coef + 2 * a->C;
  dim3 grid; int threads;
  int st = rows_launch_setup(a->rows, a->C, kNumSMs * 4, &p, &grid, &threads);
  if (st) return st;
  bn_rows_kernel<1><<<grid, threads, 0, (cudaStream_t)a->stream>>>(p);
  return finish_launch("sdf_bn_bwd_apply");
}

extern "C" int sdf_bn_apply(const sdf_bn_apply_args* a) {
  SDF_REQUIRE(a && a->u && a->out, "sdf_bn_apply: null argument");
  SDF_REQUIRE(a->rows > 0 && a->ld_u % 4 == 0 && aligned16(a->u) && aligned16(a->out) && (!a->res || aligned16(a->res)),
              "sdf_bn_apply: bad shape/alignment");
  SDF_REQUIRE((a->scale == nullptr) == (a->shift == nullptr), "sdf_bn_apply: scale and shift go together");
  RowsP p = {};
  p.a = a->u; p.ld_a = a->ld_u; p.b = a->res; p.ld_b = a->C; p.out = a->out; p.ld_out = a->C;
  p.v0 = a->scale; p.v1 = a->shift;
  dim3 grid; int threads;
  int st = rows_launch_setup(a->rows, a->C, kNumSMs * 4, &p, &grid, &threads);
  if (st) return st;
  bn_rows_kernel<0><<<grid, threads, 0, (cudaStream_t)a->stream>>>(p);
  return finish_launch("sdf_bn_apply");
}

extern "C" int sdf_split_tf32(const sdf_split_tf32_args* a) {
  SDF_REQUIRE(a && a->w && a->hi && a->lo && a->n >= 0, "sdf_split_tf32: null argument");
  if (a->n == 0) return SDF_OK;
  const int threads = 256;
  int64_t need = (a->n + threads - 1) / threads;
  split_tf32_kernel<<<(unsigned)(need < kNumSMs * 8 ? need : kNumSMs * 8), threads, 0, (cudaStream_t)a->stream>>>(a->w, a->hi, a->lo, a->n);
  return finish_launch("sdf_split_tf32");
}
