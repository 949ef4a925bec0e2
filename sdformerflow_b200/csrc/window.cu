// window.cu — 3D shifted-window index algebra folded into kernels.
//
// Replaces, in reference models/STSwinNet_SNN/Spiking_swin_transformer3D.py:
//   F.pad (:793) + torch.roll (:797) + window_partition_v2 (:100-113)           -> win2x gather
//   .view + window_reverse (swin_transformer3D_v2.py:52-65) + roll back (:815)
//     + crop (:820) + DropPath (:766) + shortcut add (:840)                     -> win2x scatter
//   compute_mask (:980-993)                                                     -> region ids
//   proj_sn over the fake time axis (:670, :425)                                -> sdf_lif_window_*
//   2x2 patch-merging gather + LIF (:952-974, :914-935)                         -> sdf_lif_merge_*
//
// The window buffer row index rho = (w*wd + dd)*P + pos is also (t*M + m)*P + pos for the
// (wd, M, wh, ww, C) reinterpretation the reference makes, so one int32 table `win2x[rho]`
// (token row in the (B,D,H,W) map, or -1 for zero padding) drives every kernel here.  The table
// costs 4 B per token against 4*C B of activations and is cached per (shape, shift) by the host.
#include "sdf_common.cuh"

namespace sdf {

struct WinIdxP {
  int64_t B, D, H, W, wd, wh, ww, sd, sh, sw;
  int64_t Dp, Hp, Wp, nD, nH, nW, N, P, nWin, total;
  int32_t* win2x;
  uint8_t* region;
};

__device__ __forceinline__ int axis_region(int64_t g, int64_t L, int64_t ws, int64_t sh) {
  // compute_mask slices: [:-ws] -> 0, [-ws:-sh] -> 1, [-sh:] -> 2; sh == 0 makes the last slice
  // cover the whole axis (slice(-0, None)), i.e. a single region.
  if (sh == 0) return 2;
  if (g < L - ws) return 0;
  if (g < L - sh) return 1;
  return 2;
}

__global__ void window_index_kernel(const WinIdxP p) {
  for (int64_t rho = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; rho < p.total;
       rho += (int64_t)gridDim.x * blockDim.x) {
    const int64_t wg = rho / p.N, n = rho - wg * p.N;
    const int64_t dd = n / p.P, pos = n - dd * p.P;
    const int64_t ph = pos / p.ww, pw = pos - ph * p.ww;
    const int64_t b = wg / p.nWin, w = wg - b * p.nWin;
    const int64_t wi_d = w / (p.nH * p.nW), rem = w - wi_d * (p.nH * p.nW);
    const int64_t wi_h = rem / p.nW, wi_w = rem - wi_h * p.nW;
    // coordinates on the shifted, padded grid
    const int64_t gd = wi_d * p.wd + dd, gh = wi_h * p.wh + ph, gw = wi_w * p.ww + pw;
    // shifted[g] = padded[(g + s) % L]   (torch.roll by -s)
    int64_t d = gd + p.sd; if (d >= p.Dp) d -= p.Dp;
    int64_t h = gh + p.sh; if (h >= p.Hp) h -= p.Hp;
    int64_t x = gw + p.sw; if (x >= p.Wp) x -= p.Wp;
    int32_t v = -1;
    if (d < p.D && h < p.H && x < p.W) v = (int32_t)(((b * p.D + d) * p.H + h) * p.W + x);
    p.win2x[rho] = v;
    if (p.region && b == 0) {
      const int rd = axis_region(gd, p.Dp, p.wd, p.sd);
      const int rh = axis_region(gh, p.Hp, p.wh, p.sh);
      const int rw = axis_region(gw, p.Wp, p.ww, p.sw);
      p.region[rho] = (uint8_t)((rd * 3 + rh) * 3 + rw);
    }
  }
}

// ---- row kernels driven by win2x ------------------------------------------------------------
struct WinRowsP {
  const float* src; float* dst; const float* res; const float* u;
  const int32_t* win2x;
  const float* scale; const float* shift; const float* alpha;
  float* partials;
  int64_t rows, rows_per_sample, C, tile_w;
  int R, k;
};

// MODE 0: gather        dst[rho] = src[win2x[rho]] or 0
// MODE 1: gather bwd    dst[win2x[rho]] = src[rho]
// MODE 2: scatter       dst[x] = res[x] + alpha[b]*(src[rho]*scale + shift)
// MODE 3: scatter bwd   dst[rho] = alpha[b]*src[x] or 0  (+ partial sums of dst, dst*u)
template <int MODE>
__global__ void __launch_bounds__(512) window_rows_kernel(const WinRowsP p) {
  extern __shared__ float smem[];
  const int rx = threadIdx.x % p.R, ry = threadIdx.x / p.R;
  const int64_t col = (int64_t)blockIdx.y * p.tile_w + (int64_t)rx * 4;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MODE == 2 && p.scale) {
    sc = *reinterpret_cast<const float4*>(p.scale + col);
    sh = *reinterpret_cast<const float4*>(p.shift + col);
  }
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const int64_t stride = (int64_t)gridDim.x * p.k;
  for (int64_t rho = (int64_t)blockIdx.x * p.k + ry; rho < p.rows; rho += stride) {
    const int64_t xr = p.win2x[rho];
    const float* wptr_c = nullptr; (void)wptr_c;
    if (MODE == 0) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (xr >= 0) v = ld_stream4(p.src + xr * p.C + col);
      st_stream4(p.dst + rho * p.C + col, v);
    } else if (MODE == 1) {
      if (xr >= 0) st_stream4(p.dst + xr * p.C + col, ld_stream4(p.src + rho * p.C + col));
    } else if (MODE == 2) {
      if (xr >= 0) {
        float4 y = ld_stream4(p.src + rho * p.C + col);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.res) r = ld_stream4(p.res + xr * p.C + col);
        const float a = p.alpha ? __ldg(p.alpha + rho / p.rows_per_sample) : 1.f;
        float4 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float t = fmaf(f4(y, i), f4(sc, i), f4(sh, i));
          f4(o, i) = p.alpha ? f4(r, i) + t * a : f4(r, i) + t;
        }
        st_stream4(p.dst + xr * p.C + col, o);
      }
    } else {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (xr >= 0) {
        g = ld_stream4(p.src + xr * p.C + col);
        if (p.alpha) {
          const float a = __ldg(p.alpha + rho / p.rows_per_sample);
          g.x *= a; g.y *= a; g.z *= a; g.w *= a;
        }
      }
      st_stream4(p.dst + rho * p.C + col, g);
      if (p.partials && xr >= 0) {
        float4 uu = ld_stream4(p.u + rho * p.C + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[0][i] += f4(g, i);
          acc[1][i] += f4(g, i) * f4(uu, i);
        }
      }
    }
  }
  if (MODE == 3 && p.partials)
    block_reduce_rows_to_partials<2>(acc, smem, p.partials, p.R, p.k, p.C, (int64_t)blockIdx.y * p.tile_w);
}

// ---- LIF over the fake time axis with the gather folded in ------------------------------
struct LifWinP {
  const float* x; void* spike; float* h_seq; const float* gs; float* gx;
  const int32_t* win2x;
  int64_t MP, C, tile_w;
  int R, k, wd;
  NeuronP nrn;
};

template <int T, int DT>
__global__ void __launch_bounds__(512) lif_window_fwd_kernel(const LifWinP p) {
  constexpr int TM = T > 0 ? T : 8;
  const int Tn = T > 0 ? T : p.wd;
  const NeuronP nrn = p.nrn;
  const int rx = threadIdx.x % p.R, ry = threadIdx.x / p.R;
  const int64_t col = (int64_t)blockIdx.y * p.tile_w + (int64_t)rx * 4;
  const int64_t stride = (int64_t)gridDim.x * p.k;
  for (int64_t nr = (int64_t)blockIdx.x * p.k + ry; nr < p.MP; nr += stride) {
    int64_t xr[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) xr[t] = p.win2x[t * p.MP + nr];
    float4 x[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) x[t] = xr[t] >= 0 ? ld_stream4(p.x + xr[t] * p.C + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = nrn.hard ? nrn.v_reset : 0.f;
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (t < Tn) {
        float4 h, s;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f4(h, i) = neuron_charge(nrn, v[i], f4(x[t], i));
          f4(s, i) = neuron_fire(nrn, f4(h, i));
          v[i] = neuron_reset(nrn, f4(h, i), f4(s, i));
        }
        const int64_t o = (t * p.MP + nr) * p.C + col;
        store_spike4<DT>(p.spike, o, s);
        if (p.h_seq) st_stream4(p.h_seq + o, h);
      }
    }
  }
}

template <int T>
__global__ void __launch_bounds__(512) lif_window_bwd_kernel(const LifWinP p) {
  constexpr int TM = T > 0 ? T : 8;
  const int Tn = T > 0 ? T : p.wd;
  const NeuronP nrn = p.nrn;
  const int rx = threadIdx.x % p.R, ry = threadIdx.x / p.R;
  const int64_t col = (int64_t)blockIdx.y * p.tile_w + (int64_t)rx * 4;
  const int64_t stride = (int64_t)gridDim.x * p.k;
  const float dh_dx = neuron_dh_dx(nrn), dh_dv = neuron_dh_dv(nrn);
  for (int64_t nr = (int64_t)blockIdx.x * p.k + ry; nr < p.MP; nr += stride) {
    int64_t xr[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) xr[t] = p.win2x[t * p.MP + nr];
    float4 x[TM], g[TM], h[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) {
        x[t] = xr[t] >= 0 ? ld_stream4(p.x + xr[t] * p.C + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        g[t] = ld_stream4(p.gs + (t * p.MP + nr) * p.C + col);
      }
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = nrn.hard ? nrn.v_reset : 0.f;
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f4(h[t], i) = neuron_charge(nrn, v[i], f4(x[t], i));
          v[i] = neuron_reset(nrn, f4(h[t], i), neuron_fire(nrn, f4(h[t], i)));
        }
      }
    float gv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = TM - 1; t >= 0; --t)
      if (t < Tn) {
        float4 dx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float gh = neuron_grad_h(nrn, f4(h[t], i), f4(g[t], i), gv[i]);
          f4(dx, i) = gh * dh_dx;
          gv[i] = gh * dh_dv;
        }
        if (xr[t] >= 0) st_stream4(p.gx + xr[t] * p.C + col, dx);
      }
  }
}

// ---- 2x2 patch-merging gather (+ LIF over D) ----------------------------------------------
struct MergeP {
  const float* x; void* spike; float* h_seq; const float* gs; float* gx;
  int64_t B, D, H, W, C, H2, W2, rows;  // rows = B*H2*W2 output token columns (time handled inside)
  int64_t tile_w;
  int R, k, apply_neuron;
  NeuronP nrn;
};

// one thread: 4 channels of one (b, h2, w2, kq) for all D steps
template <int T, int DT, bool BWD>
__global__ void __launch_bounds__(512) lif_merge_kernel(const MergeP p) {
  constexpr int TM = T > 0 ? T : 32;
  const int Tn = T > 0 ? T : (int)p.D;
  const NeuronP nrn = p.nrn;
  const int rx = threadIdx.x % p.R, ry = threadIdx.x / p.R;
  const int64_t col = (int64_t)blockIdx.y * p.tile_w + (int64_t)rx * 4;  // in [0, 4C)
  const int64_t kq = col / p.C, c = col - kq * p.C;
  const int64_t dh = kq & 1, dw = kq >> 1;
  const int64_t stride = (int64_t)gridDim.x * p.k;
  const float dh_dx = neuron_dh_dx(nrn), dh_dv = neuron_dh_dv(nrn);
  for (int64_t r = (int64_t)blockIdx.x * p.k + ry; r < p.rows; r += stride) {
    const int64_t b = r / (p.H2 * p.W2), rem = r - b * (p.H2 * p.W2);
    const int64_t h2 = rem / p.W2, w2 = rem - h2 * p.W2;
    const int64_t hh = 2 * h2 + dh, ww = 2 * w2 + dw;
    const bool inb = hh < p.H && ww < p.W;
    const int64_t in0 = (((b * p.D) * p.H + hh) * p.W + ww) * p.C + c;
    const int64_t in_st = p.H * p.W * p.C;
    const int64_t out0 = (((b * p.D) * p.H2 + h2) * p.W2 + w2) * 4 * p.C + col;
    const int64_t out_st = p.H2 * p.W2 * 4 * p.C;
    float4 x[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t)
      if (t < Tn) x[t] = inb ? ld_stream4(p.x + in0 + t * in_st) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!p.apply_neuron) {
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (t < Tn) {
          if (!BWD) st_stream4(reinterpret_cast<float*>(p.spike) + out0 + t * out_st, x[t]);
          else if (inb) st_stream4(p.gx + in0 + t * in_st, ld_stream4(p.gs + out0 + t * out_st));
        }
      continue;
    }
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = nrn.hard ? nrn.v_reset : 0.f;
    if (!BWD) {
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (t < Tn) {
          float4 h, s;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f4(h, i) = neuron_charge(nrn, v[i], f4(x[t], i));
            f4(s, i) = neuron_fire(nrn, f4(h, i));
            v[i] = neuron_reset(nrn, f4(h, i), f4(s, i));
          }
          store_spike4<DT>(p.spike, out0 + t * out_st, s);
          if (p.h_seq) st_stream4(p.h_seq + out0 + t * out_st, h);
        }
    } else {
      float4 g[TM];
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (t < Tn) g[t] = ld_stream4(p.gs + out0 + t * out_st);
      // h overwrites x
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (t < Tn) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float hv = neuron_charge(nrn, v[i], f4(x[t], i));
            f4(x[t], i) = hv;
            v[i] = neuron_reset(nrn, hv, neuron_fire(nrn, hv));
          }
        }
      float gv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = TM - 1; t >= 0; --t)
        if (t < Tn) {
          float4 dx;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float gh = neuron_grad_h(nrn, f4(x[t], i), f4(g[t], i), gv[i]);
            f4(dx, i) = gh * dh_dx;
            gv[i] = gh * dh_dv;
          }
          if (inb) st_stream4(p.gx + in0 + t * in_st, dx);
        }
    }
  }
}

static int win_rows_setup(int64_t rows, int64_t C, int target, int64_t max_blocks, RowTiling* rt) {
  SDF_REQUIRE(rows > 0 && C > 0 && C % 4 == 0, "window kernel: rows=%lld C=%lld (C must be a multiple of 4)", (long long)rows, (long long)C);
  SDF_REQUIRE(make_row_tiling(rows, C, 4, target, (int)max_blocks, rt), "window kernel: cannot tile C=%lld", (long long)C);
  return SDF_OK;
}

static void geom_derive(const sdf_window_geom& g, WinIdxP* p) {
  p->B = g.B; p->D = g.D; p->H = g.H; p->W = g.W; p->wd = g.wd; p->wh = g.wh; p->ww = g.ww;
  p->sd = g.sd; p->sh = g.sh; p->sw = g.sw;
  p->nD = (g.D + g.wd - 1) / g.wd; p->nH = (g.H + g.wh - 1) / g.wh; p->nW = (g.W + g.ww - 1) / g.ww;
  p->Dp = p->nD * g.wd; p->Hp = p->nH * g.wh; p->Wp = p->nW * g.ww;
  p->P = g.wh * g.ww; p->N = g.wd * p->P; p->nWin = p->nD * p->nH * p->nW;
  p->total = g.B * p->nWin * p->N;
}

static int geom_check(const sdf_window_geom& g) {
  SDF_REQUIRE(g.B > 0 && g.D > 0 && g.H > 0 && g.W > 0, "window geometry: empty feature map");
  SDF_REQUIRE(g.wd > 0 && g.wh > 0 && g.ww > 0, "window geometry: empty window");
  SDF_REQUIRE(g.sd >= 0 && g.sd < g.wd && g.sh >= 0 && g.sh < g.wh && g.sw >= 0 && g.sw < g.ww,
              "window geometry: shift must be in [0, window)");
  return SDF_OK;
}

}  // namespace sdf

using namespace sdf;

extern "C" int64_t sdf_window_rows(const sdf_window_geom* g) {
  if (!g || geom_check(*g)) return -1;
  WinIdxP p;
  geom_derive(*g, &p);
  return p.total;
}

extern "C" int sdf_window_index(const sdf_window_index_args* a) {
  SDF_REQUIRE(a && a->win2x, "sdf_window_index: null argument");
  int st = geom_check(a->g);
  if (st) return st;
  WinIdxP p;
  geom_derive(a->g, &p);
  SDF_REQUIRE(a->g.B * a->g.D * a->g.H * a->g.W < (int64_t)INT32_MAX && p.total < (int64_t)INT32_MAX * 64,
              "sdf_window_index: token count exceeds int32 table range");
  p.win2x = a->win2x; p.region = a->region;
  const int threads = 256;
  int64_t need = (p.total + threads - 1) / threads;
  int blocks = (int)(need < kNumSMs * 8 ? need : kNumSMs * 8);
  window_index_kernel<<<blocks, threads, 0, (cudaStream_t)a->stream>>>(p);
  return finish_launch("sdf_window_index");
}

static int launch_rows(int mode, WinRowsP& p, int64_t rows, int64_t C, cudaStream_t stream, int64_t max_blocks) {
  RowTiling rt;
  int st = win_rows_setup(rows, C, 512, max_blocks, &rt);
  if (st) return st;
  p.rows = rows; p.C = C; p.tile_w = rt.tile_w; p.R = rt.R; p.k = rt.k;
  dim3 grid(rt.blocks, rt.ncol, 1);
  const size_t smem = sizeof(float) * 4 * rt.threads;
  switch (mode) {
    case 0: window_rows_kernel<0><<<grid, rt.threads, 0, stream>>>(p); break;
    case 1: window_rows_kernel<1><<<grid, rt.threads, 0, stream>>>(p); break;
    case 2: window_rows_kernel<2><<<grid, rt.threads, 0, stream>>>(p); break;
    default: window_rows_kernel<3><<<grid, rt.threads, smem, stream>>>(p); break;
  }
  return rt.blocks;
}

extern "C" int sdf_window_gather(const sdf_window_gather_args* a) {
  SDF_REQUIRE(a && a->x && a->xw && a->win2x && aligned16(a->x) && aligned16(a->xw), "sdf_window_gather: null/unaligned argument");
  WinRowsP p = {};
  p.src = a->x; p.dst = a->xw; p.win2x = a->win2x;
  int r = launch_rows(0, p, a->rows, a->C, (cudaStream_t)a->stream, kNumSMs * 4);
  if (r < 0) return r;
  return finish_launch("sdf_window_gather");
}

extern "C" int sdf_window_gather_bwd(const sdf_window_gather_bwd_args* a) {
  SDF_REQUIRE(a && a->dxw && a->dx && a->win2x && aligned16(a->dx) && aligned16(a->dxw), "sdf_window_gather_bwd: null/unaligned argument");
  WinRowsP p = {};
  p.src = a->dxw; p.dst = a->dx; p.win2x = a->win2x;
  int r = launch_rows(1, p, a->rows, a->C, (cudaStream_t)a->stream, kNumSMs * 4);
  if (r < 0) return r;
  return finish_launch("sdf_window_gather_bwd");
}

extern "C" int sdf_window_scatter(const sdf_window_scatter_args* a) {
  SDF_REQUIRE(a && a->y && a->out && a->win2x && aligned16(a->y) && aligned16(a->out) && (!a->res || aligned16(a->res)),
              "sdf_window_scatter: null/unaligned argument");
  SDF_REQUIRE((a->scale == nullptr) == (a->shift == nullptr), "sdf_window_scatter: scale and shift go together");
  SDF_REQUIRE(a->rows_per_sample > 0, "sdf_window_scatter: rows_per_sample must be > 0");
  WinRowsP p = {};
  p.src = a->y; p.dst = a->out; p.res = a->res; p.win2x = a->win2x; p.scale = a->scale; p.shift = a->shift;
  p.alpha = a->alpha; p.rows_per_sample = a->rows_per_sample;
  int r = launch_rows(2, p, a->rows, a->C, (cudaStream_t)a->stream, kNumSMs * 4);
  if (r < 0) return r;
  return finish_launch("sdf_window_scatter");
}

extern "C" int sdf_window_scatter_bwd(const sdf_window_scatter_bwd_args* a) {
  SDF_REQUIRE(a && a->dout && a->dy && a->win2x && aligned16(a->dout) && aligned16(a->dy), "sdf_window_scatter_bwd: null/unaligned argument");
  SDF_REQUIRE(a->rows_per_sample > 0, "sdf_window_scatter_bwd: rows_per_sample must be > 0");
  if (a->bn_partials) SDF_REQUIRE(a->u && a->n_partial_blocks >= 1, "sdf_window_scatter_bwd: partials need u and capacity");
  WinRowsP p = {};
  p.src = a->dout; p.dst = a->dy; p.u = a->u; p.win2x = a->win2x; p.alpha = a->alpha; p.partials = a->bn_partials;
  p.rows_per_sample = a->rows_per_sample;
  int64_t cap = kNumSMs * 2;
  if (a->bn_partials && a->n_partial_blocks < cap) cap = a->n_partial_blocks;
  cudaStream_t stream = (cudaStream_t)a->stream;
  int r = launch_rows(3, p, a->rows, a->C, stream, cap);
  if (r < 0) return r;
  if (a->bn_partials && a->n_partial_blocks > r)
    cudaMemsetAsync(a->bn_partials + (int64_t)r * 2 * a->C, 0, sizeof(float) * (a->n_partial_blocks - r) * 2 * a->C, stream);
  return finish_launch("sdf_window_scatter_bwd");
}

#define SDF_DT3(KERNEL, TT, DT, ...)                                     \
  do {                                                                   \
    if (DT == SDF_SPIKE_F32) KERNEL<TT, SDF_SPIKE_F32> __VA_ARGS__;      \
    else if (DT == SDF_SPIKE_U8) KERNEL<TT, SDF_SPIKE_U8> __VA_ARGS__;   \
    else KERNEL<TT, SDF_SPIKE_BF16> __VA_ARGS__;                         \
  } while (0)

extern "C" int sdf_lif_window_fwd(const sdf_lif_window_fwd_args* a) {
  SDF_REQUIRE(a && a->x && a->spike && a->win2x && aligned16(a->x) && aligned16(a->spike), "sdf_lif_window_fwd: null/unaligned argument");
  SDF_REQUIRE(a->wd >= 1 && a->wd <= 8, "sdf_lif_window_fwd: window depth %lld not in [1,8]", (long long)a->wd);
  int st = validate_neuron(a->neuron);
  if (st) return st;
  const int DT = a->spike_dtype;
  SDF_REQUIRE(DT >= 0 && DT <= 2, "sdf_lif_window_fwd: bad spike_dtype");
  RowTiling rt;
  st = win_rows_setup(a->MP, a->C, 512, kNumSMs * 2, &rt);
  if (st) return st;
  LifWinP p = {};
  p.x = a->x; p.spike = a->spike; p.h_seq = a->h_seq; p.win2x = a->win2x; p.MP = a->MP; p.C = a->C;
  p.tile_w = rt.tile_w; p.R = rt.R; p.k = rt.k; p.wd = (int)a->wd; p.nrn = make_neuron(a->neuron);
  dim3 grid(rt.blocks, rt.ncol, 1);
  cudaStream_t stream = (cudaStream_t)a->stream;
  switch (a->wd) {
    case 1: SDF_DT3(lif_window_fwd_kernel, 1, DT, <<<grid, rt.threads, 0, stream>>>(p)); break;
    case 2: SDF_DT3(lif_window_fwd_kernel, 2, DT, <<<grid, rt.threads, 0, stream>>>(p)); break;
    case 4: SDF_DT3(lif_window_fwd_kernel, 4, DT, <<<grid, rt.threads, 0, stream>>>(p)); break;
    default: SDF_DT3(lif_window_fwd_kernel, 0, DT, <<<grid, rt.threads, 0, stream>>>(p)); break;
  }
  return finish_launch("sdf_lif_window_fwd");
}

extern "C" int sdf_lif_window_bwd(const sdf_lif_window_bwd_args* a) {
  SDF_REQUIRE(a && a->x && a->grad_spike && a->grad_x && a->win2x && aligned16(a->x) && aligned16(a->grad_spike) && aligned16(a->grad_x),
              "sdf_lif_window_bwd: null/unaligned argument");
  SDF_REQUIRE(a->wd >= 1 && a->wd <= 8, "sdf_lif_window_bwd: window depth %lld not in [1,8]", (long long)a->wd);
  int st = validate_neuron(a->neuron);
  if (st) return st;
  RowTiling rt;
  st = win_rows_setup(a->MP, a->C, 512, kNumSMs * 2, &rt);
  if (st) return st;
  LifWinP p = {};
  p.x = a->x; p.gs = a->grad_spike; p.gx = a->grad_x; p.win2x = a->win2x; p.MP = a->MP; p.C = a->C;
  p.tile_w = rt.tile_w; p.R = rt.R; p.k = rt.k; p.wd = (int)a->wd; p.nrn = make_neuron(a->neuron);
  dim3 grid(rt.blocks, rt.ncol, 1);
  cudaStream_t stream = (cudaStream_t)a->stream;
  switch (a->wd) {
    case 1: lif_window_bwd_kernel<1><<<grid, rt.threads, 0, stream>>>(p); break;
    case 2: lif_window_bwd_kernel<2><<<grid, rt.threads, 0, stream>>>(p); break;
    case 4: lif_window_bwd_kernel<4><<<grid, rt.threads, 0, stream>>>(p); break;
    default: lif_window_bwd_kernel<0><<<grid, rt.threads, 0, stream>>>(p); break;
  }
  return finish_launch("sdf_lif_window_bwd");
}

static int merge_setup(int64_t B, int64_t D, int64_t H, int64_t W, int64_t C, MergeP* p, dim3* grid, int* threads) {
  SDF_REQUIRE(B > 0 && D > 0 && D <= 32 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "patch merging: bad shape (D <= 32, C %% 4 == 0)");
  p->B = B; p->D = D; p->H = H; p->W = W; p->C = C; p->H2 = (H + 1) / 2; p->W2 = (W + 1) / 2;
  p->rows = B * p->H2 * p->W2;
  RowTiling rt;
  SDF_REQUIRE(make_row_tiling(p->rows, 4 * C, 4, 256, kNumSMs * 4, &rt), "patch merging: cannot tile C");
  p->tile_w = rt.tile_w; p->R = rt.R; p->k = rt.k;
  *grid = dim3(rt.blocks, rt.ncol, 1);
  *threads = rt.threads;
  return SDF_OK;
}

extern "C" int sdf_lif_merge_fwd(const sdf_lif_merge_fwd_args* a) {
  SDF_REQUIRE(a && a->x && a->spike && aligned16(a->x) && aligned16(a->spike), "sdf_lif_merge_fwd: null/unaligned argument");
  MergeP p = {};
  dim3 grid; int threads;
  int st = merge_setup(a->B, a->D, a->H, a->W, a->C, &p, &grid, &threads);
  if (st) return st;
  p.x = a->x; p.spike = a->spike; p.h_seq = a->h_seq; p.apply_neuron = a->apply_neuron;
  int DT = a->spike_dtype;
  if (a->apply_neuron) {
    st = validate_neuron(a->neuron);
    if (st) return st;
    p.nrn = make_neuron(a->neuron);
    SDF_REQUIRE(DT >= 0 && DT <= 2, "sdf_lif_merge_fwd: bad spike_dtype");
  } else {
    sdf_neuron_cfg nc = {}; nc.tau = 2.0; p.nrn = make_neuron(nc);
    DT = SDF_SPIKE_F32;
  }
  cudaStream_t stream = (cudaStream_t)a->stream;
#define MERGE_FWD(TT)                                                                                        \
  do {                                                                                                       \
    if (DT == SDF_SPIKE_F32) lif_merge_kernel<TT, SDF_SPIKE_F32, false><<<grid, threads, 0, stream>>>(p);    \
    else if (DT == SDF_SPIKE_U8) lif_merge_kernel<TT, SDF_SPIKE_U8, false><<<grid, threads, 0, stream>>>(p); \
    else lif_merge_kernel<TT, SDF_SPIKE_BF16, false><<<grid, threads, 0, stream>>>(p);                       \
  } while (0)
  if (a->D == 10) MERGE_FWD(10);
  else if (a->D == 5) MERGE_FWD(5);
  else MERGE_FWD(0);
#undef MERGE_FWD
  return finish_launch("sdf_lif_merge_fwd");
}

extern "C" int sdf_lif_merge_bwd(const sdf_lif_merge_bwd_args* a) {
  SDF_REQUIRE(a && a->x && a->grad_spike && a->grad_x && aligned16(a->x) && aligned16(a->grad_spike) && aligned16(a->grad_x),
              "sdf_lif_merge_bwd: null/unaligned argument");
  MergeP p = {};
  dim3 grid; int threads;
  int st = merge_setup(a->B, a->D, a->H, a->W, a->C, &p, &grid, &threads);
  if (st) return st;
  p.x = a->x; p.gs = a->grad_spike; p.gx = a->grad_x; p.apply_neuron = a->apply_neuron;
  if (a->apply_neuron) {
    st = validate_neuron(a->neuron);
    if (st) return st;
    p.nrn = make_neuron(a->neuron);
  } else {
    sdf_neuron_cfg nc = {}; nc.tau = 2.0; p.nrn = make_neuron(nc);
  }
  cudaStream_t stream = (cudaStream_t)a->stream;
  if (a->D == 10) lif_merge_kernel<10, SDF_SPIKE_F32, true><<<grid, threads, 0, stream>>>(p);
  else if (a->D == 5) lif_merge_kernel<5, SDF_SPIKE_F32, true><<<grid, threads, 0, stream>>>(p);
  else lif_merge_kernel<0, SDF_SPIKE_F32, true><<<grid, threads, 0, stream>>>(p);
  return finish_launch("sdf_lif_merge_bwd");
}
