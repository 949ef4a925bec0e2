/*
 * sdf_b200.h — C-ABI of libsdf_b200.so: hand-written sm_100a kernels for the SDformerFlow
 * spiking spatiotemporal Swin encoder hot path.
 *
 * The reference (yitian97/SDformerFlow) has no FFI of its own: its "native" code is what
 * spikingjelly JIT-compiles through cupy when the scripts call
 *   functional.set_backend(model, "cupy", neurontype)
 *     (train_flow_parallel_supervised_SNN.py:118-119, eval_DSEC_flow_SNN.py:118-119)
 * plus the ATen ops behind models/STSwinNet_SNN/Spiking_swin_transformer3D.py.  Every entry
 * point below cites the reference interface (file:line under /root/reference) it replaces.
 *
 * Conventions
 *  - extern "C", plain structs of raw DEVICE pointers + int64 sizes + a cudaStream_t passed
 *    as void*.  No torch types.
 *  - The caller owns every buffer including workspaces; the library never allocates or
 *    frees device memory and never synchronises the stream.
 *  - Every function returns 0 on success or a negative sdf_status; sdf_last_error() returns
 *    a thread-local message for the last failure on the calling thread.
 *  - All float scalars cross the ABI as double and are narrowed to fp32 inside (the
 *    reference promotes python floats to the tensor dtype the same way).
 *  - "channels-last rows": a tensor viewed as [rows, C] with C contiguous, C % 4 == 0,
 *    base pointers 16-byte aligned.
 */
#ifndef SDF_B200_H
#define SDF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDF_VERSION_MAJOR 0
#define SDF_VERSION_MINOR 1

typedef enum {
  SDF_OK = 0,
  SDF_ERR_INVALID_ARG = -1,
  SDF_ERR_UNSUPPORTED = -2,
  SDF_ERR_CUDA = -3,
  SDF_ERR_WORKSPACE = -4
} sdf_status;

/* neuron dynamics — reference Spiking_modules.py:26-99 (Spiking_neuron switch) over
 * spikingjelly neuron.{LIFNode, IFNode, ParametricLIFNode}. */
typedef enum {
  SDF_NEURON_LIF = 0,  /* h = v + (x - (v - v_reset))/tau   (decay_input=True)            */
  SDF_NEURON_IF = 1,   /* h = v + x                                                       */
  SDF_NEURON_PLIF = 2  /* h = v + (x - (v - v_reset)) * k,  k = sigmoid(w) passed as 1/tau */
} sdf_neuron_kind;

/* spike output element type */
typedef enum {
  SDF_SPIKE_F32 = 0,  /* {0.0f, 1.0f}: the reference's dtype (parity mode, fp32 GEMM input) */
  SDF_SPIKE_U8 = 1,   /* {0, 1} bytes: inference contract, 5 B per neuron-timestep          */
  SDF_SPIKE_BF16 = 2  /* {0, 1} bf16: exact operand of the bf16x3-split spike GEMM          */
} sdf_spike_dtype;

/* surrogate gradient — spikingjelly surrogate.ATan / surrogate.Sigmoid */
typedef enum { SDF_SG_ATAN = 0, SDF_SG_SIGMOID = 1 } sdf_surrogate_kind;

typedef struct {
  int32_t kind;         /* sdf_neuron_kind */
  int32_t hard_reset;   /* 0: soft reset v = h - s*v_th (v_reset None); 1: v = (1-s)h + s*v_reset */
  int32_t detach_reset; /* 1: spike detached in the reset path (reference default True) */
  int32_t surrogate;    /* sdf_surrogate_kind */
  double v_th;
  double v_reset;       /* used when hard_reset; also the initial v */
  double tau;           /* LIF: tau; PLIF: 1/sigmoid(w) */
  double sg_alpha;      /* surrogate alpha (ATan default 2.0) */
} sdf_neuron_cfg;

/* Addressing of a multi-step neuron tensor.  Neuron n in [0, n_neurons), time t in [0, T):
 *   b = n / inner, r = n % inner, element offset = b*stride_b + t*stride_t + r.
 * Plain [T, N]:              inner = N,       stride_b = 0,        stride_t = N.
 * (B, D, H, W, C), time = D: inner = H*W*C,   stride_b = D*H*W*C,  stride_t = H*W*C
 *   (replaces the x.permute(1,0,2,3,4) copies of Spiking_swin_transformer3D.py:845,932,970). */
typedef struct {
  int64_t T;
  int64_t n_neurons;
  int64_t inner;
  int64_t stride_b;
  int64_t stride_t;
} sdf_seq_layout;

/* ---- K1 / K2: multi-step LIF/IF/PLIF forward and surrogate-gradient backward ---------------
 * Replaces spikingjelly LIFNode.multi_step_forward (torch loop / cupy LIFNodeFPTTKernel,
 * LIFNodeBPTTKernel) reached through Spiking_neuron.forward (Spiking_modules.py:98-99),
 * fused with the preceding BatchNorm apply (SpikingNormLayer, Spiking_modules.py:101-146):
 *   x_t = u_t * scale[c] + shift[c]   (c = channel, channels-last: n % C; NCHW: (n / hw) % C)
 * One thread keeps the membrane potential of 4 neurons in registers across all T steps. */
typedef struct {
  const float* u;      /* input, fp32 */
  void* spike;         /* output spikes, spike_dtype, same addressing as u */
  float* h_seq;        /* optional: membrane potential after charge, fp32 (tests / monitors) */
  const float* v_init; /* optional [n_neurons] initial membrane (NULL: 0 or v_reset) */
  float* v_final;      /* optional [n_neurons] membrane after the last step */
  const float* scale;  /* optional [C] */
  const float* shift;  /* optional [C] */
  int64_t C;           /* channels (ignored when scale == NULL) */
  int64_t hw;          /* 1: channels-last (c = n % C); >1: NCHW, c = (n / hw) % C */
  sdf_seq_layout lay;
  sdf_neuron_cfg neuron;
  int32_t spike_dtype;
  int32_t _pad;
  void* stream;
} sdf_lif_fwd_args;

int sdf_lif_fwd(const sdf_lif_fwd_args* a);

typedef struct {
  const float* u;       /* forward input (re-read; h is recomputed, nothing else was saved) */
  const void* grad_spike; /* dL/ds, fp32 */
  float* grad_u;        /* optional: dL/du, fp32 (already multiplied by scale[c] when scale != NULL) */
  float* grad_x;        /* optional: dL/dx (before the scale multiply); needed for BN-train backward.
                           At least one of grad_u / grad_x must be given. */
  const float* v_init;
  const float* scale;
  const float* shift;
  float* bn_partials;   /* optional [n_partial_blocks, 2, C]: per-block sum(dx), sum(dx*u) */
  int64_t n_partial_blocks; /* in: capacity; the launcher uses min(capacity, its grid) and zero-fills the rest */
  float* plif_partials; /* optional [n_partial_blocks]: PLIF d(1/tau) partial sums */
  const float* bn_coef; /* optional [3, C] from sdf_bn_bwd_finalize: the BatchNorm backward is applied in this pass,
                           grad_u = a[c]*dx + b[c]*u + c0[c].  Two-phase use: call once with grad_u = grad_x = NULL and
                           bn_partials (statistics only, nothing written), finalize, call again with bn_coef + grad_u:
                           20 B per neuron-timestep instead of 24 B for K2 + a separate sdf_bn_bwd_apply. */
  int64_t C;
  int64_t hw;
  sdf_seq_layout lay;
  sdf_neuron_cfg neuron;
  void* stream;
} sdf_lif_bwd_args;

int sdf_lif_bwd(const sdf_lif_bwd_args* a);

/* number of partial blocks sdf_lif_bwd / sdf_bn_stats / ... will write for a [rows, C] problem */
int64_t sdf_partial_blocks(int64_t rows, int64_t C);

/* ---- K1p: PSN (parallel spiking neuron) — reference Spiking_submodules.py:183-211 ----------
 * h[t] = sum_k W[t,k] * x[k] + b[t]; s = (h >= 0).  T <= 32. */
typedef struct {
  const float* u;
  void* spike;
  float* h_seq;        /* optional */
  const float* weight; /* [T, T] */
  const float* bias;   /* [T] */
  const float* scale;
  const float* shift;
  int64_t C;
  int64_t hw;
  sdf_seq_layout lay;
  int32_t spike_dtype;
  int32_t _pad;
  void* stream;
} sdf_psn_fwd_args;

int sdf_psn_fwd(const sdf_psn_fwd_args* a);

typedef struct {
  const float* u;
  const float* grad_spike;
  float* grad_u;       /* optional: dL/du = dL/dx * scale[c] */
  float* grad_x;       /* optional: dL/dx (BN-train backward needs it before the scale multiply) */
  float* grad_h;       /* [T, n_neurons] contiguous: dL/dh, for sdf_psn_wgrad (optional when wgrad_partials is given) */
  float* x_out;        /* optional [T, n_neurons] contiguous: the post-affine input x (for dW) */
  const float* weight;
  const float* bias;
  const float* scale;
  const float* shift;
  float* bn_partials;  /* optional [n_partial_blocks, 2, C]: per-block sum(dx), sum(dx*u) */
  int64_t n_partial_blocks;
  int64_t C;
  int64_t hw;
  sdf_seq_layout lay;
  int32_t surrogate;
  int32_t _pad;
  double sg_alpha;
  void* stream;
  float* wgrad_partials;    /* optional [n_wgrad_blocks, T*T + T]: per-block partial sums of dW[t][k] = sum_n dh[t,n] x[k,n] and
                               db[t] = sum_n dh[t,n], accumulated in this pass (T in {2,4,5,10}, 16-byte aligned layout); with
                               it grad_h / x_out may be NULL: dh and x never go to HBM (12 B instead of 28 B per neuron-step) */
  int64_t n_wgrad_blocks;
} sdf_psn_bwd_args;

int sdf_psn_bwd(const sdf_psn_bwd_args* a);

/* PSN parameter gradients from the outputs of sdf_psn_bwd: dW[t,k] = sum_n grad_h[t,n] * x[k,n], db[t] = sum_n grad_h[t,n]
 * (autograd of the addmm in Spiking_submodules.py:207-211).  partials: [n_partial_blocks, T*T + T] per-block sums (the
 * caller adds them up; unused rows are zero-filled).  T in {2, 4, 5, 10}. */
typedef struct {
  const float* grad_h;   /* [T, n_neurons] */
  const float* x;        /* [T, n_neurons] */
  float* partials;
  int64_t n_partial_blocks;
  int64_t T, n_neurons;
  void* stream;
} sdf_psn_wgrad_args;

int sdf_psn_wgrad(const sdf_psn_wgrad_args* a);

/* ---- K6: BatchNorm statistics over channels-last rows ---------------------------------------
 * Replaces the statistics half of sj_layer.BatchNorm2d on permuted views
 * (Spiking_swin_transformer3D.py:153,159,310,314,318,367,673,677,714,933,972). */
typedef struct {
  const float* x;      /* [rows, ld] fp32, channel c of row r at x[r*ld + c] */
  int64_t rows;
  int64_t C;
  int64_t ld;          /* row stride in elements (>= C, % 4 == 0) */
  float* partials;     /* [n_partial_blocks, 2, C] workspace */
  int64_t n_partial_blocks;
  void* stream;
} sdf_bn_stats_args;

int sdf_bn_stats(const sdf_bn_stats_args* a);

/* Finalise statistics: mean/var from partials, scale = w*rstd, shift = b - mean*scale,
 * running-stat update exactly like torch (momentum, unbiased running_var). */
typedef struct {
  const float* partials; /* [n_partial_blocks, 2, C]; NULL in eval mode */
  int64_t n_partial_blocks;
  int64_t count;         /* rows that contributed */
  int64_t C;
  const float* weight;   /* [C] or NULL (=1) */
  const float* bias;     /* [C] or NULL (=0) */
  float* running_mean;   /* [C] or NULL */
  float* running_var;    /* [C] or NULL */
  double momentum;
  double eps;
  int32_t training;      /* 1: batch statistics (+ running update); 0: running statistics */
  int32_t _pad;
  float* scale;          /* out [C] */
  float* shift;          /* out [C] */
  float* mean;           /* out [C] (saved for backward), optional */
  float* rstd;           /* out [C] (saved for backward), optional */
  void* stream;
} sdf_bn_finalize_args;

int sdf_bn_finalize(const sdf_bn_finalize_args* a);

/* BN backward, second half.  Given dy = dL/d(BN output) and the forward input u:
 *   train: du = w*rstd * (dy - sum(dy)/n - xhat * sum(dy*xhat)/n),  xhat = (u - mean)*rstd
 *   eval : du = w*rstd * dy
 * plus dweight = sum(dy*xhat), dbias = sum(dy) from the partial sums (sum(dy), sum(dy*u)). */
typedef struct {
  const float* partials; /* [n_partial_blocks, 2, C]: sum(dy), sum(dy*u) per block */
  int64_t n_partial_blocks;
  int64_t count;
  int64_t C;
  const float* weight;
  const float* mean;
  const float* rstd;
  float* grad_weight;    /* out [C] */
  float* grad_bias;      /* out [C] */
  float* coef;           /* out [3, C]: a, b, c with du = a*dy + b*u + c */
  int32_t training;
  int32_t _pad;
  void* stream;
} sdf_bn_bwd_finalize_args;

int sdf_bn_bwd_finalize(const sdf_bn_bwd_finalize_args* a);

/* elementwise over channels-last rows: out = a[c]*dy + b[c]*u + c[c]  (BN-train backward apply) */
typedef struct {
  const float* dy;
  const float* u;
  int64_t ld_u;
  float* du;
  int64_t ld_du;
  const float* coef;     /* [3, C] */
  int64_t rows;
  int64_t C;
  void* stream;
} sdf_bn_bwd_apply_args;

int sdf_bn_bwd_apply(const sdf_bn_bwd_apply_args* a);

/* partial sums (sum(dy), sum(dy*u)) over channels-last rows, for BN sites not preceded by a neuron */
typedef struct {
  const float* dy;
  const float* u;
  int64_t ld_u;
  int64_t rows;
  int64_t C;
  float* partials;
  int64_t n_partial_blocks;
  void* stream;
} sdf_bn_bwd_reduce_args;

int sdf_bn_bwd_reduce(const sdf_bn_bwd_reduce_args* a);

/* out = res + alpha * (u*scale[c] + shift[c]) over channels-last rows
 * (MLP tail: Spiking_swin_transformer3D.py:177-178 + :845 residual add) */
typedef struct {
  const float* u;
  int64_t ld_u;
  const float* res;      /* optional */
  float* out;
  const float* scale;    /* optional */
  const float* shift;
  int64_t rows;
  int64_t C;
  void* stream;
} sdf_bn_apply_args;

int sdf_bn_apply(const sdf_bn_apply_args* a);

/* ---- window index algebra ---------------------------------
 * Replaces F.pad + torch.roll + window_partition_v2 + window_reverse + roll + crop
 * (Spiking_swin_transformer3D.py:781-821, :100-113; swin_transformer3D_v2.py:52-81) and the
 * region ids behind compute_mask (:980-993).  Window-buffer row rho = (w*wd + dd)*P + pos. */
typedef struct {
  int64_t B, D, H, W;    /* unpadded feature map (tokens) */
  int64_t wd, wh, ww;    /* window, already clamped by get_window_size */
  int64_t sd, sh, sw;    /* shift, already clamped (0 on unshifted blocks) */
} sdf_window_geom;

typedef struct {
  sdf_window_geom g;
  int32_t* win2x;        /* out [B*nW*N]: token row in the (B,D,H,W) map or -1 for padding */
  uint8_t* region;       /* out [nW*N] (first sample): compute_mask region id 0..26; optional */
  void* stream;
} sdf_window_index_args;

int sdf_window_index(const sdf_window_index_args* a);
/* rows of the window buffer for a geometry: B * ceil(D/wd)*wd * ceil(H/wh)*wh * ceil(W/ww)*ww */
int64_t sdf_window_rows(const sdf_window_geom* g);

/* plain gather: xw[rho, :] = x[win2x[rho], :] or 0 (SEW attention input, :804) */
typedef struct {
  const float* x;
  float* xw;
  const int32_t* win2x;
  int64_t rows;          /* window rows */
  int64_t C;
  void* stream;
} sdf_window_gather_args;

int sdf_window_gather(const sdf_window_gather_args* a);

/* out[win2x[rho], :] = res[win2x[rho], :] + alpha[b] * (y[rho,:]*scale + shift)   (skip pads)
 * = proj_bn + window_reverse + roll back + crop + DropPath scale + shortcut add (:810-820,:840) */
typedef struct {
  const float* y;        /* [rows, C] window-ordered */
  const float* res;      /* optional (B,D,H,W,C) shortcut */
  float* out;            /* (B,D,H,W,C) */
  const int32_t* win2x;
  const float* scale;    /* optional [C] */
  const float* shift;
  const float* alpha;    /* optional [B] DropPath keep/scale factors */
  int64_t rows;
  int64_t rows_per_sample; /* window rows per batch element */
  int64_t C;
  void* stream;
} sdf_window_scatter_args;

int sdf_window_scatter(const sdf_window_scatter_args* a);

/* backward of the scatter: dy[rho,:] = alpha[b] * dout[win2x[rho], :] (0 at pads), optional
 * BN partial sums (sum(dy), sum(dy*u)) for the proj_bn backward. */
typedef struct {
  const float* dout;
  float* dy;
  const float* u;        /* optional forward y (pre-BN) for the partial sums */
  const int32_t* win2x;
  const float* alpha;
  float* bn_partials;    /* optional */
  int64_t n_partial_blocks;
  int64_t rows;
  int64_t rows_per_sample;
  int64_t C;
  void* stream;
} sdf_window_scatter_bwd_args;

int sdf_window_scatter_bwd(const sdf_window_scatter_bwd_args* a);

/* gather backward: dx[win2x[rho], :] = dxw[rho, :] */
typedef struct {
  const float* dxw;
  float* dx;             /* (B,D,H,W,C), fully overwritten (every token has exactly one row) */
  const int32_t* win2x;
  int64_t rows;
  int64_t C;
  void* stream;
} sdf_window_gather_bwd_args;

int sdf_window_gather_bwd(const sdf_window_gather_bwd_args* a);

/* ---- LIF over the window "fake time" axis with the gather folded in -------------------------
 * proj_sn(x_windows) of Spiking_QK_WindowAttention3D.forward (:670) / SDSA (:425): neuron
 * time step t = slice // M (Appendix B.1), input read straight from the (B,D,H,W,C) map. */
typedef struct {
  const float* x;        /* (B,D,H,W,C) */
  void* spike;           /* [wd*M*P, C] window-ordered */
  float* h_seq;          /* optional */
  const int32_t* win2x;
  int64_t wd;            /* fake T */
  int64_t MP;            /* M * P = rows / wd */
  int64_t C;
  sdf_neuron_cfg neuron;
  int32_t spike_dtype;
  int32_t _pad;
  void* stream;
} sdf_lif_window_fwd_args;

int sdf_lif_window_fwd(const sdf_lif_window_fwd_args* a);

typedef struct {
  const float* x;
  const float* grad_spike; /* [wd*M*P, C] */
  float* grad_x;           /* (B,D,H,W,C) fully overwritten */
  const int32_t* win2x;
  int64_t wd;
  int64_t MP;
  int64_t C;
  sdf_neuron_cfg neuron;
  void* stream;
} sdf_lif_window_bwd_args;

int sdf_lif_window_bwd(const sdf_lif_window_bwd_args* a);

/* ---- LIF with the 2x2 patch-merging gather folded in (MS_SpikingPatchMerging, :952-974) ----
 * xm[b,d,h2,w2, k*C + c] = x[b,d,2*h2 + (k&1), 2*w2 + (k>>1), c] (zero beyond H/W), time = D. */
typedef struct {
  const float* x;        /* (B,D,H,W,C) */
  void* spike;           /* (B,D,H2,W2,4C) */
  float* h_seq;          /* optional */
  int64_t B, D, H, W, C;
  sdf_neuron_cfg neuron;
  int32_t spike_dtype;
  int32_t apply_neuron;  /* 0: plain gather (SEW SpikingPatchMerging, :926-930), fp32 out */
  void* stream;
} sdf_lif_merge_fwd_args;

int sdf_lif_merge_fwd(const sdf_lif_merge_fwd_args* a);

typedef struct {
  const float* x;
  const float* grad_spike; /* (B,D,H2,W2,4C) */
  float* grad_x;           /* (B,D,H,W,C) */
  int64_t B, D, H, W, C;
  sdf_neuron_cfg neuron;
  int32_t apply_neuron;
  int32_t _pad;
  void* stream;
} sdf_lif_merge_bwd_args;

int sdf_lif_merge_bwd(const sdf_lif_merge_bwd_args* a);

/* ---- K5: QK token-gate attention core (Spiking_QK_WindowAttention3D.forward :671-710) ------
 * q = LIF(bn_q(q_pre)); a = LIF2(sum_{d<32} q); k = LIF(bn_k(k_pre) + pos); g = k * a;
 * output written in the proj-input order of :709-710 (Appendix B.3).  All neurons run over
 * the fake time axis wd.  Mask is ignored by the reference (:700-703). */
typedef struct {
  const float* q_pre;    /* [wd*M*P, ld] */
  const float* k_pre;    /* [wd*M*P, ld] */
  int64_t ld;            /* row stride of q_pre / k_pre (C, or 2C when they share one GEMM output) */
  const float* q_scale;  /* [C] bn_q folded */
  const float* q_shift;
  const float* k_scale;  /* [C] bn_k folded */
  const float* k_shift;
  const float* pos;      /* [wd*P*C]: positional_encoding (1,nH,N,32) flat-reshaped to (wd,1,wh,ww,C) (:678) */
  void* gate;            /* out [wd*M*P, C] spikes g, permuted (B.3) */
  float* q_h;            /* optional debug: membrane of sn_q  [wd*M*P, C] */
  float* k_h;            /* optional debug: membrane of sn_k */
  float* a_h;            /* optional debug: membrane of sn2_q [wd*M*P*nH] */
  int64_t wd, M, P, C, nH;
  sdf_neuron_cfg neuron;
  int32_t spike_dtype;
  int32_t _pad;
  void* stream;
} sdf_attn_qkgate_fwd_args;

int sdf_attn_qkgate_fwd(const sdf_attn_qkgate_fwd_args* a);

typedef struct {
  const float* q_pre;
  const float* k_pre;
  int64_t ld;
  const float* q_scale;
  const float* q_shift;
  const float* k_scale;
  const float* k_shift;
  const float* pos;
  const float* grad_gate; /* [wd*M*P, C] in the permuted (proj-input) order */
  float* grad_q;          /* dL/d(bn_q output) [wd*M*P, C] */
  float* grad_k;          /* dL/d(bn_k output + pos) [wd*M*P, C] */
  float* bn_partials_q;   /* optional [n_partial_blocks, 2, C] sum(dq), sum(dq*q_pre) */
  float* bn_partials_k;
  int64_t n_partial_blocks;
  int64_t wd, M, P, C, nH;
  sdf_neuron_cfg neuron;
  void* stream;
} sdf_attn_qkgate_bwd_args;

int sdf_attn_qkgate_bwd(const sdf_attn_qkgate_bwd_args* a);

/* grad of positional_encoding: dpos[t*P*C + j] = sum_m grad_k[(t*M + m)*P*C + j] */
typedef struct {
  const float* grad_k;
  float* grad_pos;
  int64_t wd, M, PC;
  void* stream;
} sdf_pos_grad_args;

int sdf_pos_grad(const sdf_pos_grad_args* a);

/* ---- K3 / K4: QK^T V spiking window attention on tcgen05 tensor cores -----------------------
 * Replaces Spiking_BN_WindowAttention3D.forward :320-363 and SDSA_WindowAttention3D.forward
 * :438-485:  O = (scale * Q K^T + Bias[h'] + Mask[w]) @ V on the raw [M*nH, N, 32]
 * reinterpretation (Appendix B.4).  Q, K, V are {0,1} bytes; S = Q K^T are exact integer
 * counts (kind::i8 MMA, int32 accumulate in TMEM).  hd must be 32. */
typedef struct {
  const uint8_t* q;      /* [M*nH*N, 32] spikes as bytes (flat view of the (wd,M,wh,ww,C) buffer) */
  const uint8_t* k;
  const uint8_t* v;
  const float* bias_table; /* [(2wd-1)(2wh-1)(2ww-1), nH] relative_position_bias_table (:246-248) */
  const uint8_t* region;   /* [nW*N] region ids, NULL on unshifted blocks (mask None, :801) */
  float* out;            /* [wd*M*P, C] in proj-input order (:362-363) */
  int32_t* s_dbg;        /* optional [M*nH, N, N] integer Q K^T counts (bit-exact parity tests) */
  float* attn_dbg;       /* optional [M*nH, N, N] fp32 attn (return_attention path) */
  int64_t M, nH, nW;
  int64_t wd, wh, ww;
  double scale;
  void* stream;
} sdf_attn_qktv_fwd_args;

int sdf_attn_qktv_fwd(const sdf_attn_qktv_fwd_args* a);

typedef struct {
  const uint8_t* q;
  const uint8_t* k;
  const uint8_t* v;
  const float* bias_table;
  const uint8_t* region;
  const float* grad_out; /* [wd*M*P, C] proj-input order */
  float* grad_q;         /* [M*nH*N, 32] fp32 */
  float* grad_k;
  float* grad_v;
  float* grad_bias_table; /* [(2wd-1)(2wh-1)(2ww-1), nH] accumulated with atomics; caller zero-fills */
  int64_t M, nH, nW;
  int64_t wd, wh, ww;
  double scale;
  void* stream;
} sdf_attn_qktv_bwd_args;

int sdf_attn_qktv_bwd(const sdf_attn_qktv_bwd_args* a);

/* ---- TF32 x 2 weight split (fp32-faithful spike GEMM / conv on tensor cores) -------
 * hi = rn_tf32(w), lo = rn_tf32(w - hi): both exactly representable in TF32, |w - hi - lo| <= 2^-22 |w|. */
typedef struct {
  const float* w;
  float* hi;
  float* lo;
  int64_t n;
  void* stream;
} sdf_split_tf32_args;

int sdf_split_tf32(const sdf_split_tf32_args* a);

/* ---- direct 3x3 conv for a few input channels (patch-embed head) -------------------------------
 * y[n,h,w,:] = sum_{ky,kx,c} x[n,h+ky-1,w+kx-1,c] * w[:,c,ky,kx] (+ bias), stride 1, zero pad 1, channels-last.
 * Replaces the head conv of MS_PED_Spiking_PatchEmbed_Conv_sfn (Spiking_modules.py:1737-1745, :270-277). */
typedef struct {
  const float* x;      /* (N, H, W, Cin) */
  const float* w;      /* (Cout, Cin, 3, 3) torch layout */
  const float* bias;   /* optional [Cout] */
  float* y;            /* (N, H, W, Cout) */
  int64_t N, H, W, Cin, Cout;
  void* stream;
} sdf_conv3x3_cl_args;

int sdf_conv3x3_cl_fwd(const sdf_conv3x3_cl_args* a);

/* weight and bias gradient of the same convolution (autograd of that nn.Conv2d: the voxel input needs no gradient):
 * dw (Cout, Cin, 3, 3) = sum over pixels of g (N, H, W, Cout) x the 3x3 input patch, db [Cout] = sum of g (optional).
 * Per-block partial sums in `workspace` (>= sdf_conv3x3_cl_wgrad_workspace_bytes), reduced in block order: deterministic. */
typedef struct {
  const float* x;      /* (N, H, W, Cin) */
  const float* g;      /* (N, H, W, Cout) */
  float* dw;           /* (Cout, Cin, 3, 3), overwritten */
  float* db;           /* [Cout] or NULL */
  float* workspace;
  int64_t workspace_bytes;
  int64_t N, H, W, Cin, Cout;
  void* stream;
} sdf_conv3x3_cl_wgrad_args;

int sdf_conv3x3_cl_wgrad(const sdf_conv3x3_cl_wgrad_args* a);
int64_t sdf_conv3x3_cl_wgrad_workspace_bytes(int64_t Cin, int64_t Cout);

/* ---- G1/G2: spike GEMM / implicit-GEMM convolution on tcgen05 + TMA (csrc/spike_gemm.cu) -----------
 * Replaces the cuBLAS / cuDNN calls behind sj_layer.Linear / sj_layer.Conv2d on spike operands:
 * Spiking_swin_transformer3D.py:126-131 (fc1/fc2), :267-290 and :632-652 (linear_q/k/v, proj), :909 (reduction);
 * Spiking_modules.py:268,318,803,845-846 (3x3 convolutions).
 *
 * Weights are pre-quantised once per optimizer step into three signed 8-bit digit planes of a 23-bit fixed-point value per
 * output channel (power-of-two scale in wscale):  w ~= wscale[co] * (hi*65536 + mid*256 + lo).  The forward is then an exact
 * integer contraction (tcgen05.mma kind::i8, u8 spikes x s8 digits -> s32) recombined and rounded ONCE to fp32, so the
 * result does not depend on tiling or summation order.  |w - w_q| <= 2^-23 * 2^ceil(log2 max_k|w[co,k]|).
 * Element (co, ci, tap) of the fp32 weight is read at w[co*s_co + ci*s_ci + tap_map[t]*s_tap]:
 *   Linear (Cout, K):            taps 1, s_co = K, s_ci = 1
 *   Conv2d (Cout, Cin, kh, kw):  taps kh*kw, s_co = Cin*taps, s_ci = taps, s_tap = 1, tap_map[t] = t */
typedef struct {
  const float* w;
  int8_t* wq;          /* out: sdf_spike_gemm_wq_bytes(Cout, Cin, taps) bytes, 16-byte aligned */
  float* wscale;       /* out: [Cout] */
  float* wt;           /* optional out: fp32 transposed copy [Cin][taps*Cout], wt[ci][tap*Cout + co] = w(co, ci, tap): the B
                          operand of sdf_gemm_tf32 / sdf_conv_dgrad_tf32 in the backward pass */
  int64_t wq_bytes;    /* capacity of wq */
  int64_t Cout, Cin, taps;
  int64_t s_co, s_ci, s_tap;
  int64_t tap_map[9];
  void* stream;
} sdf_spike_gemm_pack_args;

int sdf_spike_gemm_pack(const sdf_spike_gemm_pack_args* a);
int64_t sdf_spike_gemm_wq_bytes(int64_t Cout, int64_t Cin, int64_t taps);
int64_t sdf_spike_gemm_nt(int64_t Cout);   /* output channels per N tile (layout parameter of wq) */

/* out[r, :] = a[r, :] @ W^T + bias, a = u8 spikes (or integers <= 255) [rows, K], out fp32 [rows, ld_out].
 * bn_partials (optional): per-CTA-group sum(out), sum(out^2) per channel, [n_partial_blocks, 2, Cout]; the library writes
 * min(capacity, its groups) rows and zero-fills the rest (same convention as sdf_bn_stats). */
typedef struct {
  const uint8_t* a;
  const int8_t* wq;
  const float* wscale;
  const float* bias;        /* optional [Cout] */
  float* out;
  float* bn_partials;       /* optional */
  int64_t n_partial_blocks;
  int64_t rows, K, Cout, ld_out;
  int64_t a_max;            /* largest operand value: 1 for spikes (enables the conversion-free epilogue when K*a_max < 32768);
                               0 = unknown (any u8) */
  void* stream;
} sdf_spike_gemm_fwd_args;

int sdf_spike_gemm_fwd(const sdf_spike_gemm_fwd_args* a);

/* NHWC convolution of u8 spikes x (Nimg, H, W, Cin) -> fp32 out (Nimg, Ho, Wo, Cout), kernel kh x kw, stride 1 or 2,
 * zero padding `pad`; Ho = (H + 2*pad - kh)/stride + 1.  wq from sdf_spike_gemm_pack with taps = kh*kw.  Cin % 16 == 0. */
typedef struct {
  const uint8_t* x;
  const int8_t* wq;
  const float* wscale;
  const float* bias;
  float* out;
  float* bn_partials;
  int64_t n_partial_blocks;
  int64_t Nimg, H, W, Cin, Cout, Ho, Wo;
  int64_t kh, kw, stride, pad;
  int64_t a_max;            /* as in sdf_spike_gemm_fwd_args */
  void* stream;
} sdf_spike_conv_fwd_args;

int sdf_spike_conv_fwd(const sdf_spike_conv_fwd_args* a);

/* ConvTranspose2d(kernel 3, stride 2, padding 1, output_padding 1) on 1-byte spikes — the x2 up-sampling of the decoder
 * (SpikingTransposeDecoderLayer / MS_SpikingTransposeDecoderLayer, Spiking_modules.py:398-474): x u8 NHWC (Nimg, H, W, Cin) ->
 * out fp32 NHWC (Nimg, 2H, 2W, Cout).  Computed as four stride-1 implicit GEMMs, one per output parity class
 * cls = 2*(row parity) + (column parity), with the 1 / 2 / 2 / 4 kernel taps that land on that class; wq[cls] / wscale[cls] =
 * the planes of sdf_spike_gemm_pack run with taps = sdf_spike_deconv_class_taps(cls, tap_map, dh, dw) and the IOHW strides
 * (s_ci = Cout*9, s_co = 9, s_tap = 1).  bn_partials (optional): [4 * n_partial_blocks, 2, Cout], one slab per class. */
typedef struct {
  const uint8_t* x;
  const int8_t* wq[4];
  const float* wscale[4];
  const float* bias;        /* optional [Cout] */
  float* out;
  float* bn_partials;
  int64_t n_partial_blocks;
  int64_t Nimg, H, W, Cin, Cout;
  int64_t a_max;
  void* stream;
} sdf_spike_deconv_fwd_args;

int sdf_spike_deconv_fwd(const sdf_spike_deconv_fwd_args* a);
/* taps of parity class cls (0..3): fills src_tap (kh*3 + kw of the 3x3 kernel), dh, dw (input offsets); returns their number */
int64_t sdf_spike_deconv_class_taps(int64_t cls, int64_t* src_tap, int64_t* dh, int64_t* dw);

/* out[rows, N] = a[rows, K] @ b[N, K]^T (+ bias[N]) with fp32 operands read as TF32 (tcgen05.mma kind::tf32), fp32
 * accumulate: the data-gradient GEMM dS = G @ W of every Linear above (b = W^T stored [Cin, Cout]).  lda / ldb / ld_out in
 * elements, multiples of 4. */
typedef struct {
  const float* a;
  const float* b;
  const float* bias;   /* optional [N] */
  float* out;
  int64_t rows, K, N, lda, ldb, ld_out;
  void* stream;
} sdf_gemm_tf32_args;

int sdf_gemm_tf32(const sdf_gemm_tf32_args* a);

/* Data gradient of a stride-1 NHWC convolution on tcgen05 (TF32): g fp32 (Nimg, Ho, Wo, Cout) -> out fp32 (Nimg, H, W, Cin).
 * wd = the weight re-laid as [Cin][kh*kw*Cout] with wd[ci][(kh*KW+kw)*Cout + co] = W[co][ci][kh][kw].  Cout % 32 == 0.
 * Replaces cuDNN's dgrad behind the autograd of sj_layer.Conv2d (Spiking_modules.py:845-846, the 3x3 res-block convs). */
typedef struct {
  const float* g;
  const float* wd;
  float* out;
  int64_t Nimg, H, W, Cin, Cout, Ho, Wo;
  int64_t kh, kw, pad;
  void* stream;
} sdf_conv_dgrad_tf32_args;

int sdf_conv_dgrad_tf32(const sdf_conv_dgrad_tf32_args* a);

/* Data gradient of a 3x3 / stride-2 / padding-1 NHWC convolution (the strided convolutions of the patch embedding, reference
 * Spiking_modules.py:1710-1790; autograd of sj_layer.Conv2d(stride=2)): the input pixels of each parity class are a stride-1
 * convolution of g with the taps landing on that parity, i.e. four launches of the TF32 implicit GEMM writing their quarter
 * of `out` through strided TMA tensor maps.  wd as for sdf_conv_dgrad_tf32 ([Cin][9*Cout], PackedWeight.wt); Cout % 32 == 0. */
typedef struct {
  const float* g;      /* (Nimg, Ho, Wo, Cout) */
  const float* wd;     /* [Cin][9 * Cout]: wd[ci][tap*Cout + co] = W[co, ci, kh, kw], tap = kh*3 + kw */
  float* out;          /* (Nimg, H, W, Cin) */
  int64_t Nimg, H, W, Cin, Cout, Ho, Wo;
  void* stream;
} sdf_conv_dgrad_s2_tf32_args;

int sdf_conv_dgrad_s2_tf32(const sdf_conv_dgrad_s2_tf32_args* a);

/* Data gradient of ConvTranspose2d(k 3, stride 2, padding 1, output_padding 1) (decoder up-sampling, reference
 * Spiking_modules.py:398-474): a stride-2 / padding-1 convolution of g, one launch of the TF32 implicit GEMM with the tapped
 * operand read through an element-stride-2 TMA tensor map.
 *   wd [Cin_w][9 * Cpad], Cpad = Cout rounded up to a multiple of 32: wd[ci][tap*Cpad + co] = W[ci, co, kh, kw], zero padded.
 *   out channels Cin_w .. Cin-1 (operand channels the caller appended as padding) are written as zeros. */
typedef struct {
  const float* g;      /* (Nimg, 2H, 2W, Cout) */
  const float* wd;
  float* out;          /* (Nimg, H, W, Cin) */
  int64_t Nimg, H, W, Cin, Cin_w, Cout;
  void* stream;
} sdf_deconv_dgrad_tf32_args;

int sdf_deconv_dgrad_tf32(const sdf_deconv_dgrad_tf32_args* a);

/* ---- G3: weight gradient dW = G^T S of a Linear / convolution on a spike operand (csrc/spike_wgrad.cu) -------------
 * G fp32 [rows, Cout] (split into bf16 hi + lo, A operand through tensor memory), S u8 spikes [rows, K] (expanded to bf16,
 * MN-major B operand); contraction over the rows on tcgen05, split over row slabs into `workspace`, reduced in slab order
 * (deterministic).  accumulate != 0: dw += result.  db (optional): the bias gradient sum_rows G[:, co] from the same pass over
 * G (replaces the reference's separate g.sum(0) reduction in the autograd of F.linear / conv2d with bias).
 * workspace_bytes >= sdf_spike_wgrad_workspace_bytes(rows (conv: Nimg*Ho*Wo + partial-patch slack), Cout, Cin, taps). */
typedef struct {
  const float* g;
  const uint8_t* s;
  float* dw;               /* [Cout, K] */
  float* workspace;
  int64_t workspace_bytes;
  int64_t rows, Cout, K, ldg;
  int32_t accumulate;
  int32_t s_max;       /* largest value in s: 1 = binary spikes (cheaper expansion), 0 = any u8 */
  void* stream;
  float* db;           /* [Cout] or NULL */
} sdf_spike_wgrad_args;

int sdf_spike_wgrad(const sdf_spike_wgrad_args* a);
int64_t sdf_spike_wgrad_workspace_bytes(int64_t rows_or_pixels, int64_t Cout, int64_t Cin, int64_t taps);

/* convolution: g fp32 NHWC (Nimg, Ho, Wo, Cout), x u8 NHWC (Nimg, H, W, Cin), dw OIHW (Cout, Cin, kh, kw) */
typedef struct {
  const float* g;
  const uint8_t* x;
  float* dw;
  float* workspace;
  int64_t workspace_bytes;
  int64_t Nimg, H, W, Cin, Cout, Ho, Wo;
  int64_t kh, kw, stride, pad;
  int32_t accumulate;
  int32_t s_max;       /* largest value in x: 1 = binary spikes, 0 = any u8 */
  void* stream;
  float* db;           /* [Cout] or NULL */
} sdf_spike_conv_wgrad_args;

int sdf_spike_conv_wgrad(const sdf_spike_conv_wgrad_args* a);

/* Weight (+ bias) gradient of ConvTranspose2d(k 3, stride 2, padding 1, output_padding 1) on 1-byte spikes (reference
 * Spiking_modules.py:398-474, autograd of the decoder's sj_layer.ConvTranspose2d):
 *   dW[ci, co, kh, kw] = sum S[n, i, j, ci] * G[n, 2i - 1 + kh, 2j - 1 + kw, co]
 * as four launches of the G3 kernel, one per output parity class: G of a class is a strided TMA view on the input grid, the
 * class's taps shift the spike operand by 0 / 1 pixels.  dw is the parameter layout (Cin_w, Cout, 3, 3); x may carry
 * Cin >= Cin_w channels (zero channels appended by the caller), whose gradients are dropped.  db (optional) = sum of g.
 * workspace_bytes >= sdf_spike_deconv_wgrad_workspace_bytes(Nimg, H, W, Cout, Cin). */
typedef struct {
  const float* g;      /* (Nimg, 2H, 2W, Cout) */
  const uint8_t* x;    /* (Nimg, H, W, Cin) */
  float* dw;           /* (Cin_w, Cout, 3, 3) */
  float* workspace;
  int64_t workspace_bytes;
  int64_t Nimg, H, W, Cin, Cin_w, Cout;
  int32_t s_max;       /* 1 = binary spikes, 0 = any u8 */
  int32_t reserved;
  void* stream;
  float* db;           /* [Cout] or NULL */
} sdf_spike_deconv_wgrad_args;

int sdf_spike_deconv_wgrad(const sdf_spike_deconv_wgrad_args* a);
int64_t sdf_spike_deconv_wgrad_workspace_bytes(int64_t Nimg, int64_t H, int64_t W, int64_t Cout, int64_t Cin);

/* ---- input pipeline: polarity split + min-max normalisation + bins->steps regroup, channels-last out ----------------
 * Replaces train_flow_parallel_supervised_SNN.py:261-265,278-284 (eval_DSEC_flow_SNN.py:196-212) and the regroup loop of
 * MS_PED_Spiking_PatchEmbed_Conv_sfn.forward (Spiking_modules.py:1772-1786).
 *   split = 1: x is the signed voxel grid (B, bins, H, W); pos = relu(x), neg = relu(-x)
 *   split = 0: x is already (B, bins, 2, H, W) (what the reference scripts hand to the model); normalize must be 0
 *   normalize = 1: non-zero entries -> (v - min) / (max - min), min / max over the non-zero entries (no-op when min == max)
 *   out (B, steps, H, W, 2*bins/steps): out[b,t,h,w,2g+pol] = chunk[b, g*steps + t, pol, h, w]
 * workspace: 888 floats (normalize only); minmax (optional): the two reduced values, for callers that log them. */
typedef struct {
  const float* x;
  float* out;
  float* workspace;
  float* minmax;
  int64_t B, bins, H, W, steps;
  int32_t split;
  int32_t normalize;
  void* stream;
} sdf_voxel_prepare_args;

int sdf_voxel_prepare(const sdf_voxel_prepare_args* a);

/* ---- misc ---------------------------------------------------------------------------------- */
int sdf_version(void);             /* major*100 + minor */
const char* sdf_last_error(void);  /* thread-local, never NULL */
/* kernel launches issued through this library by the calling process (bench gpu_launches) */
int64_t sdf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SDF_B200_H */
