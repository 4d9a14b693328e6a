/* fitv2_b200 — C ABI of the B200-native FiTv2 denoising hot path.
 *
 * This is the drop-in boundary: a reference-side binding (Python ctypes) calls
 * these entry points in place of fit.model.fit_model.FiT.forward / forward_with_cfg and the
 * CFG + Euler update of sample_fitv2_ddp.py.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative FITV2_E_* code on failure; the message of the
 *     last failure on the calling thread is available from fitv2_last_error();
 *   - all device pointers are CUDA device memory owned by the CALLER (the library never allocates
 *     or frees device memory; the workspace is caller-provided via fitv2_set_workspace);
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*), nothing
 *     synchronises; the calls are CUDA-graph capturable;
 *   - a handle is not thread-safe; use one handle per stream/thread.  A handle belongs to the CUDA device that was current
 *     in fitv2_create (kernel attributes, TMA descriptors and the workspace live there); calls under another current
 *     device fail with FITV2_E_INVALID.
 */
#ifndef FITV2_B200_H_
#define FITV2_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FITV2_OK              0
#define FITV2_E_INVALID      -1   /* bad argument / unsupported configuration            */
#define FITV2_E_UNBOUND      -2   /* a weight or the workspace has not been bound        */
#define FITV2_E_CUDA         -3   /* CUDA runtime / driver error (see fitv2_last_error)  */
#define FITV2_E_WORKSPACE    -4   /* workspace too small for the requested (rows, tokens) */

#define FITV2_OPERAND_BF16    0   /* 16-bit GEMM / attention operand type: bfloat16 (default) */
#define FITV2_OPERAND_FP16    1   /* float16 (same tensor-core rate, 3 more mantissa bits)   */

/* Normalisation kinds of fit/model/norms.py:35-50 (create_norm): block norms (norm1 / norm2 / norm_final, `norm_type`) accept
 * LAYERNORM / WLAYERNORM / RMSNORM; the per-head q / k norms (`q_norm`, `k_norm`, `qk_norm_weight`) also NONE. */
#define FITV2_NORM_NONE        0  /* nn.Identity                                                   */
#define FITV2_NORM_LAYERNORM   1  /* nn.LayerNorm(elementwise_affine=False, eps 1e-6): FiTv2 default */
#define FITV2_NORM_WLAYERNORM  2  /* nn.LayerNorm(bias=False): weight only                          */
#define FITV2_NORM_RMSNORM     3  /* RMSNorm with weight (norms.py:53-77)                           */

#define FITV2_ADALN_LORA       0  /* global adaLN + per-block LoRA (modules.py:259-264): FiTv2      */
#define FITV2_ADALN_NORMAL     1  /* per-block Linear(D -> 6D), no global term (modules.py:254-258): FiTv1 / DiT */
#define FITV2_ADALN_SWIGLU     2  /* per-block SwiGLU(D -> (D/4)*3 -> 6D) and final SwiGLU(D -> D/2 -> 2D) applied to c itself,
                                     no global term (modules.py:265-268,284-285); fp32 FMA kernels */

#define FITV2_MLP_SWIGLU       0  /* fc2(silu(fc1_g(x)) * fc1_x(x)), hidden (int(D*mlp_ratio)*2)//3 or int(D*mlp_ratio) */
#define FITV2_MLP_GELU         1  /* fc2(gelu_tanh(fc1(x))), hidden int(D*mlp_ratio) (a multiple of 256): use_swiglu = False,
                                     the class default of fit_model.py:38; GATEUP_W / GATEUP_B then hold fc1 alone */

/* Model geometry: the constructor contract of fit.model.fit_model.FiT (fit_model.py:25-65).  Zero-initialised trailing
 * fields select the FiTv2 family except for the three norm fields, which must be set (FITV2_NORM_LAYERNORM for FiTv2). */
typedef struct fitv2_config {
    int32_t hidden_size;      /* D      : 1152 (XL/2), 2304 (3B/2)                  */
    int32_t depth;            /* L      : 36, 40                                    */
    int32_t num_heads;        /* H      : 16, 24 (must be even)                     */
    int32_t head_dim;         /* D / H  : 72 or 96                                  */
    int32_t mlp_hidden;       /* (int(D*mlp_ratio)*2)//3 : 3072, 6144               */
    int32_t lora_dim;         /* adaln_lora_dim : 288, 576                          */
    int32_t token_channels;   /* p*p*C_in : 16                                      */
    int32_t num_embeddings;   /* rows of y_embedder.embedding_table : 1001          */
    int32_t operand_dtype;    /* FITV2_OPERAND_*                                    */
    float   time_shifting;    /* fit_model.py:202 ; 1.0 for the released configs    */
    float   rope_magnitude;   /* cos/sin magnitude (yarn mscale / ntk-pro proportion), 1.0 otherwise */
    int32_t out_channels;     /* p*p*C_out : 0 = token_channels; 32 with learn_sigma (fit_model.py:78) */
    int32_t adaln_type;       /* FITV2_ADALN_*                                      */
    int32_t block_norm;       /* FITV2_NORM_* of norm1 / norm2 / norm_final         */
    int32_t q_norm;           /* FITV2_NORM_* of the per-head q norm (modules.py:146) */
    int32_t k_norm;           /* FITV2_NORM_* of the per-head k norm (modules.py:147) */
    int32_t channels_first;   /* 1: use_sit = False callers, x (B, C, N) -> out (B, C_out, N) (fit_model.py:204,231) */
    int32_t mlp_type;         /* FITV2_MLP_*: 0 = timm SwiGLU (use_swiglu), 1 = timm Mlp with tanh-GELU (modules.py:253)   */
    int32_t rope_v;           /* 1: add_rel_pe_to_v, v is rotated like q / k (modules.py:171-172)                          */
} fitv2_config;

typedef struct fitv2_handle fitv2_handle;

/* Weight slots.  fp32 slots hold the reference parameter unchanged; OP16 slots hold the parameter
 * converted to the handle's operand dtype.  Stacked slots are the per-block parameters concatenated
 * over blocks (leading dimension = depth).
 * The four adaLN matrices (GLOBAL_ADALN_W, LORA_A_W, LORA_B_W, FINAL_ADALN_W) are fp32 in memory but are
 * consumed as TF32 operands by the tensor pipe (csrc/cond_tc.cuh) whenever the shapes tile (hidden_size
 * and lora_dim multiples of 32, 6*hidden_size / 2*hidden_size / depth*lora_dim multiples of 48): bind
 * them rounded to nearest TF32 (low 13 mantissa bits zero, as the Python packer does) - unrounded values
 * are truncated by the hardware.  FITV2_COND=simt keeps the fp32 FMA kernels.     reference parameter */
enum fitv2_weight {
    FITV2_W_X_EMBED_W = 0,     /* fp32 (D, 16)          x_embedder.proj.weight                        */
    FITV2_W_X_EMBED_B,         /* fp32 (D)              x_embedder.proj.bias                          */
    FITV2_W_T_MLP0_W,          /* fp32 (D, 256)         t_embedder.mlp.0.weight                       */
    FITV2_W_T_MLP0_B,          /* fp32 (D)                                                           */
    FITV2_W_T_MLP2_W,          /* fp32 (D, D)           t_embedder.mlp.2.weight                       */
    FITV2_W_T_MLP2_B,          /* fp32 (D)                                                           */
    FITV2_W_Y_TABLE,           /* fp32 (num_embeddings, D)  y_embedder.embedding_table.weight         */
    FITV2_W_GLOBAL_ADALN_W,    /* fp32 (6D, D)          global_adaLN_modulation.1.weight   (this and the LORA_* slots: adaln_type 'lora' only) */
    FITV2_W_GLOBAL_ADALN_B,    /* fp32 (6D)                                                          */
    FITV2_W_LORA_A_W,          /* fp32 (L*lora, D)      blocks.i.adaLN_modulation.1.weight, stacked   */
    FITV2_W_LORA_A_B,          /* fp32 (L*lora)                                                      */
    FITV2_W_LORA_B_W,          /* fp32 (L, 6D, lora)    blocks.i.adaLN_modulation.2.weight, stacked   */
    FITV2_W_LORA_B_B,          /* fp32 (L, 6D)                                                       */
    FITV2_W_FINAL_ADALN_W,     /* fp32 (2D, D)          final_layer.adaLN_modulation.1.weight         */
    FITV2_W_FINAL_ADALN_B,     /* fp32 (2D)                                                          */
    FITV2_W_FINAL_LINEAR_W,    /* fp32 (C_out, D)       final_layer.linear.weight  (C_out = 16, or 32 with learn_sigma) */
    FITV2_W_FINAL_LINEAR_B,    /* fp32 (C_out)                                                       */
    FITV2_W_QKV_W,             /* OP16 (L, 3D, D)       blocks.i.attn.qkv.weight                      */
    FITV2_W_QKV_B,             /* fp32 (L, 3D)                                                       */
    FITV2_W_PROJ_W,            /* OP16 (L, D, D)        blocks.i.attn.proj.weight                     */
    FITV2_W_PROJ_B,            /* fp32 (L, D)                                                        */
    FITV2_W_GATEUP_W,          /* OP16 (L, 2*Hm, D)     blocks.i.mlp.fc1_g / fc1_x interleaved in 128-row groups:
                                  rows [256t, 256t+128) = fc1_g rows [128t, 128t+128), next 128 = fc1_x rows */
    FITV2_W_GATEUP_B,          /* fp32 (L, 2*Hm)        same interleave.  FITV2_MLP_GELU: GATEUP_W = OP16 (L, Hm, D) blocks.i.mlp.fc1.weight,
                                  GATEUP_B = fp32 (L, Hm) blocks.i.mlp.fc1.bias (no interleave)            */
    FITV2_W_FC2_W,             /* OP16 (L, D, Hm)       blocks.i.mlp.fc2.weight                       */
    FITV2_W_FC2_B,             /* fp32 (L, D)                                                        */
    FITV2_W_ROPE_FREQS_H,      /* fp32 (head_dim/4)     VisionRotaryEmbedding.freqs_h (rope.py:162)   */
    FITV2_W_ROPE_FREQS_W,      /* fp32 (head_dim/4)     VisionRotaryEmbedding.freqs_w                 */
    /* slots of the non-default variants; a slot the configuration does not use need not (and cannot) be bound */
    FITV2_W_NORMAL_ADALN_W,    /* fp32 (L, 6D, D)       blocks.i.adaLN_modulation.1.weight, adaln_type 'normal' (tf32-rounded like the other adaLN matrices) */
    FITV2_W_NORMAL_ADALN_B,    /* fp32 (L, 6D)                                                       */
    FITV2_W_NORM1_W,           /* fp32 (L, D)           blocks.i.norm1.weight   (w_layernorm / rmsnorm) */
    FITV2_W_NORM2_W,           /* fp32 (L, D)           blocks.i.norm2.weight                           */
    FITV2_W_NORM_FINAL_W,      /* fp32 (D)              final_layer.norm_final.weight                   */
    FITV2_W_Q_NORM_W,          /* fp32 (L, head_dim)    blocks.i.attn.q_norm.weight                     */
    FITV2_W_K_NORM_W,          /* fp32 (L, head_dim)    blocks.i.attn.k_norm.weight                     */
    /* adaln_type 'swiglu' (FITV2_ADALN_SWIGLU): Hs = (D/4)*3, Hf = D/2; FINAL_ADALN_* are then unused */
    FITV2_W_SG_G_W,            /* fp32 (L, Hs, D)       blocks.i.adaLN_modulation.fc1_g.weight          */
    FITV2_W_SG_G_B,            /* fp32 (L, Hs)                                                          */
    FITV2_W_SG_X_W,            /* fp32 (L, Hs, D)       blocks.i.adaLN_modulation.fc1_x.weight          */
    FITV2_W_SG_X_B,            /* fp32 (L, Hs)                                                          */
    FITV2_W_SG_FC2_W,          /* fp32 (L, 6D, Hs)      blocks.i.adaLN_modulation.fc2.weight            */
    FITV2_W_SG_FC2_B,          /* fp32 (L, 6D)                                                          */
    FITV2_W_FSG_G_W,           /* fp32 (Hf, D)          final_layer.adaLN_modulation.fc1_g.weight       */
    FITV2_W_FSG_G_B,           /* fp32 (Hf)                                                             */
    FITV2_W_FSG_X_W,           /* fp32 (Hf, D)          final_layer.adaLN_modulation.fc1_x.weight       */
    FITV2_W_FSG_X_B,           /* fp32 (Hf)                                                             */
    FITV2_W_FSG_FC2_W,         /* fp32 (2D, Hf)         final_layer.adaLN_modulation.fc2.weight         */
    FITV2_W_FSG_FC2_B,         /* fp32 (2D)                                                             */
    FITV2_W_COUNT
};

const char* fitv2_last_error(void);
const char* fitv2_version(void);

/* Replaces FiT.__init__ (fit/model/fit_model.py:25-115) for the geometry part. */
int fitv2_create(const fitv2_config* cfg, fitv2_handle** out);
void fitv2_destroy(fitv2_handle* h);

/* Per-handle tuning switches (they replace the FITV2_* environment variables of earlier versions; the Python layer forwards
 * the environment when it creates a handle).  Names: "pdl" (1), "attn" (0 auto / 1 P-in-TMEM / 2 shared-memory-P / 3 online-max),
 * "attn_early" (1), "ln_threads" (64), "ln_wide_single" (0), "bn_resid" (0 = cost model),
 * "qkv_heads" (3), "resid_t" (-1 auto), "bn_resid_t" (0 = cost model), "cond" (0 tensor pipe / 1 fp32 FMA), "l2_persist_mb" (0),
 * "final_tc" (1), "gelu_epi" (0 slab-staged GELU epilogue / 1 plain epilogue), "ws_guard" (0), "verbose" (0).  Unknown names fail. */
int fitv2_set_option(fitv2_handle* h, const char* name, int64_t value);

/* Device-side checks report through a sticky word in pinned host memory instead of trapping: returns FITV2_E_INVALID (and
 * clears the word) when a kernel launched through this handle saw a class label outside [0, num_embeddings) - the reference
 * raises an index error there (modules.py:101-106).  Non-blocking: meaningful after the stream has been synchronised;
 * fitv2_forward also calls it on entry. */
int fitv2_poll_error(fitv2_handle* h);

/* Replaces load_state_dict / init_from_ckpt binding (fit/utils/eval_utils.py:12-71): records the device
 * pointer of one packed weight.  `numel` is checked against the slot's expected element count. */
int fitv2_bind_weight(fitv2_handle* h, int slot, const void* dev_ptr, int64_t numel);

/* online_rope mode (fit_model.py:212-214, rope.py:234-274): per-row inverse frequencies, two device arrays of
 * (rows, head_dim/4) fp32 computed by the host from each sample's `size` (dynamic NTK scale), used instead of the bound
 * ROPE_FREQS_H/W vectors by the following forward calls (which must use the same `rows`).  NULL, NULL switches back. */
int fitv2_set_online_rope(fitv2_handle* h, const float* freqs_h_rows, const float* freqs_w_rows, int rows);

/* Bytes of scratch the forward needs for (rows = batch incl. CFG duplication, tokens per row).  Between calls the contents of
 * the workspace belong to the handle (it keeps derived data there, e.g. the fp16 copy of the final-layer weight): a caller that
 * reuses the memory for something else announces it again with fitv2_set_workspace before the next forward. */
int64_t fitv2_workspace_bytes(const fitv2_handle* h, int rows, int tokens);
int fitv2_set_workspace(fitv2_handle* h, void* dev_ptr, int64_t bytes);

/* Replaces FiT.forward (fit/model/fit_model.py:189-233).  Layouts below are the use_sit ones; with cfg.channels_first
 * x is (x_rows, C, tokens) and out (rows, C_out, tokens).
 *   x      fp32 (x_rows, tokens, C)  latent tokens; x_rows == rows, or rows/2 when the caller wants the
 *                                    CFG duplication cat([z, z]) of sample_fitv2_ddp.py:299 done implicitly
 *   t      fp32 (rows)               timesteps in [0, 1]
 *   y      int64 (rows)              class labels (num_classes = null class)
 *   grid   int64 (rows, 2, tokens)   [:,0] = w index, [:,1] = h index
 *   mask   fp32 (rows, tokens)       segment ids (0 = padding): tokens attend where the ids are equal (fit/model/modules.py:176-204).
 *                                     All ids equal (the sampling scripts) and "n equal non-zero ids, then zeros" (a padded sample of a
 *                                     mixed-aspect batch) are recognised on the device and run without per-element id compares;
 *                                     any other pattern (several packed images per row) is compared element by element.
 *   out    fp32 (rows, tokens, C_out) velocity (or eps | sigma with learn_sigma); rows with mask 0 are exactly 0
 */
int fitv2_forward(fitv2_handle* h, const float* x, int x_rows, const float* t, const int64_t* y,
                  const int64_t* grid, const float* mask, float* out, int rows, int tokens, void* stream);

/* Replaces the CFG lines of FiT.forward_with_cfg (fit_model.py:253-275): in place on out (2B, tokens, C),
 * channels [0, c_cfg) of both halves <- uncond + s*(cond - uncond).  scale_per_sample (B) may be NULL. */
int fitv2_cfg_combine(float* out, const float* scale_per_sample, float scale, int half_rows, int tokens,
                      int channels, int c_cfg, void* stream);

/* Replaces sample_fitv2_ddp.py:310-314: z (B, tokens, C) += dsigma * (uncond + cfg*(cond - uncond)),
 * v2 = (2B, tokens, C) with rows [0,B) conditional.  fp32, bit-exact with the PyTorch expression.
 * dsigma_dev (nullable): when non-NULL the step size sigma_next - sigma_cur is read from this device
 * scalar instead of `dsigma`, so one captured CUDA graph can be replayed for every step. */
int fitv2_cfg_euler(float* z, const float* v2, float cfg_scale, float dsigma, const float* dsigma_dev,
                    int half_rows, int tokens, int channels, void* stream);

/* ---- transport Sampler updates (reference: fit/scheduler/transport/{transport,integrators,path}.py), velocity model,
 * linear path.  Elementwise fp32 over n elements, bit-exact with the PyTorch expressions of the reference.
 * coef_dev: 8 device floats computed by the host per step with the reference's expression order:
 *   [0] alpha_t/d_alpha_t  [1] sigma_t^2 - [0]*d_sigma_t*sigma_t  [2] diffusion(t)  [3] dt (or last-step size)
 *   [4] sqrt(2*diffusion)  [5] sqrt(dt)  [6] alpha_t  [7] sigma_t^2/alpha_t                                  */

/* Replaces sde.__Euler_Maruyama_step (integrators.py:29-37) with ONE model evaluation shared by drift and score
 * (the reference evaluates the network twice, transport.py:256-258):
 *   x <- (x + (v + D*score(v,x,t))*dt) + sqrt(2D) * (w*sqrt(dt)),   score = ([0]*v - x)/[1]   (path.py:71-85).
 * w == NULL: the noise-free "Mean" last step x <- x + drift*[3] (transport.py:277-280). */
int fitv2_sde_step(float* x, const float* v, const float* w, const float* coef_dev, int64_t n, void* stream);
/* out = v + D*score(v, x, t): the SDE drift (K1, K2 of sde.__Heun_step, integrators.py:45-47). */
int fitv2_sde_drift(float* out, const float* x, const float* v, const float* coef_dev, int64_t n, void* stream);
/* out = a + s[1]*(s[0]*b) (s_dev: 2 device floats): xhat = x + sqrt(2D)*(w*sqrt(dt)), xp = xhat + dt*K1 (integrators.py:43,46),
 * the fixed-grid ODE Euler / midpoint updates (integrators.py:109-116) and the "Euler" last step (transport.py:287-290). */
int fitv2_scaled_add(float* out, const float* a, const float* b, const float* s_dev, int64_t n, void* stream);
/* out = xhat + c*(k1 + k2), c_dev[0] = 0.5*dt (integrators.py:48). */
int fitv2_heun_combine(float* out, const float* xhat, const float* k1, const float* k2, const float* c_dev, int64_t n, void* stream);
/* "Tweedie" last step (transport.py:281-286): out = x/[6] + [7]*score(v, x, t). */
int fitv2_tweedie(float* out, const float* x, const float* v, const float* coef_dev, int64_t n, void* stream);

/* Fixed-grid Runge-Kutta stages of the reference's ODE route (Sampler.sample_ode -> torchdiffeq.odeint, integrators.py:109-116;
 * torchdiffeq is an un-vendored dependency, its fixed_grid.py / rk_common.py step functions are restated).  s_dev: 4 device
 * floats {dt, s1, s2, s3}; every operation is a separately rounded fp32 operation in PyTorch's evaluation order.
 *   mode 0: out = y + (dt*k1)*s1          mode 1: out = y + dt*(k1*s1 + k2*s2)      mode 2: out = y + dt*(k1*s1 + k2*s2 + k3*s3)
 *   mode 3: out = y + dt*(k2 - k1*s1)     mode 4: out = y + dt*(k1 - k2 + k3)       mode 5: out = y + (k1 + 3*(k2+k3) + k4)*dt*0.125 */
int fitv2_rk_stage(float* out, const float* y, const float* k1, const float* k2, const float* k3, const float* k4,
                   const float* s_dev, int mode, int64_t n, void* stream);

/* Adaptive Dormand-Prince 5(4) ("dopri5", the reference's DEFAULT --ode-sampling-method: fit/utils/sit_eval_utils.py:20 ->
 * torchdiffeq.odeint, integrators.py:109-116).  The step-size controller runs on the host (fitv2_b200/transport.py restates
 * torchdiffeq's published rk_common.py / dopri5.py / interp.py); the device side is two kernels:
 *   fitv2_lincomb:    out = c[0]*y + sum_{i<nk} c[1+i]*k[i]   (nk <= 7; c_dev: 1 + nk device floats; separately rounded, in order)
 *   fitv2_scaled_rms: out_dev[0] = sqrt(mean(((a - b) / (atol + rtol*|s|))^2)), b and s nullable (plain RMS norm when both are
 *                     NULL); one block, fixed summation order (deterministic). */
int fitv2_lincomb(float* out, const float* y, const float* const* k, const float* c_dev, int nk, int64_t n, void* stream);
int fitv2_scaled_rms(float* out_dev, const float* a, const float* b, const float* s, float atol, float rtol, int64_t n, void* stream);

/* ---- after the trajectory (sample_fitv2_ddp.py:319-324) ----
 * Replaces FiT.unpatchify (fit_model.py:171-187, use_sit layout) fused with the latent scaling `samples / vae.config.scaling_factor`:
 *   z (batch, hp*wp, channels*patch*patch) fp32 -> out (batch, channels, hp*patch, wp*patch) fp32, out = unpatchify(z) / scaling_factor
 * (scaling_factor 1 = plain unpatchify).  Bit-exact (one IEEE division per element). */
int fitv2_unpatchify_scale(const float* z, float* out, float scaling_factor, int batch, int hp, int wp, int channels, int patch, void* stream);
/* Replaces `samples.clamp(-1,1)` ... `torch.clamp(127.5*samples + 128.0, 0, 255).permute(0,2,3,1).to(torch.uint8)`
 * (sample_fitv2_ddp.py:321-323): img (batch, channels, H, W) fp32 -> out (batch, H, W, channels) uint8.  Bit-exact. */
int fitv2_pack_uint8(const float* img, unsigned char* out, int batch, int channels, int height, int width, void* stream);

/* Component entry points (same kernels the forward uses) — exercised by the parity tests. */
int fitv2_debug_gemm(fitv2_handle* h, int epilogue /*3 = plain*/, const void* a, const void* w, const float* bias,
                     float* out32, int M, int N, int K, int bn, void* stream);
/* dbg_s (128x128 fp32) / dbg_o (128 x head_dim_padded fp32), both nullable: raw S = Q K^T and P V tiles of
 * CTA (0,0,0), first key tile (served by the online-max kernel, attention_general.cuh).  The kernel follows the handle's
 * q_norm / k_norm and the "attn" option like the forward does. */
int fitv2_debug_attention(fitv2_handle* h, const void* q, const void* k, const void* vt, const float* mask,
                          void* out, int rows, int tokens, float* dbg_s, float* dbg_o, void* stream);
int fitv2_debug_tap(fitv2_handle* h, int what, void* dst, int64_t bytes, void* stream);
/* (offset, bytes) of every buffer the last forward carved out of the workspace, in allocation order; returns the count.  With
 * option "ws_guard" = G every buffer is followed by at least G bytes that no kernel may touch: the bounds test (compute-sanitizer is
 * not available on the target pool) fills the workspace with a canary, runs a forward and checks the padding. */
int fitv2_debug_layout(const fitv2_handle* h, int64_t* offsets, int64_t* sizes, int max_entries);
int64_t fitv2_kernel_launches(const fitv2_handle* h);   /* launches enqueued by this handle so far */

/* Per-kernel-class device timing with CUDA events on the launching stream (measurement support for bench.py).
 * Classes: 0 conditioning, 1 LayerNorm+modulate, 2 QKV GEMM, 3 attention, 4 proj GEMM, 5 gate/up GEMM,
 * 6 fc2 GEMM, 7 patch-embed + final layer.  class_mask bit i enables class i; 0 disables.  Not capturable. */
#define FITV2_PROFILE_CLASSES 8
int fitv2_profile_set(fitv2_handle* h, uint32_t class_mask);
int fitv2_profile_read(fitv2_handle* h, double* ms_sum /*[8]*/, int64_t* count /*[8]*/);

#ifdef __cplusplus
}
#endif
#endif  /* FITV2_B200_H_ */
