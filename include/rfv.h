/*
 * rfv.h -- C ABI of the B200 (sm_100a) rectified-flow Euler-integration engine.
 *
 * The reference (AlbertGoTri/rectified-flow-vision) has no FFI / plugin interface: its boundary for this path is
 * the Python API re-exported by models/__init__.py:5-12.  This header is the C-ABI seam UNDER that API: every
 * entry point below replaces the numerical body of one reference function, takes plain pointers and sizes
 * (no torch types), never throws, and returns 0 on success or a negative rfv_status with a message retrievable
 * from rfv_last_error().  One handle per GPU; a handle is not thread-safe.  All device pointers are on the
 * handle's device; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *
 * Layout contract at the boundary (what the reference's callers hold): contiguous NCHW fp32 images,
 * fp32 time vector t[B], fp32 OIHW parameters named exactly like the reference state_dict
 * ("velocity_net.enc_blocks.0.conv1.weight", ... 174 tensors, models/unet.py:157-227).
 * Inside the engine activations are NHWC bf16, accumulators / GroupNorm statistics / Euler state are fp32.
 */
#ifndef RFV_H_
#define RFV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFV_ABI_VERSION 1
#define RFV_MAX_LEVELS 8

typedef struct rfv_engine* rfv_handle;

typedef enum {
    RFV_OK = 0,
    RFV_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
    RFV_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed */
    RFV_ERR_STATE = -3,     /* weights not (fully) loaded, batch larger than capacity, ... */
    RFV_ERR_NOMEM = -4
} rfv_status;

/* Mirrors the constructor arguments of UNet / BaseFlowModel (models/unet.py:136-145, models/base_flow.py:37-47).
 * attention_resolutions and dropout are not here: the former is ignored by the reference (models/unet.py:143),
 * the latter is identity on the eval-mode path this engine implements. */
typedef struct {
    int32_t image_size;                     /* S: images are [C, S, S] */
    int32_t in_channels;                    /* 3 */
    int32_t out_channels;                   /* = in_channels for flow models (models/base_flow.py:58) */
    int32_t model_channels;                 /* 64; must be a multiple of 64 */
    int32_t num_levels;                     /* len(channel_mult) */
    int32_t channel_mult[RFV_MAX_LEVELS];   /* [1,2,4] */
    int32_t num_res_blocks;                 /* 2 */
    int32_t num_heads;                      /* 4 (models/unet.py:70) */
    int32_t micro_batch;                    /* images integrated together; activation arena capacity */
    int32_t device;                         /* CUDA ordinal */
    int32_t flags;                          /* RFV_FLAG_* */
} rfv_config;

#define RFV_FLAG_NO_UMMA   1   /* force the mma.sync implicit-GEMM kernel everywhere (debug / A-B testing) */
#define RFV_FLAG_ONE_LANE  2   /* rfv_euler_sample / rfv_euler_sample_host: integrate the micro-batches of a large batch one after
                                  the other on one stream instead of as two alternately enqueued chains on two streams (the second
                                  chain needs a twin engine: a second activation arena and weight copy, created on first use); A/B */
#define RFV_FLAG_KEEP_ACTS 4   /* never recycle activation buffers, so rfv_debug_activation can read any layer */

/* bits 8-10 (values 256 / 512 / 1024) are not switches but a field: (flags >> 8) & 7 = CTAs per cluster for TMA weight
 * multicast in the per-tap tcgen05 conv kernel (2 or 4; measured slower than 1 on B200: forward 5.11 / 5.29 / 5.36 ms at
 * cluster 1 / 2 / 4, micro-batch 256). */
#define RFV_FLAG_ONE_STREAM 2048 /* training: run the whole backward pass on one stream (default: weight / bias gradients on a
                                   second, lower-priority stream so the tcgen05 wgrad kernel overlaps the GroupNorm backward) */
#define RFV_FLAG_FUSE_GN   4096 /* apply GroupNorm+SiLU to the conv's operand in shared memory (conv_wa.cuh FUSE: 8 transform warps,
                                   register budgets re-split with setmaxnreg) in EVERY weights-as-A conv instead of only on the 32x32
                                   level.  Parity-green, but slower at 64x64, where the transform (issue-bound, ~60 instructions
                                   per 8-channel vector over a 1.8x halo-redundant box) outlasts the MMAs: forward 8.03 -> 8.47 ms
                                   at micro-batch 512. */
#define RFV_FLAG_NO_ATTN_UMMA 8192 /* attention core on the mma.sync kernel even where the tcgen05 one applies (A/B testing) */
#define RFV_FLAG_GN_BWD_TWO_PASS 16384 /* GroupNorm backward as two streaming passes (reduce, apply) everywhere instead of
                                   the single-pass kernel (A/B testing; the two-pass kernels remain the fallback for pixel counts
                                   that do not split into equal slices) */
#define RFV_FLAG_SILU_EXP  32768 /* GroupNorm+SiLU kernel evaluates x / (1 + exp(-x)) (2 MUFU ops) instead of h * (1 + tanh(h)),
                                   h = x/2 (1 MUFU op, ~2.5e-4 * |h| absolute error); A/B testing */
#define RFV_FLAG_OUTPUT_CONV_TAPS 65536 /* output conv as nine tap-shifted GEMMs (first formulation) instead of one multiply per
                                   staged pixel followed by a 27-term gather; A/B testing */
#define RFV_FLAG_INPUT_CONV_FMA 131072 /* input conv on the fp32 FMA pipe (fp32 x and weights) instead of mma.sync with bf16
                                   operands; A/B testing */
#define RFV_FLAG_TEMB_PER_STEP 262144 /* Euler loops evaluate the time MLP once per step (two single-row launches) instead of
                                   once per loop for all steps (A/B testing) */
#define RFV_FLAG_NO_FUSE_GN 524288 /* never apply GroupNorm+SiLU inside the consuming conv (by default the sampling plan does so
                                   on the 32x32 level, where the halo box is only 1.5x the tile and the fused kernel wins) */
#define RFV_FLAG_NO_WA     1048576 /* 3x3 stride-1 convs at the 32/64/128-pixel levels: do not use the weights-as-A kernel (conv_wa.cuh:
                                    * A = 128-row weight block, B = up to 256 pixels), fall back to the per-tap implicit-GEMM kernel
                                    * (conv_umma.cuh) that serves every other conv shape (A/B) */
#define RFV_FLAG_NO_GRAPH  4194304 /* Euler loops: enqueue every kernel of every step directly instead of replaying the captured
                                    * CUDA graph of the whole N-step loop of a micro-batch (A/B; profiling mode does so too) */
#define RFV_FLAG_NO_CTA_PAIR 8388608 /* 256-output-channel stride-1 convs: one CTA per tile (conv_umma_kernel<256>) instead of CTA pairs
                                    * executing one M = 256 tcgen05.mma.cta_group::2 (conv_umma2.cuh); A/B */
#define RFV_FLAG_NO_PDL    16777216 /* sampling chains: launch every kernel fully serialised behind its predecessor instead of as a
                                    * programmatic dependent (cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's
                                    * prologue -- barrier init, TMEM allocation, resident weight loads -- runs under the tail of the
                                    * previous one and stops in griddepcontrol.wait); A/B */
#define RFV_FLAG_TRAIN     32  /* build the backward plan too: keeps every activation, allocates gradient / Adam buffers */

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int rfv_abi_version(void);
const char* rfv_last_error(void);
int rfv_create(const rfv_config* cfg, rfv_handle* out);
int rfv_destroy(rfv_handle h);

/* ---- parameters: replaces nn.Module.load_state_dict / state_dict for the engine's packed copies
 *      (BaseFlowModel.load/save, models/base_flow.py:210-226 stay in Python and keep the .pt format) ------- */
int rfv_num_tensors(rfv_handle h);
/* name and element count of the i-th parameter tensor the engine expects (reference state_dict names). */
int rfv_tensor_info(rfv_handle h, int index, char* name_buf, int name_buf_len, int64_t* numel);
/* Upload one fp32 tensor (DEVICE pointer, reference layout: conv OIHW, linear [out,in], vectors [C]);
 * the engine repacks it (bf16 K-major [O][kh][kw][I] for tensor-core convs). */
int rfv_set_tensor(rfv_handle h, const char* name, const float* dev_ptr, int64_t numel, void* stream);
/* Read a packed tensor back as fp32 in reference layout (bf16-rounded where the engine stores bf16). */
int rfv_get_tensor(rfv_handle h, const char* name, float* dev_ptr, int64_t numel, void* stream);
/* The engine's fp32 copy of a parameter exactly as uploaded / as last written by rfv_optimizer_step (reference layout). */
int rfv_get_master(rfv_handle h, const char* name, float* dev_ptr, int64_t numel, void* stream);

/* ---- the hot path ------------------------------------------------------------------------------------- */
/* v = velocity_net(x, t): UNet.forward, models/unet.py:229-275 via BaseFlowModel.forward, models/base_flow.py:91-102.
 * x, v: [B,C,S,S] fp32 NCHW device; t: [B] fp32 device.  Any B >= 1 (processed in micro-batches). */
int rfv_velocity(rfv_handle h, const float* x, const float* t, float* v, int64_t batch, void* stream);

/* In-place N-step explicit Euler: for i < num_steps: x += dt * v(x, i*dt), dt = 1/num_steps
 * (BaseFlowModel.sample loop, models/base_flow.py:158-173; sample_with_trajectory :196-208).
 * x: [B,C,S,S] fp32 NCHW device, updated in place (the Python wrapper clones the caller's noise first).
 * traj (optional, may be NULL): [num_steps/save_every, B, C, S, S] fp32 device, receives x after every
 * save_every-th step. */
int rfv_euler_sample(rfv_handle h, float* x, int64_t batch, int num_steps, float* traj, int save_every,
                     void* stream);

/* Same integration with HOST buffers (pinned or pageable): noise_host -> out_host, [N,C,S,S] fp32.
 * Host<->device copies are pipelined with compute on internal streams; this is the call behind
 * generate_reflow_pairs (models/rectified_flow.py:127-174) and the bench's end-to-end number.
 * out_host may alias noise_host. */
int rfv_euler_sample_host(rfv_handle h, const float* noise_host, float* out_host, int64_t n, int num_steps);

/* Euler loop that also accumulates mean((v - (x1-x0))^2) per step:
 * RectifiedFlowModel.compute_straightness, models/rectified_flow.py:98-124.  x0, x1 device fp32 NCHW;
 * dev_out: [num_points] fp32 DEVICE (per-step MSE; the caller averages).  No per-step host sync. */
int rfv_straightness(rfv_handle h, const float* x0, const float* x1, int64_t batch, int num_points,
                     float* dev_out, void* stream);

/* Flow-matching forward loss: x_t = (1-t) x0 + t x1, target = x1 - x0 (BaseFlowModel.get_interpolation,
 * models/base_flow.py:81-89), loss = mean((v(x_t,t) - target)^2) (models/rectified_flow.py:222-231), eval-mode
 * network.  loss_out: 1 fp32 DEVICE value. */
int rfv_fm_loss(rfv_handle h, const float* x0, const float* x1, const float* t, int64_t batch,
                float* loss_out, void* stream);

/* ---- training: the body of train_rectified_flow / train_base_flow (models/rectified_flow.py:217-238,
 *      models/base_flow.py:266-275): loss = mse(v(x_t,t), x1-x0); loss.backward(); clip_grad_norm_(1.0); AdamW.step().
 *      Needs RFV_FLAG_TRAIN.  The pieces are separate so a data-parallel caller can all-reduce the flat gradient
 *      buffer between rfv_train_accumulate and rfv_optimizer_step (models have no collective in the reference). -- */
typedef struct {
    float lr, beta1, beta2, eps, weight_decay;  /* torch.optim.AdamW: 1e-4 (caller), 0.9, 0.999, 1e-8, 0.01 */
    float max_grad_norm;                        /* clip_grad_norm_ threshold (1.0 in the reference); <= 0: no clipping */
    float grad_scale;                           /* applied to every gradient first (1/world_size after a SUM all-reduce) */
    int64_t step;                               /* 1-based optimizer step (bias correction) */
} rfv_adamw;

/* Zero the flat gradient buffer (optimizer.zero_grad()). */
int rfv_zero_grad(rfv_handle h, void* stream);
/* Forward + backward of mean((v((1-t) x0 + t x1, t) - (x1 - x0))^2) over `batch` rows, training-mode network
 * (dropout_p as in the model constructor, 0.1 in the reference; mask from a counter-based generator keyed by
 * `seed`).  Parameter gradients are ADDED into the flat buffer; loss_out (1 fp32 DEVICE value) receives the mean
 * loss of this call.  Batches larger than the micro-batch are processed in chunks (gradient accumulation). */
int rfv_train_accumulate(rfv_handle h, const float* x0, const float* x1, const float* t, int64_t batch,
                         float dropout_p, uint64_t seed, float* loss_out, void* stream);
/* The two halves of rfv_train_accumulate for callers that own the loss (the autograd surface of the Python mirror:
 * UNet.forward in training mode, models/unet.py:229-275 with nn.Dropout active, followed by loss.backward(),
 * models/base_flow.py:268-270).  rfv_train_forward evaluates v = velocity_net(x, t) in training mode and KEEPS the
 * activations of this one micro-batch (batch <= micro_batch, else RFV_ERR_STATE); rfv_train_backward takes dL/dv
 * ([batch,C,S,S] fp32 NCHW device) and ADDS dL/dparam into the flat gradient buffer.  x and t must stay valid until the
 * backward call (the input conv's weight gradient re-reads them).  Any other compute call on the handle in between
 * invalidates the kept activations (rfv_train_backward then fails with RFV_ERR_STATE instead of using them). */
int rfv_train_forward(rfv_handle h, const float* x, const float* t, int64_t batch, float dropout_p, uint64_t seed,
                      float* v_out, void* stream);
int rfv_train_backward(rfv_handle h, const float* dv, int64_t batch, void* stream);
/* Zero the AdamW moments (a fresh torch.optim.AdamW, as every train_* call of the reference builds:
 * models/rectified_flow.py:208, models/base_flow.py:255).  The step counter lives with the caller (rfv_adamw.step). */
int rfv_reset_optimizer(rfv_handle h, void* stream);
/* The flat fp32 gradient buffer (device memory owned by the engine; one slot per parameter tensor, laid out in the order the
 * gradients become final during the backward pass; conv-weight slots are [O][kh*kw][I]; rfv_get_grad reads one tensor). */
int rfv_grad_buffer(rfv_handle h, float** dev_ptr, int64_t* numel);
/* Data-parallel overlap: the buffer is cut into rfv_grad_bucket_count contiguous ranges in completion order.
 * rfv_grad_bucket_wait makes `stream` wait until range `index` of the LAST rfv_train_accumulate call is final, so a caller
 * can enqueue the all-reduce of a finished range on a side stream while the rest of the backward pass still runs
 * (models have no collective in the reference; this is the hook for "bucketed, overlapped" gradient all-reduce). */
int rfv_grad_bucket_count(rfv_handle h);
int rfv_grad_bucket_info(rfv_handle h, int index, int64_t* offset, int64_t* numel);
int rfv_grad_bucket_wait(rfv_handle h, int index, void* stream);
/* One parameter's gradient, scaled, in reference layout (tests / debugging). */
int rfv_get_grad(rfv_handle h, const char* name, float* dev_ptr, int64_t numel, float scale, void* stream);
/* Make the optimizer also write updated values into caller-owned fp32 storage (the torch Parameter). */
int rfv_bind_param(rfv_handle h, const char* name, float* dev_ptr);
/* Global-norm clip + AdamW update of every parameter from the flat gradient buffer, then refresh of the packed
 * bf16 copies.  grad_norm_out (optional, DEVICE): the pre-clip global gradient norm. */
int rfv_optimizer_step(rfv_handle h, const rfv_adamw* hyper, float* grad_norm_out, void* stream);

/* ---- quality metrics on device tensors ------------------------------------------------------------------------
 * Stateless (no handle); all pointers are DEVICE pointers.  They replace the numpy / scipy / skimage arithmetic of
 * MetricsCalculator (utils/metrics.py:39-116) for tensors that already live in HBM (the sampler's output). */
/* mu[j] = mean over the n rows of x [n][d] (fp32); mu: [d] fp64.  np.mean(images_flat, axis=0), utils/metrics.py:85. */
int rfv_metrics_mean(const float* x, int64_t n, int64_t d, double* mu, void* stream);
/* sigma = (x - mu)^T (x - mu) / (n - 1), [d][d] fp64: np.cov(images_flat, rowvar=False), utils/metrics.py:86.  n >= 2. */
int rfv_metrics_covariance(const float* x, const double* mu, int64_t n, int64_t d, double* sigma, void* stream);
/* The pieces of MetricsCalculator.compute_fid (utils/metrics.py:89-116) without forming a d x d matrix:
 *   terms[0] = |mu1 - mu2|^2, terms[1] = tr sigma1, terms[2] = tr sigma2,
 *   gram [n1][n2] fp64 = (x1 - mu1)(x2 - mu2)^T / sqrt((n1-1)(n2-1)): its singular values sum to tr sqrtm(sigma1 sigma2),
 * so FID = terms[0] + terms[1] + terms[2] - 2 * sum(svdvals(gram)).  n1, n2 >= 2. */
int rfv_metrics_fid_terms(const float* x1, const double* mu1, int64_t n1, const float* x2, const double* mu2, int64_t n2, int64_t d,
                          double* gram, double* terms, void* stream);
/* Mean structural similarity of `batch` image pairs a, b [batch][channels][height][width] fp32 -> out [batch] fp64:
 * skimage.metrics.structural_similarity(channel_axis, data_range) with its defaults (7x7 uniform window, K1 = 0.01, K2 = 0.03,
 * sample covariance, mean over the image cropped by 3 pixels), as called at utils/metrics.py:52.  height, width >= 7. */
int rfv_metrics_ssim(const float* a, const float* b, int64_t batch, int channels, int height, int width, float data_range,
                     double* out, void* stream);

/* ---- introspection for tests / bench -------------------------------------------------------------------- */
/* Number of engine kernels launched (or graph-replayed) since the last call with reset != 0. */
int64_t rfv_launch_count(rfv_handle h, int reset);
/* Algorithmic FLOPs of one velocity evaluation per image (2*MAC convs/linears + 4*C*N^2 attention). */
double rfv_flops_per_image(rfv_handle h);
/* Copy a named intermediate activation of the LAST micro-batch of the last rfv_velocity call, converted to fp32
 * NCHW (names: "input_conv", "enc_blocks.0", "downsamples.0", "mid_block1", "mid_attn", "dec_blocks.3",
 * "upsamples.1", ...).  Returns the element count, or a negative status.  Test hook. */
int64_t rfv_debug_activation(rfv_handle h, const char* name, float* dev_out, int64_t capacity, void* stream);
/* Event-timed duration (ms) of each kernel class inside the last call made with profiling enabled.  The report is one line
 * per launch site: "<kind> <label>\t<total ms>\t<launches>\t<algorithmic FLOPs per image>\t<algorithmic HBM bytes per image>"
 * (bytes are set for the HBM-bound GroupNorm kernels, 0 elsewhere). */
int rfv_set_profiling(rfv_handle h, int enabled);
int rfv_profile_report(rfv_handle h, char* buf, int buf_len);

#ifdef __cplusplus
}
#endif
#endif /* RFV_H_ */
