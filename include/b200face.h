/* b200face.h -- C ABI of libb200face.so (B200 / sm_100a).
 *
 * The reference (henryhcooperr/FaceRecognition-MultiArchitecture-Pipeline) is pure Python and
 * exposes no FFI; the drop-in boundary is the Python surface
 *     ArcMarginProduct.forward / update_epoch / get_margin_stats   src/face_models.py:297-445
 *     criterion = nn.CrossEntropyLoss(label_smoothing=eps)          src/training.py:341,515
 *     loss.backward() + ArcFaceNet backward hook                    src/face_models.py:538-570
 *     compare_faces(emb, refs, thresh)                              src/app.py:50-64
 *     cosine class-centre match                                     src/hyperparameter_tuning.py:1039-1046,1076
 * and this header is what a ctypes binding of that surface calls (INTEGRATION.md shows the
 * binding).  Conventions for every entry point:
 *   - all data pointers are DEVICE pointers, row-major, borrowed for the call; the caller
 *     (PyTorch) owns every buffer including the workspace;
 *   - the last argument is the cudaStream_t to launch on (void* so no CUDA header is needed);
 *   - no allocation, no implicit synchronisation; results depend only on the arguments.  The only process-wide
 *     state is the set of performance tunables (b200f_set_tunable): they select between kernel variants with the
 *     same results and exist for tests and bench sweeps.  Set them before the first call or between calls, not while
 *     another thread is inside the library: every call reads each tunable once, so a call is consistent in itself,
 *     but "pair" and "g_chunk_mb" change the workspace layout between a b200f_arcface_bwd_phase 1 / 2 pair.  With
 *     the defaults untouched any number of threads may call in concurrently (the Streamlit UI thread + webcam
 *     thread of src/app.py:331-335,639), each on its own stream with its own workspace;
 *   - return 0 on success, <0 on error; b200f_last_error() gives the thread-local message;
 *   - a pipeline wait inside a tcgen05 kernel that expires (seconds; a bug, a stalled peer CTA) ABORTS the kernel
 *     with a trap: the next CUDA call on the stream returns an error, nothing computed from partial accumulators
 *     is ever handed back as a result.
 * Element types: B200F_F32 / B200F_BF16 run the head on the fp32 CUDA-core engine (fp32 products: the
 * 1e-5 bar); B200F_F16N (K1's normalised fp16 output, made from bf16 or fp32 inputs) runs it on the
 * tcgen05/TMEM/TMA engine.  All arithmetic accumulates in fp32.
 */
#ifndef B200FACE_H_
#define B200FACE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200F_F32   0
#define B200F_BF16  1
/* fp16 rows that are ALREADY L2-normalised and multiplied by a power of two (operand_scale below):
 * what K1 emits for the tcgen05 engine.  x_hat / w_hat elements are <= 1 in magnitude, so fp16's 11-bit
 * significand is used at full precision (bf16 inputs are widened, never narrowed).  Only valid as the
 * out_dtype of b200f_l2norm_rows and as the operand dtype of the head calls on the tcgen05 engine. */
#define B200F_F16N  2

#define B200F_METRIC_L2EPS 0   /* || q - g + 1e-6 ||_2, ascending   (src/app.py:59)            */
#define B200F_METRIC_COS   1   /* <q,g> * inv|q| * inv|g|, descending (hyperparameter_tuning.py:1046) */

#define B200F_OK            0
#define B200F_ERR_ARG      -1
#define B200F_ERR_CUDA     -2
#define B200F_ERR_WORKSPACE -3
#define B200F_ERR_UNSUPPORTED -4

/* Engines for the GEMM-shaped stages. AUTO picks the tcgen05/TMEM/TMA path for bf16 inputs
 * whose shapes it supports and the fp32 CUDA-core path otherwise (fp32 inputs: the 1e-5 bar
 * needs fp32 products). */
#define B200F_ENGINE_AUTO    0
#define B200F_ENGINE_SIMT    1
#define B200F_ENGINE_TCGEN05 2

/* Effective head parameters for ONE forward/backward call.  The epoch-driven schedule
 * (src/face_models.py:336-348), the 24.0 scale cap and the m>0.4 damping (:401-409) are host
 * logic; the kernels see only their result. */
typedef struct b200f_head_cfg {
  float   m_eff;             /* m * margin_factor (training) or m (eval)      :369        */
  float   s_eff;             /* effective scale applied to the logits        :401-412    */
  float   label_smoothing;   /* eps of nn.CrossEntropyLoss                   training.py:341 */
  int32_t easy_margin;       /* 0: cos(min(pi-1e-4, theta+m)), 1: easy branch :372-397   */
  int64_t num_classes_total; /* C over all shards (label smoothing uses eps/C)            */
  int32_t engine;            /* B200F_ENGINE_*                                             */
  float   operand_scale;     /* B200F_F16N operands: x = x_hat * operand_scale, w = w_hat * operand_scale */
} b200f_head_cfg;

/* Per-row forward statistics, [B, B200F_STAT_COLS] fp32, additive over class shards
 * (the softmax shift is the constant s_eff: |logit| <= s_eff):               */
#define B200F_STAT_SUMEXP   0   /* sum_j exp(z_ij - s_eff)                                 */
#define B200F_STAT_SUMEXP2  1   /* sum_j exp(2 (z_ij - s_eff))  -> ||p-q||_F for the hook  */
#define B200F_STAT_ZTARGET  2   /* z_{i,y_i} (0 on shards that do not own y_i)             */
#define B200F_STAT_SUMZ     3   /* sum_j z_ij                  (label smoothing)           */
#define B200F_STAT_COLS     4

int         b200f_version(void);
const char* b200f_last_error(void);
/* number of kernels this library has launched in the process so far (diagnostic; bench.py's
 * gpu_launches is a difference of two reads) */
unsigned long long b200f_launch_count(void);
/* 1 when the library was built with the tcgen05 kernels and the current device is sm_100 */
int         b200f_has_tcgen05(void);

/* K1 -- fused row L2-normalise: inv_norm[r] = 1 / max(||in[r,:]||_2, eps)  (F.normalize,
 * src/face_models.py:351-352,525).  out (optional) = in * inv_norm * out_scale in out_dtype
 * (B200F_F32 / B200F_BF16 / B200F_F16N). */
int b200f_l2norm_rows(const void* in, int in_dtype, int64_t rows, int dim, float eps,
                      float* inv_norm, void* out_or_null, int out_dtype, float out_scale, void* stream);

/* K1 over TWO row sets in one launch (the head's batch rows and class weights: saves the launch of the small set).
 * Same arithmetic as two b200f_l2norm_rows calls with the shared in_dtype / dim / eps / out_dtype / out_scale. */
int b200f_l2norm_rows_pair(const void* in0, int64_t rows0, float* inv0, void* out0,
                           const void* in1, int64_t rows1, float* inv1, void* out1,
                           int in_dtype, int dim, float eps, int out_dtype, float out_scale, void* stream);

/* Bytes of workspace the head calls need for (B, C_local, D). */
size_t b200f_head_workspace_bytes(int64_t B, int64_t C_local, int D, int dtype, int engine);

/* K2 -- cosine logits x_hat . w_hat^T for this shard's classes [class_offset, class_offset+C_local),
 * clamp, angular margin on the target column, scale, NaN/Inf scrub, and the per-row softmax /
 * cross-entropy statistics, without writing the B x C logits (src/face_models.py:351-427 +
 * training.py:515).  Outputs:
 *   row_stats  [B,4] fp32 (see B200F_STAT_*), row_best [B] fp32 + row_argmax [B] int64
 *   (max logit and its GLOBAL class index, first index on ties), cos_minmax [2] fp32
 *   ({min,max} raw cosine over the shard, :358-360), nan_flag [1] int32 (set to 1 when a
 *   non-finite logit was scrubbed to 0, :423-427).
 *   logits_or_null: when non-NULL the scaled logits are ALSO stored ([B, ld_logits] fp32) --
 *   the compatibility path of ArcMarginProduct.forward at small C. */
int b200f_arcface_fwd(const void* x, const void* w, int dtype,
                      const float* inv_nx, const float* inv_nw, const int64_t* label,
                      int64_t B, int64_t C_local, int64_t class_offset, int D,
                      const b200f_head_cfg* cfg,
                      float* row_stats, float* row_best, int64_t* row_argmax,
                      float* cos_minmax, int32_t* nan_flag,
                      float* logits_or_null, int64_t ld_logits,
                      void* workspace, size_t workspace_bytes, void* stream);

/* State of the ArcFaceNet backward hook (src/face_models.py:538-570) as its closure would read it at backward time. */
typedef struct b200f_hook_cfg {
  int32_t enabled;        /* 0 until the second training forward (the reference registers the hook after the first) */
  float   max_grad_norm;  /* model.max_grad_norm (1.0)                                                      :544 */
  int32_t phase;          /* 1 -> thr = min(0.5, max_grad_norm)                                             :546 */
  int32_t epoch;          /* current_epoch: thr = min(thr, 0.5 + 0.05 epoch) while < 10                     :549 */
} b200f_hook_cfg;

/* K2 + K2b in one call for an UNSHARDED head (C_local == cfg->num_classes_total): everything b200f_arcface_fwd
 * computes, and -- out of the last block of the statistics reduction, no extra launch on the tcgen05 engine -- what
 * b200f_arcface_loss and b200f_arcface_hook_scale(upstream = NULL) would: lse [B], loss [1], pq_norm2 [1], out4 [4].
 * A backward whose upstream gradient is the constant 1 (loss.backward()) can take out4 as its grad_scale directly. */
int b200f_arcface_fwd_loss(const void* x, const void* w, int dtype,
                           const float* inv_nx, const float* inv_nw, const int64_t* label,
                           int64_t B, int64_t C_local, int64_t class_offset, int D,
                           const b200f_head_cfg* cfg, const b200f_hook_cfg* hook,
                           float* row_stats, float* row_best, int64_t* row_argmax,
                           float* cos_minmax, int32_t* nan_flag,
                           float* lse, float* loss, float* pq_norm2, float* out4,
                           void* workspace, size_t workspace_bytes, void* stream);

/* K1 + K2 (+ K2b) from the RAW rows in one call, tcgen05 engine only (F.normalize of input and weight, src/face_models.py:
 * 351-352, then everything b200f_arcface_fwd computes).  x_raw [B,D] (bf16 / fp32; NULL when x_f16n was already made, e.g.
 * by b200f_tail_fwd) -> x_f16n [B,D] fp16 rows * cfg->operand_scale + inv_nx [B]; w_raw [C_local,D] (bf16 / fp32) ->
 * w_f16n [C_local,D] + inv_nw [C_local]: all four are OUTPUTS, to be handed to b200f_arcface_bwd.  With D = 512 and 32-byte
 * aligned rows the class weights are normalised INSIDE K2 (two warps of every CTA produce the fp16 rows in the order the
 * TMA producers consume them, handing over through per-128-row counters in the workspace), so the 2 x C x D x 2 bytes of
 * the stand-alone K1 pass move under K2's MMAs; other shapes run K1 as its own pass first (tunable "k2_prep" = 0 forces
 * that).  hook NULL: statistics only (class shards: all-reduce row_stats, then b200f_arcface_loss_hook); non-NULL
 * (unsharded head): lse / loss / pq_norm2 / out4 as b200f_arcface_fwd_loss. */
int b200f_arcface_fwd_raw(const void* x_raw_or_null, int x_dtype, void* x_f16n, float* inv_nx,
                          const void* w_raw, int w_dtype, void* w_f16n, float* inv_nw, float eps,
                          const int64_t* label, int64_t B, int64_t C_local, int64_t class_offset, int D,
                          const b200f_head_cfg* cfg, const b200f_hook_cfg* hook_or_null,
                          float* row_stats, float* row_best, int64_t* row_argmax,
                          float* cos_minmax, int32_t* nan_flag,
                          float* lse, float* loss, float* pq_norm2, float* out4,
                          void* workspace, size_t workspace_bytes, void* stream);

/* K2b -- after the (optional) cross-shard SUM of row_stats: lse[i] = s_eff + log(sumexp_i),
 * loss = mean_i[ lse_i - (1-eps) z_target_i - (eps/C) sum_z_i ]  (CrossEntropyLoss with
 * label smoothing, mean reduction), pq_norm2 = sum_ij (p_ij - q_ij)^2  (the Frobenius norm the
 * ArcFaceNet hook clips on, src/face_models.py:541). */
int b200f_arcface_loss(const float* row_stats, int64_t B, const b200f_head_cfg* cfg,
                       float* lse, float* loss, float* pq_norm2, void* stream);

/* b200f_arcface_loss and b200f_arcface_hook_scale(upstream = NULL, i.e. 1) in ONE launch: the class-sharded forward
 * calls it after the all-reduce of row_stats. */
int b200f_arcface_loss_hook(const float* row_stats, int64_t B, const b200f_head_cfg* cfg, const b200f_hook_cfg* hook,
                            float* lse, float* loss, float* pq_norm2, float* out4, void* stream);

/* Hook scalar (src/face_models.py:538-567), on the device, no host sync:
 *   n = |upstream| * s_eff / B * sqrt(pq_norm2);  kappa = thr/(n+1e-8) if n > thr else 1
 *   with thr from (max_grad_norm, phase, epoch) and the n > 3 rule.
 * out[0] = grad_scale = upstream * kappa * s_eff / B   (what K3 multiplies (p-q) by)
 * out[1] = n (last_grad_norm), out[2] = kappa, out[3] = g_scale (power of two with |grad_scale| * g_scale in
 * (512, 1024]: range centring of the tcgen05 engine's fp16 logit-gradient buffer).  out has FOUR floats and
 * is what b200f_arcface_bwd takes as grad_scale.   hook_enabled = 0 -> kappa = 1. */
int b200f_arcface_hook_scale(const float* pq_norm2, const float* upstream, int64_t B,
                             float s_eff, int hook_enabled, float max_grad_norm, int phase,
                             int epoch, float* out4, void* stream);

/* K3 -- backward for this shard: recompute the logits tile by tile, form
 *   G_ij = grad_scale * (p_ij - q_ij) * dphi/dc * 1[lo <= cos <= hi]
 * and produce
 *   dxhat [B,D] fp32   = G . w_hat            (this shard's partial; SUM over shards)
 *   dw    [C_local,D] fp32 = normalise-backward of (G^T . x_hat) w.r.t. the raw weight rows
 * lse from K2b (global), grad_scale = out3[0] of b200f_arcface_hook_scale (device scalar).
 * dlogits_or_null: compatibility path of ArcMarginProduct.forward -> logits: when non-NULL
 * ([B, ld_dlogits] fp32, upstream dL/dlogits of this shard) G_ij = grad_scale * dlogits_ij * dphi/dc *
 * clamp-mask instead, and lse is ignored. */
int b200f_arcface_bwd(const void* x, const void* w, int dtype,
                      const float* inv_nx, const float* inv_nw, const int64_t* label,
                      const float* lse, const float* grad_scale,
                      const float* dlogits_or_null, int64_t ld_dlogits,
                      int64_t B, int64_t C_local, int64_t class_offset, int D,
                      const b200f_head_cfg* cfg,
                      float* dxhat, float* dw,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ||dW||^2 without a pass over dW (clip_grad_norm_, src/training.py:528-533, needs the gradient norm before the optimizer
 * step; reading the 205 MB of dW again for it costs ~35 us at cfg3).  After this call the calling THREAD's next backward
 * (b200f_arcface_bwd, b200f_arcface_bwd_dx, or the b200f_arcface_bwd_phase 1 + 2 pair) also writes sum(dW[c,d]^2) over its
 * class rows to out[0] (device memory, fp32): the dW epilogues add up what they store, in a fixed order (bitwise
 * reproducible), and one small kernel folds the per-warp partials.  One-shot: the request is consumed by that backward;
 * NULL cancels it.  Class shards: all-reduce (SUM) the word with the other ranks' before taking the square root. */
int b200f_head_request_dw_sqnorm(float* out_or_null);

/* b200f_arcface_bwd in two calls, for class shards that overlap the cross-rank all-reduce of dx_hat with the dW GEMM.
 * phase 1: per class chunk K3a -> K3c -> split reduction -> K3b, WITHOUT the last chunk's K3b: dxhat is complete when it
 * returns (in stream order) and the caller starts its all-reduce on ANOTHER stream; phase 2 (same arguments, same
 * workspace, nothing of this library in between on that workspace): the last chunk's K3b from the G^T and r partials phase
 * 1 left in the workspace -- dw is complete after it.  Same results as the single call, bit for bit (the kernels and
 * their inputs are the same; only the launch order differs).  CUDA-core engine: phase 1 does everything, phase 2 nothing. */
int b200f_arcface_bwd_phase(const void* x, const void* w, int dtype,
                            const float* inv_nx, const float* inv_nw, const int64_t* label,
                            const float* lse, const float* grad_scale,
                            int64_t B, int64_t C_local, int64_t class_offset, int D,
                            const b200f_head_cfg* cfg, float* dxhat, float* dw, int phase,
                            void* workspace, size_t workspace_bytes, void* stream);

/* The backward as its three GEMM stages, for a caller that runs two of them SIDE BY SIDE (single class chunk, batch <= 512,
 * tcgen05 engine: b200f_arcface_bwd_parts_ok says whether the shape qualifies).
 *   part 1: K3a -- the logit gradient G^T and the r partials, into the workspace;
 *   part 2: K3b -- dW from them, on at most max_clusters CTA pairs (0 = the whole chip);
 *   part 3: K3c + split reduction -- dxhat (and, when dx_or_null is given, dL/dx and its bf16 copy as in
 *           b200f_arcface_bwd_dx), on at most max_clusters CTA pairs.
 * Parts 2 and 3 only READ what part 1 left in the workspace and write disjoint outputs, so behind part 1 they may run on
 * two streams at once.  Why: at 512 x 100 k x 512 the dW stage is bound by the latency of its epilogue on every SM (87 us
 * on 74 pairs, 133 us on 44) and the dx stage by the tensor pipe (51 us on 74 pairs, 81 us on 30): one after the other they
 * take 145 us, side by side on 44 + 30 pairs less.  Results: dW as from the single call, bit for bit; dxhat
 * sums fewer, longer K splits (same inputs, last-bit differences).  A b200f_head_request_dw_sqnorm request is consumed
 * by part 2. */
int b200f_arcface_bwd_parts_ok(int64_t B, int64_t C_local, int D, int dtype);
int b200f_arcface_bwd_part(const void* x, const void* w, int dtype,
                           const float* inv_nx, const float* inv_nw, const int64_t* label,
                           const float* lse, const float* grad_scale,
                           int64_t B, int64_t C_local, int64_t class_offset, int D,
                           const b200f_head_cfg* cfg,
                           float* dxhat, float* dw,
                           const void* x_raw_or_null, int x_raw_dtype, float* dx_or_null, void* dx_bf16_or_null,
                           int part, int max_clusters,
                           void* workspace, size_t workspace_bytes, void* stream);

/* K3 for an unsharded head, finished: b200f_arcface_bwd, then dL/dx = normalise-backward of dxhat (below) written to
 * dx [B,D] fp32 and, optionally, as bf16 (the cast autograd applies for a bf16 input) -- on the tcgen05 engine inside
 * the split reduction of dxhat when the call is one class chunk (no extra launch, no round trip of dxhat).
 * x_raw_or_null: the RAW input rows (B200F_F32 / B200F_BF16) the projection should use, x_hat = x_raw * inv_nx;
 * NULL = the operand rows x.  With B200F_F16N operands pass the raw rows: their fp16 copy carries a 2^-12 rounding
 * that the projection of an embedding nearly parallel to its class centre amplifies (measured: dx error 9.8e-4 ->
 * 3e-4 of the bf16 bar at cfg3).  inv_nx is required. */
int b200f_arcface_bwd_dx(const void* x, const void* w, int dtype,
                         const float* inv_nx, const float* inv_nw, const int64_t* label,
                         const float* lse, const float* grad_scale,
                         const float* dlogits_or_null, int64_t ld_dlogits,
                         int64_t B, int64_t C_local, int64_t class_offset, int D,
                         const b200f_head_cfg* cfg,
                         float* dxhat, float* dw,
                         const void* x_raw_or_null, int x_raw_dtype, float* dx, void* dx_bf16_or_null,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Normalise-backward for rows: dv = inv_n * (dvhat - vhat * <vhat, dvhat>), vhat = v * inv_n
 * (autograd of F.normalize, src/face_models.py:351,525).  dv fp32 [rows, dim]; dv_bf16_or_null: the same rows
 * rounded to bf16.  With dtype B200F_F16N v holds v_hat * v_scale already (v_scale ignored otherwise).
 * dv may alias dvhat. */
int b200f_l2norm_bwd(const void* v, int dtype, float v_scale, const float* inv_norm, const float* dvhat,
                     int64_t rows, int dim, float* dv, void* dv_bf16_or_null, void* stream);

/* ---- the embedding tail in front of the head (src/face_models.py:515-525,584-590) -------------------
 *   z = embedding(features) [rows, dim]  ->  BatchNorm1d  ->  dropout (train)  ->  F.normalize  ->  head
 * b200f_bn_stats (train mode): batch mean and 1 / sqrt(biased var + eps) per column, and torch's running-stat update
 *   in place (momentum; the running variance takes the unbiased batch variance).  running_* may be NULL.
 * b200f_tail_fwd: ONE pass over z: y = dropout(BN(z)) with (mean, stat) = (batch mean, invstd) from b200f_bn_stats, or
 *   (running_mean, running_var) with stat_is_var = 1 (eval); mask_or_null = uint8 keep mask [rows, dim] (the caller owns
 *   the random generator), keep_scale = 1 / (1 - p).  Outputs, each optional except inv_norm: y fp32 (the rows the
 *   head's backward projects with), yhat16 = B200F_F16N operand rows y / |y| * out_scale (what b200f_l2norm_rows would
 *   emit for y: the head takes them as is), emb = y / |y| fp32 (get_embedding), inv_norm = 1 / max(|y|, norm_eps).
 * b200f_tail_bwd: gradient of dropout + BatchNorm: dy = dL/dy [rows, dim] fp32 (what b200f_arcface_bwd_dx returns as dx)
 *   -> dgamma, dbeta [dim] and dz [rows, dim] fp32.  batch_stats = 1: train-mode BatchNorm (statistics depend on z). */
int b200f_bn_stats(const void* z, int dtype, int64_t rows, int dim, float eps, float momentum, float* running_mean,
                   float* running_var, float* mean_out, float* invstd_out, void* stream);
int b200f_tail_fwd(const void* z, int dtype, int64_t rows, int dim, const float* gamma, const float* beta, const float* mean,
                   const float* stat, int stat_is_var, float bn_eps, const uint8_t* mask_or_null, float keep_scale, float norm_eps,
                   float out_scale, float* y_or_null, void* yhat16_or_null, float* emb_or_null, float* inv_norm, void* stream);
int b200f_tail_bwd(const float* dy, const uint8_t* mask_or_null, float keep_scale, const void* z, int dtype, const float* mean,
                   const float* stat, int stat_is_var, float bn_eps, const float* gamma, int batch_stats, int64_t rows, int dim,
                   float* dgamma, float* dbeta, float* dz, void* stream);

/* K5 -- optimizer step of the class-weight rows: torch.optim.AdamW (amsgrad when vmax != NULL) exactly as the
 * reference trains the head (src/training.py:343-348; default betas (0.9, 0.999), eps 1e-8), after the optional
 * clip_grad_norm_ of :528-533, whose coefficient the caller passes as the device scalar grad_scale (NULL = 1):
 *   g = dw * grad_scale;  w *= 1 - lr*wd;  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;  vmax = max(vmax, v)
 *   w -= lr / (1 - b1^step) * m / (sqrt(vmax or v) / sqrt(1 - b2^step) + eps)              (step counts from 1)
 * Fused with NEXT step's K1: when w_hat_out / inv_norm are given they receive, for the updated rows,
 * inv_norm[r] = 1 / max(||w[r,:]||, norm_eps) and w_hat_out = w * inv_norm * out_scale as B200F_F16N rows -- the
 * operands b200f_arcface_fwd / _bwd take -- so the head's next forward needs no b200f_l2norm_rows over W.
 * Hyper-parameters are doubles like torch's Python scalars (1 - beta2 etc. are formed in double, then rounded once).
 * All tensors fp32, [rows, dim] row-major, dim % 4 == 0, dim <= 1024, 16-byte aligned; w, m, v, vmax updated in place. */
int b200f_head_adamw(float* w, const float* dw, float* m, float* v, float* vmax, int64_t rows, int dim, double lr,
                     double beta1, double beta2, double eps, double weight_decay, int64_t step, const float* grad_scale,
                     void* w_hat_out, float out_scale, float norm_eps, float* inv_norm, void* stream);

/* K4 -- gallery match: for each query the k best rows of this gallery shard.
 *   metric L2EPS: score = || q - g + 1e-6 ||_2 ascending, accept = best <= thresh (src/app.py:50-64)
 *   metric COS  : score = <q,g> * q_inv[i] * g_inv[j] descending, accept = best >= thresh
 *                 (q_inv / g_inv from b200f_l2norm_rows; NULL = 1)
 * idx [Q,k] int64 holds GLOBAL indices (index_offset + local row), -1 for missing slots;
 * ties go to the lowest index (strict '<' at src/app.py:60). k <= 16. */
size_t b200f_gallery_workspace_bytes(int64_t Q, int64_t N_local, int D, int k, int dtype, int engine);
int b200f_gallery_topk(const void* q, const void* g, int dtype,
                       const float* q_inv, const float* g_inv,
                       int64_t Q, int64_t N_local, int64_t index_offset, int D,
                       int k, int metric, float thresh, int engine,
                       int64_t* idx, float* score, uint8_t* accept,
                       void* workspace, size_t workspace_bytes, void* stream);

/* K4 on the tensor cores (fp32 queries / gallery, D % 8 == 0, D <= 512, sm_100):
 *   b200f_gallery_prepare   builds the scan operand of a gallery ONCE: g16 [N,D] bf16 or fp16 (the rows for L2EPS, the
 *                           L2-normalised rows for COS) and bias [N+2] fp32 (|g|^2 - 2e-6 sum g per row for L2EPS, 0 for
 *                           COS; two extra slots: the largest row norm, used by the error bound, and the number of
 *                           rows the 16-bit operand cannot represent -- if non-zero every query goes to the exact engine).
 *   b200f_gallery_topk_tc   bf16 tcgen05 scan of g16 (queries resident in shared memory, gallery streamed once: HBM-bound
 *                           for Q <= ~250) keeping 8/16/32 candidates per query, exact fp32 re-rank of the candidates
 *                           against g with the reference formula, and a proof per query that nothing outside the
 *                           candidates can reach the top-k (|approx - exact| <= bf16 rounding bound).  Queries without a
 *                           proof are recomputed by the exact CUDA-core engine in the same stream; redo_count (optional
 *                           device int, caller zeroes it) counts them.  Results are those of b200f_gallery_topk. */
int    b200f_gallery_has_tc(int D);
#define B200F_OPERAND_BF16 0   /* any value range; proof margin 2 * 3.97e-3 |q||g|                              */
#define B200F_OPERAND_FP16 1   /* 8x tighter proof margin; for embeddings / class centres with |values| << 65504 */
int    b200f_gallery_prepare(const void* g, int dtype, int64_t N, int D, int metric, int operand_fmt, void* g16, float* bias,
                             void* stream);
size_t b200f_gallery_tc_workspace_bytes(int64_t Q, int64_t N_local, int D, int k);
int    b200f_gallery_topk_tc(const void* q, const void* g, const void* g16, const float* bias,
                             const float* q_inv, const float* g_inv,
                             int64_t Q, int64_t N_local, int64_t index_offset, int D,
                             int k, int metric, int operand_fmt, float thresh,
                             int64_t* idx, float* score, uint8_t* accept, int32_t* redo_count,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Merge P per-shard top-k lists (after an all-gather): idx_all/score_all [P,Q,k] ->
 * idx/score [Q,k], accept [Q]; lowest global index wins ties. */
int b200f_gallery_merge(const int64_t* idx_all, const float* score_all, int P, int64_t Q, int k,
                        int metric, float thresh,
                        int64_t* idx, float* score, uint8_t* accept, void* stream);

/* ---- tcgen05 engine: self-test and diagnostics (tests/test_gpu_umma.py, bench.py) ----------------
 * b200f_umma_selftest: out[M,N] fp32 = sum_k A(m,k) B(n,k) through the TMA + tcgen05 + TMEM GEMM core,
 *   for K-major (x_mn = 0: [rows,K] row-major) and MN-major (x_mn = 1: [K,rows] row-major) operands,
 *   fmt 0 = bf16 x bf16, 2 = fp16 x fp16 (1 = fp16 x bf16 faults: not a hardware format pair); with k_splits > 1 out is [k_splits, M, N] partial sums.
 *   Descriptor byte offsets < 0 select the defaults.
 * b200f_umma_timeout_flag: 1 if a bounded pipeline wait ever expired (synchronises; reset clears it).  The kernel
 *   that raised it has trapped, so the context reports a launch failure as well; the flag only says why.
 */
int b200f_umma_selftest(const void* a, const void* b, float* out, int M, int N, int K, int a_mn, int b_mn,
                        int fmt, int k_splits, int a_lbo, int a_sbo, int a_kstep, int b_lbo, int b_sbo,
                        int b_kstep, void* stream);
int b200f_umma_timeout_flag(int reset);
/* b200f_umma_xw_selftest: out[B,C] fp32 = x[B,D] . w[C,D]^T (fp16 operands, D % 8 == 0, D <= 512) through the
 *   X-stationary kernel that carries K2 / K3a, on single CTAs (pair = 1) or cta_group::2 CTA pairs (pair = 2).
 * b200f_umma_set_pair: process-wide choice between the two for the head calls (default 2); returns the old one. */
int b200f_umma_xw_selftest(const void* x, const void* w, float* out, int B, int C, int D, int pair, void* stream);
int b200f_umma_set_pair(int pair);
/* pipeline probe: rowsum[b] += sum_c (x . w^T)[b,c] with a do-nothing epilogue (rowsum zeroed by the caller) */
int b200f_umma_xw_probe(const void* x, const void* w, float* rowsum, int B, int C, int D, int pair, void* stream);
/* Tunables for tests and bench sweeps: "pair" (1 | 2), "g_chunk_mb" (budget in MB of the fp16 logit-gradient buffer
 * per class chunk of the backward; default 112), "pdl" (0 | 1), "k3b_reverse" (0 | 1),
 * "xw_prefetch" (tiles of the streamed operand pulled into L2 ahead of the TMA ring; default 0: measured +-0), "stage_events" (0 | 1),
 * "k2_groups" / "epi_groups" / "k3b_groups" (epilogue geometry of K2 / K3a / K3b: 1 = one group of 8 warps on 32-column slices,
 * 2 = two groups on alternating tiles, 16-column slices; "epi_groups" also takes 4 = one group of 16 warps on column quarters of
 * every tile; defaults 1 / 1 / 1), "k3b_tma_store" (0 | 1: dW through shared-memory staging and TMA tensor stores; default 0, it
 * measured slower), "early" (0 | 1: K3a / K3c start before their predecessor grid has drained), "l2_hints" (bit mask of
 * L2 cache-policy hints, default 6: G^T stores evict_last, dW stores evict_first),
 * "dw_n_fastest" (0 | 1, default 1: the streamed dW GEMM at batch > 512 runs the feature tiles of a class block side by side, so
 * G^T comes from HBM once), "k3c_follow" (0 | 1, default 0: the dx part beside the dW part walks the class rows in the dW kernel's
 * order; measured within noise), "k3a_reverse" (0 | 1, default 0: K3a walks its chunks last tile first; within noise),
 * "gt_blocked" (0 | 1, default 1: at batch > 512 the logit gradient is kept in [32 classes x 64 batch rows] blocks and read back
 * through rank-4 tensor maps; bit-identical),
 * "k3a_tma_store" (0 | 1, default 0: G^T through shared-memory staging and TMA tensor stores; bit-identical, measured slower),
 * "stream_k" (0 | 1, default 0: stream-K of the dx GEMM where split-K leaves clusters idle; measured slower, see umma_head.cu).
 * Returns the previous value, -1 for an unknown name.  Every setting computes the same results ("g_chunk_mb" changes the
 * workspace size: query b200f_head_workspace_bytes again after changing it).  The measurement probes that skip memory
 * traffic ("k3a_ablate" / "k3b_ablate", WRONG gradients) exist only in -DB200F_PROBES builds made by tools/: the
 * shipped library answers -1 for them. */
int b200f_set_tunable(const char* name, int value);
/* With the tunable "stage_events" = 1 the head calls record a CUDA event pair around each of their GEMM kernels on
 * the caller's stream (eager launches only, never under graph capture).  b200f_stage_ms returns the duration in ms
 * of the "k2" (fused forward), "k3a" (logit gradient), "k3b" (dW) or "k3c" (dX) kernels of the calling thread's last
 * head call, summed over its class chunks; it waits for those kernels.  Measurement aid for bench.py's roofline line. */
int b200f_stage_ms(const char* name, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* B200FACE_H_ */
