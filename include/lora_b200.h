/* lora_b200 C-ABI: the drop-in boundary for the LoRA hot path of cloneofsimo/lora on B200 (sm_100a).
 *
 * Plain C: raw device pointers, sizes, dtype enums and a cudaStream_t passed as void*. No torch
 * or ATen types cross this boundary. Every entry point enqueues work on `stream` and returns an
 * int status immediately (0 = LB_OK, negative = refused, nothing was launched); kernels never
 * synchronise the host. The reference has no FFI of its own (it is pure Python on top of torch
 * eager); each entry point below names the reference code it replaces, and INTEGRATION.md shows
 * the ctypes stub that binds it under the reference's own module classes.
 *
 * Conventions
 *   - "16-bit" operands are bf16 (LB_BF16) or fp16 (LB_F16), chosen per call by in_dtype.
 *   - LoRA master factors are fp32:   down A[r,K] row-major,   up B[N,r] row-major
 *     (reference: lora_down.weight / lora_up.weight, lora_diffusion/lora.py:44-46).
 *   - "down16" operands are 16-bit copies zero-padded to 16 rows: [16,K] row-major
 *     (built by lb_cast_rows_pad16). Rank r <= 16 in this version.
 *   - T buffers are fp32 [M,16] row-major, columns >= r are zero.
 *   - All base pointers 16-byte aligned; K % 8 == 0; N % 8 == 0 (N % 4 for fp32 outputs).
 */
#ifndef LORA_B200_H
#define LORA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

enum lb_status {
  LB_OK = 0,
  LB_ERR_SHAPE = -1, /* unsupported / misaligned extent */
  LB_ERR_RANK = -2,  /* r outside [1,16] */
  LB_ERR_DTYPE = -3,
  LB_ERR_ALIGN = -4, /* pointer not 16-byte aligned */
  LB_ERR_TMAP = -5,  /* cuTensorMapEncodeTiled unavailable or failed */
  LB_ERR_CUDA = -6   /* launch error (cudaGetLastError) */
};

enum lb_dtype { LB_BF16 = 0, LB_F16 = 1, LB_F32 = 2 };
/* OR-ed into `in_dtype` of the fused linear entries: the frozen weight argument (W, or Wt of the dX
 * entry) is NOT row-major [N, K] but the 64 x 64-block layout written by lb_tile_weight -- block
 * (n64, kb) occupies rows [(n64 * ceil(K/64) + kb) * 64, +64) of a [rows, 64] 16-bit tensor, zero padded
 * -- so that every TMA box of the weight stream is one contiguous 8 KB run of HBM (a row-major box is
 * 64-192 rows of 128 bytes a whole K apart). Optional: measured on B200 against row-major on every
 * SD1.5 site shape it is neither faster nor slower (profiles/r2h_site_table_tiled_vs_rowmajor.md). */
#define LB_W_TILED 0x100

/* ABI version of this header (bumped on any signature change). */
int lb_abi_version(void);

/* Fused frozen linear + LoRA forward (tcgen05/TMEM/TMA), one launch:
 *     Y[M,N] = X[M,K] . W[N,K]^T (+ bias[N]) + ((X . down16^T) * (scale * diag)) . up^T
 * up element (n, j) is read from the fp32 master at up[n*up_rs + j*up_cs]; diag is the optional
 * selector diagonal [r] (NULL = identity); T_out (NULL = skip) receives X . down16^T (unscaled);
 * T_in (normally NULL): when given, the rank-r activations are read from this fp32 [M,16] buffer
 * instead of being computed (dropout path: T = (mask o gY) . B comes from lb_lora_dropout_dt).
 * Replaces LoraInjectedLinear.forward, lora_diffusion/lora.py:53-58 (F.linear + lora_down +
 * selector + lora_up + scale + add), dropout handled by the caller.
 *
 * The SAME entry point is the backward-dX kernel: with W := W^T[K,N] (pre-transposed frozen
 * weight), down16 := B^T padded [16,N], up := A read transposed (up_rs = 1, up_cs = K) it returns
 *     dX[M,K] = gY . W + ((gY . B) * (scale * diag)) . A       and  T_out = gY . B.
 */
int lb_lora_linear_fwd(const void* X, const void* W, const float* bias, const void* down16,
                       const float* up, long long up_rs, long long up_cs, const float* diag,
                       float scale, void* Y, float* T_out, const float* T_in, int M, int K, int N,
                       int r, int in_dtype, int out_dtype, void* stream);

/* lb_lora_linear_fwd with an ACTIVE nn.Dropout(p) on the LoRA branch (lora.py:45,56: module in
 * training mode, p > 0 -- the state of every site built by inject_trainable_lora_extended and the
 * monkeypatch_* loaders), still ONE launch:
 *     Y = X . W^T (+ bias) + keep(m,n)/(1-p) * (((X . down16^T) * (scale * diag)) . up^T)[m,n]
 * The LoRA product keeps its own TMEM columns and the keep-mask (csrc/dropmask.cuh: a counter hash
 * of (*seed_dev, m*N + n); seed_dev points to one uint64 in DEVICE memory so that a captured graph
 * draws a new mask per replay) is applied while the tile is drained. Backward recomputes the same
 * mask: lb_lora_dropout_dt, lb_lora_wgrad_masked / lb_lora_wgrad_pair(drop_p). */
int lb_lora_linear_fwd_dropout(const void* X, const void* W, const float* bias, const void* down16,
                               const float* up, long long up_rs, long long up_cs, const float* diag,
                               float scale, void* Y, float* T_out, int M, int K, int N, int r,
                               int in_dtype, int out_dtype, float drop_p, const void* seed_dev,
                               void* stream);

/* Input gradient of lb_lora_linear_fwd_dropout in ONE launch:
 *     dX[M, K_in] = gY . W + (((mask o gY) . B) * (scale * diag / (1-p))) . A ,  T_out = (mask o gY) . B
 * with gY [M, N_out], Wt = W^T [K_in, N_out] (pre-transposed frozen weight), upT16 = B^T padded [16, N_out],
 * `down` = A read transposed (element (k, j) at down[k*down_rs + j*down_cs]: down_rs = 1, down_cs = K_in).
 * The mask (same counter hash as the forward, element index m*N_out + n) is applied to a shared-memory
 * copy of every gY tile by the otherwise idle epilogue warps and feeds only the rank-r MMA -- no
 * separate pass over gY (round 1: lb_lora_dropout_dt + T_in). dA then is lb_lora_wgrad(X, T_out) with
 * scale / (1-p); dB the masked reduction. */
int lb_lora_linear_dx_dropout(const void* gY, const void* Wt, const void* upT16, const float* down,
                              long long down_rs, long long down_cs, const float* diag, float scale,
                              void* dX, float* T_out, int M, int N_out, int K_in, int r, int in_dtype,
                              int out_dtype, float drop_p, const void* seed_dev, void* stream);

/* Up to 4 independent lb_lora_linear_fwd problems of the same operand/output dtype in ONE launch
 * (sites that share an input -- q/k/v of a self-attention, k/v of a cross-attention, CLIP's k/v/q --
 * each under-fill 148 SMs on their own). All array arguments are HOST arrays of length n; per-problem
 * semantics are exactly lb_lora_linear_fwd's (T_in unsupported). */
int lb_lora_linear_fwd_grouped(int n, const void* const* X, const void* const* W,
                               const float* const* bias, const void* const* down16,
                               const float* const* up, const long long* up_rs, const long long* up_cs,
                               const float* const* diag, const float* scale, void* const* Y,
                               float* const* T_out, const int* M, const int* K, const int* N,
                               const int* r, int in_dtype, int out_dtype, void* stream);

/* Benchmark/profiling knob: tile schedule of lb_lora_linear_fwd.
 *   mode = schedule + 4 * block_n + 16 * split
 *   schedule: 0 = auto (default), 1 = one tile per CTA, 2 = persistent CTAs with double-buffered
 *             TMEM accumulators, 3 = EXPERIMENTAL cluster split-K (csrc/fused_splitk.cuh; never
 *             chosen by `auto`)
 *   block_n:  0 = auto, 1 = 64, 2 = 128
 *   split:    (schedule 3 only) 0 = auto (4), 1..3 = 2..4 CTAs of one cluster share a tile's K loop */
int lb_debug_set_linear_mode(int mode);
/* Profiling knob: device buffer (16 x uint64) receiving %globaltimer phase stamps of CTA (0,0) of
 * subsequent lb_lora_linear_fwd launches; NULL switches it off. */
int lb_debug_set_stamp_buffer(void* dev_buf);
/* Programmatic dependent launch for the fused kernels (cudaLaunchAttributeProgrammaticStream-
 * Serialization): the prologue of a launch (barrier init, TMEM allocation, descriptor prefetch)
 * overlaps the tail of the previous kernel in the stream. 0 = off, 1 = on; initial value from the
 * environment variable LB_PDL. */
int lb_debug_set_pdl(int on);

/* Skinny weight-gradient reduction (streams S once, fp32 atomics into out):
 *     out[j*out_js + c*out_cs] += scale * diag[j] * sum_m V[m,j] * S[m,c]     j < r, c < C
 * S may also be fp32 (in_dtype = LB_F32; the fp32-faithful mode keeps X and gY in fp32).
 * dA[r,K]: S = X[M,K],  V = gY.B (T_out of the dX call), out_js = K, out_cs = 1
 * dB[N,r]: S = gY[M,N], V = X.A^T (T_out of the forward), out_js = 1, out_cs = r
 * Replaces the autograd-generated dA/dB GEMMs of lora.py:53-58 (W frozen: no dW).
 */
int lb_lora_wgrad(const void* S, const float* V, const float* diag, float scale, float* out,
                  long long out_js, long long out_cs, int M, int C, int r, int in_dtype,
                  void* stream);

/* Both factor gradients of one linear site in ONE launch (the two reductions are independent):
 *   dA[j*dA_js + k*dA_cs] += scale*diag[j] * sum_m dTs[m,j] * X[m,k]      k < K
 *   dB[j*dB_js + n*dB_cs] += scale*diag[j] * sum_m T[m,j]  * gY[m,n]      n < N
 * drop_p > 0: gY is masked like in lb_lora_wgrad_masked (mask index m*N + n). */
int lb_lora_wgrad_pair(const void* X, const float* dTs, float* dA, long long dA_js, long long dA_cs,
                       int K, const void* gY, const float* T, float* dB, long long dB_js,
                       long long dB_cs, int N, const float* diag, float scale, int M, int r,
                       float drop_p, const void* seed_dev, int in_dtype, void* stream);

/* One reduction  out[j*out_js + c*out_cs] += scale*diag[j] * sum_{m<M} V[m,j] * S[m,c]  (see
 * lb_lora_wgrad; drop_p > 0: S masked like lb_lora_wgrad_masked with the uint64 at seed_dev). */
typedef struct lb_wgrad_problem {
  const void* S;        /* [M, C] rows of in_dtype                         */
  const float* V;       /* [M, 16] fp32                                    */
  float* out;
  long long out_js, out_cs;
  int M, C, r;
  float scale;
  const float* diag;    /* selector diagonal [r] or NULL                   */
  float drop_p;
  const void* seed_dev; /* device uint64 (drop_p > 0) or NULL              */
  /* conv_H > 0: the lora_down weight-gradient of a conv site (lb_lora_wgrad_conv as one problem):
   * S = NHWC input rows [M/(H*W), H, W, C], every tap (ty, tx) of the kh x kw filter reads the pixel
   * shifted by (ty - pad_h, tx - pad_w), out = dA[r, C, kh, kw] flat with out_js = C*kh*kw,
   * out_cs = kh*kw (+ tap). 0: plain rows. */
  int conv_H, conv_W, kh, kw, pad_h, pad_w;
} lb_wgrad_problem;

/* n independent reductions (HOST array) in ceil(n/24) launches. The step engine queues the dA / dB
 * of every linear site during backward and flushes them here once: the factor gradients feed only
 * the optimizer (autograd's dA/dB GEMMs of lora.py:53-58 for all sites at once). */
int lb_lora_wgrad_batch(const lb_wgrad_problem* probs, int n, int in_dtype, void* stream);

/* lb_lora_wgrad_pair for up to 4 sites that share X (a grouped family), one launch; HOST arrays. */
int lb_lora_wgrad_multi(int n, const void* X, const float* const* dTs, float* const* dA,
                        const void* const* gY, const float* const* T, float* const* dB, const int* N,
                        const float* const* diag, const float* scale, const int* r, int M, int K,
                        int in_dtype, void* stream);

/* Conv tap of the same reduction: rows of S are the pixels of NHWC images [M/(H*W), H, W, C]; S is
 * read at pixel (h+dy, w+dx), zero outside the image (the convolution's zero padding):
 *     out[...] += scale * diag[j] * sum_p V[p,j] * S[p shifted by (dy,dx), c]
 * dA[r,Cin,kh,kw] of a LoraInjectedConv2d: one call per tap t = (ty,tx) with dy = ty - pad_h,
 * dx = tx - pad_w, out = dA + t, out_js = Cin*kh*kw, out_cs = kh*kw, V = gY.B (per pixel).
 * H = W = 0 degenerates to lb_lora_wgrad. Replaces the lora_down wgrad of lora.py:130-135. */
int lb_lora_wgrad_shift(const void* S, const float* V, const float* diag, float scale, float* out,
                        long long out_js, long long out_cs, int M, int C, int r, int H, int W,
                        int dy, int dx, int in_dtype, void* stream);

/* ---- dropout on the LoRA branch (nn.Dropout inside the operator, lora.py:45,56,115,133; active
 * only when module.training and p > 0). keep(m,n) is a counter-based hash of (*seed_dev, m*N+n),
 * recomputed by every kernel; seed_dev is a DEVICE uint64 so CUDA-graph replays draw new masks.
 * Forward:  Y = lb_lora_linear_fwd(scale = 0)  ->  lb_lora_up_dropout adds the masked branch.
 * Backward: dTs = lb_lora_dropout_dt(gY) -> lb_lora_linear_fwd(T_in = dTs) for dX ->
 *           lb_lora_wgrad(X, dTs) for dA, lb_lora_wgrad_masked(gY, T) for dB. */
/* Y[m,n] += scale/(1-p) * keep(m,n) * sum_j T[m,j]*diag[j]*up[n*up_rs + j*up_cs]   (in place) */
int lb_lora_up_dropout(void* Y, int y_dtype, const float* T, const float* up, long long up_rs,
                       long long up_cs, const float* diag, float scale, float drop_p,
                       const void* seed_dev, int M, int N, int r, void* stream);
/* dTs[m,j] = sum_n keep(m,n)/(1-p) * gY[m,n] * up[n*up_rs + j*up_cs]    (fp32 [M,16]) */
int lb_lora_dropout_dt(const void* gY, int in_dtype, const float* up, long long up_rs,
                       long long up_cs, float drop_p, const void* seed_dev, float* dTs, int M, int N,
                       int r, void* stream);
/* lb_lora_wgrad with S := keep o S / (1-p)  (S = gY [M,C], mask index m*C + c) */
int lb_lora_wgrad_masked(const void* S, const float* V, const float* diag, float scale, float* out,
                         long long out_js, long long out_cs, int M, int C, int r, float drop_p,
                         const void* seed_dev, int in_dtype, void* stream);

/* dA[r,Cin,kh,kw] of a LoraInjectedConv2d in ONE launch (all taps; equals kh*kw calls of
 * lb_lora_wgrad_shift): S = X NHWC rows [M = n_img*H*W, C = Cin], V = gY.B per pixel [M,16]. */
int lb_lora_wgrad_conv(const void* S, const float* V, const float* diag, float scale, float* out, int M,
                       int C, int r, int H, int W, int kh, int kw, int pad_h, int pad_w, int in_dtype,
                       void* stream);

/* Fused frozen Conv2d + LoRA (NHWC implicit GEMM on tcgen05; stride 1, dilation 1, groups 1,
 * 1x1 or 3x3 "same" padding -- the ResnetBlock2D sites of SD1.5):
 *     Y[n,h,w,:] = sum_tap X[n,h+ty-pad,w+tx-pad,:] . W[:, tap, :]^T (+ bias)
 *                  + ((sum_tap X[..] . down16[:, tap, :]^T) * (scale*diag)) . up^T
 * X: NHWC 16-bit [n_img,H,W,Cin]; W: 16-bit [Cout, kh*kw*Cin] with K ordered tap-major,
 * channel-minor (built once from the frozen [Cout,Cin,kh,kw] weight); down16: [16, kh*kw*Cin] in
 * the same K order; up element (n, j) at up[n*up_rs + j*up_cs]; Y: NHWC [n_img,H,W,Cout];
 * T_out: fp32 [n_img*H*W, 16] = the rank-r activations (unscaled).
 * Replaces LoraInjectedConv2d.forward, lora_diffusion/lora.py:130-135.
 *
 * per_tap_T = 1 is the input-gradient form (dX = conv_T(gY, W) + conv_T((gY.B)*s*d, A)): call it
 * with X := gY [n_img,H,W,Cout'], W := the flipped+transposed frozen weight [Cin', taps*Cout'],
 * down16 := B^T padded [16, Cout'] (restarts every tap), up := A read flipped:
 * element (c, tap g, j) at up[c*up_rs + j*up_cs + g*up_gs] (up_gs = -1 from the last tap).
 * T_out then receives gY.B of the unshifted tap. */
int lb_lora_conv2d_fwd(const void* X, const void* W, const float* bias, const void* down16,
                       const float* up, long long up_rs, long long up_cs, long long up_gs,
                       const float* diag, float scale, void* Y, float* T_out, const float* T_in,
                       int n_img, int H, int Wd, int Cin, int Cout, int kh, int kw, int pad_h,
                       int pad_w, int r, int per_tap_T, int in_dtype, int out_dtype, void* stream);

/* lb_lora_conv2d_fwd (forward direction) with an active nn.Dropout(p) on the LoRA branch
 * (lora.py:115,133; every conv site of inject_trainable_lora_extended): one launch, mask applied in
 * the drain exactly as in lb_lora_linear_fwd_dropout; mask index = pixel * Cout + n with
 * pixel = (img*H + h)*W + w. */
int lb_lora_conv2d_fwd_dropout(const void* X, const void* W, const float* bias, const void* down16,
                               const float* up, long long up_rs, long long up_cs, const float* diag,
                               float scale, void* Y, float* T_out, int n_img, int H, int Wd, int Cin,
                               int Cout, int kh, int kw, int pad_h, int pad_w, int r, int in_dtype,
                               int out_dtype, float drop_p, const void* seed_dev, void* stream);

/* Input gradient of lb_lora_conv2d_fwd_dropout in one launch (the conv analogue of
 * lb_lora_linear_dx_dropout): gY NHWC [n_img, H, Wd, Cout], Wb = flipped/transposed frozen weight
 * [Cin, kh*kw*Cout], upT16 = B^T padded [16, Cout], `down` = A read flipped (see lb_lora_conv2d_fwd,
 * per_tap_T = 1), pad_* = kh-1-pad of the forward. T_out = (mask o gY) . B at the unshifted tap. */
int lb_lora_conv2d_dx_dropout(const void* gY, const void* Wb, const void* upT16, const float* down,
                              long long down_rs, long long down_cs, long long down_gs, const float* diag,
                              float scale, void* dX, float* T_out, int n_img, int H, int Wd, int Cout,
                              int Cin, int kh, int kw, int pad_h, int pad_w, int r, int in_dtype,
                              int out_dtype, float drop_p, const void* seed_dev, void* stream);

/* Frozen conv-weight preparation: src [Cout,Cin,kh,kw] (LB_F32/LB_BF16/LB_F16) ->
 *   dst16  [Cout, kh*kw*Cin]  (tap-major K; forward operand)              and/or
 *   dstT16 [Cin, kh*kw*Cout]  with taps flipped (input-gradient operand). Either may be NULL. */
int lb_cast_conv_weight(const void* src, int src_dtype, void* dst16, void* dstT16, int Cout,
                        int Cin, int kh, int kw, int out_dtype, void* stream);

/* dst16[j, c] = (j < r) ? src[j*src_rs + c*src_cs] : 0   for j < 16, c < C  (16-bit, [16,C]).
 * A[r,K] -> down16:  src_rs = K, src_cs = 1.    B[N,r] -> B^T padded: src_rs = 1, src_cs = r. */
int lb_cast_rows_pad16(const float* src, long long src_rs, long long src_cs, void* dst16, int r,
                       int C, int out_dtype, void* stream);

/* Pivotal-tuning phase 1 (textual inversion) update of the n trained token rows, one launch:
 * AdamW on rows [n,D] (fp32 masters; grad is zeroed), t = ++(*step_dev), then, if clip_decay,
 *   w <- w/||w|| * (||w|| + min(1, 100*lr) * (target_norm - ||w||))        (target_norm = 0.4)
 * and the row is written into table[token_ids[j], :] (LB_F32/LB_BF16/LB_F16). Equivalent to the
 * reference's AdamW over the WHOLE embedding table followed by restoring every untouched row,
 * lora_diffusion/cli_lora_pti.py:448-479. */
int lb_ti_embed_step(float* rows, float* grad, float* m, float* v, const long long* token_ids, void* table,
                     int table_dtype, int n_rows, int D, const float* lr_dev, float beta1, float beta2,
                     float eps, float weight_decay, int* step_dev, int clip_decay, float target_norm,
                     void* stream);

/* Fold a LoRA into its frozen weight: out[n,k] = W[n,k] + alpha * sum_j up[n,j]*down[j,k]
 * (W, out: LB_F32/LB_BF16/LB_F16 [N,K]; conv weights flattened to [Cout, Cin*kh*kw]; out may alias W).
 * Replaces the up@down GEMM + add of collapse_lora, lora_diffusion/lora.py:635-669. */
int lb_lora_merge(const void* W, int w_dtype, const float* up, const float* down, float alpha, void* out,
                  int N, int K, int r, void* stream);

/* fp32-faithful mode ("split-bf16"): dst16 [R, 3C] bf16 = three bf16 terms of src [R,C] fp32 side
 * by side along K; pattern 0 (activations) [hi|lo|hi], pattern 1 (weights, factors) [hi|hi|lo], so
 * an ordinary K-major bf16 GEMM over 3C columns evaluates hi*hi + lo*hi + hi*lo (error 2^-16).
 * The fused kernels are then called with K := 3K; nothing else changes. */
int lb_split_bf16x3(const float* src, long long src_rs, void* dst16, int R, int C, int pattern,
                    void* stream);

/* Frozen-weight preparation for LB_W_TILED: the logical [N, K] operand, element (n, k) read from
 * src[n*src_rs + k*src_cs] (W itself: src_rs = K, src_cs = 1; the dX operand W^T of a [N_out, K_in]
 * weight: N = K_in, K = N_out, src_rs = 1, src_cs = K_in), cast to 16 bits and written as 64 x 64 blocks
 * (see LB_W_TILED), zero padded. dst16 holds lb_tiled_weight_elems(N, K) elements, 16-byte aligned. */
long long lb_tiled_weight_elems(int N, int K);
int lb_tile_weight(const void* src, int src_dtype, long long src_rs, long long src_cs, int N, int K,
                   void* dst16, int out_dtype, void* stream);

/* Frozen-weight preparation: src [R,C] (LB_F32/LB_BF16/LB_F16) -> dst16 [R,C] and/or
 * dstT16 [C,R] (either may be NULL). One-time cost per frozen weight. */
int lb_cast_weight(const void* src, int src_dtype, void* dst16, void* dstT16, int R, int C,
                   int out_dtype, void* stream);

/* Global-norm clip + AdamW over one flat fp32 arena, two launches, no host sync.
 *   total = inv_world * ||g||_2 ; coef = min(1, max_norm / (total + 1e-6))  (max_norm <= 0: no clip)
 *   g' = g * inv_world * coef ; decoupled weight decay ; Adam moments ; bias correction with
 *   t = ++(*step_dev) ; g is zeroed for the next step ; *gnorm_out = total.
 * Elements [group_off[i], group_off[i+1]) use lr_dev[i]; group_off is a HOST array of
 * n_groups+1 offsets (n_groups <= 8), lr_dev a DEVICE array (so a captured graph sees scheduler
 * updates), step_dev a DEVICE int. partials: device scratch of >= 1024 floats.
 * Replaces clip_grad_norm_ + torch.optim.AdamW.step + zero_grad,
 * training_scripts/train_lora_dreambooth.py:878-888 and lora_diffusion/cli_lora_pti.py:606-609.
 */
int lb_adamw_clip_step(float* p, float* g, float* m, float* v, long long n,
                       const long long* group_off, int n_groups, const float* lr_dev, float beta1,
                       float beta2, float eps, float weight_decay, float max_norm, float inv_world,
                       int* step_dev, float* partials, float* gnorm_out, void* stream);

/* The whole optimizer step in ONE cooperative launch: lb_adamw_clip_step's two passes plus
 * lb_refresh_shadows (the re-cast of every LoRA factor into the 16-bit [16, C] operands of the
 * fused kernels; `table` = the rows lb_refresh_shadows walks, n_entries may be 0), separated by
 * grid barriers on `barrier2` (2 x uint32 of device memory, zero-initialised once by the caller).
 * train_lora_dreambooth.py:878-888 (clip_grad_norm_, optimizer.step, zero_grad). */
int lb_optim_step_fused(float* p, float* g, float* m, float* v, long long n,
                        const long long* group_off, int n_groups, const float* lr_dev, float beta1,
                        float beta2, float eps, float weight_decay, float max_norm, float inv_world,
                        int* step_dev, float* partials, float* gnorm_out, const long long* table,
                        int n_entries, int max_C, void* shadow16, int shadow_dtype,
                        unsigned int* barrier2, void* stream);

/* lb_optim_step_fused for `world` data-parallel replicas on one NVLink/NVSwitch node, with the
 * gradient all-reduce INSIDE the launch (no NCCL call; replaces DDP's bucketed all-reduce implied by
 * accelerator.prepare/backward, train_lora_dreambooth.py:744-757,877): after a flag barrier over
 * NVLink every rank sums the flat gradient buffers of all ranks by direct peer reads (fixed rank
 * order => identical sums everywhere), clips on the averaged gradient (DDP order), runs AdamW on its
 * replica, and zeroes its own gradients once every peer has finished reading them.
 * peer_g / peer_flags: HOST arrays of `world` device pointers in rank order (the entries of the
 * other ranks are CUDA-IPC mappings of their `g` / flag buffers; flags = 2*world uint32, zeroed
 * once); gsum: n floats of scratch; epoch_dev: one zero-initialised uint32. Must be launched by
 * every rank the same number of times. Waits are bounded (~20 s) and trap. */
int lb_optim_step_dp(float* p, float* g, float* gsum, float* m, float* v, long long n,
                     const long long* group_off, int n_groups, const float* lr_dev, float beta1,
                     float beta2, float eps, float weight_decay, float max_norm, int* step_dev,
                     float* partials, float* gnorm_out, const long long* table, int n_entries,
                     int max_C, void* shadow16, int shadow_dtype, unsigned int* barrier2,
                     const void* const* peer_g, void* const* peer_flags, int world, int rank,
                     unsigned int* epoch_dev, void* stream);

/* CUDA-IPC helpers for the peer mappings of lb_optim_step_dp. lb_ipc_export: `ptr` (anywhere inside
 * a cudaMalloc'ed block) -> 64-byte IPC handle of the block + byte offset of ptr inside it.
 * lb_ipc_open (in another process of the node, once per distinct handle): -> the block's base address
 * in this process; enables peer access from the current device to the exporter's. */
int lb_ipc_export(const void* ptr, void* handle64, long long* offset);
int lb_ipc_open(const void* handle64, void** base_out);

/* Batched 16-bit shadow refresh after an optimizer step: for every table entry e, j < 16, c < e.C
 *   dst16_base[e.dst_off + j*e.dst_rs + c] = (j < e.r) ? p[e.src_off + j*e.src_rs + c*e.src_cs] : 0
 * table: DEVICE array of n_entries x 7 long long {src_off, src_rs, src_cs, r, C, dst_off, dst_rs}.
 * (conv down factors use one entry per filter tap so that K is ordered tap-major, channel-minor) */
int lb_refresh_shadows(const float* p, const long long* table, int n_entries, int max_C,
                       void* dst16_base, int out_dtype, void* stream);

/* ---- batched truncated SVD for LoRA distillation (replaces the serial torch.linalg.svd loop of
 * lora_diffusion/cli_svd.py:24-92). Randomized range finder with 32 probe vectors on
 * dW[b] = Wt[b] - Wb[b] (N x K, row-major, LB_F32/LB_BF16/LB_F16). Wt/Wb are DEVICE arrays of `batch` device pointers to same-shape
 * matrices. Tall-skinny operands are fp32 row-major [batch, rows, 32]. See lora_b200/svd.py for
 * the driver (probe -> power iterations with CholeskyQR2 -> 32x32 Jacobi -> factors). */
/* SVD distillation of cli_svd.py:24-92 (`overwrite_base`) for ALL sites of a model in one call,
 * ragged over shapes (SURVEY.md 8b `lora_svd_truncated_batched`): for b < batch,
 *   dW = Wt[b] - Wb[b]  ([N[b], K[b]] row-major, w_dtype in {LB_F32, LB_BF16, LB_F16}; Wb NULL or
 *   Wb[b] ignored when the array is NULL: dW = Wt[b]) -> top-`rank` singular triplets by a
 *   randomized range finder (32 probes, `power_iters` power iterations = 2*power_iters+2 streaming
 *   passes over both weights, contraction on the tensor cores in 3-term split-bf16) ->
 *   out[off_b ...] = [ up = U_r diag(S_r) as [N, rank] | down = Vh_r as [rank, K] ] fp32, with
 *   off_b = rank * sum_{c<b} (N[c] + K[c]); sigma_out[b, 0..31] (descending; may be NULL);
 *   clamp_q > 0: both factors clamped to +-torch.quantile(cat(up, down), clamp_q) (exact order
 *   statistics, cli_svd.py:43-47), the threshold reported in hi_out[b] (may be NULL).
 * All array arguments are HOST arrays; workspace is device memory of at least
 * lb_svd_workspace_bytes(N, K, batch, w_dtype) bytes. K[b] must keep rows 16-byte aligned. */
long long lb_svd_workspace_bytes(const int* N, const int* K, int batch, int w_dtype);
int lb_svd_truncated_batched(const void* const* Wt, const void* const* Wb, const int* N, const int* K,
                             int batch, int w_dtype, int rank, int power_iters, float clamp_q,
                             unsigned long long seed, float* out, float* sigma_out, float* hi_out,
                             void* workspace, long long workspace_bytes, void* stream);

/* transpose = 0: out[b] (N x 32) = dW[b] . in[b] (K x 32);  1: out[b] (K x 32) = dW[b]^T . in[b] (N x 32) */
int lb_svd_mul(const void* const* Wt, const void* const* Wb, int w_dtype, const float* in, float* out,
               int N, int K, int batch, int transpose, void* stream);
/* G[b] (32x32) = Y[b]^T Y[b] */
int lb_svd_gram(const float* Y, float* G, int rows, int batch, void* stream);
/* G = R^T R (upper R, tiny ridge) -> Rinv[b] = R^-1 (32x32) */
int lb_svd_chol_inv(const float* G, float* Rinv, int batch, void* stream);
/* out[b] = Y[b] (rows x 32) . M[b] (32x32), first out_cols columns, optionally column-scaled
 * (scale_mode 1: * colscale[b][j], 2: / colscale[b][j]); out may alias Y; transposed = 1 writes
 * out[b][j*out_pitch + row]. */
int lb_svd_apply(const float* Y, const float* M, const float* colscale, int scale_mode, float* out,
                 int rows, int out_cols, long long out_pitch, int transposed,
                 long long out_batch_stride, int batch, void* stream);
/* symmetric 32x32 eigen-decomposition (cyclic Jacobi): V columns sorted by descending eigenvalue,
 * sigma = sqrt(max(eig, 0)) */
int lb_svd_jacobi(const float* G, float* V, float* sigma, int batch, int sweeps, void* stream);
/* standard-normal probes from a counter hash */
int lb_svd_randn(float* out, long long n, unsigned long long seed, void* stream);

/* Step prologue, one launch (cli_lora_pti.py:295-313, train_lora_dreambooth.py:822-840):
 *   out[b, h, w, c] = sqrt_acp[t_b] * latents[b, c, h, w] + sqrt_one_minus_acp[t_b] * noise[b, c, h, w]
 * latents / noise fp32 NCHW, timesteps int64 [B], out NHWC (a channels_last [B, C', H, W] tensor) of
 * out_dtype. With inpaint_mask [B,1,H,W] and masked_latents [B,C,H,W] (both or neither): C' = 2C + 1,
 * out = cat([noisy, mask, masked_latents], channel) -- the 9-channel inpainting input. */
int lb_step_prologue(const float* latents, const float* noise, const long long* timesteps,
                     const float* sqrt_acp, const float* sqrt_one_minus_acp, int n_timesteps,
                     const float* inpaint_mask, const float* masked_latents, void* out, int out_dtype,
                     int B, int C, int H, int W, void* stream);

/* Loss epilogue, one launch (cli_lora_pti.py:340-370, train_lora_dreambooth.py:855-875):
 *   loss = sum_b weights[b] * mean_{c,h,w} (mask[b,h,w] * (pred - target))^2      (weights NULL: 1/B)
 *   grad = d loss / d pred = 2 weights[b] mask^2 (pred - target) / (C H W)        (pred's dtype/layout)
 * pred element (b, c, pixel) at b*stride_b + c*stride_c + pixel*stride_p (NCHW or channels_last);
 * target fp32 NCHW; mask [B,1,H,W] fp32 (already normalised as in cli_lora_pti.py:356-364) or NULL.
 * partials64: >= 64 floats of scratch; counter: one zero-initialised uint32 (self-resetting). */
int lb_masked_mse_fwd_bwd(const void* pred, int pred_dtype, long long stride_b, long long stride_c,
                          long long stride_p, const float* target, const float* mask,
                          const float* weights, void* grad, float* loss, float* partials64,
                          unsigned int* counter, int B, int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LORA_B200_H */
