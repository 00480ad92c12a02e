/* octcube_b200 — C ABI of the B200 (sm_100a) kernels behind the OCTCube 3D-MAE pre-training step.
 *
 * The reference (ZucksLiu/OCTCubeM) has no FFI of its own: its operator surface is the Python nn.Module
 * `MaskedAutoencoderViT` (Pre-training/models_mae_joint_res_flash_attn.py:29-790) whose arithmetic is delegated to
 * ATen / cuDNN / cuBLASLt / the flash-attn pip package.  Every entry point below replaces one of those library
 * call sites (cited per function; `models:` = Pre-training/models_mae_joint_res_flash_attn.py, `vv:` =
 * Pre-training/custom_util/video_vit.py, `FA:` = flash_attn 2.8.3 python modules).  INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions:
 *  - plain pointers + sizes; every pointer is a DEVICE pointer on the current device unless stated otherwise;
 *  - the caller owns all memory (inputs, outputs, workspaces) and keeps it alive until `stream` passes the call;
 *  - nothing here allocates, frees, synchronises the device or touches another stream;
 *  - return 0 on success; <0 = argument/shape error (OCT_ERR_*), >0 = cudaError_t; message via oct_last_error();
 *  - `stream` is a cudaStream_t passed as void*;
 *  - dtype codes: OCT_F32 / OCT_BF16 for tensors that exist in both precisions; indices are int64 (bit-identical
 *    to what torch.argsort hands the reference's callers); masks are float {0,1}.
 */
#ifndef OCTCUBE_B200_H
#define OCTCUBE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* oct_stream_t;

enum { OCT_OK = 0, OCT_ERR_INVALID = -1, OCT_ERR_UNSUPPORTED = -2, OCT_ERR_WORKSPACE = -3 };
enum { OCT_F32 = 0, OCT_BF16 = 1, OCT_SIMT_BF16 = 2 /* attention only: CUDA-core math on bf16 tensors (validation aid) */ };

/* GEMM operand layouts.  All matrices are row-major with explicit leading dimensions (in elements).
 *   OCT_GEMM_NT : D[M,N] = A[M,K] · B[N,K]^T   (nn.Linear forward,  FA:mha.py:635,703  FA:mlp.py:48-50  models:511,595)
 *   OCT_GEMM_NN : D[M,N] = A[M,K] · B[K,N]     (its dgrad:  dX = dY · W)
 *   OCT_GEMM_TN : D[M,N] = A[K,M]^T · B[K,N]   (its wgrad:  dW = dY^T · X) */
enum { OCT_GEMM_NT = 0, OCT_GEMM_NN = 1, OCT_GEMM_TN = 2 };

/* GEMM epilogues (applied to the fp32 accumulator acc[m,n]):
 *   OCT_EPI_NONE       D = acc (+ beta·D when beta==1, fp32 D only)
 *   OCT_EPI_BIAS       D = acc + bias[n]
 *   OCT_EPI_BIAS_GELU  aux = acc + bias[n] (pre-activation, same dtype as D);  D = gelu_erf(round(aux))
 *   OCT_EPI_DGELU      D = acc · gelu_erf'(aux[m,n])          (fc2 dgrad fused with the GELU backward)
 * The bf16 tensor-core path (GELU epilogues need a bf16 D) evaluates erf-GELU through a fitted tanh form: |gelu error|
 * <= 3e-5 + 2.5e-4 |x| (tanh.approx), |derivative error| <= 1.2e-4 + 5e-4 — below the bf16 rounding of D; the fp32
 * entry points (oct_gelu_fwd / oct_gelu_bwd) use erfc to 1.5e-7. */
enum { OCT_EPI_NONE = 0, OCT_EPI_BIAS = 1, OCT_EPI_BIAS_GELU = 2, OCT_EPI_DGELU = 3 };

/* compute paths: OCT_F32 = fp32 CUDA-core kernels (the 1e-4 parity mode);
 *                OCT_BF16 = tcgen05 / TMEM / TMA tensor-core kernels (bf16 operands, fp32 accumulate). */

const char* oct_version(void);
const char* oct_last_error(void);
/* number of kernels this library has launched in this process so far (evidence for bench.py's gpu_launches) */
uint64_t oct_launch_count(void);
/* Programmatic dependent launch for the launches that follow (process-wide): 1 = the hot kernels are launched with the
 * programmatic-stream-serialization attribute (their prologues — barrier init, TMEM allocation, tensor-map prefetch — overlap
 * the tail of the previous kernel; every kernel executes griddepcontrol.wait before touching memory), 0 = off, -1 = follow the
 * OCT_PDL environment variable (default off).  The module turns it on for the FORWARD pass only: in the backward pass early-
 * resident CTAs of the dgrad chain would take the SMs the side-stream weight-gradient GEMMs fill. */
void oct_set_pdl(int mode);
/* sm_count / compute capability of the current device; returns OCT_ERR_UNSUPPORTED unless cc == 10.x */
int oct_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- random_masking (models:336-372; replaces torch.rand-argsort-argsort-gather-ones-gather) ----------------
 * Stable ascending sort of each noise row by (key, index) == torch.argsort on CUDA (radix sort).
 * ids_restore[b, ids_shuffle[b,r]] = r ; ids_keep[b, r] = ids_shuffle[b, r] for r < keep ; mask = (ids_restore >= keep).
 * noise [B,L] f32 (finite or +-inf, no NaN), ids_restore [B,L] i64, ids_keep [B,keep] i64, mask [B,L] f32. L <= 16384. */
int oct_mask_sort(const float* noise, int64_t B, int64_t L, int64_t keep, int64_t* ids_restore, int64_t* ids_keep,
                  float* mask, oct_stream_t stream);

/* ---- patchify (models:289-314) and the kept-token variant --------------------------------------------------
 * imgs [B,1,T,H,W] f32 -> out [B, L, u*p*p] (ids_keep == NULL) or out [B*keep, u*p*p] rows ids_keep[b,i];
 * per-patch order (kt,kh,kw) == Conv3d weight order (vv:69-72), token order (t',h,w). out dtype f32|bf16.
 * frame_idx (nullable, [T_sel] i64 on device): temporal index_select of models:630-640 (T_sel frames are used). */
int oct_patchify(const float* imgs, void* out, int out_dtype, const int64_t* ids_keep, const int64_t* frame_idx,
                 int64_t B, int64_t T, int64_t H, int64_t W, int64_t p, int64_t u, int64_t T_sel, int64_t keep,
                 oct_stream_t stream);

/* ---- PatchEmbed.forward (vv:74-83; replaces cuDNN Conv3d + permute copy) -----------------------------------
 * im2col-free: the volume is addressed through a 5-D TMA box (kw,kh,w,h,frame), the contraction runs on tcgen05
 * (kind::tf32, fp32 operands straight from HBM, fp32 accumulate in TMEM), bias added in the epilogue, output written
 * token-major [B, T'*h*w, E] (the einsum 'ncts->ntsc' of vv:82 costs nothing).
 * imgs [B,1,T,H,W] f32; weight [E, u*p*p] f32; bias [E] f32; out bf16|f32.  Requires p == 16, W % 16 == 0. */
int oct_patch_embed_fwd(const float* imgs, const float* weight, const float* bias, void* out, int out_dtype, int64_t B,
                        int64_t T, int64_t H, int64_t W, int64_t p, int64_t u, int64_t E, oct_stream_t stream);

/* ---- kept-token gather + cls + separable pos-embed add (models:406-478) ------------------------------------
 * out[b,0,:] = cls_row ; out[b,1+i,:] = x[b, ids_keep[b,i], :] + pos_sp[ids % G] + pos_tmp[ids / G]   (fp32 out)
 * x [B,L,C] f32|bf16; pos_sp [G,C]; pos_tmp [L/G, C] or NULL (T'==1, models:437-440); cls_row [C] or NULL. */
int oct_gather_tokens_fwd(const void* x, int x_dtype, const int64_t* ids_keep, const float* pos_sp, const float* pos_tmp,
                          const float* cls_row, float* out, int64_t B, int64_t L, int64_t keep, int64_t G, int64_t C,
                          oct_stream_t stream);
/* Same output as oct_gather_tokens_fwd for an x that holds the kept rows only (x_keep [B,keep,C], the gather-first patch
 * embedding: patchify the kept tokens, one GEMM over B*keep rows): out[b,0] = cls_row, out[b,1+i] = x_keep[b,i] +
 * pos_sp[ids_keep[b,i] % G] + pos_tmp[ids_keep[b,i] / G]. */
int oct_posadd_tokens_fwd(const void* x_keep, int x_dtype, const int64_t* ids_keep, const float* pos_sp, const float* pos_tmp,
                          const float* cls_row, float* out, int64_t B, int64_t L, int64_t keep, int64_t G, int64_t C,
                          oct_stream_t stream);

/* backward: dx_keep [B*keep, C] (f32|bf16) = dout rows 1.. ; d_pos_sp [G,C], d_pos_tmp [L/G,C], d_cls_row [C] are
 * deterministic segmented sums (overwritten). */
int oct_gather_tokens_bwd(const float* dout, const int64_t* ids_keep, void* dx_keep, int dx_dtype, float* d_pos_sp,
                          float* d_pos_tmp, float* d_cls_row, int64_t B, int64_t L, int64_t keep, int64_t G, int64_t C,
                          oct_stream_t stream);

/* ---- residual add + LayerNorm (FA:block.py:126-130,163-167; models:489,592) --------------------------------
 * r = h (+ res_in) ; res_out = r (fp32, optional) ; y = LN(r)·gamma + beta ; mean/rstd saved ([M] fp32).
 * h [M,C] f32|bf16 ; res_in/res_out [M,C] f32 or NULL ; y f32|bf16.  C % 4 == 0, C <= 8192. */
int oct_add_ln_fwd(const void* h, int h_dtype, const float* res_in, float* res_out, const float* gamma,
                   const float* beta, void* y, int y_dtype, float* mean, float* rstd, int64_t M, int64_t C, float eps,
                   oct_stream_t stream);
/* backward of the above.  x = the LN input that was normalised (res_out if it was written, else h).
 * dx = LNbwd(dy) (+ dres_in) written as fp32 (dx_f32, optional) and/or low precision (dx_lp, optional);
 * dgamma/dbeta [C] overwritten; ws: >= oct_add_ln_bwd_ws_bytes(M,C). */
size_t oct_add_ln_bwd_ws_bytes(int64_t M, int64_t C);
int oct_add_ln_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                   const float* gamma, const float* dres_in, float* dx_f32, void* dx_lp, int dx_lp_dtype, float* dgamma,
                   float* dbeta, void* ws, size_t ws_bytes, int64_t M, int64_t C, oct_stream_t stream);
/* oct_add_ln_bwd in two launches: `main` writes dx and leaves per-CTA partial sums of dgamma / dbeta in ws (*nblocks = number of
 * partial rows, a HOST int), `finish` reduces them in a fixed order — on any stream, e.g. off the dgrad chain. */
int oct_add_ln_bwd_main(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                        const float* gamma, const float* dres_in, float* dx_f32, void* dx_lp, int dx_lp_dtype, void* ws,
                        size_t ws_bytes, int64_t M, int64_t C, int* nblocks, oct_stream_t stream);
int oct_add_ln_bwd_finish(const void* ws, int nblocks, int64_t C, float* dgamma, float* dbeta, oct_stream_t stream);

/* ---- GEMM (replaces cuBLASLt for Wqkv/out_proj/fc1/fc2/decoder_embed/decoder_pred and their dgrad/wgrad) ----
 * compute = OCT_BF16: A,B bf16, tcgen05.mma kind::f16, fp32 accumulate in TMEM, TMA-fed; D bf16|f32.
 * compute = OCT_F32 : A,B,D f32, CUDA-core FFMA (parity mode).
 * bias [N] f32 ; aux [M,ldd] same dtype as D (see epilogues) ; beta in {0,1} (1 only with fp32 D, EPI_NONE). */
int oct_gemm(int compute, int layout, const void* A, const void* B, void* D, int d_dtype, int64_t M, int64_t N,
             int64_t K, int64_t lda, int64_t ldb, int64_t ldd, int epilogue, const float* bias, void* aux, int beta,
             oct_stream_t stream);

/* ---- nn.Linear weight + bias gradient in ONE tensor-core kernel (replaces cuBLASLt wgrad + the sum-over-tokens reduction
 * autograd runs for every Linear: FA:mha.py:635,703, FA:mlp.py:48-50, models:511,595) -------------------------------
 * dw [n_out, ld_dw] f32 (= | +=) dy[tokens, n_out]^T x[tokens, k_in] ; db [n_out] f32 (= | +=) sum_t dy[t, :]
 * dy, x bf16 row-major (ld_dy, ld_x elements per row).  The bias gradient is one extra 128x16x16 tcgen05.mma per k-step
 * against a constant tile of ones (dy is already in shared memory as the A operand).  compute must be OCT_BF16. */
int oct_gemm_wgrad_bias(int compute, const void* dy, const void* x, float* dw, float* db, int64_t n_out, int64_t k_in,
                        int64_t tokens, int64_t ld_dy, int64_t ld_x, int64_t ld_dw, int beta, oct_stream_t stream);

/* ---- self-attention on packed qkv (FA:mha.py:122-130 flash_attn_qkvpacked_func, non-causal, p=0) -----------
 * qkv [B,S,3,H,d] ; out [B,S,H,d] ; lse [B,H,S] f32 (natural-log-sum-exp of scaled scores) ; scale = d^-0.5.
 * compute = OCT_BF16: tcgen05 flash kernel (d in {32,64}); OCT_F32: fp32 CUDA-core kernel (d <= 128). */
int oct_attn_fwd(int compute, const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H, int64_t d,
                 float scale, oct_stream_t stream);
size_t oct_attn_bwd_ws_bytes(int compute, int64_t B, int64_t S, int64_t H, int64_t d);
int oct_attn_bwd(int compute, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* ws,
                 size_t ws_bytes, int64_t B, int64_t S, int64_t H, int64_t d, float scale, oct_stream_t stream);

/* ---- GELU (exact erf; FA:mlp.py:49) for the fp32 path; the bf16 path fuses it into the GEMM epilogues ------- */
int oct_gelu_fwd(const void* x, void* y, int dtype, int64_t n, oct_stream_t stream);
int oct_gelu_bwd(const void* dy, const void* x, void* dx, int dtype, int64_t n, oct_stream_t stream);

/* ---- bias gradient: out[n] (= or +=) sum_m x[m,n] ; x [M,ldx] f32|bf16 ; deterministic two-stage (N, ldx multiples of 4;
 * any other width takes a one-thread-per-column path) ------------------------------------------------------------ */
size_t oct_colsum_ws_bytes(int64_t M, int64_t N);
int oct_colsum(const void* x, int x_dtype, float* out, int64_t M, int64_t N, int64_t ldx, int beta, void* ws,
               size_t ws_bytes, oct_stream_t stream);

/* ---- decoder un-shuffle + mask tokens + cls + separable pos add (models:515-573) --------------------------
 * out[b,0] = cls_row ; out[b,1+j] = (r<keep ? y[b,y_row0+r] : mask_token) + pos_sp[j % G] + pos_tmp[j / G], r = ids_restore[b,j]
 * y [B,y_row0+keep,D] f32|bf16 ; out [B,1+L,D] f32 (cls_row NULL -> no cls row, out [B,L,D]).
 * y_row0 = 0: the 3D model (its encoder cls is dropped, models:491-493; the decoder cls is a parameter).
 * y_row0 = 1: the 2D model (OCTCube/models_mae_flash_attn.py:299-312): row 0 of y is the sample's own cls token after
 * decoder_embed, and out[b,0] = y[b,0] + cls_row (cls_row = decoder_pos_embed[0]). */
int oct_unshuffle_fwd(const void* y, int y_dtype, const int64_t* ids_restore, const float* mask_token,
                      const float* pos_sp, const float* pos_tmp, const float* cls_row, float* out, int64_t B, int64_t L,
                      int64_t keep, int64_t G, int64_t D, int64_t y_row0, oct_stream_t stream);
/* backward: dy [B,y_row0+keep,D] (row 0 = dout[b,0] when y_row0 = 1) ; d_mask_token [D], d_pos_sp [G,D], d_pos_tmp [L/G,D]
 * (nullable), d_cls_row [D] (nullable) are deterministic sums (overwritten).  ws >= oct_unshuffle_bwd_ws_bytes(B,L,G,D). */
size_t oct_unshuffle_bwd_ws_bytes(int64_t B, int64_t L, int64_t G, int64_t D);
int oct_unshuffle_bwd(const float* dout, const int64_t* ids_restore, void* dy, int dy_dtype, float* d_mask_token,
                      float* d_pos_sp, float* d_pos_tmp, float* d_cls_row, void* ws, size_t ws_bytes, int64_t B,
                      int64_t L, int64_t keep, int64_t G, int64_t D, int has_cls, int64_t y_row0, oct_stream_t stream);

/* ---- forward_loss (models:613-667): masked MSE on (optionally per-patch normalised) pixels ------------------
 * Reads the volume in place through patch indexing (no patchify copy).  imgs [B,1,T,H,W]; T_sel frames enter the loss
 * (T_sel == T, or frame_idx [T_sel] = linspace(0,T-1,pred_t_dim).long() of models:630-640); u = t_pred_patch_size.  pred [B, pred_rows, P] with the token j at
 * row (pred_row0 + j) (pred_row0 = 1 skips the cls row) ; mask [B,L] ; loss_tok [B,L] (workspace/out: per-patch mean
 * squared error, 0 where mask==0) ; loss [1] ; mask_sum [1] ; frame_losses [B,T'] ; sums are reduced in a fixed order.
 * flags: OCT_LOSS_NORM_PIX (1) per-patch normalised target (unbiased variance, eps 1e-6) ;
 *        OCT_LOSS_CHANNEL_LAST (2) the 2D model's patch order 'nchpwq->nhwpqc' (OCTCube/models_mae_flash_attn.py:214-226):
 *        imgs is [B,u,H,W] (u channels, T = T_sel = u) and element e of a patch is (kh, kw, c) with c fastest ;
 *        OCT_LOSS_ALL_TOKENS (4) loss_tok is written for kept tokens too (the 2D model's return_frame_loss, :343-345). */
#define OCT_LOSS_NORM_PIX 1
#define OCT_LOSS_CHANNEL_LAST 2
#define OCT_LOSS_ALL_TOKENS 4
int oct_mse_loss_fwd(const float* imgs, const int64_t* frame_idx, const void* pred, int pred_dtype, const float* mask,
                     float* loss_tok, float* loss, float* mask_sum, float* frame_losses, int64_t B, int64_t T, int64_t T_sel,
                     int64_t H, int64_t W, int64_t p, int64_t u, int64_t pred_rows, int64_t pred_row0, int flags,
                     oct_stream_t stream);
/* dpred [B,pred_rows,P] (rows < pred_row0 and kept tokens zero) = dloss · 2 (pred − target) · mask / (P · Σmask) */
int oct_mse_loss_bwd(const float* imgs, const int64_t* frame_idx, const void* pred, int pred_dtype, const float* mask,
                     const float* mask_sum, const float* dloss, void* dpred, int dpred_dtype, int64_t B, int64_t T,
                     int64_t T_sel, int64_t H, int64_t W, int64_t p, int64_t u, int64_t pred_rows, int64_t pred_row0,
                     int flags, oct_stream_t stream);

/* ---- token pooling of the encoder-only ViT (OCTCube/models_vit_st_flash_attn.py:247-251) ------------------------
 * out[b,:] = mean over s in [row0, row1) of x[b,s,:] : `x[:, 1:, :].mean(dim=1)` (row0 = 1, row1 = S: global pool without the
 * cls token) or `x[:, 0]` (row0 = 0, row1 = 1).  x [B,S,C] f32|bf16, out [B,C] f32|bf16, fp32 accumulation in a fixed order.
 * ws >= oct_mean_pool_ws_bytes(B,C,row0,row1). */
size_t oct_mean_pool_ws_bytes(int64_t B, int64_t C, int64_t row0, int64_t row1);
int oct_mean_pool_fwd(const void* x, int x_dtype, void* out, int out_dtype, int64_t B, int64_t S, int64_t C, int64_t row0,
                      int64_t row1, void* ws, size_t ws_bytes, oct_stream_t stream);
/* dx[b,s,:] = dout[b,:] / (row1 - row0) for s in [row0, row1), 0 for the other rows */
int oct_mean_pool_bwd(const void* dout, int dout_dtype, void* dx, int dx_dtype, int64_t B, int64_t S, int64_t C, int64_t row0,
                      int64_t row1, oct_stream_t stream);

/* ---- fixed sparse linear map on a table (ELL format): y[r,:] = sum_k w[r,k] * x[idx[r,k],:]  (fp32, fixed order) ----------
 * The bicubic 32x32 -> 16x16 resampling of the learnable spatial pos tables (models:419-421, 537-539; F.interpolate bicubic,
 * align_corners=False) is a fixed linear map with 16 taps per output cell; its transpose (the backward) has <= 9 per input cell.
 * idx int32 [R,K], w f32 [R,K] (padding entries carry w = 0), x f32 [rows,C], y f32 [R,C]. */
int oct_ell_spmm(const int* idx, const float* w, const float* x, float* y, int64_t R, int64_t K, int64_t C, oct_stream_t stream);

/* ---- volume ingest: uint8 cube -> the step's fp32 input, replacing the loader's CPU work ------------------
 * dst [B,1,T,H,W] f32 <- src [B,T_src,H,W] u8 : value / divisor (ToTensor's /255, PatientDataset_inhouse.py:420), centre
 * zero-padding ((T - T_src) // 2 frames on the left) or centre cropping (frames [(T_src - T) // 2, ... + T)) to T frames
 * (:436-450), then per-sample flips along the frame axis / the width where flip_t[b] / flip_w[b] != 0 (RandFlipd spatial_axis
 * 0 / 2 of create_3d_transforms :59-62; NULL = no flips).  Bit-identical to the CPU pipeline (true division). */
int oct_ingest_u8(const uint8_t* src, float* dst, const uint8_t* flip_t, const uint8_t* flip_w, int64_t B, int64_t T_src,
                  int64_t T, int64_t H, int64_t W, float divisor, oct_stream_t stream);

/* CropForegroundd + Resized(trilinear) + RandFlipd of create_3d_transforms (PatientDataset_inhouse.py:56-63), per cube, on the
 * device.  The loader first pads / centre-crops the cube to T_pad frames (:436-450); coordinates below are in that padded cube.
 * oct_fg_bbox_u8: bounding box of the voxels > 0 (monai CropForegroundd, margin 0) -> box[6] = {t0,t1,h0,h1,w0,w1} (device ints,
 * ends exclusive; t1 == -1 when the cube has no foreground).  oct_resize_trilinear_u8: dst [T,H,W] f32 <- the box (or the whole
 * padded cube when box == NULL or empty) of src [T_src,H_src,W_src] u8 scaled by 1/divisor, resampled like
 * F.interpolate(mode="trilinear", align_corners=False) (ATen's source-index rule and blend order), then flipped along the frame
 * axis / the width when flip_t / flip_w != 0. */
int oct_fg_bbox_u8(const uint8_t* src, int* box, int64_t T_src, int64_t H, int64_t W, int64_t T_pad, oct_stream_t stream);
int oct_resize_trilinear_u8(const uint8_t* src, float* dst, const int* box, int64_t T_src, int64_t H_src, int64_t W_src,
                            int64_t T_pad, int64_t T, int64_t H, int64_t W, int flip_t, int flip_w, float divisor,
                            oct_stream_t stream);

/* ---- fp32 -> bf16 shadow copy of parameters (the autocast weight cast, done once per step) ------------------- */
int oct_cast_f32_to_bf16(const float* src, void* dst, int64_t n, oct_stream_t stream);
/* every shadow in one launch: `table` = device int64 [n_chunks][3] = {src pointer (16-byte aligned fp32), dst pointer (8-byte
 * aligned bf16), element count <= 16384}; chunks never cross a tensor */
int oct_cast_f32_to_bf16_multi(const int64_t* table, int64_t n_chunks, oct_stream_t stream);

/* ---- fused multi-tensor AdamW (replaces torch.optim._multi_tensor.AdamW, main_pretrain...:451-455, the GradScaler unscale
 * before it and the next forward's bf16 weight casts) -----------------------------------------------------------------
 * table [n_chunks][6] int64 on the device = {param f32*, grad f32*, exp_avg f32*, exp_avg_sq f32*, bf16 shadow* or 0, count}
 * per chunk (count <= 16384; chunks never cross a tensor; pointers 16-byte aligned, shadow 8-byte).  One call = one parameter
 * group (its lr / weight_decay); step counts from 1; the gradient is first multiplied by grad_scale (1/loss_scale, or 1) and,
 * if grad_scale_dev is not NULL, by that device scalar (the clip coefficient oct_grad_norm leaves in out[1]). */
int oct_adamw_step(const int64_t* table, int64_t n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                   int64_t step, float grad_scale, const float* grad_scale_dev, oct_stream_t stream);

/* Device-resident optimizer clock for CUDA-graph replay: `clock` = 16 bytes on the device {int32 step, f32 lr, f32 1-beta1^step,
 * f32 sqrt(1-beta2^step)}, zero-initialised by the caller.  oct_adamw_clock_advance (one thread, inside the captured step)
 * increments step and evaluates the reference's per-iteration half-cycle cosine (custom_util/lr_sched.py:10-28) at the
 * fractional epoch e = (step-1) * epochs_per_step (epochs_per_step = accum_iter / len(data_loader), engine_pretrain.py:87-91):
 * lr = base_lr * e / warmup_epochs while e < warmup_epochs, else min_lr + (base_lr - min_lr) * (1 + cos(pi (e - warmup) /
 * (epochs - warmup))) / 2.  oct_adamw_step_clocked is oct_adamw_step with lr = lr_scale * clock.lr and the bias corrections
 * read from the clock, so that replaying the same launch parameters performs consecutive optimizer steps. */
int oct_adamw_clock_advance(void* clock, float base_lr, float min_lr, float warmup_epochs, float epochs, float epochs_per_step,
                            float beta1, float beta2, oct_stream_t stream);
int oct_adamw_step_clocked(const int64_t* table, int64_t n_chunks, const void* clock, float lr_scale, float beta1, float beta2,
                           float eps, float weight_decay, float grad_scale, const float* grad_scale_dev, oct_stream_t stream);

/* global L2 norm of the gradients in `table` (same rows as oct_adamw_step; several groups = several tables concatenated by the
 * caller) — misc.py:356-373: out[0] = grad_scale * ||g||_2 ; out[1] = max_norm > 0 ? min(1, max_norm / (out[0] + 1e-6)) : 1.
 * partial: n_chunks floats of workspace; deterministic two-stage sum; nothing is copied to the host. */
int oct_grad_norm(const int64_t* table, int64_t n_chunks, float grad_scale, float max_norm, float* partial, float* out,
                  oct_stream_t stream);

/* ---- contrastive step of OCTCube-IR (retinal-COEM/src/open_clip/model.py:661-683, loss.py:21-63,148-229) ----
 * oct_l2norm_*: F.normalize(x, dim=-1) (model.py:663,667): y[b,:] = x[b,:] / max(||x[b,:]||_2, eps), y fp32, inv_norm [B]
 * saved for the backward (its sign bit marks rows that hit the eps clamp); dx = inv_norm * (dy - y <dy, y>).
 *
 * oct_clip_loss_*: ClipLoss.forward in the recipe's configuration (local_loss, gather_with_grad, labels = arange(B) + B*rank):
 *   loss = ( CE(scale * image  @ all_enface^T, labels) + CE(scale * enface @ all_image^T, labels) ) / 2
 * with the feature all-gather (loss.py:51-52) and, in the backward, the reduce-scatter its autograd implies folded INTO the
 * kernels: `peer_bufs` is a HOST array of `world` device pointers — entry s is rank s's exchange buffer
 * (oct_clip_xchg_bytes(B, D) bytes, zero-initialised, peer-mapped on this device: oct_peer_* below or any symmetric-memory
 * allocator).  The forward publishes this rank's features into its own buffer, raises an epoch flag on every peer and reads
 * the peers' features over NVLink tile by tile as their flags arrive; the backward additionally reads the peers' log-sum-exp
 * vectors and emits d image, d enface ([B,D] fp32) and d logit_scale — no collective is launched.  Every rank must call fwd
 * (then optionally bwd) once per step in the same order.  `state`: oct_clip_state_bytes(B) bytes of zero-initialised device
 * memory owned by the caller (device-resident epoch: the launches replay from a CUDA graph).  image / enface [B,D] fp32,
 * L2-normalised; logit_scale / dloss / loss / d_scale: device scalars; D % 4 == 0, D <= 1024, world <= 64. */
int oct_l2norm_fwd(const void* x, int x_dtype, float* y, float* inv_norm, int64_t B, int64_t D, float eps, oct_stream_t stream);
int oct_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, void* dx, int dx_dtype, int64_t B, int64_t D,
                   oct_stream_t stream);
size_t oct_clip_xchg_bytes(int64_t B, int64_t D);
size_t oct_clip_state_bytes(int64_t B);
int oct_clip_loss_fwd(const float* image, const float* enface, const float* logit_scale, const void* const* peer_bufs, void* state,
                      float* loss, int rank, int world, int64_t B, int64_t D, oct_stream_t stream);
int oct_clip_loss_bwd(const float* image, const float* enface, const float* logit_scale, const float* dloss,
                      const void* const* peer_bufs, void* state, float* d_image, float* d_enface, float* d_scale, int rank, int world,
                      int64_t B, int64_t D, oct_stream_t stream);

/* ---- gradient all-reduce over NVLink / NVSwitch without a collective library (the all-reduce behind
 * DistributedDataParallel, main_pretrain...:435-439) ----------------------------------------------------------------------
 * In-place SUM (times `scale`) of n_elems fp32 values at element offset off_elems of a SYMMETRIC buffer: every rank holds one
 * copy of the allocation, all copies are mapped into every process (peer_bufs: HOST array [world] of device pointers, entry s =
 * rank s's copy) and, on NVSwitch fabrics, behind ONE multicast address mc_ptr (NULL selects the peer path).  Multicast path:
 * rank r reduces its 1/world shard with multimem.ld_reduce (the switch adds the copies) and broadcasts it with multimem.st; peer
 * path: the same two-shot schedule with peer loads / stores.  Ranks synchronise through epoch flags inside the allocation
 * (flag_off_bytes: oct_allreduce_flag_bytes() zero-initialised bytes at the same offset in every copy); `state`: local device
 * memory (oct_allreduce_state_bytes(), zero-initialised; word 2 of slot `slot` is raised if a peer never arrived).  The kernel
 * uses no shared memory, so its CTAs co-reside with the persistent GEMM / attention CTAs of the backward pass it overlaps.
 * Every rank issues the same sequence of calls; calls in flight at the same time use different slots (0..7). */
size_t oct_allreduce_flag_bytes(void);
size_t oct_allreduce_state_bytes(void);
int oct_allreduce_sym(void* mc_ptr, const void* const* peer_bufs, int64_t flag_off_bytes, void* state, int64_t off_elems,
                      int64_t n_elems, int rank, int world, int slot, float scale, int ctas, oct_stream_t stream);

/* Peer-mapped device memory, one process per GPU (CUDA IPC; NVLink 5 / NVSwitch carries the loads and stores).  Set-up time
 * only — these are the library's only allocating entry points.  oct_peer_alloc: zero-filled cudaMalloc; oct_peer_export: 64-byte
 * handle (HOST memory) to hand to the other ranks by any host channel; oct_peer_open: maps a peer's allocation, returns its
 * address in this process; oct_peer_close / oct_peer_free undo them. */
int oct_peer_alloc(void** ptr, int64_t bytes);
int oct_peer_free(void* ptr);
int oct_peer_export(void* ptr, void* handle64);
int oct_peer_open(const void* handle64, void** ptr);
int oct_peer_close(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* OCTCUBE_B200_H */
