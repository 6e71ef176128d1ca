/* clm_b200.h — C-ABI of the B200-native CLIP+LoRA retrieval hot path.
 *
 * This is the drop-in boundary.  The reference (youngalip/clip-lora-match) has no FFI
 * of its own: its hot path is plain Python that calls transformers / peft / torch
 * (SURVEY.md §8b).  Each entry point below therefore cites the reference *call site*
 * whose arithmetic it replaces.  The Python wrappers in clip_lora_match_b200/ bind
 * these with ctypes (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns every buffer; the library allocates nothing after *_create;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - bf16 buffers are raw uint16 storage (__nv_bfloat16);
 *   - return value 0 = ok, non-zero = error (clm_last_error() gives the text);
 *   - all functions are stream-ordered and re-entrant (no global mutable state except
 *     the thread-local last-error string).
 */
#ifndef CLM_B200_H
#define CLM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLM_OK 0
#define CLM_ERR_INVALID 1
#define CLM_ERR_CUDA 2
#define CLM_ERR_UNSUPPORTED 3

/* epilogue selector for clm_gemm_epi (bit flags) */
#define CLM_EPI_NONE 0
#define CLM_EPI_QUICKGELU 1 /* x * sigmoid(1.702 x); transformers/activations.py QuickGELU */
/* Allow split-K for an in-place fp32 accumulation (out == residual) that has fewer tiles than half the SMs and
 * K >= 2048: the k range of a tile is cut into work units that add their partial products into out through the
 * L2 (TMA reduce-add).  The additions arrive in no fixed order: the result is reproducible to fp32 rounding, not
 * bit for bit.  Used by the training step's LoRA weight gradients (out [features, 64] += dy^T t over all token
 * rows).  No effect on any other call. */
#define CLM_EPI_SPLIT_K 2
/* Scheduling hint: walk the tile list from its end (the LAST rows of the output first).  A kernel that starts on the
 * rows its predecessor touched last finds them in the L2; the tower alternates the direction from kernel to kernel.
 * No effect on the result. */
#define CLM_EPI_REVERSE 4
#define CLM_OUT_BF16 0
#define CLM_OUT_F32 1

const char* clm_last_error(void);
int clm_version(void);
/* 0 when a CUDA device of compute capability 10.x is usable, error otherwise. */
int clm_device_check(void);

/* ------------------------------------------------------------------------------------
 * Elementwise / normalisation kernels (HBM-bound; warp-shuffle, 128-bit access)
 * ---------------------------------------------------------------------------------- */

/* nn.LayerNorm(eps) over the last dim: y = (x-mean)/sqrt(var+eps)*gamma+beta.
 * Replaces transformers modeling_clip.py:359,361,562,677,686 as reached from
 * models/clip_model.py:115,144.  x fp32 [rows, dim] (the fp32 residual stream),
 * y bf16 [rows, dim] (the next GEMM's operand).  dim in {512, 768, 1024}. */
int clm_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16,
                  int rows, int dim, float eps, void* stream);

/* The same kernels over a residual stream held in bf16 (x_dtype / h_dtype = CLM_OUT_BF16) instead of fp32
 * (CLM_OUT_F32: exactly the functions without the suffix).  A bf16 stream is what clm_tower_set_residual_dtype
 * selects: every residual update then rounds once to bf16 (see there).  No reference counterpart: the reference's
 * stream is whatever dtype the checkpoint was loaded in (models/clip_model.py:62-66, fp32). */
int clm_layernorm_ex(const void* x, int x_dtype, const float* gamma, const float* beta, void* y_bf16,
                     int rows, int dim, float eps, void* stream);
int clm_embed_text_ex(const int32_t* ids, const float* tok_emb, const float* pos_emb, void* h, int h_dtype,
                      int32_t* eos_pos, int batch, int tokens, int dim, int vocab, int eos_id, void* stream);
int clm_vision_embed_ln_ex(const float* patch_out, const float* class_emb, const float* pos_emb,
                           const float* gamma, const float* beta, void* h, int h_dtype, int batch, int np,
                           int dim, float eps, void* stream);
int clm_pool_ln_ex(const void* h, int h_dtype, const int32_t* row_idx_or_null, const float* gamma,
                   const float* beta, void* y_bf16, int batch, int tokens, int dim, float eps, void* stream);

/* stats[row] = (mean, 1/sqrt(var + eps)) of every row of a bf16 residual stream h [rows, dim]; stats fp32
 * [rows, 2].  The row statistics of nn.LayerNorm (modeling_clip.py:359,361) for a GEMM that has the LayerNorm
 * folded in (clm_gemm_ln_epi). */
int clm_row_stats(const void* h_bf16, float* stats, int rows, int dim, float eps, void* stream);

/* Row L2 normalisation, no epsilon: x / ||x||  (models/clip_model.py:116,148;
 * src/embedding/search.py:68,93).  fp32 in, fp32 out (may alias), optional bf16 copy. */
int clm_l2norm(const float* x, float* y, void* y_bf16_or_null, int rows, int dim, void* stream);

/* Query fusion of the seeker path (src/embedding/seeker_service.py:146-157):
 *   v = wa * a (+ wb * b);  out = v / ||v||   (no epsilon, as the reference)
 * a, b fp32 [rows, dim] L2-normalised text / image embeddings; b may be NULL (single modality:
 * :148-151, pass wa = 1).  out fp32 (may alias a or b), optional bf16 copy for the scan. */
int clm_fuse_normalize(const float* a, float wa, const float* b_or_null, float wb, float* out,
                       void* out_bf16_or_null, int rows, int dim, void* stream);

/* Text embeddings: h[b,t,:] = tok_emb[ids[b,t]] + pos_emb[t]  (modeling_clip.py:249-256).
 * ids int32 [batch, tokens]; tables fp32; h fp32 [batch*tokens, dim].
 * Also writes eos_pos[b] = first t with ids[b,t] == eos_id (0 if none), the pooling
 * row of modeling_clip.py:577-584. */
int clm_embed_text(const int32_t* ids, const float* tok_emb, const float* pos_emb, float* h,
                   int32_t* eos_pos, int batch, int tokens, int dim, int vocab, int eos_id,
                   void* stream);

/* Patch extraction (the im2col view of nn.Conv2d(stride=patch, bias=False),
 * modeling_clip.py:148-154,208-209).  pixel_values fp32 [batch,3,image,image] ->
 * patches bf16 [batch*grid*grid, kpad] with column order (c, py, px), zero padded to
 * kpad (multiple of 64). */
int clm_patch_im2col(const float* pixel_values, void* patches_bf16, int batch, int image,
                     int patch, int kpad, void* stream);

/* Vision embeddings + pre_layrnorm (modeling_clip.py:202-219, 677):
 * h[b,0,:] = LN(class_emb + pos[0]); h[b,1+p,:] = LN(patch_out[b*np+p,:] + pos[1+p]).
 * patch_out fp32 [batch*np, dim]; h fp32 [batch*(np+1), dim]. */
int clm_vision_embed_ln(const float* patch_out, const float* class_emb, const float* pos_emb,
                        const float* gamma, const float* beta, float* h, int batch, int np,
                        int dim, float eps, void* stream);

/* Pooling: y[b,:] = LN(h[b*tokens + (row_idx ? row_idx[b] : 0), :])  -> bf16 [batch, dim]
 * (modeling_clip.py:685-686 CLS + post_layernorm; :562,577-584 final_layer_norm + EOS row). */
int clm_pool_ln(const float* h, const int32_t* row_idx_or_null, const float* gamma,
                const float* beta, void* y_bf16, int batch, int tokens, int dim, float eps,
                void* stream);

/* CLIP image preprocessing on the GPU: decoded uint8 RGB (HWC) images of any size ->
 * pixel_values fp32 [batch, 3, S, S].  Replaces processor(images=...) of models/clip_model.py:105-107
 * (resize shortest edge to S with Pillow's antialiased bicubic — bit exact on the uint8 stage —
 * centre crop S x S, rescale 1/255, (x - mean) / std; config/clip_config.yaml:7-14).
 * descs_host is a HOST array; .data are DEVICE pointers; the workspace is caller-owned device memory. */
typedef struct {
  const uint8_t* data;      /* device pointer, RGB interleaved */
  int32_t height, width;
  int32_t row_stride_bytes; /* >= 3 * width */
} clm_image_desc;
size_t clm_preprocess_workspace_bytes(const clm_image_desc* descs_host, int batch, int out_size);
int clm_preprocess_images(const clm_image_desc* descs_host, int batch, int out_size,
                          const float* mean3_host, const float* std3_host, float* pixel_values,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * tcgen05 / TMEM GEMM fed by TMA, with fused epilogue
 * ---------------------------------------------------------------------------------- */

/* out[M,N] = epi( A[M,K]·W[N,K]^T  (+ A2[M,K2]·W2[N,K2]^T)  + bias ) (+ residual)
 * A, W, A2, W2 bf16 row-major with leading dims lda.. (elements; multiples of 8);
 * K and K2 are padded by the hardware (TMA zero fill) to multiples of 64.
 * nn.Linear of modeling_clip.py:295-298,310-312,334 (q/k/v/out), :344-350 (fc1+QuickGELU,
 * fc2), :822-823,860-861 (projections).  The (A2,W2) pair is the unmerged PEFT LoRA
 * update y += (x A^T) (s B)^T  (models/lora_adapter.py:35-42; Appendix B of SURVEY.md)
 * folded in as a K-extension of the same accumulator.
 * bias fp32 [N] or NULL; residual fp32 [M, ldr] or NULL (added after the activation);
 * out is bf16 or fp32 per out_dtype; out may alias residual.
 * In-place residual update (residual == out, ldr == ldo): the stream is updated in its own type through a TMA
 * reduce-add inside the L2 -- fp32 with out_dtype = CLM_OUT_F32, and with out_dtype = CLM_OUT_BF16 the buffer is
 * a bf16 residual stream (pass the same pointer for both): out = bf16(out + bf16(acc + bias)). */
int clm_gemm_epi(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                 const void* A2, int lda2, const void* W2, int ldw2, int K2,
                 void* out, int ldo, int out_dtype, const float* bias,
                 const float* residual, int ldr, int epilogue, void* stream);

/* nn.LayerNorm folded into the nn.Linear that consumes it (layer_norm1 -> q/k/v, layer_norm2 -> fc1;
 * modeling_clip.py:359-362,368-369), for a bf16 residual stream:
 *   LN(h) W^T + b  =  rstd (h Wg^T - mean * col_sums) + bias',   Wg = W diag(gamma) (bf16), col_sums[n] = sum_k Wg[n,k]
 *   (of the bf16 values, in fp32), bias' = b + W beta.
 * H is the raw bf16 stream [M, K] (no normalised copy exists), row_stats = clm_row_stats(H).
 * ln_mode 1: out = act( rstd (acc - mean col_sums) + bias ).  The optional (A2, W2) K-extension is inside acc, so A2
 *   must be the LoRA down-projection DIVIDED by rstd, which is what ln_mode 2 produces:
 * ln_mode 2: out = (acc - mean col_sums) + bias.  With Wg = A diag(gamma) and bias = NULL this is
 *   u = (LN(h) A^T - A beta) / rstd; the constant (A beta) (sB)^T the adapter adds to every row belongs into the bias'
 *   of the ln_mode 1 GEMM (kernels.py fold_layernorm / models/clip_model.py do this on the host).
 * No residual; out bf16 or fp32; everything else as clm_gemm_epi. */
int clm_gemm_ln_epi(const void* H, int ldh, const void* Wg, int ldw, int M, int N, int K,
                    const void* A2, int lda2, const void* W2, int ldw2, int K2,
                    void* out, int ldo, int out_dtype, const float* bias, const float* row_stats,
                    const float* col_sums, int ln_mode, int epilogue, void* stream);

/* Debug / test hook: the kernel instantiation the calling thread's last clm_gemm_epi selected, encoded as
 * BN*100 + ctas*10 + epilogue (ctas: 1 = single CTA, 2 = cta_group::2 pair; epilogue: 0 = per-thread,
 * 1 = TMA store bf16, 2 = TMA store fp32, 3 = TMA reduce-add fp32, 4 = TMA reduce-add bf16).  25621 =
 * gemm_kernel<256,2,1>.
 * 0 before the first call.  No reference counterpart (the reference calls ATen's matmul). */
int clm_last_gemm_variant(void);

/* ------------------------------------------------------------------------------------
 * Fused attention (tcgen05: S=QK^T in TMEM, fp32 softmax in registers, O=PV)
 * ---------------------------------------------------------------------------------- */

/* qkv bf16 [batch*tokens, 3*dim] (q | k | v, head h at columns h*64..h*64+63 of each
 * third); out bf16 [batch*tokens, dim].  softmax(q k^T / sqrt(64) + mask) v with the
 * softmax in fp32 (modeling_clip.py:261-279); causal != 0 applies the text tower's
 * causal mask (modeling_clip.py:546-557).  head_dim is 64 for every CLIP ViT. */
int clm_attention(const void* qkv_bf16, void* out_bf16, int batch, int tokens, int heads,
                  int causal, void* stream);

/* ------------------------------------------------------------------------------------
 * Whole-tower convenience (what models/clip_model.py encode_image / encode_text run)
 * ---------------------------------------------------------------------------------- */

typedef struct clm_tower clm_tower; /* opaque */

typedef struct {
  int32_t kind;   /* 0 = vision, 1 = text */
  int32_t width;  /* D */
  int32_t layers;
  int32_t heads;
  int32_t mlp;      /* intermediate size */
  int32_t proj_dim; /* P */
  int32_t tokens;   /* vision: 1 + (image/patch)^2 ; text: context length (77) */
  int32_t image;    /* vision only */
  int32_t patch;    /* vision only */
  int32_t vocab;    /* text only */
  int32_t eos_id;   /* text only */
  /* Padded total LoRA rank folded into each fused GEMM as extra K blocks: 0 (no adapter on that GEMM) or a
   * multiple of 64 (any rank / any subset of q,k,v for the QKV GEMM; models/lora_adapter.py:33-41 passes any
   * r and target_modules to peft). */
  int32_t lora_cols_qkv;
  int32_t lora_cols_out;
  int32_t lora_cols_fc1;
  int32_t lora_cols_fc2;
  float ln_eps;
} clm_tower_config;

/* Per-layer device pointers.  Weights bf16 [out,in]; biases / LN params fp32. */
typedef struct {
  const float* ln1_g; const float* ln1_b;
  const void* w_qkv; const float* b_qkv;       /* [3D, D], [3D] */
  const void* lora_a_qkv; const void* lora_b_qkv; /* [cols_qkv, D], [3D, cols_qkv] or NULL */
  const void* w_o; const float* b_o;           /* [D, D], [D] */
  const void* lora_a_o; const void* lora_b_o;  /* [cols_out, D], [D, cols_out] or NULL */
  const float* ln2_g; const float* ln2_b;
  const void* w_fc1; const float* b_fc1;       /* [mlp, D] */
  const void* w_fc2; const float* b_fc2;       /* [D, mlp] */
  const void* lora_a_fc1; const void* lora_b_fc1; /* [cols_fc1, D], [mlp, cols_fc1] or NULL */
  const void* lora_a_fc2; const void* lora_b_fc2; /* [cols_fc2, mlp], [D, cols_fc2] or NULL */
} clm_layer_weights;

typedef struct {
  /* vision */
  const void* patch_w;      /* bf16 [D, kpad] conv weight flattened (c,py,px), zero padded */
  const float* class_emb;   /* [D] */
  const float* pre_ln_g; const float* pre_ln_b;
  /* text */
  const float* tok_emb;     /* fp32 [vocab, D] */
  /* both */
  const float* pos_emb;     /* fp32 [tokens, D] */
  const float* final_ln_g; const float* final_ln_b; /* post_layernorm / final_layer_norm */
  const void* proj_w;       /* bf16 [P, D] visual_projection / text_projection (no bias) */
} clm_tower_weights;

int clm_tower_create(const clm_tower_config* cfg, const clm_tower_weights* w,
                     const clm_layer_weights* layers, clm_tower** out);
void clm_tower_destroy(clm_tower* t);
/* bytes of caller-provided device workspace needed for a micro-batch of `batch` items */
size_t clm_tower_workspace_bytes(const clm_tower* t, int batch);
/* Type of the residual stream h between the layers: CLM_OUT_F32 (the default after clm_tower_create: the
 * reference's fp32 stream, every update an fp32 add) or CLM_OUT_BF16 (h stored in bf16: LayerNorm reads 2 instead
 * of 4 bytes per element and the out-projection / fc2 epilogues add bf16 tiles inside the L2; each of the 2 x layers
 * updates rounds once to bf16 -- embeddings stay within north_star's cosine >= 0.999 of the fp32 path, measured
 * >= 0.9999, tests/test_residual_bf16_gpu.py).  Applies to the following encode calls; the workspace need shrinks.
 * clm_tower_residual_dtype returns the current setting. */
int clm_tower_set_residual_dtype(clm_tower* t, int dtype);
int clm_tower_residual_dtype(const clm_tower* t);

/* LayerNorm folding for the bf16 residual stream: per layer, the weights of the two GEMMs that consume a
 * LayerNorm (fused QKV, fc1) and of their LoRA down-projections with gamma folded in, see clm_gemm_ln_epi.  With
 * these set (and the stream in bf16) the per-layer LayerNorm passes are replaced by clm_row_stats (half the bytes)
 * and the normalisation happens in the GEMM epilogues.  Device pointers; the caller keeps them alive.
 * folds = NULL switches folding off again. */
typedef struct {
  const void* w_qkv_g; const float* s_qkv; const float* b_qkv_f; /* [3D, D] bf16, [3D], [3D] (incl. the adapter's constant) */
  const void* lora_a_qkv_g; const float* s_a_qkv;                /* [cols_qkv, D] bf16, [cols_qkv], or NULL */
  const void* w_fc1_g; const float* s_fc1; const float* b_fc1_f; /* [mlp, D] bf16, [mlp], [mlp] */
  const void* lora_a_fc1_g; const float* s_a_fc1;                /* [cols_fc1, D] bf16, [cols_fc1], or NULL */
} clm_layer_ln_fold;
int clm_tower_set_ln_fold(clm_tower* t, const clm_layer_ln_fold* folds /* [layers] or NULL */);

/* models/clip_model.py:89-118 without the PIL step: pixel_values fp32 [batch,3,H,W]
 * -> embeddings fp32 [batch, P]; normalize != 0 applies x/||x|| (clip_model.py:116),
 * normalize == 0 returns the raw projected features (embed_image.py:27 normalize=False). */
int clm_encode_image(clm_tower* t, const float* pixel_values, int batch, float* out_emb,
                     int normalize, void* workspace, size_t workspace_bytes, void* stream);
/* models/clip_model.py:121-150 without the tokenizer: ids int32 [batch, tokens]
 * (right padded with eos) -> L2-normalised embeddings fp32 [batch, P]. */
int clm_encode_text(clm_tower* t, const int32_t* ids, int batch, float* out_emb,
                    int normalize, void* workspace, size_t workspace_bytes, void* stream);
/* The same on the first `tokens` <= context positions only: ids int32 [batch, tokens], every row's first
 * eos inside.  The text tower is causal and pooled at the first eos (modeling_clip.py:546-557,577-584), so
 * positions after it never reach the output; the reference itself encodes a single caption unpadded
 * (models/clip_model.py:133-138).  Lets the host mirror run length-bucketed batches. */
int clm_encode_text_len(clm_tower* t, const int32_t* ids, int batch, int tokens, float* out_emb,
                        int normalize, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * LoRA training step (reference scripts/train_lora.py:83-108,170-211)
 *
 * The forward of a training step is the same clm_layernorm / clm_gemm_epi / clm_attention sequence as the
 * encoder, run layer by layer by the host mirror (models/lora_trainer.py) so that it can keep the activations;
 * every GEMM-shaped piece of the backward (dgrad through the frozen weights with the LoRA term as a K extension,
 * the LoRA projections, their weight gradients) is clm_gemm_epi on transposed operands.  The entry points below
 * are the rest.  Only the LoRA factors receive gradients (models/lora_adapter.py:46-56: the base is frozen).
 * ---------------------------------------------------------------------------------- */

/* g = z * sigmoid(1.702 z) on a stored bf16 pre-activation (the training forward keeps z for the backward; the
 * encoder fuses the activation into the fc1 epilogue instead), and dz = dg * d/dz quickgelu(z).  n % 8 == 0.
 * transformers/activations.py QuickGELUActivation as reached from modeling_clip.py:344-350. */
int clm_quickgelu_fwd(const void* z_bf16, void* g_bf16, long long n, void* stream);
int clm_quickgelu_bwd(const void* dg_bf16, const void* z_bf16, void* dz_bf16, long long n, void* stream);

/* Input gradient of nn.LayerNorm (gamma, beta frozen): x fp32 [rows, dim] is the saved LN input, dy the gradient
 * of the LN output (bf16, or fp32 when dy_is_f32); dres fp32 [rows, dim] is the gradient of the residual stream:
 * dres = (accumulate ? dres : 0) + dx, and dres_bf16 (optional) receives its bf16 copy -- the next dgrad GEMM's
 * operand.  gather != 0: the pooled rows (modeling_clip.py:685-686 class token, :577-584 first EOS): item b
 * reads dy row b and works on row b * tokens + (row_idx ? row_idx[b] : 0) of x / dres / dres_bf16. */
int clm_layernorm_bwd(const void* dy, int dy_is_f32, const float* x, const float* gamma, float* dres,
                      void* dres_bf16_or_null, int rows, int dim, float eps, int accumulate,
                      const int32_t* row_idx_or_null, int tokens, int gather, void* stream);

/* out[b][c][r] = bf16(scale * in[b][r][c]) for b < batch: in is bf16 or fp32 [rows, cols] (leading dim ld_in,
 * batch stride in elements), out bf16 [cols, rows] (ld_out >= rows).  Operands of the weight-gradient GEMMs
 * (dB = dy^T t, dA^T = x^T u contract over the token rows) and the bf16 [cols, in] copy of the LoRA masters. */
int clm_transpose_to_bf16(const void* in, int in_is_f32, long long ld_in, long long batch_stride_in, int rows,
                          int cols, void* out_bf16, long long ld_out, long long batch_stride_out, int batch,
                          float scale, void* stream);
/* out = bf16(scale * in), n % 4 == 0 */
int clm_cast_to_bf16(const float* in, void* out_bf16, long long n, float scale, void* stream);

/* Both LoRA weight gradients of one adapted GEMM in ONE launch, for few token rows (the reference's batch of 8:
 * 400 / 616 rows -- there the tensor-core route, four transposes and two deep-K clm_gemm_epi calls, is launch bound):
 *   grad_b  [n_out, cols] += dy^T t      dy bf16 [rows, n_out], t = x A_cat^T bf16 [rows, cols]
 *   grad_a_t [n_in, cols] += x^T u       x  bf16 [rows, n_in],  u = dy (sB_cat) bf16 [rows, cols]
 * fp32 accumulation on the CUDA cores.  deterministic != 0: one CTA per output tile, bit-for-bit reproducible;
 * 0: when there are fewer tiles than SMs up to four CTAs share a tile's rows and add their partial sums with atomics
 * (reproducible to fp32 rounding).  cols a multiple of 64.  The gradients PEFT's autograd
 * produces for lora_B / lora_A (models/lora_adapter.py:35-42) in the fused layouts of models/lora_trainer.py. */
int clm_lora_wgrad_small(const void* dy_bf16, int ld_dy, int n_out, const void* t_bf16, int ld_t, const void* x_bf16,
                         int ld_x, int n_in, const void* u_bf16, int ld_u, int cols, int rows, float* grad_b,
                         float* grad_a_t, int deterministic, void* stream);

/* Backward of clm_attention: qkv as in the forward, dout bf16 [batch*tokens, dim] -> dqkv bf16
 * [batch*tokens, 3*dim] (dq | dk | dv).  P is recomputed from q and k (fp32 softmax, modeling_clip.py:261-279).
 * scratch: clm_attention_bwd_scratch_bytes() of device memory (P and dS, transposed, between the two kernels).
 * tokens <= 384.  Deterministic (no atomics). */
size_t clm_attention_bwd_scratch_bytes(int batch, int tokens, int heads);
int clm_attention_bwd(const void* qkv_bf16, const void* dout_bf16, void* dqkv_bf16, void* scratch,
                      size_t scratch_bytes, int batch, int tokens, int heads, int causal, void* stream);

/* Symmetric InfoNCE of scripts/train_lora.py:83-108 on UN-normalised features fp32 [n, dim]:
 *   n_* = feat_* / ||feat_*||;  L = n_i n_t^T / temperature;  loss = (CE(L, arange) + CE(L^T, arange)) / 2.
 * loss_out (device fp32 scalar) = loss * loss_scale (loss_scale = 1 / gradient_accumulation_steps, :186);
 * dfeat_* fp32 [n, dim] (both or neither; optional bf16 copies) = d(loss * loss_scale) / d feat_*.  fp32 throughout. */
size_t clm_clip_loss_workspace_bytes(int n, int dim);
int clm_clip_loss(const float* feat_i, const float* feat_t, int n, int dim, float temperature, float loss_scale,
                  float* loss_out, float* dfeat_i, float* dfeat_t, void* dfeat_i_bf16, void* dfeat_t_bf16,
                  void* workspace, size_t workspace_bytes, void* stream);

/* torch.nn.utils.clip_grad_norm_(max_grad_norm) + torch.optim.AdamW.step() over one flat fp32 buffer
 * (scripts/train_lora.py:141,190-193).  The effective gradient of entry i is grads[i] * grad_mult[i]; entries
 * with grad_mult == 0 (padding columns / off-diagonal blocks of the fused LoRA layouts) are never touched.
 * hyper_dev (device fp32[4]) = {lr, 1/(1-beta1^t), 1/sqrt(1-beta2^t), -}: device-resident so that a captured
 * CUDA graph of the step can be replayed with a new learning rate.  max_grad_norm <= 0 disables clipping.
 * sumsq_scratch: device fp32[CLM_ADAMW_SCRATCH_FLOATS]; element 0 receives the squared global gradient norm (before
 * clipping), the rest holds per-block partial sums (summed in a fixed order: the update is reproducible bit for bit). */
#define CLM_ADAMW_SCRATCH_FLOATS 1185
int clm_adamw_step(float* params, const float* grads, const float* grad_mult, float* exp_avg, float* exp_avg_sq,
                   long long n, const float* hyper_dev, float* sumsq_scratch, float max_grad_norm, float beta1,
                   float beta2, float eps, float weight_decay, void* stream);

/* ------------------------------------------------------------------------------------
 * Search: similarity GEMM with fused per-tile top-k, merge, exact fp32 re-score
 * ---------------------------------------------------------------------------------- */

/* Number of index splits (partial candidate lists per query) the scan will use. */
int clm_search_num_splits(int num_queries, int num_rows);

/* First pass (src/embedding/search.py:96 + :99 fused): scores = Q·E^T on tensor cores
 * from the bf16 shadow of the index; the score matrix stays in TMEM; every (query,
 * split) keeps its kc best candidates.  q_bf16 [nq, dim], index_bf16 [n, dim],
 * cand_score fp32 / cand_id int32 [nq, splits, kc].  dim multiple of 64, kc <= 64.
 * thr_io (fp32 [nq], may be NULL): per-query lower bound L of the kb-th best bf16 score over the whole
 * index (kb = the k the caller will ask clm_topk_merge for), shared by all work units of the scan through
 * atomic max.  The caller initialises it (-inf, or any valid lower bound such as the kb-th best score of a
 * sample of rows); a unit publishes the minimum of its list as soon as the list holds kb entries; on return it
 * holds the tightest bound the scan established.  Scores <= L - margin are never kept, so per-split lists may
 * hold fewer than kc entries; empty slots are (-inf, -1).  Lists are unordered.
 * margin: bf16 scores only NOMINATE rows for the exact fp32 re-score of clm_topk_merge.  With |bf16 score -
 * fp32 score| <= eps for every row (unit-norm rows rounded to bf16: eps <= 2^-8), every row of the fp32
 * top-k has a bf16 score >= t_k - 2 eps, t_k the k-th best bf16 score; since L <= t_k, margin = 2 eps keeps
 * all of them -- unless a list overflows, which clm_topk_merge detects.  kb > kc is allowed: such lists never
 * publish their own minimum, the bound then comes from the caller's seed and the histogram only.
 * hist_base (fp32 [nq]) / hist (uint32 [nq, 32], zeroed by the caller, 16-byte aligned), both optional: a
 * per-query histogram of the kept candidates' scores above hist_base[q] in steps of 1/256, shared by all work
 * units: kb candidates counted at or above a bin edge prove t_kb >= that edge over ALL rows seen so far, which
 * tightens L much earlier than any single unit's list can (a unit only sees 1/splits of the rows).  Queries
 * with hist_base = -inf do not use it. */
int clm_search_topk(const void* q_bf16, const void* index_bf16, int nq, int n, int dim, int kc, int kb,
                    int splits, float* thr_io, const float* hist_base, uint32_t* hist, float margin,
                    float* cand_score, int32_t* cand_id, void* stream);

/* out[r] = (kth largest of x[r, 0..n)) - guard, for each of `rows` rows of a row-major fp32 matrix
 * (n <= 16384).  Seeds clm_search_topk's thr_io from the exact scores of a sample of index rows
 * (computed with clm_gemm_epi): the kc-th best over a subset never exceeds the kc-th best over the
 * whole index, so it is a valid bound; `guard` keeps ties with the bound alive. */
int clm_kth_largest(const float* x, int rows, int n, int kth, float guard, float* out, void* stream);

/* out[r] = a LOWER BOUND of the kth largest of x[r, 0..n) minus guard (n >= 1024, kth <= 256, any n): the row is
 * cut into 1024 strided groups and the kth largest of the 1024 group maxima is selected exactly; kth distinct
 * elements lie at or above it, so it never exceeds the true kth largest and for kth << 1024 it is within a
 * rank or two of it.  One streaming pass per row: what lets the scan's seed come from a 16 K-row sample (0.06 ms
 * for 4096 queries where the exact select needs 0.5 ms).  Same use as clm_kth_largest. */
int clm_kth_lower_bound(const float* x, int rows, int n, int kth, float guard, float* out, void* stream);

/* Second pass, per query over its lists*kc candidates: t = k-th best first-pass score, select EVERY candidate
 * with score >= t - margin (not a fixed number), re-score the selected rows exactly in fp32 against the fp32
 * master rows (q_f32·E_f32[id]; skipped when index_f32 is NULL), sort descending (ties: lower id first) and
 * emit the top k with id_offset added (the shard's first global row); fewer than k candidates pad with
 * (-inf, -1).  Reproduces torch.topk(largest=True, sorted=True) of search.py:99.  k <= 1024.
 * overflow (int32 [nq] or NULL): set to 1 for a query whose selection may be incomplete -- a list that is
 * full and whose minimum is still >= t - margin may have dropped qualifying rows, or more than 2048 rows
 * qualified -- so the caller can redo that query exactly (clm_cosine_gemv + clm_topk_row); 0 otherwise. */
int clm_topk_merge(const float* cand_score, const int32_t* cand_id, int nq, int lists, int kc, float margin,
                   const float* q_f32, const float* index_f32, int dim, int k, int64_t id_offset,
                   float* out_score, int64_t* out_id, int32_t* overflow, void* stream);

/* Cross-shard merge after the allgather: in [nq, lists, k] (score, global id) ->
 * top k sorted.  No re-scoring (scores are already exact).  k <= 1024. */
int clm_topk_merge_sorted(const float* in_score, const int64_t* in_id, int nq, int lists, int k,
                          float* out_score, int64_t* out_id, void* stream);

/* The exchange step of the row-sharded search in ONE collective: every rank packs its local
 * top-k into a chunk of clm_topk_gather_chunk_bytes(nq, k) bytes -- int64 global ids [nq*k] (padded to an even
 * count), then fp32 scores [nq*k] -- and all-gathers the chunks rank-major into one buffer;
 * clm_topk_merge_gathered reads that buffer in place ([world] chunks) and emits the global top k, sorted.
 * Unused slots are (id -1, score -inf).  k <= 1024. */
size_t clm_topk_gather_chunk_bytes(int nq, int k);
int clm_topk_merge_gathered(const void* gathered, int world, int nq, int k, float* out_score,
                            int64_t* out_id, void* stream);

/* Exact top-k of one row of n fp32 scores, sorted descending (ties: lower id first among the rows kept;
 * which of several rows EQUAL to the k-th score are kept is unspecified, as for torch.topk): the reference's
 * torch.topk(sims, k) of src/embedding/search.py:99 / similarity.py:57 for the queries the fused scan hands
 * back and for k beyond the fused path.  k <= min(n, 2048). */
int clm_topk_row(const float* scores, int n, int k, int64_t id_offset, float* out_score, int64_t* out_id,
                 void* stream);

/* out[i] = q · E[i] in exact fp32 for ONE query: the reference's batch-1 scoring
 * (src/embedding/search.py:96, src/embedding/similarity.py:32).  Rows are expected to be
 * L2-normalised already (clm_l2norm).  HBM-streaming kernel, one warp per row. */
int clm_cosine_gemv(const float* q_f32, const float* index_f32, int n, int dim, float* out,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * Launch accounting and per-launch timing (measurement support for bench.py)
 * ---------------------------------------------------------------------------------- */
#define CLM_K_GEMM 0
#define CLM_K_ATTENTION 1
#define CLM_K_ELEMENTWISE 2
#define CLM_K_SEARCH 3
#define CLM_K_MERGE 4
/* total number of kernels this library has launched in this process */
unsigned long long clm_launch_count(void);
/* adds n (may be negative) to the total: the host mirror adds a captured graph's launch count on every
 * replay (replayed kernels do not pass through the library) and takes the capture pass itself back out */
void clm_launch_count_add(long long n);
/* 1 while per-launch timing is on (graph replay is bypassed then, so every launch gets its events) */
int clm_prof_is_enabled(void);
/* on != 0: bracket every subsequent launch with CUDA events on its stream (clears old records) */
int clm_prof_enable(int on);
/* device-synchronises, then sums the recorded launches of one kind: total ms, the algorithmic
 * FLOPs and bytes the launches were issued for, and their count */
int clm_prof_summary(int kind, double* ms, double* flops, double* bytes, int* launches);
/* the launch list: out[4i..4i+3] = {kind, flops, bytes, ms} in launch order; returns the number of
 * records available (only max_records are written), -1 on a CUDA error */
int clm_prof_records(double* out, int max_records);
/* the same list with start stamps: out[3i..3i+2] = {kind, start offset in ms from the first recorded launch,
 * duration in ms}; the gaps between consecutive launches are what the sum of kernel durations leaves out */
int clm_prof_timeline(double* out, int max_records);

#ifdef __cplusplus
}
#endif
#endif /* CLM_B200_H */
