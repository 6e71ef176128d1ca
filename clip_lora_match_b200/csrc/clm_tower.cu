// clm_tower.cu — host-side orchestration of one CLIP tower (vision or text) over the
// sm_100a kernels.  This is the native replacement for what models/clip_model.py:115,144
// reach through transformers' CLIPModel.get_image_features / get_text_features plus the
// peft LoRA wrappers: a fixed sequence of stream-ordered launches, no allocation, no host
// synchronisation (CUDA-graph capturable by the caller).
//
// Residual stream: fp32 [rows, D] ("h"), or bf16 after clm_tower_set_residual_dtype(CLM_OUT_BF16).  GEMM operands:
// bf16.  Per layer:
//   x   = LN1(h)                                   (bf16)
//   t   = x A_qkv^T                                (bf16 [rows,64], LoRA down-projection)
//   qkv = x W_qkv^T + t (sB)_qkv^T + b             (bf16 [rows,3D])   <- LoRA fused as K-extension
//   ao  = attention(qkv)                           (bf16 [rows,D])
//   h  += ao W_o^T (+ LoRA) + b_o                  (in place, in the stream's type: TMA reduce-add)
//   x   = LN2(h)
//   g   = quickgelu(x W_1^T + b_1)                 (bf16 [rows,mlp])
//   h  += g W_2^T + b_2
#include <stdlib.h>

#include <new>
#include <vector>

#include "clm_common.cuh"

int clm_gemm_launch(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                    const void* A2, int lda2, const void* W2, int ldw2, int K2, void* out, int ldo,
                    int out_dtype, const float* bias, const float* residual, int ldr, int epilogue,
                    cudaStream_t stream, const float* row_stats = nullptr, const float* col_sums = nullptr,
                    int ln_mode = 0);
int clm_attention_launch(const void* qkv, void* out, int batch, int tokens, int heads, int causal,
                         cudaStream_t stream, int reverse = 0);
int clm_row_stats_dir(const void* h_bf16, float* stats, int rows, int dim, float eps, void* stream, int reverse);

struct clm_tower {
  clm_tower_config cfg;
  clm_tower_weights w;
  std::vector<clm_layer_weights> layers;
  int kpad;  // padded im2col width (vision)
  int np;    // patches per image (vision)
  int h_dtype = CLM_OUT_F32;  // residual stream: fp32 (reference) or bf16 (clm_tower_set_residual_dtype)
  std::vector<clm_layer_ln_fold> folds;  // LayerNorm folded into QKV / fc1 (bf16 stream only); empty = off
};

namespace {

constexpr int kLoraBlock = 64;  // LoRA column counts are multiples of one K block of the GEMM

inline int max_lora_cols(const clm_tower_config& c) {
  int m = c.lora_cols_qkv;
  if (c.lora_cols_out > m) m = c.lora_cols_out;
  if (c.lora_cols_fc1 > m) m = c.lora_cols_fc1;
  if (c.lora_cols_fc2 > m) m = c.lora_cols_fc2;
  return m > 0 ? m : kLoraBlock;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  void* h;  // residual stream, fp32 or bf16 (clm_tower::h_dtype)
  __nv_bfloat16* x;
  __nv_bfloat16* ao;
  __nv_bfloat16* qkv;
  __nv_bfloat16* g;
  __nv_bfloat16* t;
  __nv_bfloat16* pooled;
  float* emb;
  int32_t* eos;
  float* stats;  // (mean, rstd) per row, folded-LayerNorm mode
  size_t total;
};

Workspace carve(const clm_tower* tw, int batch, uint8_t* base, int tokens = 0) {
  const clm_tower_config& c = tw->cfg;
  if (tokens <= 0) tokens = c.tokens;  // text passes may run on fewer positions than the context length
  const size_t rows = static_cast<size_t>(batch) * tokens;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  Workspace ws;
  ws.h = take(rows * c.width * (tw->h_dtype == CLM_OUT_BF16 ? 2 : 4));
  ws.x = reinterpret_cast<__nv_bfloat16*>(take(rows * c.width * 2));
  ws.ao = reinterpret_cast<__nv_bfloat16*>(take(rows * c.width * 2));
  size_t qkv_bytes = rows * 3 * c.width * 2;
  size_t g_bytes = rows * c.mlp * 2;
  if (c.kind == 0) {
    // vision prologue aliases: patch GEMM output (fp32) lives in the qkv region, the im2col
    // matrix in the fc1-output region
    const size_t prow = static_cast<size_t>(batch) * tw->np;
    if (prow * c.width * 4 > qkv_bytes) qkv_bytes = prow * c.width * 4;
    if (prow * tw->kpad * 2 > g_bytes) g_bytes = prow * tw->kpad * 2;
  }
  ws.qkv = reinterpret_cast<__nv_bfloat16*>(take(qkv_bytes));
  ws.g = reinterpret_cast<__nv_bfloat16*>(take(g_bytes));
  ws.t = reinterpret_cast<__nv_bfloat16*>(take(rows * max_lora_cols(c) * 2));
  ws.pooled = reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(batch) * c.width * 2));
  ws.emb = reinterpret_cast<float*>(take(static_cast<size_t>(batch) * c.proj_dim * 4));
  ws.eos = reinterpret_cast<int32_t*>(take(static_cast<size_t>(batch) * 4));
  ws.stats = reinterpret_cast<float*>(take(rows * 8));
  ws.total = off;
  return ws;
}

int max_batch_for(const clm_tower* tw, size_t bytes, int want, int tokens = 0) {
  if (carve(tw, want, nullptr, tokens).total <= bytes) return want;
  int lo = 0, hi = want;  // largest batch that fits
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (carve(tw, mid, nullptr, tokens).total <= bytes) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

#define CLM_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != CLM_OK) return _rc; \
  } while (0)

// the transformer layers + pooling + projection + L2 normalise, on rows already embedded in ws.h
int run_layers(clm_tower* tw, const Workspace& ws, int batch, const int32_t* pool_idx, float* out_emb,
               int normalize, cudaStream_t s, int tokens = 0) {
  const clm_tower_config& c = tw->cfg;
  if (tokens <= 0) tokens = c.tokens;
  const int rows = batch * tokens;
  const int D = c.width;
  void* sv = static_cast<void*>(s);
  const int hd = tw->h_dtype;
  const float* hres = static_cast<const float*>(ws.h);  // in-place residual: same pointer as `out`, type per hd
  const bool fold = hd == CLM_OUT_BF16 && !tw->folds.empty();
  // Serpentine order: every kernel of the layer streams more bytes than the L2 holds, so a kernel that walks its rows
  // in the same direction as its predecessor starts on rows that were evicted long ago.  Alternating the direction
  // (GEMM tile list / attention items / row-statistics blocks from the end) makes each kernel start on the ~100 MB
  // its predecessor touched last.  Scheduling only: results are unchanged.  Measured neutral (ViT-L/14 83.8 -> 83.4 ms,
  // ViT-B/16 39.9 -> 40.2 ms, both inside the run-to-run spread; the row-statistics pass does not get faster at all:
  // profiles/r2_serpentine_ab.log), so it is OFF unless CLM_SERPENTINE=1.
  const char* serp_env = getenv("CLM_SERPENTINE");  // read per call: same-process A/B
  const bool serp = fold && serp_env && serp_env[0] == '1';
  int dir = 0;
  auto next_rev = [&]() { const int r = dir; if (serp) dir ^= 1; return serp ? r : 0; };
  auto epi = [&](int flags) { return flags | (next_rev() ? CLM_EPI_REVERSE : 0); };
  for (int l = 0; l < c.layers; ++l) {
    const clm_layer_weights& L = tw->layers[l];
    // LoRA (unmerged): t = x A_cat^T is a skinny GEMM, then (t, (s B)_cat) ride along as extra K blocks
    const int cq = (c.lora_cols_qkv > 0 && L.lora_a_qkv && L.lora_b_qkv) ? c.lora_cols_qkv : 0;
    const clm_layer_ln_fold* F = fold ? &tw->folds[l] : nullptr;
    if (F) {
      // LayerNorm folded into the GEMMs: the operand is the raw bf16 stream, the epilogue normalises (clm_gemm_ln_epi)
      CLM_TRY(clm_row_stats_dir(ws.h, ws.stats, rows, D, c.ln_eps, sv, next_rev()));
      if (cq)
        CLM_TRY(clm_gemm_launch(ws.h, D, F->lora_a_qkv_g, D, rows, cq, D, nullptr, 0, nullptr, 0, 0, ws.t, cq,
                                CLM_OUT_BF16, nullptr, nullptr, 0, epi(CLM_EPI_NONE), s, ws.stats, F->s_a_qkv, 2));
      CLM_TRY(clm_gemm_launch(ws.h, D, F->w_qkv_g, D, rows, 3 * D, D, cq ? ws.t : nullptr, cq,
                              cq ? L.lora_b_qkv : nullptr, cq, cq, ws.qkv, 3 * D, CLM_OUT_BF16, F->b_qkv_f, nullptr,
                              0, epi(CLM_EPI_NONE), s, ws.stats, F->s_qkv, 1));
    } else {
      CLM_TRY(clm_layernorm_ex(ws.h, hd, L.ln1_g, L.ln1_b, ws.x, rows, D, c.ln_eps, sv));
      if (cq)
        CLM_TRY(clm_gemm_launch(ws.x, D, L.lora_a_qkv, D, rows, cq, D, nullptr, 0, nullptr, 0, 0,
                                ws.t, cq, CLM_OUT_BF16, nullptr, nullptr, 0, CLM_EPI_NONE, s));
      CLM_TRY(clm_gemm_launch(ws.x, D, L.w_qkv, D, rows, 3 * D, D, cq ? ws.t : nullptr, cq,
                              cq ? L.lora_b_qkv : nullptr, cq, cq, ws.qkv, 3 * D,
                              CLM_OUT_BF16, L.b_qkv, nullptr, 0, CLM_EPI_NONE, s));
    }
    CLM_TRY(clm_attention_launch(ws.qkv, ws.ao, batch, tokens, c.heads, c.kind == 1, s, next_rev()));
    const int co = (c.lora_cols_out > 0 && L.lora_a_o && L.lora_b_o) ? c.lora_cols_out : 0;
    if (co)
      CLM_TRY(clm_gemm_launch(ws.ao, D, L.lora_a_o, D, rows, co, D, nullptr, 0, nullptr, 0, 0,
                              ws.t, co, CLM_OUT_BF16, nullptr, nullptr, 0, CLM_EPI_NONE, s));
    CLM_TRY(clm_gemm_launch(ws.ao, D, L.w_o, D, rows, D, D, co ? ws.t : nullptr, co,
                            co ? L.lora_b_o : nullptr, co, co, ws.h, D,
                            hd, L.b_o, hres, D, epi(CLM_EPI_NONE), s));
    const int c1 = (c.lora_cols_fc1 > 0 && L.lora_a_fc1 && L.lora_b_fc1) ? c.lora_cols_fc1 : 0;
    if (F) {
      CLM_TRY(clm_row_stats_dir(ws.h, ws.stats, rows, D, c.ln_eps, sv, next_rev()));
      if (c1)
        CLM_TRY(clm_gemm_launch(ws.h, D, F->lora_a_fc1_g, D, rows, c1, D, nullptr, 0, nullptr, 0, 0, ws.t, c1,
                                CLM_OUT_BF16, nullptr, nullptr, 0, epi(CLM_EPI_NONE), s, ws.stats, F->s_a_fc1, 2));
      CLM_TRY(clm_gemm_launch(ws.h, D, F->w_fc1_g, D, rows, c.mlp, D, c1 ? ws.t : nullptr, c1,
                              c1 ? L.lora_b_fc1 : nullptr, c1, c1, ws.g, c.mlp, CLM_OUT_BF16, F->b_fc1_f, nullptr, 0,
                              epi(CLM_EPI_QUICKGELU), s, ws.stats, F->s_fc1, 1));
    } else {
      CLM_TRY(clm_layernorm_ex(ws.h, hd, L.ln2_g, L.ln2_b, ws.x, rows, D, c.ln_eps, sv));
      if (c1)
        CLM_TRY(clm_gemm_launch(ws.x, D, L.lora_a_fc1, D, rows, c1, D, nullptr, 0, nullptr, 0, 0,
                                ws.t, c1, CLM_OUT_BF16, nullptr, nullptr, 0, CLM_EPI_NONE, s));
      CLM_TRY(clm_gemm_launch(ws.x, D, L.w_fc1, D, rows, c.mlp, D, c1 ? ws.t : nullptr, c1,
                              c1 ? L.lora_b_fc1 : nullptr, c1, c1, ws.g,
                              c.mlp, CLM_OUT_BF16, L.b_fc1, nullptr, 0, CLM_EPI_QUICKGELU, s));
    }
    const int c2 = (c.lora_cols_fc2 > 0 && L.lora_a_fc2 && L.lora_b_fc2) ? c.lora_cols_fc2 : 0;
    if (c2)
      CLM_TRY(clm_gemm_launch(ws.g, c.mlp, L.lora_a_fc2, c.mlp, rows, c2, c.mlp, nullptr, 0, nullptr, 0, 0,
                              ws.t, c2, CLM_OUT_BF16, nullptr, nullptr, 0, CLM_EPI_NONE, s));
    CLM_TRY(clm_gemm_launch(ws.g, c.mlp, L.w_fc2, c.mlp, rows, D, c.mlp, c2 ? ws.t : nullptr, c2,
                            c2 ? L.lora_b_fc2 : nullptr, c2, c2, ws.h, D, hd, L.b_fc2, hres, D,
                            epi(CLM_EPI_NONE), s));
  }
  CLM_TRY(clm_pool_ln_ex(ws.h, hd, pool_idx, tw->w.final_ln_g, tw->w.final_ln_b, ws.pooled, batch, tokens,
                         D, c.ln_eps, sv));
  float* proj_out = normalize ? ws.emb : out_emb;
  CLM_TRY(clm_gemm_launch(ws.pooled, D, tw->w.proj_w, D, batch, c.proj_dim, D, nullptr, 0, nullptr,
                          0, 0, proj_out, c.proj_dim, CLM_OUT_F32, nullptr, nullptr, 0, CLM_EPI_NONE, s));
  if (normalize) CLM_TRY(clm_l2norm(ws.emb, out_emb, nullptr, batch, c.proj_dim, sv));
  return CLM_OK;
}

}  // namespace

extern "C" int clm_tower_create(const clm_tower_config* cfg, const clm_tower_weights* w,
                                const clm_layer_weights* layers, clm_tower** out) {
  CLM_REQUIRE(cfg && w && layers && out, "clm_tower_create: null argument");
  CLM_REQUIRE(cfg->kind == 0 || cfg->kind == 1, "clm_tower_create: kind must be 0 (vision) or 1 (text)");
  CLM_REQUIRE(cfg->heads * 64 == cfg->width, "clm_tower_create: head_dim must be 64 (width=%d heads=%d)",
              cfg->width, cfg->heads);
  CLM_REQUIRE(cfg->width % 128 == 0 && cfg->mlp % 64 == 0 && cfg->proj_dim % 8 == 0,
              "clm_tower_create: unsupported dims width=%d mlp=%d proj=%d", cfg->width, cfg->mlp,
              cfg->proj_dim);
  CLM_REQUIRE(cfg->layers > 0 && cfg->tokens > 0 && cfg->tokens <= 384,
              "clm_tower_create: bad layers/tokens (tokens <= 384: the attention kernel's limit)");
  CLM_REQUIRE(cfg->lora_cols_qkv >= 0 && cfg->lora_cols_qkv % kLoraBlock == 0 && cfg->lora_cols_out >= 0 &&
                  cfg->lora_cols_out % kLoraBlock == 0 && cfg->lora_cols_fc1 >= 0 &&
                  cfg->lora_cols_fc1 % kLoraBlock == 0 && cfg->lora_cols_fc2 >= 0 &&
                  cfg->lora_cols_fc2 % kLoraBlock == 0,
              "clm_tower_create: lora_cols must be 0 or multiples of %d", kLoraBlock);
  clm_tower* t = new (std::nothrow) clm_tower();
  CLM_REQUIRE(t != nullptr, "clm_tower_create: out of host memory");
  t->cfg = *cfg;
  t->w = *w;
  t->layers.assign(layers, layers + cfg->layers);
  t->kpad = 0;
  t->np = 0;
  if (cfg->kind == 0) {
    if (cfg->patch <= 0 || cfg->image % cfg->patch != 0) {
      delete t;
      clm_set_error("clm_tower_create: bad image/patch %d/%d", cfg->image, cfg->patch);
      return CLM_ERR_INVALID;
    }
    const int g = cfg->image / cfg->patch;
    t->np = g * g;
    if (t->np + 1 != cfg->tokens) {
      delete t;
      clm_set_error("clm_tower_create: tokens=%d != 1 + (image/patch)^2 = %d", cfg->tokens, t->np + 1);
      return CLM_ERR_INVALID;
    }
    t->kpad = (3 * cfg->patch * cfg->patch + 63) / 64 * 64;
  }
  *out = t;
  return CLM_OK;
}

extern "C" void clm_tower_destroy(clm_tower* t) { delete t; }

extern "C" int clm_tower_set_residual_dtype(clm_tower* t, int dtype) {
  CLM_REQUIRE(t != nullptr, "clm_tower_set_residual_dtype: null tower");
  CLM_REQUIRE(dtype == CLM_OUT_F32 || dtype == CLM_OUT_BF16,
              "clm_tower_set_residual_dtype: dtype must be CLM_OUT_F32 (%d) or CLM_OUT_BF16 (%d)", CLM_OUT_F32,
              CLM_OUT_BF16);
  t->h_dtype = dtype;
  return CLM_OK;
}

extern "C" int clm_tower_residual_dtype(const clm_tower* t) { return t ? t->h_dtype : -1; }

extern "C" int clm_tower_set_ln_fold(clm_tower* t, const clm_layer_ln_fold* folds) {
  CLM_REQUIRE(t != nullptr, "clm_tower_set_ln_fold: null tower");
  if (!folds) {
    t->folds.clear();
    return CLM_OK;
  }
  const clm_tower_config& c = t->cfg;
  for (int l = 0; l < c.layers; ++l) {
    const clm_layer_ln_fold& f = folds[l];
    const clm_layer_weights& L = t->layers[l];
    CLM_REQUIRE(f.w_qkv_g && f.s_qkv && f.b_qkv_f && f.w_fc1_g && f.s_fc1 && f.b_fc1_f,
                "clm_tower_set_ln_fold: layer %d: folded QKV / fc1 weights, column sums and biases are required", l);
    const bool need_q = c.lora_cols_qkv > 0 && L.lora_a_qkv && L.lora_b_qkv;
    const bool need_1 = c.lora_cols_fc1 > 0 && L.lora_a_fc1 && L.lora_b_fc1;
    CLM_REQUIRE(!need_q || (f.lora_a_qkv_g && f.s_a_qkv),
                "clm_tower_set_ln_fold: layer %d has a QKV adapter but no folded down-projection", l);
    CLM_REQUIRE(!need_1 || (f.lora_a_fc1_g && f.s_a_fc1),
                "clm_tower_set_ln_fold: layer %d has an fc1 adapter but no folded down-projection", l);
  }
  t->folds.assign(folds, folds + c.layers);
  return CLM_OK;
}

extern "C" size_t clm_tower_workspace_bytes(const clm_tower* t, int batch) {
  if (!t || batch <= 0) return 0;
  return carve(t, batch, nullptr).total;
}

extern "C" int clm_encode_image(clm_tower* t, const float* pixel_values, int batch, float* out_emb,
                                int normalize, void* workspace, size_t workspace_bytes, void* stream) {
  CLM_REQUIRE(t && t->cfg.kind == 0, "clm_encode_image: not a vision tower");
  CLM_REQUIRE(batch >= 0 && (batch == 0 || (pixel_values && out_emb && workspace)),
              "clm_encode_image: null argument");
  if (batch == 0) return CLM_OK;
  const clm_tower_config& c = t->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int mb = max_batch_for(t, workspace_bytes, batch);
  CLM_REQUIRE(mb > 0, "clm_encode_image: workspace of %zu bytes too small for one image (need %zu)",
              workspace_bytes, carve(t, 1, nullptr).total);
  const size_t img_elems = 3ull * c.image * c.image;
  for (int b0 = 0; b0 < batch; b0 += mb) {
    const int nb = (batch - b0) < mb ? (batch - b0) : mb;
    Workspace ws = carve(t, nb, static_cast<uint8_t*>(workspace));
    float* patch_out = reinterpret_cast<float*>(ws.qkv);
    __nv_bfloat16* patches = ws.g;
    CLM_TRY(clm_patch_im2col(pixel_values + b0 * img_elems, patches, nb, c.image, c.patch, t->kpad, s));
    CLM_TRY(clm_gemm_launch(patches, t->kpad, t->w.patch_w, t->kpad, nb * t->np, c.width, t->kpad,
                            nullptr, 0, nullptr, 0, 0, patch_out, c.width, CLM_OUT_F32, nullptr,
                            nullptr, 0, CLM_EPI_NONE, s));
    CLM_TRY(clm_vision_embed_ln_ex(patch_out, t->w.class_emb, t->w.pos_emb, t->w.pre_ln_g, t->w.pre_ln_b,
                                   ws.h, t->h_dtype, nb, t->np, c.width, c.ln_eps, s));
    CLM_TRY(run_layers(t, ws, nb, nullptr, out_emb + static_cast<size_t>(b0) * c.proj_dim, normalize, s));
  }
  return CLM_OK;
}

extern "C" int clm_encode_text_len(clm_tower* t, const int32_t* ids, int batch, int tokens, float* out_emb,
                                   int normalize, void* workspace, size_t workspace_bytes, void* stream) {
  CLM_REQUIRE(t && t->cfg.kind == 1, "clm_encode_text: not a text tower");
  CLM_REQUIRE(batch >= 0 && (batch == 0 || (ids && out_emb && workspace)),
              "clm_encode_text: null argument");
  CLM_REQUIRE(tokens >= 1 && tokens <= t->cfg.tokens, "clm_encode_text: tokens=%d must be in [1,%d]", tokens,
              t->cfg.tokens);
  if (batch == 0) return CLM_OK;
  const clm_tower_config& c = t->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int mb = max_batch_for(t, workspace_bytes, batch, tokens);
  CLM_REQUIRE(mb > 0, "clm_encode_text: workspace of %zu bytes too small for one caption (need %zu)",
              workspace_bytes, carve(t, 1, nullptr, tokens).total);
  for (int b0 = 0; b0 < batch; b0 += mb) {
    const int nb = (batch - b0) < mb ? (batch - b0) : mb;
    Workspace ws = carve(t, nb, static_cast<uint8_t*>(workspace), tokens);
    CLM_TRY(clm_embed_text_ex(ids + static_cast<size_t>(b0) * tokens, t->w.tok_emb, t->w.pos_emb, ws.h, t->h_dtype,
                              ws.eos, nb, tokens, c.width, c.vocab, c.eos_id, s));
    CLM_TRY(run_layers(t, ws, nb, ws.eos, out_emb + static_cast<size_t>(b0) * c.proj_dim, normalize, s, tokens));
  }
  return CLM_OK;
}

extern "C" int clm_encode_text(clm_tower* t, const int32_t* ids, int batch, float* out_emb,
                               int normalize, void* workspace, size_t workspace_bytes, void* stream) {
  CLM_REQUIRE(t && t->cfg.kind == 1, "clm_encode_text: not a text tower");
  return clm_encode_text_len(t, ids, batch, t->cfg.tokens, out_emb, normalize, workspace, workspace_bytes, stream);
}
