/* pg_b200 — C ABI of the B200-native PaliGemma hot path.
 *
 * The reference (PhilipWilliamVentura/multimodal-financial-analysis-tool-using-paligemma)
 * is pure Python/PyTorch and has no FFI of its own: its boundary is the
 * Python class surface of modeling_gemma.py / modeling_siglip.py.  Our drop-in modules of
 * the same names call these entry points (ctypes); every entry names
 * the reference code it replaces.
 *
 * Conventions: plain device pointers and sizes, no torch types; `stream` is a
 * cudaStream_t passed as void*; every call is asynchronous on that stream, allocates
 * nothing and never synchronises; return 0 on success, non-zero otherwise with a message
 * in pg_last_error().  `dtype` selects the model dtype of activations and weights
 * (PG_F32 is the fp32 verification mode: true fp32 FMA, no TF32).  Matrices are row-major;
 * weights are [out_features, in_features] exactly as torch.nn.Linear stores them.
 */
#ifndef PG_B200_H
#define PG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PG_F32 = 0, PG_BF16 = 1, PG_F16 = 2 };
enum { PG_OK = 0, PG_ERR_INVALID = 1, PG_ERR_CUDA = 2 };

/* pg_gemm epilogues (all round to the model dtype where the reference materialises a tensor) */
enum {
  PG_EPI_NONE = 0,      /* C = A W^T                                  */
  PG_EPI_BIAS = 1,      /* C = A W^T + b                  nn.Linear   */
  PG_EPI_BIAS_GELU = 2, /* C = gelu_tanh(A W^T + b)       SiglipMLP fc1 */
  PG_EPI_BIAS_RES = 3,  /* C = (A W^T + b) + R            out_proj / fc2 / patch-embed+pos */
  PG_EPI_RES = 4,       /* C = (A W^T) + R                o_proj / down_proj */
  PG_EPI_GEGLU = 5      /* C = gelu_tanh(A Wg^T) * (A Wu^T), W = [Wg; Wu]  GemmaMLP */
};

/* One exchange of the tensor-parallel decoder over NVLink peer memory (csrc/tp_exchange.cuh; SURVEY.md §8e: the
 * reference is single-device, the sharding is the north star's).  The PRODUCER kernel stores this rank's fp32 partial
 * as {value, sequence number} words into slot (parity, rank) of EVERY rank's buffer; the CONSUMER kernel waits for all
 * `size` slots of its local buffer and sums them in rank order.  Passed by pointer (host memory, read at launch);
 * NULL means "not tensor parallel" everywhere a parameter of this type appears. */
typedef struct pg_tp_exchange {
  const void* peers;    /* device array of `size` pointers: base address of every rank's exchange buffer        */
  const int* epoch;     /* device int: decode-step counter of this rank, advanced by pg_tp_begin_step           */
  int* err_dev;         /* device int: set when a wait timed out (lost peer); later waits return at once         */
  int* err_host;        /* pinned host int mirrored from err_dev (may be NULL)                                   */
  long long region_off; /* byte offset of this exchange's region inside the buffers                              */
  long long slot_bytes; /* bytes of one (parity, source rank) slot: 8 per fp32 word                              */
  int rank, size;       /* this rank, number of ranks (2..8)                                                     */
  int index, stride;    /* sequence number = *epoch * stride + index, 1 <= index < stride; exchanges that follow  */
                        /* one another in the same region must alternate the parity of `index`                   */
} pg_tp_exchange;

const char* pg_last_error(void);
int pg_abi_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
unsigned long long pg_launch_count(void);

/* ---- embedding + image/text merge + sqrt(D) normaliser --------------------------------
 * Replaces nn.Embedding lookup (modeling_gemma.py:565), _merge_input_ids_with_image_features
 * (:476-500) and the `hidden_states * normalizer` of GemmaModel.forward (:367-368).
 * out[t] = 0 for pad ids; rnd(rnd(img[k]/img_div) * normalizer) for the k-th image token in
 * row-major order; rnd(emb[id] * normalizer) otherwise.  ids may be int64 device memory.
 * If an image token has no image row left (reference: masked_scatter raises) or an id is out
 * of range, *err_flag (device int, may be NULL) is set to 1 and the row is zero. */
int pg_embed_merge(void* out, const int64_t* ids, const void* emb, const void* img_feats,
                   int n_tokens, int D, int64_t vocab, int64_t image_token_id, int64_t pad_id,
                   int n_img_rows, float img_div, float normalizer, int* err_flag,
                   int dtype, void* stream);

/* GemmaRMSNorm.forward (modeling_gemma.py:114-120) */
int pg_rmsnorm(void* out, const void* x, const void* w, int rows, int D, float eps,
               int dtype, void* stream);
/* nn.LayerNorm (modeling_siglip.py:175,177,234) */
int pg_layernorm(void* out, const void* x, const void* w, const void* b, int rows, int D,
                 float eps, int dtype, void* stream);

/* Image pre-processing on the device (processing_paligemma.py:13-50).  pg_resample_u8 is one pass of Pillow's 8-bit
 * bicubic resize (Image.resize(..., BICUBIC), libImaging/Resample.c) along one axis of an interleaved uint8 image:
 * vertical == 0: in (rows, n_in, C) -> out (rows, n_out, C); vertical != 0: in (n_in, rows, C) -> out (n_out, rows, C).
 * bounds int32 [n_out][2] = (first input index, tap count), kk int32 [n_out][ksize] = 22-bit fixed-point taps, both
 * from the host (pg_b200/preprocess.py::resample_coeffs).  Horizontal pass first, then vertical, as Pillow does.
 * pg_u8_to_chw: out[c][y][x] = lut[in[y][x][c]] in the model dtype; lut float32 [256] = ((b/255) - mean) / std with
 * the reference's dtypes (pg_b200/preprocess.py::value_table). */
int pg_resample_u8(uint8_t* out, const uint8_t* in, const int32_t* bounds, const int32_t* kk, int ksize,
                   int rows, int n_in, int n_out, int channels, int vertical, void* stream);
int pg_u8_to_chw(void* out, const uint8_t* in, const float* lut, int H, int W, int channels, int dtype,
                 void* stream);

/* SiglipVisionEmbeddings patch conv as im2col (modeling_siglip.py:45-51,67-73): stride == kernel
 * so it is a pure permutation: out[(b*P + py*G + px), c*p*p + ky*p + kx], row stride ld_out
 * (columns beyond C*p*p are zero-filled up to ld_out). */
int pg_im2col(void* out, const void* pixels, int B, int C, int H, int W, int p, int ld_out,
              int dtype, void* stream);

/* C[M,N] = A[M,K] W[N,K]^T with a fused epilogue; fp32 accumulation.  R row index is
 * m % res_mod when res_mod > 0 (position-embedding broadcast over the batch).  out_f32 != 0
 * stores fp32 (lm_head `.float()`, modeling_gemma.py:417-418).  impl: 0 = auto, 1 = SIMT
 * (any dtype), 2 = tcgen05/TMA (bf16/f16 only; picks the CTA-pair, single-CTA, skinny or split-K kernel by shape),
 * 3 = the CTA-pair swap-AB kernel for 129..512 rows (weights as the M operand, split-K through pg_set_workspace;
 * auto takes it for the projections where it measured faster).  Replaces every nn.Linear / matmul call site
 * of SURVEY.md §2.3 on the prefill and vision paths. */
int pg_gemm(void* C, const void* A, const void* W, const void* bias, const void* R,
            int M, int N, int K, int lda, int ldw, int ldc, int ldr, int res_mod,
            int epilogue, int out_f32, int impl, int dtype, void* stream);

/* Scratch memory for GEMMs that split K over CTA pairs (prompt-sized row counts: csrc/gemm_tcgen05_swap.cu): a
 * caller-owned device buffer the library may overwrite during any pg_gemm call on this device (fp32 partials
 * [splits][M][N]); without one (or with one that is too small for a problem) pg_gemm takes its other kernels.
 * The library never allocates. */
int pg_set_workspace(void* ptr, long long bytes);

/* RoPE on q and k + append of K,V to the paged cache (modeling_gemma.py:155-199, 23-36).
 * qkv: [B*q_len, (nq+2*nkv)*hd] from the fused projection; q_out: [B*q_len, nq*hd].
 * Token (b,i) gets position positions[b*q_len+i] (clamped to [0,max_pos-1]) and cache slot
 * slot_base[b]+i.  Pools are [num_pages, page_size, nkv*hd]. */
int pg_rope_append(void* q_out, const void* qkv, const float* inv_freq, const int32_t* positions,
                   void* k_pool, void* v_pool, const int32_t* page_table, int pt_stride,
                   int page_size, const int32_t* slot_base, int B, int q_len, int nq, int nkv,
                   int hd, int max_pos, int dtype, void* stream);

/* Unmasked softmax attention (the reference's additive mask is all zeros: modeling_gemma.py:
 * 506-514; SigLIP has none: modeling_siglip.py:116-131).  fp32 softmax.  q: [B*q_len, ld_q],
 * head h at column h*hd.  K/V either contiguous (page_table NULL: row (b,j) at
 * base + b*kv_batch_stride + j*ld_kv elements) or paged.  kv_len: device int32[B] or NULL
 * (then kv_len_const).  kv_len_add is added to the device value.  scale_mode 0: s*scale
 * (SigLIP), 1: s/scale (Gemma). */
int pg_attention(void* out, int ld_out, const void* q, int ld_q, const void* k, const void* v,
                 int ld_kv, long long kv_batch_stride, const int32_t* page_table, int pt_stride,
                 int page_size, const int32_t* kv_len, int kv_len_const, int kv_len_add, int B,
                 int q_len, int n_heads, int n_kv_heads, int hd, float scale, int scale_mode,
                 int dtype, void* stream);

/* The same attention on the tcgen05 tensor cores (16-bit dtypes; prefill, cache-off recompute, SigLIP):
 * QK^T and PV as tcgen05.mma with TMEM accumulators, Q/K/V tiles by TMA, two-pass softmax (row max, then
 * exp / PV) with one thread per query row.  q, k, v are row-major matrices (`*_rows` x `ld_*`); head h sits
 * at columns *_col0 + h*hd_stride .. +hd, and a head row must be padded (with zeros) to a multiple of 64
 * columns unless hd is 128 or 256.  Paged K/V needs page_size == key tile (128 for hd <= 128, 64 above).
 * Output is packed: head h at column h*hd of [B*q_len, ld_out]. */
int pg_attention_tc(void* out, int ld_out, const void* q, long long q_rows, int ld_q, int q_col0,
                    const void* k, const void* v, long long kv_rows, int ld_kv, int k_col0, int v_col0,
                    int hd_stride, long long kv_batch_rows, const int32_t* page_table, int pt_stride,
                    int page_size, const int32_t* kv_len, int kv_len_const, int kv_len_add, int B, int q_len,
                    int n_heads, int n_kv_heads, int hd, float scale, int scale_mode, int dtype, void* stream);

/* ---- decode path (q_len == 1 per sequence), B <= PG_MAX_DECODE_BATCH per call ----------- */
#define PG_MAX_DECODE_BATCH 8

/* L2 prefetch chain: registers [ptr, ptr+bytes) (16-byte aligned; normally the NEXT kernel's
 * weights) with the calling thread; the next pg_decode_* / pg_gemv_res launch from this thread
 * consumes it and starts by issuing cp.async.bulk.prefetch.L2 for that region, one slice per CTA.
 * Weights never depend on activations, so HBM keeps streaming through the small latency-bound
 * kernels and across kernel boundaries.  Purely a performance hint: results never change. */
int pg_set_next_prefetch(const void* ptr, long long bytes);

/* input RMSNorm + fused q/k/v projection + RoPE + KV append (GemmaDecoderLayer.forward
 * :314-316, GemmaAttention.forward :241-259).  x: [B,D] residual stream. positions/kv_len:
 * device int32[B]; K,V go to slot kv_len[b]. */
int pg_decode_qkv(void* q_out, const void* x, const void* norm_w, const void* w_qkv,
                  const float* inv_freq, const int32_t* positions, void* k_pool, void* v_pool,
                  const int32_t* page_table, int pt_stride, int page_size, const int32_t* kv_len,
                  int B, int D, int nq, int nkv, int hd, float eps, int max_pos,
                  const pg_tp_exchange* ex, void* x_out, int dtype, void* stream);
/* ex != NULL (tensor parallel; same meaning in pg_decode_gateup / pg_decode_lmhead): the residual stream this
 * kernel normalises is x_new = rnd(x + rnd(sum over ranks of the fp32 partials of exchange `ex`)), i.e. the
 * all-reduce of the previous o_proj / down_proj and the `residual + hidden_states` of modeling_gemma.py:327,336
 * happen in this kernel's prologue; x_new is also written to x_out [B,D] (must not alias x). */

/* split-K MQA attention over the paged cache for one new token per sequence
 * (modeling_gemma.py:262-288).  Attends kv_len[b]+kv_len_add entries.  ws: fp32 workspace of
 * pg_decode_attention_ws_floats() floats; counters: int32[B*nkv], zero before first use. */
long long pg_decode_attention_ws_floats(int B, int nq, int hd, int max_splits);
int pg_decode_attention(void* out, const void* q, const void* k_pool, const void* v_pool,
                        const int32_t* page_table, int pt_stride, int page_size,
                        const int32_t* kv_len, int kv_len_add, int B, int nq, int nkv, int hd,
                        float scale_div, float* ws, int* counters, int max_splits, int dtype,
                        void* stream);

/* out[B,N] = (x[B,K] W[N,K]^T) + R  — o_proj / down_proj with the residual add
 * (modeling_gemma.py:291,327 and :134,336).  R may be NULL.  ex != NULL (tensor parallel: W holds this rank's K
 * columns): the unrounded fp32 partial x W^T goes to word b*N+n of this rank's slot in every rank's exchange buffer
 * instead; out and R are ignored (the consumer adds the residual). */
int pg_gemv_res(void* out, const void* x, const void* W, const void* R, int B, int N, int K,
                const pg_tp_exchange* ex, int dtype, void* stream);

/* post-attention RMSNorm + gate/up projections + GeGLU (modeling_gemma.py:332,134).
 * w_gu = [Wgate; Wup] : [2F, D];  out: [B,F]. */
int pg_decode_gateup(void* out, const void* x, const void* norm_w, const void* w_gu, int B,
                     int D, int F, float eps, const pg_tp_exchange* ex, void* x_out, int dtype, void* stream);

/* final RMSNorm + tied lm_head + fp32 logits + greedy argmax (modeling_gemma.py:379,417-418;
 * inference.py:68).  logits: fp32 [B,V] (values rounded to the model dtype first, as the
 * reference's `.float()` of a model-dtype tensor).  argmax_keys: device u64[B], must hold 0
 * on entry; decoded by pg_step_advance.  Ties go to the lowest index. */
int pg_decode_lmhead(float* logits, const void* x, const void* norm_w, const void* w_emb,
                     int B, int D, int64_t V, float eps, unsigned long long* argmax_keys,
                     const pg_tp_exchange* ex, void* x_out, int dtype, void* stream);

/* End of a decode step, all on device (replaces the host side of inference.py:68-78):
 * token[b] = argmax(keys[b]) (or sampled[b] if sampled != NULL); next_ids[b] = token;
 * history[b*hist_stride + step_counter] = token; kv_len[b] += 1; positions[b] += 1;
 * keys[b] = 0; *step_counter += 1 (by thread 0). */
int pg_step_advance(int64_t* next_ids, int64_t* history, int hist_stride, int* step_counter,
                    unsigned long long* keys, const int64_t* sampled, int32_t* kv_len,
                    int32_t* positions, int B, const pg_tp_exchange* keys_ex, void* stream);
/* keys_ex != NULL (tensor parallel, greedy): token[b] is decoded from the largest key over the ranks' vocabulary
 * shards, waited for in the key exchange filled by pg_tp_keys_push (ties: lowest global index). */

/* Head of a cached decode step driven through the reference API (inference.py:56-63,
 * modeling_gemma.py:557-559): ids_dst[b] = ids_src[b]; positions[b] = position; and the reference's
 * "The input cannot be padded" check without a host synchronisation: *bad_flag (any device-visible int,
 * e.g. pinned host memory) is set to 1 when one of the mask_n attention-mask elements is not 1.
 * mask_kind: 0 int64, 1 float32, 2 int32, 3 bfloat16, 4 float16, 5 uint8/bool, 6 float64; mask may be NULL. */
int pg_decode_inputs(int64_t* ids_dst, const int64_t* ids_src, int32_t* positions, int position,
                     const void* mask, int mask_kind, long long mask_n, int* bad_flag, int B, void* stream);

/* torch.argmax(logits, -1) over fp32 [B,V] (inference.py:68).  keys_ws: device u64[B] holding 0
 * on entry (left 0 on exit).  out == NULL: stop after the packed (value, index) keys, which stay in keys_ws
 * (tensor parallel: the shard's keys go to pg_tp_keys_push). */
int pg_argmax(int64_t* out, const float* logits, unsigned long long* keys_ws, int B, int64_t V,
              void* stream);

/* softmax(logits/temperature) + nucleus (top-p) sampling (inference.py:15-24,65-66).
 * The nucleus is the reference's: descending order, keep while (cumsum - p_i) <= top_p;
 * the draw uses counter-based Philox (seed, *rng_offset + b) rather than torch's RNG stream.
 * probs_ws: fp32 [B,V] scratch. nucleus_size (int32[B], may be NULL) reports the kept count. */
int pg_top_p_sample(int64_t* out, const float* logits, float* probs_ws, int B, int64_t V,
                    float temperature, float top_p, unsigned long long seed,
                    const int* rng_offset, int* nucleus_size, void* stream);

/* ---- tensor-parallel decoder plumbing (pg_tp_exchange above) --------------------------------------------------
 * pg_tp_begin_step: *epoch += 1, once at the head of every decode step (before its first producer).
 * pg_tp_push: producer for a partial that a GEMM left in local memory (batched decode step): partial[n] fp32 ->
 *   words 0..n-1 of this rank's slot in every rank's buffer.
 * pg_rmsnorm_reduce: consumer for that step, one CTA per row: x_new = rnd(x_in + rnd(sum of partials)) -> x_out,
 *   out = GemmaRMSNorm(x_new) (modeling_gemma.py:114-120).  x_out must not alias x_in.
 * pg_tp_keys_push: vocabulary-split lm_head: keys[b] (value, LOCAL index) -> (value, index + rank * v_local) in
 *   every rank's key slot; pg_step_advance(keys_ex) picks the winner. */
int pg_tp_begin_step(int* epoch, void* stream);
int pg_tp_push(const float* partial, long long n, const pg_tp_exchange* ex, void* stream);
int pg_rmsnorm_reduce(void* out, void* x_out, const void* x_in, const void* w, int rows, int D, float eps,
                      const pg_tp_exchange* ex, int dtype, void* stream);
int pg_tp_keys_push(const unsigned long long* keys, int B, long long v_local, const pg_tp_exchange* ex,
                    void* stream);

/* KVCache.key_cache[layer] / value_cache[layer] view: gather pages into a contiguous
 * [B, nkv, T, hd] tensor (modeling_gemma.py:12-36 attribute parity). */
int pg_kv_gather(void* out, const void* pool, const int32_t* page_table, int pt_stride,
                 int page_size, int B, int T, int nkv, int hd, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PG_B200_H */
