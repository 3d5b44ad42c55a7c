/*
 * sdeo.h — C ABI of the B200-native ControlNet-SD1.5 denoising kernels (libsdeo.so).
 *
 * Every entry point is `extern "C"`, takes plain device pointers / sizes / a cudaStream_t passed as void*,
 * launches asynchronously on that stream, never allocates, never throws, and returns 0 on success or a
 * negative SDEO_E* code. The reference has exactly one native interface on this path — the TensorRT plugin
 * `GroupNormPlugin::enqueue(inputDesc, outputDesc, inputs, outputs, workspace, stream) noexcept -> int32`
 * (plugin/groupNormPlugin/groupNormPlugin.cpp:179-228) — and every other op is an ATen call made from the
 * Python modules (cited per function below). These functions are what a cgo/ctypes/pybind binding for the
 * path binds instead.
 *
 * Activations are NHWC ("channels-last") bf16 unless noted; tokens are [rows, channels] bf16 row-major.
 */
#ifndef SDEO_H_
#define SDEO_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDEO_OK 0
#define SDEO_EINVAL (-22)   /* bad argument / unsupported shape */
#define SDEO_ENOSYS (-38)   /* driver entry point missing (no CUDA driver / no sm_100 device) */
#define SDEO_ECUDA (-5)     /* a CUDA runtime/driver call failed; see sdeo_last_error() */

/* Human-readable description of the last failure on this thread. */
const char* sdeo_last_error(void);
/* Library ABI version (bumped on any signature change). */
int sdeo_version(void);
/* Programmatic dependent launch (default on; also SDEO_NO_PDL=1): every kernel of the library is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and orders itself behind its stream predecessor with
 * griddepcontrol.wait, so launch latency, the kernel prologue and the weight prefetch of call N+1 overlap the tail of
 * call N (also inside captured CUDA graphs). Stream semantics seen by the caller are unchanged. */
int sdeo_set_pdl(int enable);
/* Optional kernel timeline (profiling aid, tools/step_timeline.py): `buf` = device array of uint64, buf[0] = 0,
 * buf[1] = capacity in records; block (0,0,0) of every library kernel appends [tag | grid << 8, t_start,
 * t_dependency_resolved, t_end] (globaltimer ns) from buf[4]. NULL turns it off (default). Synchronises the device. */
int sdeo_set_trace(void* buf);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * Replaces: nn.Conv2d 3x3 / 1x1 in ResBlock, Downsample, Upsample, zero convs, hint block
 * (ldm/modules/diffusionmodules/openaimodel.py:200-240,150-152,106; cldm/cldm.py:147-163,281-282),
 * nn.Linear in CrossAttention / GEGLU / FeedForward / time_embed (ldm/modules/attention.py:49-76,154-179;
 * openaimodel.py:528-533), VAE convs (ldm/modules/diffusionmodules/model.py:100-127,571-617).
 *
 * y[pix, n] = epilogue( sum_{tap, c} x[pix*stride + tap - pad, c] * w[n, tap, c] )
 * Linear layers are the 1x1 case with N=1, H=1, W=rows.
 * ---------------------------------------------------------------------------------------------- */
enum {
  SDEO_EPI_NORMAL = 0, /* y = act(acc + bias + emb[batch]) * scale + residual                     */
  SDEO_EPI_GEGLU = 1,  /* packed N tile = [x | gate]; y[:, j] = (x + bx) * gelu_erf(gate + bg)      */
  SDEO_EPI_QKV = 2     /* scatter to head-major q,k [B,heads,tok,d] and transposed v [B,heads,d,ldv] */
};
enum { SDEO_ACT_NONE = 0, SDEO_ACT_SILU = 1, SDEO_ACT_QUICK_GELU = 2 /* x * sigmoid(1.702 x): CLIP text MLP */ };

typedef struct sdeo_conv_args {
  /* input activation(s): x1 carries channels [0,c1), optional x2 channels [c1,c1+c2) (fused torch.cat, dim=1) */
  const void* x1;   /* bf16, [n, h, w, ld1] with ld1 >= c1 (elements per pixel)                        */
  const void* x2;   /* bf16 or NULL                                                                    */
  int32_t n, h, w;  /* input spatial geometry                                                          */
  int32_t c1, ld1;  /* channels used from x1 / its pixel stride in elements (multiple of 8)            */
  int32_t c2, ld2;
  /* filter */
  const void* w_packed; /* bf16 [n_rows_packed, k_packed] from sdeo_pack_conv_weight (K-major)         */
  int32_t cout;         /* logical output channels (rows of w before packing/padding)                  */
  int32_t ksize;        /* 1 or 3 (2 only with up2_phase)                                              */
  int32_t stride;       /* 1 or 2                                                                      */
  int32_t pad;          /* 0 or 1                                                                      */
  /* epilogue */
  int32_t epi_mode;     /* SDEO_EPI_*                                                                  */
  int32_t act;          /* SDEO_ACT_*                                                                  */
  const float* bias;    /* fp32 [cout] (packed order for GEGLU) or NULL                                */
  const float* emb;     /* fp32 [n, cout] per-sample additive term (ResBlock emb_layers) or NULL        */
  const int32_t* emb_step; /* optional device pointer: table mode, emb is [S, cout] and EVERY sample adds row
                              *emb_step (the per-image table of emb_layers outputs for all S DDIM timesteps,
                              indexed by the device-side step counter). NULL: row = sample index             */
  const void* residual; /* [n, ho, wo, ldr] added after scaling, or NULL; bf16, or fp32 if residual_f32 */
  int32_t ldr;
  int32_t residual_f32;
  float scale;          /* multiplies act(acc+bias+emb) (ControlNet control_scales); 1.0f otherwise    */
  void* y;              /* bf16 (or fp32 if y_fp32) [n, ho, wo, ldy]                                   */
  int32_t ldy;
  int32_t y_fp32;
  void* y2;             /* optional bf16 twin of an fp32 y (residual-stream tensors kept in fp32 are also   */
  int32_t ldy2;         /* written in bf16 for consumers that read them through TMA); NULL otherwise        */
  /* SDEO_EPI_QKV only: packed column n -> which = n / (heads*dhead) + qkv_first (0=q,1=k,2=v) */
  void* q; void* k; void* vt;
  int32_t heads, dhead, tokens, ldv, qkv_first;
  /* split-K scratch (unused since split-K partials travel through distributed shared memory; may be NULL) */
  void* workspace;
  size_t workspace_bytes;
  /* optional: per-channel partial statistics of the FINAL fp32 output for the GroupNorm that consumes it (replaces
   * that GroupNorm's own pass over the tensor, groupNormKernel.cu:49-120 / F.group_norm's statistics). fp32
   * [slots][cout][2] = (sum, sum of squares) over the rows of one (M tile, K slice) slot; a sample's slots are
   * contiguous. Size and layout come from sdeo_conv_gn_stats_slots(); consumed by sdeo_groupnorm_apply_stats(). */
  float* gn_stats;
  /* optional, producer of a LayerNorm input (nn.LayerNorm, attention.py:372-374): per output row the (sum, sum of
   * squares) over each N tile's columns, fp32 [parts][row_stats_ld][2] (row_stats_ld >= n*ho*wo rows). Geometry from
   * sdeo_conv_row_stats_parts(). Mutually exclusive with gn_stats. */
  float* row_stats;
  int32_t row_stats_ld;
  /* optional, LayerNorm FOLDED into this GEMM (any epilogue mode): x1 is the raw (un-normalised) bf16 input, the packed
   * weight carries gamma (W' = W * gamma along K), `bias` carries beta @ W^T. With mean / rstd of every input row taken
   * from the producer's row statistics (ln_parts partials of ln_ld rows; ln_c = normalised width; ln_eps), the epilogue
   * computes rstd[row] * (acc - mean[row] * ln_csum[n]) + bias[n]; ln_csum[n] = sum_k of the PACKED bf16 weight row n. */
  const float* ln_stats;
  int32_t ln_parts, ln_ld, ln_c;
  float ln_eps;
  const float* ln_csum;
  /* extra zero rows / columns AFTER the last input row / column (0, or 1 with ksize 3, stride 2, pad 0): the VAE
   * encoder's Downsample, F.pad(x, (0,1,0,1)) + conv(stride 2, padding 0) (ldm/modules/diffusionmodules/model.py:80-84) */
  int32_t pad_hi;
  /* optional, GroupNorm (+ SiLU) FOLDED into this convolution's operand path (GroupNorm32 -> SiLU -> conv of ResBlock
   * in_layers / out_layers, openaimodel.py:200-240; Normalize -> proj_in of SpatialTransformer, attention.py:406-418;
   * norm1/norm2 -> conv1/conv2 of the VAE ResnetBlock, model.py:125-139; the plugin's bSwish contract,
   * groupNormPlugin.cpp:291-304): x1 / x2 are the RAW bf16 tensors; each operand tile is normalised (and SiLU'd) in shared
   * memory before the MMA reads it, so the normalised tensor never exists in global memory and the GroupNorm costs no
   * launch. Statistics come from the producers' epilogues (gn_stats of the convolutions that wrote x1 / x2):
   * gnf_stats1 = fp32 [n][gnf_parts1][c1][2], gnf_stats2 likewise for x2 (NULL without x2). gnf_gamma / gnf_beta: fp32
   * [c1 + c2]; groups of (c1 + c2) / gnf_groups consecutive channels of the virtual concat (a group may straddle the
   * seam). Zero padding stays zero (it pads the NORMALISED tensor). Not with ln_stats. NULL gnf_stats1 = off. */
  const float* gnf_stats1;
  const float* gnf_stats2;
  int32_t gnf_parts1, gnf_parts2;
  const float* gnf_gamma;
  const float* gnf_beta;
  int32_t gnf_groups;
  float gnf_eps;
  int32_t gnf_silu;
  /* optional, one sub-pixel phase of nearest-x2 upsampling followed by a 3x3 "same" convolution (Upsample:
   * openaimodel.py:108-118, model.py:50-65): output pixel (2i + a, 2j + b) of that pair only sees the 2x2 input pixels
   * (i - 1 + a .. i + a, j - 1 + b .. j + b), with the 3x3 taps that coincide after upsampling summed -- four 2x2 filters
   * over the LOW resolution input (16 tap-pixels instead of 36: 2.25x fewer multiply-adds, and the 4x tensor is never
   * written). up2_phase = 1 + 2a + b (0 = off); ksize = 2, stride 1; w_packed holds that phase's [cout, cin, 2, 2] filter;
   * y is the FULL [n, 2h, 2w, ldy] output, of which this call writes one pixel in four. No gn_stats / row_stats. */
  int32_t up2_phase;
} sdeo_conv_args;

/* Bytes of workspace the planner may use for these args (fp32 partial tiles + tile counters).
 * The first sdeo_conv_counter_bytes() bytes hold the tile counters and must be zero before first use
 * (the kernel leaves them zero again). */
size_t sdeo_conv_workspace_bytes(const sdeo_conv_args* a);
size_t sdeo_conv_counter_bytes(void);
int sdeo_conv2d(const sdeo_conv_args* a, void* stream);
/* GroupNorm partial statistics geometry for these args: *max_slots_total = upper bound of slots (allocate
 * max_slots_total * cout * 2 floats; independent of tuning); *parts_per_sample = slots per sample under the plan
 * sdeo_conv2d currently uses for this shape (call it AFTER sdeo_conv2d; 0 = this call produces no statistics: bf16
 * output, unaligned pitches, or an M tile that spans two samples). */
int sdeo_conv_gn_stats_slots(const sdeo_conv_args* a, int32_t* max_slots_total, int32_t* parts_per_sample);
/* Row statistics geometry: *max_parts = upper bound of N tiles (allocate max_parts * row_stats_ld * 2 floats);
 * *parts = N tiles under the plan sdeo_conv2d currently uses (call it AFTER sdeo_conv2d; 0 = not produced). */
int sdeo_conv_row_stats_parts(const sdeo_conv_args* a, int32_t* max_parts, int32_t* parts);
/* The launch plan for these args (host logic only, no device work): halo < 0 = the plan sdeo_conv2d currently uses (tuned
 * if the shape has been tuned), 0 / 1 = the heuristic plan of the tap-by-tap / HALO tiling. out[16] = {N tile, K slices,
 * halo | pair << 1 | occ2 << 2 | producers << 4 (pair: two CTAs run one M=256 cta_group::2 MMA, each staging half of the
 * weight tile; occ2: the plan uses at most 112 KB of shared memory so that two CTAs share an SM; producers: TMA producer
 * threads, a divisor of the ring depth), tile samples, tile rows (bh), tile columns (bw), M tiles, N tiles, B (or A+B) pipeline stages, halo A stages, halo
 * A stage bytes, dynamic shared memory bytes, MMA rows in use, TMEM columns, halo pixel pitch, K steps per slice}.
 * HALO tiling (3x3, stride 1): one (bh+2) x (bw+2) input box per 64-channel chunk serves all nine filter taps. */
int sdeo_conv_plan_describe(const sdeo_conv_args* a, int32_t halo, int32_t* out, int32_t n_out);
/* Enable (1) / disable (0) per-shape autotuning of the N tile and the number of K slices: the first eager call of a
 * layer shape times the candidates on the caller's stream and caches the winner (calls made while the stream is
 * being captured into a CUDA graph only read the cache). */
int sdeo_conv_autotune(int enable);
/* CTA budget of subsequent sdeo_conv2d launches (0 = whole GPU, the default). A caller that runs two independent
 * branches on two streams sets ~half the SM count while it enqueues them, so that both branches' kernels fit on the
 * GPU side by side (one CTA per SM at ~200 KB of shared memory); tile / split-K choices are tuned per budget. */
int sdeo_conv_set_cta_budget(int max_ctas);

/* Repack an fp32 filter [cout, cin, k, k] (PyTorch layout, device memory) into the K-major bf16 layout the
 * kernel streams: [rows_packed, k*k*(chunks(c1)+chunks(c2))*64]. `geglu_bn` > 0 interleaves the two GEGLU
 * halves per N tile of that width. Query sizes with sdeo_packed_rows / sdeo_packed_k. */
int32_t sdeo_packed_rows(int32_t cout);
int32_t sdeo_packed_k(int32_t c1, int32_t c2, int32_t ksize);
int32_t sdeo_pick_bn(int32_t rows_packed, int32_t epi_mode, int32_t dhead);
int sdeo_pack_conv_weight(const float* w, int32_t cout, int32_t c1, int32_t c2, int32_t ksize, int32_t geglu_bn,
                          void* w_packed, void* stream);
/* Same interleave for a GEGLU bias: [2*inner] -> packed order. */
int sdeo_pack_geglu_bias(const float* b, int32_t n2, int32_t geglu_bn, float* b_packed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm(32)(+SiLU), NHWC. Replaces GroupNormPlugin::enqueue (groupNormPlugin.cpp:179-228,
 * groupNormKernel.cu:49-266) and torch GroupNorm32 / Normalize (+ nn.SiLU)
 * (ldm/modules/diffusionmodules/util.py:217-219; attention.py:88-89; model.py:46-47).
 * Follows PyTorch numerics (eps applied, fp32 statistics), not the plugin's eps-less variance.
 * x2 != NULL normalises torch.cat([x1, x2], dim=1) and writes the concatenated result.
 * ---------------------------------------------------------------------------------------------- */
/* One workspace per stream: calls that may run concurrently must not share it. */
size_t sdeo_groupnorm_workspace_bytes(int32_t n, int32_t hw, int32_t groups);
/* Host-side view of the two-launch grid (tests, tuning): plan[0] = CTAs (chunks) per sample, plan[1] = pixel rows per chunk
 * for a tensor of n samples x hw pixels x row_bytes bytes per pixel row. One CTA per 64 KB of the batch, at least ~1.5 CTAs
 * per SM, at most two waves of the three resident CTAs per SM over the batch and 384 chunks per sample. */
int sdeo_groupnorm_plan(int32_t n, int32_t hw, int64_t row_bytes, int32_t* plan);
/* x1/x2: bf16, or fp32 when x_f32 != 0 (fp32 residual-stream tensors); y is always bf16. */
int sdeo_groupnorm_nhwc(const void* x1, const void* x2, int32_t x_f32, const float* gamma, const float* beta, void* y,
                        int32_t n, int32_t hw, int32_t c1, int32_t c2, int32_t groups, float eps,
                        int32_t with_silu, void* workspace, size_t workspace_bytes, void* stream);
/* The plugin's exact I/O contract (groupNormPlugin.cpp:136-160, enqueue :179-228): x / y fp16 NHWC (kHWC8), gamma / beta
 * fp32, one tensor, optional Swish (bSwish). Three kernels behind one call (csrc/groupnorm_stream.cu; which one runs:
 * sdeo_groupnorm_f16_variant): RESIDENT when a sample fits the shared memory of one thread-block cluster (one read + one
 * write, workspace untouched); STREAMED otherwise (one persistent kernel: tiles staged in shared memory by bulk TMA copies,
 * a statistics visit and an apply visit per tile, the second read served from L2 while it lasts); two launches for
 * geometries neither takes (C > 4096 or more than 32 groups). The workspace (64-bit partial slots per CTA, sample and
 * group, plus one (mean, rstd) slot per sample and group; replaces GroupNormPlugin::getWorkspaceSize :173-177) is sized by
 * sdeo_groupnorm_f16_workspace_bytes; the call presets the slots itself. Requires C % 8 == 0, C % groups == 0,
 * groups <= 64. eps IS applied, unlike groupNormKernel.cu:190-194. Results are deterministic (fixed summation order). */
size_t sdeo_groupnorm_f16_workspace_bytes(int32_t n, int32_t hw, int32_t c, int32_t groups);
/* Host-side view of that kernel's schedule (tests, tuning): plan[0..6] = tiles per sample, pixels per tile, apply lag in
 * tiles, grid size, dynamic shared memory bytes, tile buffer stride, tile buffers for `sms` SMs (<= 0: 148); returns 1 when the shape
 * falls back to the two-launch variant. sdeo_groupnorm_f16_visits: the visits CTA `cta` of `grid` makes, in order, as
 * (kind, tile) pairs (kind 0 statistics, 1 apply); returns their number, writes at most `cap` pairs. Visits belong to
 * units: unit u = statistics of tile u, then apply of tile u - lag; CTA b takes units b, b + grid, ... */
int sdeo_groupnorm_f16_plan(int32_t n, int32_t hw, int32_t c, int32_t groups, int32_t sms, int32_t* plan);
int32_t sdeo_groupnorm_f16_visits(int32_t cta, int32_t grid, int32_t tiles, int32_t lag, int32_t* out, int32_t cap);
/* Which kernel sdeo_groupnorm_nhwc_f16 runs for a geometry on a device with `sms` SMs and clusters of up to `max_cluster`
 * CTAs (<= 0: 148 / 8): 3 = slab (one CTA per sample and slab of whole groups, the slab parked in shared memory: one read,
 * one write, nothing exchanged between CTAs, no workspace; info[0..2] = groups per slab, 16-byte vectors per slab row,
 * shared memory bytes; slabs up to SDEO_GN_F16_SLAB_KB, default 128), 2 = resident (the sample lives in the shared memory of one thread-block cluster: one read and one
 * write from anywhere, no workspace; info[0..2] = cluster size, pixel rows per CTA, shared memory bytes), 1 = two launches
 * (statistics + apply over the whole GPU: the default for samples beyond a cluster), 0 = streamed (one persistent kernel;
 * only when the environment asks for it with SDEO_GN_F16_VARIANT=stream). */
int sdeo_groupnorm_f16_variant(int32_t n, int32_t hw, int32_t c, int32_t groups, int32_t sms, int32_t max_cluster, int32_t* info);
/* Host-side view of the slab kernel's geometry (tests): plan[0..6] = groups per slab, 16-byte vectors per slab row, slabs per
 * sample, CTAs per slab, pixel rows each of them normalises, threads per CTA, dynamic shared memory bytes. Returns 1 when
 * the shape does not suit the slab kernel within max_kb (<= 0: the default 128) KB of shared memory per slab. */
int sdeo_groupnorm_f16_slab_plan(int32_t n, int32_t hw, int32_t c, int32_t groups, int32_t max_kb, int32_t* plan);
int sdeo_groupnorm_nhwc_f16(const void* x, const float* gamma, const float* beta, void* y, int32_t n, int32_t hw, int32_t c,
                            int32_t groups, float eps, int32_t with_silu, void* workspace, size_t workspace_bytes,
                            void* stream);
/* Same normalisation with the statistics taken from the partials the producing convolutions left
 * (sdeo_conv_args::gn_stats): stats1 = fp32 [n][parts1][c1][2] for x1, stats2 likewise for x2 (NULL without x2). The
 * tensor is read once. Results match sdeo_groupnorm_nhwc up to fp32 summation order. */
int sdeo_groupnorm_apply_stats(const void* x1, const void* x2, int32_t x_f32, const float* stats1, int32_t parts1,
                               const float* stats2, int32_t parts2, const float* gamma, const float* beta, void* y,
                               int32_t n, int32_t hw, int32_t c1, int32_t c2, int32_t groups, float eps, int32_t with_silu,
                               void* stream);

/* Folds partial statistics [n][parts][c][2] down to [n][*out_parts][c][2], *out_parts = ceil(parts / 256) (fixed order):
 * large feature maps leave one slot per 128-pixel tile, too many for every CTA of a consumer with a folded GroupNorm
 * (sdeo_conv_args::gnf_*) to add up on its own. stats == NULL or out == NULL: geometry query only. */
int sdeo_gn_stats_fold(const float* stats, float* out, int32_t n, int32_t parts, int32_t c, int32_t* out_parts, void* stream);

/* LayerNorm over the last dim of [rows, c] bf16 (nn.LayerNorm, attention.py:372-374), eps 1e-5. */
int sdeo_layernorm(const void* x, int32_t x_f32, const float* gamma, const float* beta, void* y, int32_t rows,
                   int32_t c, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Flash-style attention on tcgen05: o = softmax(q k^T * scale) v   (attention.py:227-249; model.py:186-199)
 * q [B*heads, nq, d], k [B*heads, nkv, d], vt [B*heads, d, ldv] (ldv >= nkv, multiple of 8), o [B, nq, heads*d].
 * d in {40, 80, 160, 512?}: multiples of 8, <= 256 on this build.
 * ---------------------------------------------------------------------------------------------- */
int sdeo_attention(const void* q, const void* k, const void* vt, void* o, int32_t batch, int32_t heads,
                   int32_t nq, int32_t nkv, int32_t d, int32_t ldv, float scale, void* stream);
/* Same with a causal mask (query i attends keys 0..i; nq == nkv == n): the CLIP text encoder's self-attention
 * (transformers CLIPTextModel, used by FrozenCLIPEmbedder, ldm/modules/encoders/modules.py:123-141). */
int sdeo_attention_causal(const void* q, const void* k, const void* vt, void* o, int32_t batch, int32_t heads, int32_t n,
                          int32_t d, int32_t ldv, float scale, void* stream);
/* y[r] = tok[ids[r]] + pos[r % t] (fp32 [rows, c]) + optional bf16 copy y2: CLIP token + position embeddings. */
int sdeo_embedding_add(const int64_t* ids, const float* tok, const float* pos, float* y, void* y2, int32_t rows, int32_t t,
                       int32_t c, int32_t vocab, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise / layout passes
 * ---------------------------------------------------------------------------------------------- */
/* CFG combine + DDIM x_{t-1} update (cldm/ddim_hacked.py:192,208-230) in one pass over the latent:
 *   e = eu + s (ec - eu);  x0 = (x - sqrt(1-a_t) e) * rsqrt(a_t);
 *   x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev - sigma^2) e + sigma * noise
 * coef_table is DEVICE memory, 8 floats per step row:
 *   {s, sqrt_1m_at, rsqrt_at, sqrt_aprev, dir_coef, sigma, 0, 0}; row = step_idx ? *step_idx : 0
 * (device-resident so that a captured CUDA graph replays for every step without host patching).
 * eps_c / eps_u: fp32, NHWC with ld_eps floats per pixel when eps_nhwc != 0 (the UNet out-conv layout), else NCHW.
 * eps_u may be NULL (no guidance: e = ec). x, noise, x_prev, pred_x0: fp32 NCHW [n, c, hw]; noise/pred_x0 may be NULL.
 * (a row whose sigma is 0 never reads noise).
 * x_next (optional): bf16 NHWC [dup*n, hw, ldn] copy of x_prev, channels >= c zero-filled, written dup times
 * (the next step's cond+uncond network input). */
int sdeo_cfg_ddim_step(const float* eps_c, const float* eps_u, int32_t eps_nhwc, int32_t ld_eps, const float* x,
                       const float* noise, float* x_prev, float* pred_x0, void* x_next, int32_t dup, int32_t ldn,
                       const float* coef_table, const int32_t* step_idx, int32_t n, int32_t c, int32_t hw,
                       void* stream);
/* The same with the noise as a device TABLE [steps][n, c, hw] of which row *step_idx is read: eta > 0 inside a captured step
 * graph (the sampler engine draws the per-step noise up front, ddim_hacked.py:227-230). step_idx must not be NULL. */
int sdeo_cfg_ddim_step_noise_table(const float* eps_c, const float* eps_u, int32_t eps_nhwc, int32_t ld_eps, const float* x,
                                   const float* noise_table, float* x_prev, float* pred_x0, void* x_next, int32_t dup,
                                   int32_t ldn, const float* coef_table, const int32_t* step_idx, int32_t n, int32_t c,
                                   int32_t hw, void* stream);
/* *ctr += delta (single thread) — advances the device-side step index between graph replays. */
int sdeo_counter_add(int32_t* ctr, int32_t delta, void* stream);
/* fp32 NCHW -> bf16 NHWC (y = scale * x) with channel padding to ldy (zeros), and back (first c channels).
 * `scale` folds decode_first_stage's 1/scale_factor (canny2image_torch.py:64-67) into the layout pass. */
int sdeo_nchw_to_nhwc_bf16(const float* x, void* y, int32_t n, int32_t c, int32_t hw, int32_t ldy, float scale,
                           void* stream);
int sdeo_nhwc_bf16_to_nchw(const void* x, float* y, int32_t n, int32_t c, int32_t hw, int32_t ldx, void* stream);
int sdeo_nhwc_f32_to_nchw(const float* x, float* y, int32_t n, int32_t c, int32_t hw, int32_t ldx, void* stream);
/* nearest x2 upsample, NHWC bf16 (F.interpolate(scale_factor=2, mode="nearest"), openaimodel.py:115) */
int sdeo_upsample_nearest2x(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, void* stream);
/* y = a + alpha*b (bf16, same shape) — ControlNet residual injection (cldm/cldm.py:35,41) */
int sdeo_add_scaled(const void* a, const void* b, float alpha, void* y, int64_t count, void* stream);
/* sinusoidal timestep embedding [cos | sin] (util.py:154-174) -> bf16 [n, ldy] (cols >= dim zero-filled).
 * t is int64 device memory: t[i] for sample i, or, when step_idx != NULL, t[*step_idx] for every sample. */
int sdeo_timestep_embedding(const int64_t* t, const int32_t* step_idx, void* y, int32_t n, int32_t dim, int32_t ldy,
                            float max_period, void* stream);
/* row-wise softmax(x * scale): fp32 scores [rows, cols] (ldx) -> bf16 probabilities (ldy)  (VAE AttnBlock, model.py:192-193) */
int sdeo_softmax_rows(const float* x, void* y, int32_t rows, int32_t cols, int32_t ldx, int32_t ldy, float scale,
                      void* stream);
/* y = silu(x) bf16 elementwise; y = bf16(x) from fp32; y = fp32(x) from bf16 */
int sdeo_silu(const void* x, void* y, int64_t count, void* stream);
int sdeo_f32_to_bf16(const float* x, void* y, int64_t count, void* stream);
int sdeo_bf16_to_f32(const void* x, float* y, int64_t count, void* stream);
/* VAE output: uint8 NHWC image = clip(x*127.5+127.5, 0, 255) from bf16 NHWC (canny2image_torch.py:68) */
int sdeo_image_to_u8(const void* x, uint8_t* y, int32_t npix, int32_t c, int32_t ldx, void* stream);
int sdeo_memset_async(void* p, int value, size_t bytes, void* stream);
/* y[i] = a[i / per_sample] * x[i] + b[i / per_sample] * z[i] (fp32; a, b device arrays, one entry per sample): the DDIM
 * encode step x_next = sqrt(a_next/a) x + sqrt(a_next) (sqrt(1/a_next - 1) - sqrt(1/a - 1)) eps (cldm/ddim_hacked.py:262-265),
 * stochastic_encode / q_sample sqrt(abar_t) x0 + sqrt(1 - abar_t) noise (:278-292), a CFG combine outside the fused step. */
int sdeo_axpby_f32(const float* x, const float* z, const float* a, const float* b, float* y, int64_t count,
                   int64_t per_sample, void* stream);
/* Inpainting blend of ddim_sampling (cldm/ddim_hacked.py:154-157): y = mask * (a x0 + b noise) + (1 - mask) * img with the
 * parenthesis = q_sample(x0, t). img / x0 / noise / y fp32 [n, c, hw]; mask fp32 [n, mask_c, hw], mask_c = 1 or c. */
/* The same blend as a node of the captured step graph: orig_table fp32 [S][n, c, hw] holds q_sample(x0, t_s) of every DDIM
 * step (drawn up front, in the step-by-step path's order); row *step_idx (device step counter) is blended into img.
 * y may alias img. */
int sdeo_mask_blend_table_f32(const float* orig_table, const float* img, const float* mask, float* y, const int32_t* step_idx,
                              int32_t n, int32_t c, int32_t mask_c, int64_t hw, void* stream);
int sdeo_mask_blend_f32(const float* x0, const float* noise, const float* img, const float* mask, const float* a,
                        const float* b, float* y, int32_t n, int32_t c, int32_t mask_c, int64_t hw, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 ("precise") mode: the 1e-4 parity configuration of the path (reference = PyTorch fp32 throughout,
 * ldm/modules/attention.py, openaimodel.py, cldm/cldm.py). Contractions still run on sdeo_conv2d: fp32 operands are
 * split into bf16 terms concatenated along K (x.w ~= hi.hi + mid.hi + hi.mid), everything else stays fp32.
 * ---------------------------------------------------------------------------------------------- */
/* x fp32 [rows, ldx] (or NCHW with nchw_hw = H*W > 0: rows = N*H*W) -> y bf16 [rows, terms*cp]; block t of a row holds
 * level ((pattern >> 2t) & 3) of every channel (0 = bf16(v), 1 = bf16 of what level 0 left, 2 = of what 0+1 left);
 * channels c..cp-1 (cp multiple of 8) are zero. */
int sdeo_split_terms(const float* x, void* y, int64_t rows, int32_t c, int32_t cp, int64_t ldx, int32_t nchw_hw,
                     int32_t terms, uint32_t pattern, void* stream);
/* Filter side: w fp32 [cout, src_cin, kk], channels cin0..cin0+cin-1 -> fp32 [cout, terms*cp, kk] of exactly
 * bf16-representable values (feed it to sdeo_pack_conv_weight with c = terms*cp). */
int sdeo_split_terms_weight(const float* w, float* y, int32_t cout, int32_t src_cin, int32_t cin0, int32_t cin, int32_t cp,
                            int32_t kk, int32_t terms, uint32_t pattern, void* stream);
/* GroupNorm(+SiLU) on fp32 NHWC (optionally the concat of x1, x2) -> fp32; statistics accumulated in double. */
int sdeo_groupnorm_f32(const float* x1, const float* x2, const float* gamma, const float* beta, float* y, int32_t n,
                       int32_t hw, int32_t c1, int32_t c2, int32_t groups, float eps, int32_t with_silu, void* stream);
int sdeo_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, int32_t rows, int32_t c, float eps,
                       void* stream);
/* o = softmax(q k^T * scale) v in fp32 on CUDA cores. q [B, nq, ldq], k [B, nkv, ldk], v [B, nkv, ldv], o [B, nq, ldo];
 * head h occupies columns h*d .. h*d+d-1 of each; d <= 160. */
int sdeo_attention_f32(const float* q, const float* k, const float* v, float* o, int32_t batch, int32_t heads, int32_t nq,
                       int32_t nkv, int32_t d, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo, float scale, void* stream);
/* y[r, j] = x[r, j] * gelu_erf(x[r, inner + j]) for x fp32 [rows, 2*inner] (attention.py:49-56). */
int sdeo_geglu_f32(const float* x, float* y, int64_t rows, int32_t inner, void* stream);
int sdeo_silu_f32(const float* x, float* y, int64_t count, void* stream);
/* fp32 [n, dim] sinusoidal embedding [cos | sin] of int64 timesteps (util.py:154-174). */
int sdeo_timestep_embedding_f32(const int64_t* t, float* y, int32_t n, int32_t dim, float max_period, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Hint preprocessing: cv2.Canny on the device, bit-exact with OpenCV for 8-bit 1- or 3-channel images,
 * aperture 3, L1 gradient (annotator/canny/__init__.py:4-6 -> cv2.Canny; canny2image_torch.py:30-38).
 * img uint8 HWC [h, w, c]; edges uint8 [h, w] (0 / 255). Workspace from sdeo_canny_workspace_bytes.
 * ---------------------------------------------------------------------------------------------- */
size_t sdeo_canny_workspace_bytes(int32_t h, int32_t w);
int sdeo_canny_u8(const uint8_t* img, int32_t h, int32_t w, int32_t c, double low_threshold, double high_threshold,
                  uint8_t* edges, void* workspace, size_t workspace_bytes, void* stream);
/* hint fp32 NCHW [n, 3, h*w] = HWC3(edges) / 255, repeated n times (canny2image_torch.py:34-38). */
int sdeo_edges_to_hint(const uint8_t* edges, float* hint, int32_t n, int32_t hw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDEO_H_ */
