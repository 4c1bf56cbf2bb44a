/*
 * foley_b200.h — C ABI of libfoley_b200.so, the B200 (sm_100a) denoising engine that replaces the
 * hot path of phazei/ComfyUI-HunyuanVideo-Foley behind its ComfyUI node surface.
 *
 * Every entry point returns a foley_status (0 = OK); no exceptions cross the boundary.  All tensor
 * arguments are plain pointers + sizes.  Unless stated otherwise pointers are DEVICE pointers owned
 * by the caller, and work is enqueued on the cudaStream_t passed as `stream` (void*; NULL = default
 * stream).  Reference citations are relative to the reference repository root.
 *
 * Path replaced:
 *   utils.py:125-258           denoise_process_with_generator      -> foley_denoise
 *   hifi_foley.py:707-924      HunyuanVideoFoley.forward           -> foley_dit_forward
 *   dac.py:280-303             DAC.decode                          -> foley_dac_decode
 *   nodes.py:85-104, utils.py:61-87  state-dict loading            -> foley_engine_load_tensor
 */
#ifndef FOLEY_B200_H_
#define FOLEY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int foley_status;
enum {
    FOLEY_OK = 0,
    FOLEY_ERR_INVALID = 1,   /* bad argument / shape (reference raises RuntimeError / AssertionError) */
    FOLEY_ERR_MISSING = 2,   /* weight missing at finalize (reference: load_state_dict(strict) error) */
    FOLEY_ERR_CUDA = 3,      /* CUDA runtime / driver error, see foley_last_error */
    FOLEY_ERR_STATE = 4,     /* call order violated (e.g. forward before conditions are set) */
    FOLEY_ERR_UNSUPPORTED = 5
};

enum { FOLEY_DT_BF16 = 0, FOLEY_DT_F32 = 1, FOLEY_DT_F16 = 2, FOLEY_DT_F8_E4M3FN = 3, FOLEY_DT_F8_E5M2 = 4 };

/* Hyper-parameters of configs/hunyuanvideo-foley-{xl,xxl}.yaml (model_config.model_kwargs) that the
 * hot path reads (hifi_foley.py:403-450), plus the torch semantics knobs the reference inherits. */
typedef struct foley_config {
    int32_t hidden_size;          /* 1536 (xxl) / 1408 (xl) */
    int32_t num_heads;            /* 12 / 11; head_dim must be 128 */
    int32_t depth_triple_blocks;  /* 18 / 12 */
    int32_t depth_single_blocks;  /* 36 / 24 */
    int32_t mlp_hidden_triple;    /* int(hidden*mlp_ratio) = 6144 / 5632 */
    int32_t mlp_hidden_single;    /* ConvMLP hidden, 4096 / 3840 (mlp_layers.py:133-134) */
    int32_t sync_hidden;          /* sync_in ConvMLP hidden (hifi_foley.py:482) */
    int32_t latent_dim;           /* audio_vae_latent_dim = 128 */
    int32_t clip_dim;             /* 768 */
    int32_t sync_dim;             /* 768 */
    int32_t text_dim;             /* condition_dim = 768 */
    int32_t freq_dim;             /* TimestepEmbedder frequency_embedding_size = 256 */
    float   rope_theta;           /* 10000 */
    float   single_rms_eps;       /* eps torch's nn.RMSNorm(eps=None) uses for the model dtype */
    int32_t max_batch;            /* largest number of variations per call (per GPU) */
    int32_t max_seconds;          /* largest clip duration the buffers are sized for (node cap: 60) */
    int32_t with_dac;             /* 1: allocate the DAC-VAE decoder too */
} foley_config;

typedef struct foley_engine foley_engine;

/* Thread-local description of the last error returned by any call on this thread. */
const char* foley_last_error(void);
/* Library / build identification, e.g. "foley_b200 0.1 sm_100a". */
const char* foley_version(void);

/* ---- lifetime ------------------------------------------------------------------------------ */
foley_status foley_engine_create(const foley_config* cfg, int device, foley_engine** out);
void         foley_engine_destroy(foley_engine* e);

/* ---- weights (nodes.py:85-104 load_state_dict; utils.py:61-87 load_dac_any) ------------------
 * `name` is the reference state-dict key, e.g. "triple_blocks.3.audio_self_attn_qkv.weight",
 * "single_blocks.7.linear2.w1.weight" ([4096,1536,3]) or, for the DAC-VAE,
 * "dac.decoder.model.1.block.1.parametrizations.weight.original0".  `data` is a HOST pointer to a
 * contiguous tensor of `dtype` with `ndim` dims `shape`.  The engine converts / repacks into its own
 * device layout (conv k=3 -> tap-major GEMM rows, single-block QKV (H D K)->(K H D), SwiGLU w1/w3
 * interleave, weight-norm folding).  Unknown names that belong to unused reference modules
 * (DAC encoder, final_layer.adaLN_modulation, ...) are accepted and ignored, like strict=False. */
foley_status foley_engine_load_tensor(foley_engine* e, const char* name, const void* data,
                                      const int64_t* shape, int32_t ndim, int32_t dtype);
/* The whole checkpoint in one call (nodes.py:85-104: load_torch_file -> init_empty_weights -> to_empty ->
 * load_state_dict -> .to(dtype), and utils.py:61-87 for the DAC with prefix "dac."): `path` is a .safetensors file;
 * it is mapped, its header parsed, and every tensor the hot path uses is copied from the mapping straight to the
 * device under the name `prefix` + key.  FP8 checkpoints (utils.py:492-503) are de-quantised at finalize.  n_loaded
 * (may be NULL) receives the number of tensors taken. */
foley_status foley_engine_load_safetensors(foley_engine* e, const char* path, const char* prefix,
                                           int64_t* n_loaded);
/* Header-only inspection of a .safetensors file, no GPU needed: tensor count, payload bytes and element counts per
 * FOLEY_DT_* (what _detect_ckpt_fp8 / _detect_ckpt_major_precision, utils.py:492-515, need). */
foley_status foley_safetensors_probe(const char* path, int64_t* n_tensors, int64_t* data_bytes,
                                     int64_t numel_by_dtype[5]);
/* 1 if the reference's FP8 weight-only storage (quantization != "none", utils.py:410-485) replaces the module that
 * owns this tensor — i.e. if option "fp8_weight_storage" rounds it through FP8 at load. */
int32_t      foley_fp8_wraps(const char* tensor_name, int32_t ndim);
/* Verifies that every tensor the hot path needs was provided; folds weight norm. */
foley_status foley_engine_finalize(foley_engine* e);

/* ---- conditions (utils.py:158-199; hifi_foley.py:744-770 step-invariant part) ----------------
 * clip [n_cond, Lv, clip_dim], sync [n_cond, S, sync_dim], text [n_cond, T, text_dim], bf16 or f32,
 * DEVICE pointers.  n_cond is 2 with CFG (row 0 = unconditional, row 1 = conditional, the order
 * of torch.cat([uncond, cond]) in utils.py:192-199) or 1 without.  All variations of a batch share
 * them (utils.py:159-162 `.repeat`).  `batch` variations x n_cond rows form the fused CFG batch.
 * L = number of audio latent frames (= int(duration * 50)). */
foley_status foley_set_conditions(foley_engine* e, const void* clip, const void* sync, const void* text,
                                  int32_t dtype, int32_t n_cond, int32_t Lv, int32_t S, int32_t T,
                                  int32_t L, int32_t batch, void* stream);

/* ---- one velocity prediction (hifi_foley.py:707-924) ------------------------------------------
 * x: [batch*n_cond, latent_dim, L] f32 (rows ordered like the reference's CFG batch: all uncond
 * samples first), t: n_t host floats (timesteps in [0,1000]); n_t is 1 (shared) or batch*n_cond.
 * out: [batch*n_cond, latent_dim, L] f32 holding the bf16-rounded model output. */
foley_status foley_dit_forward(foley_engine* e, const float* x, const float* t, int32_t n_t, float* out,
                               void* stream);

/* ---- the whole Euler loop (utils.py:203-247 + scheduling_flow_match_discrete.py:210-297) ------
 * latents: [batch, latent_dim, L] f32, in: initial noise, out: final latents.  sigmas: n_steps+1
 * host floats (torch.linspace(1,0,n_steps+1) for shift=1).  guidance: CFG scale (used when
 * n_cond==2).  progress(step, user) (ComfyUI ProgressBar.update, utils.py:247) is called once per finished step on
 * the CALLING thread when non-NULL: all steps are enqueued first, the device publishes the count of finished steps in
 * mapped host memory and the call polls it while the GPU keeps running (no per-step stream synchronize); it returns
 * after the last step was reported.  NULL keeps the call fully asynchronous. */
typedef void (*foley_progress_fn)(int32_t step, void* user);
foley_status foley_denoise(foley_engine* e, float* latents, const float* sigmas, int32_t n_steps,
                           float guidance, foley_progress_fn progress, void* user, void* stream);

/* ---- the same loop with the reference scheduler's other solvers (scheduling_flow_match_discrete.py:299-373) ---
 * solver: FOLEY_SOLVER_*; EULER is identical to foley_denoise.  n_calls model evaluations are made, call i with
 * timestep sigmas[i]*1000 exactly as the reference's loop does (utils.py:215-246): the scheduler treats consecutive
 * calls as inner stages of one step and only advances sigma -> sigma_next after a solver's last stage, so
 * heun-2 / midpoint-2 cover n_calls/2 sigma intervals and kutta-4 n_calls/4.  That behaviour is reproduced, not
 * corrected.  sigmas: n_calls+1 host floats. */
enum { FOLEY_SOLVER_EULER = 0, FOLEY_SOLVER_HEUN2 = 1, FOLEY_SOLVER_MIDPOINT2 = 2, FOLEY_SOLVER_KUTTA4 = 3 };
foley_status foley_denoise_solver(foley_engine* e, float* latents, const float* sigmas, int32_t n_calls,
                                  float guidance, int32_t solver, foley_progress_fn progress, void* user,
                                  void* stream);

/* Host only (tests): the per-call stage table foley_denoise_solver runs from — 9 floats per model call: dt, the four
 * derivative coefficients c0 c1 c2 cm, kind (0 model output / 1 Heun / 2 Kutta combination), derivative slot to keep
 * (-1 none), save-base-sample flag, update-from-saved-base flag. */
foley_status foley_solver_table(int32_t solver, const float* sigmas, int32_t n_calls, float* out9);

/* ---- DAC-VAE decode (dac.py:280-303) ----------------------------------------------------------
 * z: [batch, latent_dim, L] f32 -> wav: [batch, 1, L*hop] f32 (hop = 960). */
foley_status foley_dac_decode(foley_engine* e, const float* z, int32_t batch, int32_t L, float* wav,
                              void* stream);

/* ---- frame preprocessing for the condition encoders (nodes.py:293-317 + 184-196, utils.py:270-273) -------------
 * image: DEVICE fp32 [n_frames, H, W, 3] in [0,1] (ComfyUI IMAGE).  frame_idx: T HOST indices into it — the 8 fps /
 * 25 fps picks `torch.linspace(0, n-1, int(duration*fps)).long()` with the last frame held for short inputs.  Every
 * picked frame is quantised like `(image*255).byte()`, resized to (resize_h, resize_w) with torchvision's uint8
 * antialiased bicubic (ATen int16 fixed-point, horizontal pass first), cropped to the window [crop_top, +out_h) x
 * [crop_left, +out_w), scaled by 1/255 and normalised with mean = std = 0.5.  out: DEVICE fp32 [T, 3, out_h, out_w].
 * SigLIP2: resize 512x512, no crop.  Synchformer: short side to 224, centre crop 224.  Bit-exact with the reference's
 * CPU path. */
foley_status foley_preprocess_frames(const float* image, int32_t n_frames, int32_t H, int32_t W,
                                     const int32_t* frame_idx, int32_t T, int32_t resize_h, int32_t resize_w,
                                     int32_t crop_top, int32_t crop_left, int32_t out_h, int32_t out_w, float* out,
                                     void* stream);
/* Host-only: the fixed-point filter bank of one resize axis (xmin / xsize: out_size entries; w: out_size*max_interp
 * int16, pass NULL to query max_interp first). */
foley_status foley_resize_weights(int32_t in_size, int32_t out_size, int32_t* xmin, int32_t* xsize, int16_t* w,
                                  int64_t w_cap, int32_t* max_interp, int32_t* precision);

/* ---- introspection ---------------------------------------------------------------------------- */
/* Number of kernel launches the engine issued (graph-replayed launches included) since create. */
int64_t      foley_launch_count(const foley_engine* e);
/* Copies an internal activation buffer to `dst` (f32) for per-block parity tests; `what` names it
 * (e.g. "audio", "v_cond", "vec"); returns the element count via n_out. */
foley_status foley_debug_read(foley_engine* e, const char* what, float* dst, int64_t cap, int64_t* n_out);

/* Reads and clears device debug words: out4[0] = code of the first pipeline wait that timed out (0 = none). */
foley_status foley_debug_flags(uint32_t* out4);
/* Timeline stamps (SM clock cycles; [15]/[14] = %globaltimer ns at entry/exit) of CTA 0 of the last foley_gemm launched
 * with the bring-up selector 8 in the upper bits of `bn`: [0] entry, [1] prologue done, [2] activations released by
 * the predecessor, [3] first stage landed, [4] last MMA issued, [5] accumulator complete, [6] epilogue stores issued,
 * [7] CTA joined, [8] TMEM released. */
foley_status foley_debug_times(uint64_t* out16);
/* Runtime switches: "cuda_graph" (0/1, default 1), "max_splits" (1..8, default 8), "fp8_weight_storage" (0 none /
 * 1 e4m3fn / 2 e5m2; applies to tensors loaded AFTER the call: the Linear / Conv weights the reference would keep in
 * FP8 are rounded through that format, so results match the reference's quantization setting; compute stays bf16). */
foley_status foley_engine_set_option(foley_engine* e, const char* key, int64_t value);

/* ---- low-level kernels exported for unit tests and micro-benchmarks --------------------------- */
/* C[b,r,n] = sum_tap sum_k A[b, r+off0+tap*stride, k] * W[n, tap*K+k]; dtype bf16 -> tcgen05 kind::f16,
 * f32 -> kind::tf32.  mode: 0 bf16 out (+bias,+act), 1 SwiGLU pairs, 2 f32 partials (splits).  */
foley_status foley_gemm(const void* a, int32_t dtype, int64_t batch, int64_t rows, int64_t k,
                        int64_t lda, int64_t a_batch_stride, const void* w, int64_t n,
                        int32_t taps, int32_t tap_off0, int32_t tap_stride, int32_t splits, int32_t bn,
                        int32_t mode, int32_t act, const void* bias, void* out, int64_t ldo,
                        int64_t out_batch_stride, int64_t split_stride, void* stream);

/* ---- the row-wise kernels of the step, exported one by one for unit tests (tests/test_gpu_rowwise.py) ---- */
/* q/k RMSNorm + RoPE + scatter to the attention layout (csrc/rowwise.cuh qk_norm_rope_kernel; reference
 * attn_layers.py:112-148, norm_layers.py:49-51, hifi_foley.py:376-381).  Source rows: bf16 [batch*L, src_ld] with the
 * parts laid out (part, head, 128), or — partials != NULL — the fp32 K-split partials [splits, batch*L, src_ld] of the
 * projection GEMM plus its bf16 bias.  Part p of row (b, l) lands at dst[p][b, head, seq_offset + l, :] of a
 * [batch, heads, S_total, 128] bf16 tensor; norm_w[p] == NULL copies the part (v).  cos / sin: fp32 [L, 128]. */
foley_status foley_qk_norm_rope(const void* src, const float* partials, int32_t splits, const void* bias, int64_t src_ld,
                                int32_t n_parts, int32_t batch, int32_t L, int32_t heads, int32_t norm_kind, float eps,
                                const void* const* norm_w, const float* cos_t, const float* sin_t, void* const* dst,
                                int32_t S_total, int32_t seq_offset, void* stream);
/* K-split reduce + bias + gate + residual + LayerNorm + modulate (combine_ln_mod_kernel; reference hifi_foley.py:216-331,
 * 364-390 and modulate_layers.py): for every row (b, l) of [batch*L, C]
 *   y = bf16(sum_s partials[s] + bias); y = bf16(y * gate) if gate_chunk >= 0; x += y (x rounded to bf16 if round_x);
 *   h = bf16(LayerNorm(x) * bf16(1 + scale) + shift)      (plain LayerNorm when shift_chunk < 0)
 * Modulation vectors (bf16): chunk k of row (b, l) at mod + b*mod_sample_stride + l*mod_tok_stride + k*C.
 * partials == NULL skips the residual update; x_init (bf16 [batch*L, C]) replaces the incoming x; h == NULL skips the norm. */
foley_status foley_combine_ln_mod(const float* partials, int32_t splits, const void* bias, const void* mod,
                                  int64_t mod_sample_stride, int64_t mod_tok_stride, int32_t gate_chunk, int32_t shift_chunk,
                                  int32_t scale_chunk, float* x, const void* x_init, int32_t round_x, void* h, float eps,
                                  int32_t batch, int32_t L, int32_t C, void* stream);
/* bf16 CFG combine + fp32 Euler update + next bf16 model input (cfg_euler_kernel; reference utils.py:241-243,
 * scheduling_flow_match_discrete.py:262-297).  y: bf16 [n_cond*B, L, ch] (unconditional rows first); lat: fp32 [B, ch, L]
 * in/out; x_next: bf16 [n_cond*B, L, ch]; the step size is sigmas[step+1] - sigmas[step] (device arrays). */
foley_status foley_cfg_euler(const void* y, float* lat, void* x_next, int32_t B, int32_t n_cond, int32_t ch, int32_t L,
                             float guidance, const float* sigmas_dev, const int32_t* step_dev, void* stream);

/* softmax(Q K^T * scale) V for head_dim 128 (replaces F.scaled_dot_product_attention at attn_layers.py:422 and
 * hifi_foley.py:383), with the q/k RMSNorm + RoPE of the reference's attention modules (attn_layers.py:112-148,
 * norm_layers.py:49-51, hifi_foley.py:376-381) optionally folded into the operand load.  Element (b, h, r, d) of an
 * operand lives at ptr + b*batch_stride + h*head_stride + r*row_stride + d (bf16, device, 16-byte aligned).  For q and
 * k, rows [0, rows0) and [rows0, rows) form two groups (joint attention: visual tokens, then audio tokens): a group with
 * norm_w[i] != NULL is RMS-normalised with that [128] weight and rotated with the fp32 [group rows][64][2] (cos, sin)
 * table rope[i], indexed by the row inside the group.  out: bf16 [batch, Sq, heads*128].
 * impl 0: tcgen05 / TMEM kernel; impl 1: the round-1 mma.sync kernel (prepared [B,H,S,128] inputs without norm only). */
typedef struct foley_attn_src {
    const void* ptr;
    int64_t batch_stride, head_stride, row_stride;   /* elements */
    int32_t rows;                     /* rows per (sample, head) */
    int32_t rows0;                    /* rows of the first norm group (rows: one group) */
    int32_t batch;                    /* samples behind ptr (kv operands: the number of condition sets) */
    int32_t reserved;
    const void* norm_w[2];
    const float* rope[2];
} foley_attn_src;
typedef struct foley_attn_args {
    foley_attn_src q, k, v;
    void* out;
    int64_t out_batch_stride;
    int32_t batch, heads;
    const int32_t* kv_batch_map;      /* device, [batch]: kv sample of query sample b; NULL = b */
    float scale;                      /* softmax scale (1/sqrt(128)) */
    int32_t norm_kind;                /* 0: custom RMSNorm, two roundings (triple blocks); 1: nn.RMSNorm on bf16 (single blocks) */
    float eps;
    int32_t impl;
    int32_t dbg[4];                   /* [0] = 1 + chunk whose phases CTA 0 stamps with clock64 (foley_debug_times); 0 = off */
} foley_attn_args;
foley_status foley_attention(const foley_attn_args* args, void* stream);

/* ---- condition encoders (feature_utils.py:64-79, 132-138; nodes.py:283-284) ----------------
 * The reference moves its HF extractor modules to the device IN THE DiT's DTYPE before use, so these are the bf16
 * modules: bf16 weights, fp32 accumulation, one bf16 rounding per op.  Weights are taken under their HF state-dict names
 * ("vision_model.encoder.layers.3.self_attn.q_proj.weight", "text_model.encoder.layer.0.attention.self.query.weight",
 * ...) from host tensors or straight from the model.safetensors of the HF snapshot; tensors of modules the path does
 * not run (SigLIP text tower, CLAP pooler / projection) are accepted and ignored. */
enum { FOLEY_ENC_SIGLIP_VISION = 0, FOLEY_ENC_CLAP_TEXT = 1, FOLEY_ENC_SYNCHFORMER = 2 };
typedef struct foley_encoder_config {
    int32_t kind;                 /* FOLEY_ENC_* */
    int32_t hidden_size;          /* 768 */
    int32_t num_heads;            /* 12; head_dim must be 64 */
    int32_t num_layers;           /* 12 */
    int32_t intermediate_size;    /* 3072 */
    float   layer_norm_eps;       /* 1e-6 (google/siglip2-base-patch16-512) / 1e-12 (laion/larger_clap_general text) */
    int32_t image_size;           /* SigLIP: 512 */
    int32_t patch_size;           /* SigLIP: 16 */
    int32_t vocab_size;           /* CLAP: 50265 */
    int32_t max_positions;        /* CLAP: 514 */
    int32_t pad_token_id;         /* CLAP: 1 */
    int32_t max_frames_per_pass;  /* SigLIP: frames encoded together (activation memory ~17 MB per frame); 0 = 48 */
} foley_encoder_config;
typedef struct foley_encoder foley_encoder;
foley_status foley_encoder_create(const foley_encoder_config* cfg, int device, foley_encoder** out);
void         foley_encoder_destroy(foley_encoder* e);
foley_status foley_encoder_load_tensor(foley_encoder* e, const char* name, const void* data, const int64_t* shape,
                                       int32_t ndim, int32_t dtype);
foley_status foley_encoder_load_safetensors(foley_encoder* e, const char* path, const char* prefix, int64_t* n_loaded);
foley_status foley_encoder_finalize(foley_encoder* e);
/* encode_video_with_siglip2 (feature_utils.py:64-79): pixels = DEVICE fp32 [n_frames, 3, image, image] (the output of
 * foley_preprocess_frames) -> out = DEVICE bf16 [n_frames, hidden]: `get_image_features(...).pooler_output`. */
foley_status foley_siglip_encode(foley_encoder* e, const float* pixels, int32_t n_frames, void* out, void* stream);
/* encode_text_feat (feature_utils.py:132-138): ids / mask = HOST int32 [batch, T] (tokenizer input_ids and
 * attention_mask, padding=True; mask NULL = all ones) -> out = DEVICE bf16 [batch, T, hidden]: `last_hidden_state`,
 * padded positions included exactly as the reference passes them on. */
foley_status foley_clap_text_encode(foley_encoder* e, const int32_t* ids, const int32_t* mask, int32_t batch, int32_t T,
                                    void* out, void* stream);
/* encode_video_with_sync (feature_utils.py:81-106) -> Synchformer.forward -> MotionFormer (models/synchformer/motionformer.py,
 * video_model_builder.py, vit_helper.py; config divided_224_16x4): frames = DEVICE fp32 [n_frames, 3, 224, 224] (the 25 fps
 * output of foley_preprocess_frames); windows of 16 frames every 8 frames, segments = (n_frames - 16) / 8 + 1;
 * out = DEVICE fp32 [segments * 8, hidden] (the reference's `(b s) 1 t d -> b (s t) d`).  Arithmetic as the reference runs it:
 * fp16 autocast around a module whose parameters are in the DiT's dtype (fp16 GEMMs, fp32 LayerNorm / softmax / residual).
 * Weights under the Synchformer state-dict names ("vfeat_extractor.blocks.0.timeattn.qkv.weight", ...; the audio extractor and
 * the sync head are ignored).  kind FOLEY_ENC_SYNCHFORMER: hidden 768, 12 heads, 12 layers, intermediate 3072, eps 1e-6,
 * image_size 224, patch_size 16; max_frames_per_pass = SEGMENTS per pass here (0 = 16). */
foley_status foley_synchformer_encode(foley_encoder* e, const float* frames, int32_t n_frames, float* out, void* stream);
/* "layers_run": stop after this many transformer layers (per-layer parity taps; -1 = all); "att_tc": 1 = tcgen05 / TMEM
 * self-attention kernel in the vision tower (default), 0 = the mma.sync flash kernel. */
foley_status foley_encoder_set_option(foley_encoder* e, const char* key, int64_t value);
int64_t      foley_encoder_launch_count(const foley_encoder* e);
/* Activation buffers of the last call as HOST fp32: "x" residual stream, "h" last LayerNorm output, "qkv", "att", "y", "mlp". */
foley_status foley_encoder_debug_read(foley_encoder* e, const char* what, float* dst, int64_t cap, int64_t* n_out);
/* softmax(Q K^T * scale) V for head_dim 64 (HF SiglipAttention / ClapTextSelfAttention / nn.MultiheadAttention of the
 * pooling head).  Element (b, h, r, d) of an operand at ptr + b*batch_stride + r*row_stride + h*64 + d (bf16, device).
 * impl 0: flash-style mma.sync kernel, no mask; impl 2: tcgen05 / TMEM kernel, no mask (the vision tower's default); impl 1: one warp per query row, key_mask = DEVICE int32 [batch, Sk] (0 = padded
 * key) and optional bf16 rounding of scores and probabilities (the bmm + softmax path of nn.MultiheadAttention); impl 3: the
 * same arithmetic for ONE query per (sample, head) over many keys, one CTA per unit (pooling probe, class-token queries). */
foley_status foley_attention_d64(const void* q, const void* k, const void* v, void* out, int32_t batch, int32_t heads,
                                 int32_t Sq, int32_t Sk, int64_t q_batch_stride, int64_t q_row_stride,
                                 int64_t kv_batch_stride, int64_t kv_row_stride, int64_t o_batch_stride,
                                 int64_t o_row_stride, float scale, const int32_t* key_mask, int32_t round_scores,
                                 int32_t impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FOLEY_B200_H_ */
