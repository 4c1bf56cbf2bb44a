// Engine state: weights in GEMM-ready device layout, activation buffers sized per "plan"
// (batch, lengths), and the launch sequence of one DiT step.  See DESIGN.md for the data layout.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "gemm_host.cuh"
#include "rowwise.cuh"
#include "attention_tc.cuh"

namespace foley {

using bf16 = __nv_bfloat16;

struct RawTensor {
    void* dev = nullptr;          // device copy in source dtype
    std::vector<int64_t> shape;
    int dtype = FOLEY_DT_BF16;
    int64_t numel = 0;
    int fp8_mode = 0;             // round through this FP8 format at repack (0 none, 1 e4m3fn, 2 e5m2)
};

struct LinearW {                  // K-major weight [N, Ktot] + optional bias [N]
    bf16* w = nullptr;
    bf16* b = nullptr;
    int n = 0, k = 0, taps = 1;   // k = per-tap K
};

struct TripleW {
    LinearW qkv[2], self_proj[2], cross_q[2], cross_proj[2], fc1[2], fc2[2];   // [0]=audio, [1]=v_cond
    bf16* self_q_norm[2] = {nullptr, nullptr};
    bf16* self_k_norm[2] = {nullptr, nullptr};
    bf16* cross_q_norm[2] = {nullptr, nullptr};
    bf16* text_k_norm = nullptr;
};
struct SingleW {
    LinearW qkv, linear1, w13, w2;
    bf16* q_norm = nullptr;
    bf16* k_norm = nullptr;
};

struct DacLayer;  // dac.cu
void delete_dac_layer(DacLayer* l);

struct Plan {                     // shapes of the current generation
    int B = 0, U = 0, B2 = 0, G = 0, L = 0, Lv = 0, S = 0, T = 0, n_t = 0;
    bool valid = false;
};

// Per-call stage table of the multi-stage solvers (engine.cu); exported through foley_solver_table for host-side tests.
std::vector<SolverCall> solver_table(int solver, const float* sigmas, int n_calls);

class Engine {
  public:
    foley_config cfg{};
    int device = 0;
    int C = 0, H = 0, NT = 0, NS = 0, F = 0, Hs = 0, Hy = 0, LAT = 0;
    int64_t launches = 0;
    bool finalized = false;
    bool packed = false;          // packed DiT weights exist (possibly stale / partial when !finalized)
    int debug_skip = 0;           // option "debug_skip": ablation bitmask for tools/ablate_step.py (results are garbage)
    bool skip_gemm_once = false;
    int fp8_storage = 0;          // option "fp8_weight_storage": the reference's quantization != none for tensors loaded next
    std::unordered_map<std::string, RawTensor> raw;

    // ---- DiT weights
    LinearW audio_embed, vis_w13, vis_w2, cond1, cond2, time1, time2, sync0, sync_w13, sync_w2, final_lin;
    LinearW mod_triple_all;       // [NT*2*9C, C]: block i audio rows [i*18C, +9C), v_cond rows [i*18C+9C, +9C)
    LinearW mod_single_all;       // [NS*6C, C]
    LinearW text_kv_all;          // [NT*2C, C]
    bf16* sync_pos_emb = nullptr; // [8, sync_dim]
    bf16* empty_clip = nullptr;   // [clip_dim]
    bf16* empty_sync = nullptr;   // [sync_dim]
    std::vector<TripleW> triple;
    std::vector<SingleW> single;

    // ---- plan + buffers
    Plan plan;
    std::vector<void*> plan_allocs;
    // conditions / per-generation
    bf16 *a_sync = nullptr, *vcond0 = nullptr, *text_k = nullptr, *text_v = nullptr;
    float *rope_av_a_cos = nullptr, *rope_av_a_sin = nullptr, *rope_av_v_cos = nullptr, *rope_av_v_sin = nullptr;
    float *rope_plain_cos = nullptr, *rope_plain_sin = nullptr;
    float2 *rope2_av_a = nullptr, *rope2_av_v = nullptr, *rope2_plain = nullptr;   // the same tables as [rows][64] (cos, sin) pairs (attention_tc.cuh)
    bf16* qkv_j = nullptr;                // [B2][Lv + L][3C]: QKV (or cross-Q, [..][C]) of both streams, visual rows first, read by the fused attention
    int *grp_of_sample = nullptr, *trow_of_grp = nullptr, *cond_of_grp = nullptr, *step_dev = nullptr;
    float *sigmas_dev = nullptr, *t_dev = nullptr;
    bf16 *vec_all = nullptr, *mod_triple = nullptr;
    // per-step activations
    bf16 *x_in = nullptr, *h_a = nullptr, *h_v = nullptr, *qkv_a = nullptr, *qkv_v = nullptr;
    bf16 *Qj = nullptr, *Kj = nullptr, *Vj = nullptr, *attn_out = nullptr, *mlp_a = nullptr, *mlp_v = nullptr;
    bf16 *vectok_act = nullptr, *mod_single = nullptr, *y_out = nullptr;
    float *audio = nullptr, *vcond = nullptr, *part_a = nullptr, *part_v = nullptr, *lat_dev = nullptr;
    int max_splits = 8;          // workspace capacity
    int max_splits_used = 8;     // runtime cap (<= max_splits)
    cudaGraphExec_t step_graph = nullptr;
    bool graph_valid = false;
    bool tables_ready = false;   // RoPE / nearest-exact tables of the current plan uploaded
    bool use_cuda_graph = true;
    float graph_guidance = 0.f;
    bool graph_solver_kind = false;      // the captured step ends in cfg_solver_kernel instead of cfg_euler_kernel
    float *sol_d[3] = {nullptr, nullptr, nullptr}, *sol_samp = nullptr;   // multi-stage solver state [B, latent, L] fp32
    SolverCall* sol_table = nullptr;
    int sol_table_cap = 0;
    int64_t graph_launches_per_step = 0;
    int* host_step = nullptr;            // pinned + mapped: steps completed, written by advance_step_kernel, polled by denoise()
    int* host_step_dev = nullptr;        // its device alias
    int cur_G = 1;
    bool cur_per_sample = false;         // timesteps differ per sample (foley_dit_forward with n_t > 1)
    int *uq_first = nullptr, *uq_src = nullptr, *uq_tok_row = nullptr;   // distinct rows of the sync-token table (set_conditions)
    int n_uq = 0;
    size_t mod_rows_cap = 0;             // rows the single-block modulation buffers are allocated for
    int num_sms = 148;
    cudaStream_t side_stream = nullptr;  // visual-stream branch of the step
    cudaStream_t mod_stream = nullptr;   // single-block modulation GEMM branch
    cudaEvent_t ev_mod = nullptr;
    bool qkv_split = true;               // audio QKV / cross-Q GEMMs may split K (partials summed by the q/k-norm kernel)
    bool att_kv_split = true;            // small attention grids: in-CTA split-KV variant (FOLEY_ATT_KVSPLIT=0 disables)
    // Attention kernel policy.  att_tc: -1 (default) = tcgen05 / TMEM kernel (attention_tc.cuh) for sequences of more than
    // 320 keys (measured 1.6-1.9x the mma.sync kernel at 1500 / 1740 keys), the mma.sync kernel below that (a 5 s clip:
    // the tcgen05 kernel alone is as fast, 9.6 vs 9.8 us, but it holds a whole SM — 195 KB of shared memory, all of TMEM —
    // so its neighbours in the graph no longer overlap with it: step 4.24 vs 4.06 ms); FOLEY_ATT_TC=1 / 0 forces one kernel.
    // att_fused (FOLEY_ATT_FUSED=1, off by default): q/k-norm + RoPE inside the tcgen05 kernel's operand load for
    // sequences of <= 320 keys; correct (tests) but the norm work is then concentrated on the 44-66 SMs of the attention
    // grid and repeated per query tile: +7 us per call, step 4.51 ms.
    int att_tc = -1;
    bool att_fused = false;
    int side_split_cap = 0;              // > 0: cap on the K splits of the visual-branch GEMMs (FOLEY_SIDE_SPLITS): they run beside the
                                         // audio-stream GEMMs and every CTA they occupy is taken from those
    int mod_ctas = 24;                   // grid cap of the (persistent) modulation GEMM when it runs on its branch (FOLEY_MOD_CTAS; 0 = all SMs):
                                         // it has the whole triple-stream phase (1.4 ms) to finish, and every SM it holds is taken from the
                                         // latency-critical kernels beside it: 148 CTAs 3.971 ms, 24 CTAs 3.932 ms (profiles/r02_mod_ctas_sweep.log)
    bool mod_on_branch = true;           // single-block modulation GEMM on its own graph branch under the triple-stream phase: no gain with
                                         // the round-1 plans, 4.012 -> 3.991 ms with the plans below (profiles/r02_plan_sweep_*.log)
    // planner cost model (us per k-block of a 128- / 256-wide tile, fixed cost per CTA wave, cost per extra K split).  Swept
    // on the whole step at the end of round 2 (FOLEY_PLAN, profiles/r02_plan_sweep_*.log): pricing the 256-wide tile below the
    // 128-wide one and a split at 0.85 us moves the visual-branch GEMMs (80 rows) from 128-wide tiles x 4 splits (132 CTAs)
    // to 256-wide x 3 (51 CTAs) and the audio proj / cross-q GEMMs from 4 to 3 splits: the branch that runs BESIDE the audio
    // stream takes fewer SMs from it.  Step 4.124 -> 4.012 ms on the same box.
    double plan_tkb128 = 0.36, plan_tkb256 = 0.30, plan_tfix = 5.0, plan_tsplit = 0.85;
    // per-shape overrides of the planner: (rows, batch, n, k-blocks) -> (tile width, K splits).  The built-in table below holds
    // what tools/plan_search.py found on the whole step (coordinate descent, objective = measured ms per Euler step);
    // FOLEY_PLAN_OVERRIDE="rows:batch:n:kb=bn:s;..." adds / replaces entries (the search tool drives it).
    std::map<std::tuple<int, int, int, int>, std::pair<int, int>> plan_override;
    mutable std::map<std::tuple<int, int, int, int>, std::pair<int, int>> plan_seen;   // every (shape -> plan) the planner returned
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t own_stream = nullptr;   // blocking stream used when the caller passes NULL (legacy stream cannot be captured)
    // scratch of set_conditions / prepare_timesteps (plan-owned)
    bf16 *sc_clip = nullptr, *sc_sync = nullptr, *sc_text = nullptr, *sc_s0 = nullptr, *sc_s1 = nullptr, *sc_s2 = nullptr;
    bf16 *sc_s3 = nullptr, *sc_c1 = nullptr, *sc_c2 = nullptr, *sc_kv = nullptr, *sc_v1 = nullptr;
    bf16 *sc_e = nullptr, *sc_h1 = nullptr, *sc_vs = nullptr;
    int* sc_idx = nullptr;
    cudaStream_t pick_stream(void* st) { return st ? static_cast<cudaStream_t>(st) : own_stream; }
    // A caller on the legacy default stream (torch's default) gets legacy-stream ordering: the work went to this engine's own
    // blocking stream (the legacy stream cannot be captured), so the legacy stream is made to wait for it — otherwise the
    // next consumer on ANOTHER blocking stream (e.g. the DAC engine's own stream decoding the latents this engine has just
    // enqueued) would not be ordered behind it.
    cudaEvent_t ev_null = nullptr;
    foley_status order_after(void* caller_stream) {
        if (caller_stream) return FOLEY_OK;
        FOLEY_CUDA_OK(cudaSetDevice(device));
        FOLEY_CUDA_OK(cudaEventRecord(ev_null, own_stream));
        FOLEY_CUDA_OK(cudaStreamWaitEvent(cudaStreamLegacy, ev_null, 0));
        return FOLEY_OK;
    }

    // ---- DAC
    std::vector<DacLayer*> dac_layers;
    bool dac_ready = false;
    std::vector<void*> dac_allocs;
    int64_t dac_cap_T = 0;
    float *dac_buf[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t dac_buf_elems = 0;

    ~Engine();
    foley_status create(const foley_config* c, int dev);
    foley_status load_tensor(const char* name, const void* data, const int64_t* shape, int ndim, int dtype);
    foley_status load_safetensors(const char* path, const char* prefix, int64_t* n_loaded);
    foley_status finalize();
    foley_status set_conditions(const void* clip, const void* sync, const void* text, int dtype, int U, int Lv,
                                int S, int T, int L, int B, cudaStream_t st);
    foley_status prepare_timesteps(const float* t_host, int n_t, bool per_sample, cudaStream_t st);
    foley_status step(cudaStream_t st);                 // one forward: x_in -> y_out
    foley_status forward(const float* x, const float* t, int n_t, float* out, cudaStream_t st);
    foley_status denoise(float* latents, const float* sigmas, int n_steps, float guidance, int solver,
                         foley_progress_fn progress, void* user, cudaStream_t st);
    foley_status dac_finalize();
    foley_status dac_decode(const float* z, int batch, int L, float* wav, cudaStream_t st);
    foley_status debug_read(const char* what, float* dst, int64_t cap, int64_t* n_out);

  private:
    foley_status take_linear(const std::string& name, LinearW* out, bool bias, int taps_expected);
    foley_status take_vec(const std::string& name, bf16** out, int64_t n_expected);
    foley_status raw_as_bf16(const std::string& name, bf16** out, RawTensor** rt);
    foley_status gemm(cudaStream_t st, const bf16* A, int rows, int batch, long long lda, long long a_bs,
                      const LinearW& W, int n_off, int n_cnt, GemmEpi epi, int splits, int bn, int max_ctas = 0);
    foley_status alloc_plan(int B, int U, int L, int Lv, int S, int T);
    void free_plan();
    void free_packed();
    void free_dac();
    template <typename T> foley_status palloc(T** p, size_t count);
    template <typename T> void pfree(T*& p);
    int pick_splits(int rows, int batch, int n, int kblocks, int bn) const;
    int pick_bn(int rows, int batch, int n, int kblocks) const;
    void plan_gemm(int rows, int batch, int n, int kblocks, bool can_split, int* bn_out, int* splits_out, int split_cap = 0) const;
    foley_status proj_combine(cudaStream_t st, const bf16* A, int rows, int batch, long long lda, long long a_bs,
                              const LinearW& W, float* partials, CombineArgs ca);
};

}  // namespace foley
