// C-ABI entry points of libfoley_b200.so (include/foley_b200.h): thin, exception-free wrappers over Engine.
#include <new>

#include "common.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "engine.cuh"
#include "gemm_host.cuh"
#include "preprocess.cuh"
#include "safetensors.cuh"

namespace foley {
std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
}  // namespace foley

using namespace foley;

extern "C" const char* foley_last_error(void) { return last_error_ref().c_str(); }
#ifndef FOLEY_SRC_HASH
#define FOLEY_SRC_HASH "unknown"
#endif
extern "C" const char* foley_version(void) { return "foley_b200 0.2 sm_100a (tcgen05+TMA) src:" FOLEY_SRC_HASH; }

extern "C" foley_status foley_gemm(const void* a, int32_t dtype, int64_t batch, int64_t rows, int64_t k,
                                   int64_t lda, int64_t a_batch_stride, const void* w, int64_t n,
                                   int32_t taps, int32_t tap_off0, int32_t tap_stride, int32_t splits,
                                   int32_t bn, int32_t mode, int32_t act, const void* bias, void* out,
                                   int64_t ldo, int64_t out_batch_stride, int64_t split_stride,
                                   void* stream) {
    if (!a || !w || !out) return fail(FOLEY_ERR_INVALID, "foley_gemm: null pointer");
    if (dtype != FOLEY_DT_BF16 && dtype != FOLEY_DT_F32 && dtype != FOLEY_DT_F16) return fail(FOLEY_ERR_INVALID, "foley_gemm: dtype");
    if (dtype == FOLEY_DT_F16 && mode != 0) return fail(FOLEY_ERR_UNSUPPORTED, "foley_gemm: fp16 operands exist for mode 0 (16-bit output) only");
    if (mode < 0 || mode > 2) return fail(FOLEY_ERR_INVALID, "foley_gemm: mode must be 0,1,2");
    GemmLaunch L;
    L.a.ptr = a;
    L.a.dtype = dtype == FOLEY_DT_F32 ? DT_F32 : DT_BF16;     // fp16 shares the 16-bit operand layout and tensor maps
    L.f16 = dtype == FOLEY_DT_F16 ? 1 : 0;
    L.a.k = k; L.a.rows = rows; L.a.batch = batch; L.a.ld = lda; L.a.batch_stride = a_batch_stride;
    L.w = w; L.n = n;
    L.taps = taps; L.tap_off0 = tap_off0; L.tap_stride = tap_stride;
    L.splits = splits; L.bn = bn & 0xFFFF;
    L.dbg_stop = (bn >> 16) & 0xF;  // bring-up aid: upper bits of bn select a partial pipeline
    L.pair = ((bn >> 20) & 3) - 1;  // 0: default policy, 1: single-CTA tiles, 2: CTA-pair (cta_group::2) tiles
    L.cluster_m = ((bn >> 22) & 7) ? ((bn >> 22) & 7) : -1;   // 0: default policy, else weight-tile multicast over 1 / 2 / 4 m-tiles
    L.epi.mode = mode; L.epi.act = act; L.epi.bias = bias; L.epi.out = out; L.epi.ldo = ldo;
    L.epi.out_batch_stride = out_batch_stride; L.epi.split_stride = split_stride;
    std::string err;
    if (!launch_gemm(L, static_cast<cudaStream_t>(stream), &err)) return fail(FOLEY_ERR_CUDA, err);
    return FOLEY_OK;
}

// ---- row-wise kernels, one by one (unit tests)
static const int* identity_map(int n) {   // device iota: the group maps of a stand-alone call (sample b = group b = row b)
    static int* dev[16] = {};
    static int cap[16] = {};
    int d = 0;
    cudaGetDevice(&d);
    d &= 15;
    if (cap[d] < n) {
        if (dev[d]) cudaFree(dev[d]);
        const int c = std::max(n, 256);
        std::vector<int> h(c);
        for (int i = 0; i < c; ++i) h[i] = i;
        if (cudaMalloc(&dev[d], c * sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemcpy(dev[d], h.data(), c * sizeof(int), cudaMemcpyHostToDevice);
        cap[d] = c;
    }
    return dev[d];
}

extern "C" foley_status foley_qk_norm_rope(const void* src, const float* partials, int32_t splits, const void* bias, int64_t src_ld,
                                           int32_t n_parts, int32_t batch, int32_t L, int32_t heads, int32_t norm_kind, float eps,
                                           const void* const* norm_w, const float* cos_t, const float* sin_t, void* const* dst,
                                           int32_t S_total, int32_t seq_offset, void* stream) {
    if ((!src && !partials) || !norm_w || !dst || n_parts < 1 || n_parts > 3 || batch < 1 || L < 1 || heads < 1)
        return fail(FOLEY_ERR_INVALID, "foley_qk_norm_rope: bad argument");
    QkvArgs q;
    q.src = static_cast<const __nv_bfloat16*>(src); q.src_ld = static_cast<int>(src_ld); q.n_parts = n_parts; q.H = heads; q.L = L;
    q.rows_total = batch * L; q.norm_kind = norm_kind; q.eps = eps; q.cos = cos_t; q.sin = sin_t;
    if (partials) {
        q.partials = partials; q.splits = splits; q.split_stride = static_cast<long long>(batch) * L * src_ld;
        q.bias = static_cast<const __nv_bfloat16*>(bias);
    }
    for (int p = 0; p < n_parts; ++p) {
        if (!dst[p]) return fail(FOLEY_ERR_INVALID, "foley_qk_norm_rope: null destination");
        if (norm_w[p] && (!cos_t || !sin_t)) return fail(FOLEY_ERR_INVALID, "foley_qk_norm_rope: norm needs cos / sin tables");
        q.part[p].dst = static_cast<__nv_bfloat16*>(dst[p]);
        q.part[p].dst_batch_stride = static_cast<long long>(heads) * S_total * 128;
        q.part[p].dst_head_stride = static_cast<long long>(S_total) * 128;
        q.part[p].seq_offset = seq_offset;
        q.part[p].norm_w = static_cast<const __nv_bfloat16*>(norm_w[p]);
        q.part[p].src_col = p * heads * 128;
    }
    const long long warps = static_cast<long long>(q.rows_total) * n_parts * heads;
    FOLEY_CUDA_OK(launch_k(qk_norm_rope_kernel, dim3(static_cast<unsigned>((warps + 3) / 4)), dim3(128), 0,
                           static_cast<cudaStream_t>(stream), q));
    return FOLEY_OK;
}

extern "C" foley_status foley_combine_ln_mod(const float* partials, int32_t splits, const void* bias, const void* mod,
                                             int64_t mod_sample_stride, int64_t mod_tok_stride, int32_t gate_chunk,
                                             int32_t shift_chunk, int32_t scale_chunk, float* x, const void* x_init, int32_t round_x,
                                             void* h, float eps, int32_t batch, int32_t L, int32_t C, void* stream) {
    if (!x || batch < 1 || L < 1 || C < 128 || C % 128 != 0 || C > 2048)
        return fail(FOLEY_ERR_INVALID, "foley_combine_ln_mod: bad argument (C must be a multiple of 128 <= 2048)");
    if ((gate_chunk >= 0 || shift_chunk >= 0) && !mod) return fail(FOLEY_ERR_INVALID, "foley_combine_ln_mod: modulation vectors missing");
    const int* iota = identity_map(batch);
    if (!iota) return fail(FOLEY_ERR_CUDA, "foley_combine_ln_mod: out of memory");
    CombineArgs a;
    a.partials = partials; a.splits = splits; a.split_stride = static_cast<long long>(batch) * L * C;
    a.bias = static_cast<const __nv_bfloat16*>(bias);
    ModRef m;
    m.base = static_cast<const __nv_bfloat16*>(mod); m.sample_stride = mod_sample_stride; m.tok_stride = mod_tok_stride; m.by_trow = 0;
    if (gate_chunk >= 0) { a.gate = m; a.gate_chunk = gate_chunk; }
    a.x = x; a.x_init = static_cast<const __nv_bfloat16*>(x_init); a.round_x = round_x;
    a.h = static_cast<__nv_bfloat16*>(h); a.eps = eps;
    if (shift_chunk >= 0) { a.mod = m; a.shift_chunk = shift_chunk; a.scale_chunk = scale_chunk; }
    a.C = C; a.rows_total = batch * L;
    a.rm = RowMap{iota, iota, iota, L};
    FOLEY_CUDA_OK(launch_k(combine_ln_mod_kernel, dim3(a.rows_total), dim3(C / 4), 0, static_cast<cudaStream_t>(stream), a));
    return FOLEY_OK;
}

extern "C" foley_status foley_cfg_euler(const void* y, float* lat, void* x_next, int32_t B, int32_t n_cond, int32_t ch, int32_t L,
                                        float guidance, const float* sigmas_dev, const int32_t* step_dev, void* stream) {
    if (!y || !lat || !x_next || !sigmas_dev || !step_dev || B < 1 || n_cond < 1 || n_cond > 2 || ch < 1 || L < 1)
        return fail(FOLEY_ERR_INVALID, "foley_cfg_euler: bad argument");
    dim3 blk(32, 8), grid((L + 31) / 32, (ch + 31) / 32, B);
    FOLEY_CUDA_OK(launch_k(cfg_euler_kernel, grid, blk, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(y),
                           lat, static_cast<__nv_bfloat16*>(x_next), B, n_cond, ch, L, guidance, sigmas_dev, step_dev));
    return FOLEY_OK;
}

static AttOperand to_att_operand(const foley_attn_src& s, int heads) {
    AttOperand o;
    o.ptr = static_cast<const __nv_bfloat16*>(s.ptr);
    o.batch_stride = s.batch_stride; o.head_stride = s.head_stride; o.row_stride = s.row_stride;
    o.rows = s.rows; o.heads = heads; o.batch = s.batch;
    return o;
}
static AttNorm to_att_norm(const foley_attn_src& s) {
    AttNorm n;
    n.rows0 = s.rows0;
    for (int i = 0; i < 2; ++i) {
        n.w[i] = static_cast<const __nv_bfloat16*>(s.norm_w[i]);
        n.rope[i] = reinterpret_cast<const float2*>(s.rope[i]);
    }
    return n;
}

extern "C" foley_status foley_attention(const foley_attn_args* g, void* stream) {
    if (!g || !g->out) return fail(FOLEY_ERR_INVALID, "foley_attention: null pointer");
    if (g->batch < 1 || g->heads < 1) return fail(FOLEY_ERR_INVALID, "foley_attention: bad shape");
    for (const foley_attn_src* s : {&g->q, &g->k, &g->v}) {
        if (!s->ptr || s->rows < 1 || s->batch < 1 || s->rows0 < 0 || s->rows0 > s->rows)
            return fail(FOLEY_ERR_INVALID, "foley_attention: bad operand");
        if ((reinterpret_cast<uintptr_t>(s->ptr) & 15) || (s->batch_stride | s->head_stride | s->row_stride) % 8)
            return fail(FOLEY_ERR_INVALID, "foley_attention: operands must be 16-byte aligned (pointer and strides)");
        for (int i = 0; i < 2; ++i)
            if (s->norm_w[i] && !s->rope[i]) return fail(FOLEY_ERR_INVALID, "foley_attention: norm_w needs the (cos, sin) table");
    }
    if (g->k.rows != g->v.rows) return fail(FOLEY_ERR_INVALID, "foley_attention: k and v must have the same length");
    if (g->v.norm_w[0] || g->v.norm_w[1]) return fail(FOLEY_ERR_INVALID, "foley_attention: v takes no norm");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g->impl == 1) {
        if (g->q.norm_w[0] || g->q.norm_w[1] || g->k.norm_w[0] || g->k.norm_w[1] || g->q.row_stride != 128 ||
            g->k.row_stride != 128 || g->v.row_stride != 128)
            return fail(FOLEY_ERR_UNSUPPORTED, "foley_attention: impl 1 takes prepared [B,H,S,128] inputs");
        AttnArgs a;
        a.q = static_cast<const __nv_bfloat16*>(g->q.ptr); a.k = static_cast<const __nv_bfloat16*>(g->k.ptr);
        a.v = static_cast<const __nv_bfloat16*>(g->v.ptr); a.o = static_cast<__nv_bfloat16*>(g->out);
        a.H = g->heads; a.Sq = g->q.rows; a.Sk = g->k.rows;
        a.q_batch_stride = g->q.batch_stride; a.q_head_stride = g->q.head_stride;
        a.kv_batch_stride = g->k.batch_stride; a.kv_head_stride = g->k.head_stride;
        a.o_batch_stride = g->out_batch_stride; a.kv_batch_map = g->kv_batch_map;
        a.scale_log2 = g->scale * 1.4426950408889634f;
        { static bool once = false; if (!once) { FOLEY_CUDA_OK(cudaFuncSetAttribute(attention_kernel<8, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<8, 3, 2>::SMEM)); once = true; } }
        dim3 grid((a.Sq + 63) / 64, g->heads, g->batch);
        FOLEY_CUDA_OK(launch_k(attention_kernel<8, 3, 2>, grid, dim3(256), AttCfg<8, 3, 2>::SMEM, st, a));
        return FOLEY_OK;
    }
    AttTcArgs a;
    a.qn = to_att_norm(g->q); a.kn = to_att_norm(g->k);
    a.o = static_cast<__nv_bfloat16*>(g->out); a.o_batch_stride = g->out_batch_stride;
    a.H = g->heads; a.Sq = g->q.rows; a.Sk = g->k.rows; a.kv_batch_map = g->kv_batch_map;
    a.scale_log2 = g->scale * 1.4426950408889634f; a.norm_kind = g->norm_kind; a.eps = g->eps;
    a.probe_chunk = g->dbg[0] - 1;
    { static bool once = false; if (!once) { FOLEY_CUDA_OK(attention_tc_init()); once = true; } }
    std::string err;
    if (!launch_attention_tc(to_att_operand(g->q, g->heads), to_att_operand(g->k, g->heads), to_att_operand(g->v, g->heads), a,
                             g->batch, st, &err))
        return fail(FOLEY_ERR_CUDA, err);
    return FOLEY_OK;
}

// Reads and clears the device debug words ([0] = first timed-out mbarrier wait code).
extern "C" foley_status foley_debug_flags(uint32_t* out4) {
    unsigned int h[4] = {0, 0, 0, 0};
    FOLEY_CUDA_OK(cudaMemcpyFromSymbol(h, g_foley_dbg, sizeof h));
    unsigned int z[4] = {0, 0, 0, 0};
    FOLEY_CUDA_OK(cudaMemcpyToSymbol(g_foley_dbg, z, sizeof z));
    for (int i = 0; i < 4; ++i) out4[i] = h[i];
    return FOLEY_OK;
}

// In-kernel timeline of the last GEMM launched with dbg_stop == 8 (tools/gemm_micro.py --dbg 8): clock64 stamps.
extern "C" foley_status foley_debug_times(uint64_t* out16) {
    unsigned long long h[16];
    FOLEY_CUDA_OK(cudaMemcpyFromSymbol(h, g_foley_times, sizeof h));
    for (int i = 0; i < 16; ++i) out16[i] = h[i];
    return FOLEY_OK;
}

// ------------------------------------------------------------------------------------------------ engine
struct foley_engine {
    Engine impl;
};

#define API_GUARD_BEGIN try {
#define API_GUARD_END                                                          \
    } catch (const std::bad_alloc&) {                                          \
        return fail(FOLEY_ERR_CUDA, "out of host memory");                     \
    } catch (const std::exception& ex) {                                       \
        return fail(FOLEY_ERR_CUDA, std::string("internal error: ") + ex.what()); \
    }

extern "C" foley_status foley_preprocess_frames(const float* image, int32_t n_frames, int32_t H, int32_t W,
                                                const int32_t* frame_idx, int32_t T, int32_t resize_h, int32_t resize_w,
                                                int32_t crop_top, int32_t crop_left, int32_t out_h, int32_t out_w,
                                                float* out, void* stream) {
    API_GUARD_BEGIN
    return preprocess_frames(image, n_frames, H, W, frame_idx, T, resize_h, resize_w, crop_top, crop_left, out_h, out_w, out,
                             static_cast<cudaStream_t>(stream));
    API_GUARD_END
}

// The int16 filter bank of one axis, host only (tests compare it with the oracle's / ATen's).
extern "C" foley_status foley_resize_weights(int32_t in_size, int32_t out_size, int32_t* xmin, int32_t* xsize, int16_t* w,
                                             int64_t w_cap, int32_t* max_interp, int32_t* precision) {
    if (in_size < 1 || out_size < 1) return fail(FOLEY_ERR_INVALID, "foley_resize_weights: empty axis");
    API_GUARD_BEGIN
    const ResizeBank b = make_resize_bank(in_size, out_size);
    if (max_interp) *max_interp = b.max_interp;
    if (precision) *precision = b.precision;
    if (w && static_cast<int64_t>(b.w.size()) > w_cap) return fail(FOLEY_ERR_INVALID, "foley_resize_weights: weight buffer too small");
    for (int i = 0; i < out_size; ++i) {
        if (xmin) xmin[i] = b.xmin[i];
        if (xsize) xsize[i] = b.xsize[i];
    }
    if (w) std::copy(b.w.begin(), b.w.end(), w);
    return FOLEY_OK;
    API_GUARD_END
}

extern "C" foley_status foley_engine_create(const foley_config* cfg, int device, foley_engine** out) {
    if (!cfg || !out) return fail(FOLEY_ERR_INVALID, "foley_engine_create: null argument");
    API_GUARD_BEGIN
    foley_engine* e = new foley_engine();
    foley_status s = e->impl.create(cfg, device);
    if (s != FOLEY_OK) { delete e; return s; }
    *out = e;
    return FOLEY_OK;
    API_GUARD_END
}

extern "C" void foley_engine_destroy(foley_engine* e) { delete e; }

extern "C" foley_status foley_engine_load_tensor(foley_engine* e, const char* name, const void* data,
                                                 const int64_t* shape, int32_t ndim, int32_t dtype) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null engine");
    API_GUARD_BEGIN
    return e->impl.load_tensor(name, data, shape, ndim, dtype);
    API_GUARD_END
}

extern "C" foley_status foley_engine_load_safetensors(foley_engine* e, const char* path, const char* prefix,
                                                      int64_t* n_loaded) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null engine");
    API_GUARD_BEGIN
    return e->impl.load_safetensors(path, prefix, n_loaded);
    API_GUARD_END
}

// Header-only inspection (no CUDA): tensor count, payload bytes, element counts per floating dtype.
extern "C" foley_status foley_safetensors_probe(const char* path, int64_t* n_tensors, int64_t* data_bytes,
                                                int64_t numel_by_dtype[5]) {
    if (!path) return fail(FOLEY_ERR_INVALID, "foley_safetensors_probe: null path");
    API_GUARD_BEGIN
    StFile f;
    std::string err;
    if (!f.open_file(path, &err)) return fail(FOLEY_ERR_INVALID, err);
    if (n_tensors) *n_tensors = static_cast<int64_t>(f.entries.size());
    if (data_bytes) *data_bytes = static_cast<int64_t>(f.data_bytes);
    if (numel_by_dtype) {
        for (int i = 0; i < 5; ++i) numel_by_dtype[i] = 0;
        for (const StEntry& en : f.entries) {
            const int dt = st_dtype_to_foley(en.dtype);
            if (dt < 0) continue;
            int64_t numel = 1;
            for (int64_t d : en.shape) numel *= d;
            numel_by_dtype[dt] += numel;
        }
    }
    return FOLEY_OK;
    API_GUARD_END
}

extern "C" int32_t foley_fp8_wraps(const char* tensor_name, int32_t ndim) {
    return tensor_name && fp8_wraps(tensor_name, ndim) ? 1 : 0;
}

extern "C" foley_status foley_engine_finalize(foley_engine* e) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null engine");
    API_GUARD_BEGIN
    return e->impl.finalize();
    API_GUARD_END
}

extern "C" foley_status foley_set_conditions(foley_engine* e, const void* clip, const void* sync, const void* text,
                                             int32_t dtype, int32_t n_cond, int32_t Lv, int32_t S, int32_t T,
                                             int32_t L, int32_t batch, void* stream) {
    if (!e || !clip || !sync || !text) return fail(FOLEY_ERR_INVALID, "foley_set_conditions: null argument");
    if (dtype != FOLEY_DT_BF16 && dtype != FOLEY_DT_F32 && dtype != FOLEY_DT_F16)
        return fail(FOLEY_ERR_INVALID, "foley_set_conditions: dtype");
    API_GUARD_BEGIN
    foley_status s_ = e->impl.set_conditions(clip, sync, text, dtype, n_cond, Lv, S, T, L, batch, e->impl.pick_stream(stream));
    return s_ != FOLEY_OK ? s_ : e->impl.order_after(stream);
    API_GUARD_END
}

extern "C" foley_status foley_dit_forward(foley_engine* e, const float* x, const float* t, int32_t n_t, float* out,
                                          void* stream) {
    if (!e || !x || !t || !out) return fail(FOLEY_ERR_INVALID, "foley_dit_forward: null argument");
    API_GUARD_BEGIN
    foley_status s_ = e->impl.forward(x, t, n_t, out, e->impl.pick_stream(stream));
    return s_ != FOLEY_OK ? s_ : e->impl.order_after(stream);
    API_GUARD_END
}

extern "C" foley_status foley_denoise(foley_engine* e, float* latents, const float* sigmas, int32_t n_steps,
                                      float guidance, foley_progress_fn progress, void* user, void* stream) {
    if (!e || !latents || !sigmas) return fail(FOLEY_ERR_INVALID, "foley_denoise: null argument");
    API_GUARD_BEGIN
    foley_status s_ = e->impl.denoise(latents, sigmas, n_steps, guidance, FOLEY_SOLVER_EULER, progress, user, e->impl.pick_stream(stream));
    return s_ != FOLEY_OK ? s_ : e->impl.order_after(stream);
    API_GUARD_END
}

extern "C" foley_status foley_denoise_solver(foley_engine* e, float* latents, const float* sigmas, int32_t n_calls,
                                             float guidance, int32_t solver, foley_progress_fn progress, void* user,
                                             void* stream) {
    if (!e || !latents || !sigmas) return fail(FOLEY_ERR_INVALID, "foley_denoise_solver: null argument");
    API_GUARD_BEGIN
    foley_status s_ = e->impl.denoise(latents, sigmas, n_calls, guidance, solver, progress, user, e->impl.pick_stream(stream));
    return s_ != FOLEY_OK ? s_ : e->impl.order_after(stream);
    API_GUARD_END
}

// Host only: the stage table foley_denoise_solver uploads (one row of 9 numbers per model call):
// dt, c0, c1, c2, cm, kind, store_slot, save_sample, base_saved.
extern "C" foley_status foley_solver_table(int32_t solver, const float* sigmas, int32_t n_calls, float* out9) {
    if (!sigmas || !out9 || n_calls < 1) return fail(FOLEY_ERR_INVALID, "foley_solver_table: bad argument");
    if (solver < FOLEY_SOLVER_HEUN2 || solver > FOLEY_SOLVER_KUTTA4) return fail(FOLEY_ERR_INVALID, "foley_solver_table: multi-stage solvers only");
    API_GUARD_BEGIN
    const std::vector<SolverCall> t = solver_table(solver, sigmas, n_calls);
    for (int i = 0; i < n_calls; ++i) {
        float* o = out9 + 9 * i;
        o[0] = t[i].dt; o[1] = t[i].c0; o[2] = t[i].c1; o[3] = t[i].c2; o[4] = t[i].cm;
        o[5] = static_cast<float>(t[i].kind); o[6] = static_cast<float>(t[i].store_slot);
        o[7] = static_cast<float>(t[i].save_sample); o[8] = static_cast<float>(t[i].base_saved);
    }
    return FOLEY_OK;
    API_GUARD_END
}

extern "C" foley_status foley_dac_decode(foley_engine* e, const float* z, int32_t batch, int32_t L, float* wav,
                                         void* stream) {
    if (!e || !z || !wav) return fail(FOLEY_ERR_INVALID, "foley_dac_decode: null argument");
    API_GUARD_BEGIN
    if (!e->impl.dac_ready) {
        foley_status s = e->impl.dac_finalize();
        if (s != FOLEY_OK) return s;
    }
    foley_status s_ = e->impl.dac_decode(z, batch, L, wav, e->impl.pick_stream(stream));
    return s_ != FOLEY_OK ? s_ : e->impl.order_after(stream);
    API_GUARD_END
}

extern "C" int64_t foley_launch_count(const foley_engine* e) { return e ? e->impl.launches : 0; }

extern "C" foley_status foley_debug_read(foley_engine* e, const char* what, float* dst, int64_t cap, int64_t* n_out) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null engine");
    API_GUARD_BEGIN
    return e->impl.debug_read(what, dst, cap, n_out);
    API_GUARD_END
}

// Runtime switches for tests / profiling: "cuda_graph" (0/1), "max_splits" (1..8).
extern "C" foley_status foley_engine_set_option(foley_engine* e, const char* key, int64_t value) {
    if (!e || !key) return fail(FOLEY_ERR_INVALID, "null argument");
    const std::string k(key);
    if (k == "cuda_graph") { e->impl.use_cuda_graph = value != 0; e->impl.graph_valid = false; }
    else if (k == "max_splits") {
        if (value < 1 || value > 8) return fail(FOLEY_ERR_INVALID, "max_splits must be in [1,8]");
        if (e->impl.plan.valid && value > e->impl.max_splits) return fail(FOLEY_ERR_STATE, "raise max_splits before set_conditions");
        e->impl.max_splits_used = static_cast<int>(value);
        e->impl.graph_valid = false;
    } else if (k == "debug_skip") {
        e->impl.debug_skip = static_cast<int>(value);
        e->impl.graph_valid = false;
    } else if (k == "fp8_weight_storage") {
        if (value < 0 || value > 2) return fail(FOLEY_ERR_INVALID, "fp8_weight_storage must be 0 (none), 1 (e4m3fn) or 2 (e5m2)");
        e->impl.fp8_storage = static_cast<int>(value);
    } else return fail(FOLEY_ERR_INVALID, "unknown option " + k);
    return FOLEY_OK;
}
