// Engine implementation: weight intake/repack, per-generation precompute, the DiT step and the Euler loop.
#include "engine.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>

#include "attention.cuh"
#include "safetensors.cuh"

namespace foley {

#ifndef ST_OK
#define ST_OK(expr)                         \
    do {                                    \
        foley_status _s = (expr);           \
        if (_s != FOLEY_OK) return _s;      \
    } while (0)
#endif

static inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ lifetime
// Releases the packed (GEMM-layout) DiT weights: before a re-finalize and at destruction.
void Engine::free_packed() {
    auto fl = [](LinearW& w) { if (w.w) cudaFree(w.w); if (w.b) cudaFree(w.b); w = LinearW(); };
    for (LinearW* w : {&audio_embed, &vis_w13, &vis_w2, &cond1, &cond2, &time1, &time2, &sync0, &sync_w13, &sync_w2,
                       &final_lin, &mod_triple_all, &mod_single_all, &text_kv_all})
        fl(*w);
    auto fv = [](bf16*& p) { if (p) cudaFree(p); p = nullptr; };
    for (auto& t : triple) {
        for (int s = 0; s < 2; ++s) {
            fl(t.qkv[s]); fl(t.self_proj[s]); fl(t.cross_q[s]); fl(t.cross_proj[s]); fl(t.fc1[s]); fl(t.fc2[s]);
            fv(t.self_q_norm[s]); fv(t.self_k_norm[s]); fv(t.cross_q_norm[s]);
        }
        fv(t.text_k_norm);
    }
    for (auto& s : single) {
        fl(s.qkv); fl(s.linear1); fl(s.w13); fl(s.w2);
        fv(s.q_norm); fv(s.k_norm);
    }
    fv(sync_pos_emb); fv(empty_clip); fv(empty_sync);
    packed = false;
}

void Engine::free_dac() {
    for (DacLayer* l : dac_layers) delete_dac_layer(l);
    dac_layers.clear();
    for (void* p : dac_allocs) cudaFree(p);
    dac_allocs.clear();
    for (int i = 0; i < 4; ++i) dac_buf[i] = nullptr;
    dac_buf_elems = 0;
    dac_ready = false;
}

Engine::~Engine() {
    cudaSetDevice(device);
    if (step_graph) cudaGraphExecDestroy(step_graph);
    if (own_stream) cudaStreamDestroy(own_stream);
    if (side_stream) cudaStreamDestroy(side_stream);
    if (mod_stream) cudaStreamDestroy(mod_stream);
    if (ev_mod) cudaEventDestroy(ev_mod);
    if (ev_null) cudaEventDestroy(ev_null);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    free_plan();
    if (host_step) cudaFreeHost(host_step);
    for (auto& kv : raw)
        if (kv.second.dev) cudaFree(kv.second.dev);
    free_packed();
    free_dac();
}

foley_status Engine::create(const foley_config* c, int dev) {
    cfg = *c;
    device = dev;
    C = cfg.hidden_size; H = cfg.num_heads; NT = cfg.depth_triple_blocks; NS = cfg.depth_single_blocks;
    F = cfg.mlp_hidden_triple; Hs = cfg.mlp_hidden_single; Hy = cfg.sync_hidden; LAT = cfg.latent_dim;
    if (H <= 0 || C % H != 0 || C / H != 128) return fail(FOLEY_ERR_UNSUPPORTED, "head_dim must be 128");
    if (C % 128 != 0 || C > 2048) return fail(FOLEY_ERR_UNSUPPORTED, "hidden_size must be a multiple of 128 and <= 2048");
    if (F % 64 || Hs % 64 || Hy % 64 || LAT % 64 || cfg.clip_dim % 64 || cfg.sync_dim % 64 || cfg.text_dim % 64 ||
        cfg.freq_dim % 64)
        return fail(FOLEY_ERR_UNSUPPORTED, "all feature widths must be multiples of 64");
    FOLEY_CUDA_OK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    FOLEY_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(FOLEY_ERR_UNSUPPORTED, "foley_b200 requires an sm_100 (Blackwell B200) device");
    num_sms = prop.multiProcessorCount;
    FOLEY_CUDA_OK(cudaStreamCreate(&own_stream));
    FOLEY_CUDA_OK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    FOLEY_CUDA_OK(cudaStreamCreateWithFlags(&mod_stream, cudaStreamNonBlocking));
    FOLEY_CUDA_OK(cudaEventCreateWithFlags(&ev_mod, cudaEventDisableTiming));
    FOLEY_CUDA_OK(cudaEventCreateWithFlags(&ev_null, cudaEventDisableTiming));
    if (const char* e = getenv("FOLEY_MOD_BRANCH")) mod_on_branch = atoi(e) != 0;
    if (const char* e = getenv("FOLEY_MOD_CTAS")) mod_ctas = atoi(e);
    if (const char* e = getenv("FOLEY_QKV_SPLIT")) qkv_split = atoi(e) != 0;
    if (const char* e = getenv("FOLEY_PLAN")) sscanf(e, "%lf,%lf,%lf,%lf", &plan_tkb128, &plan_tkb256, &plan_tfix, &plan_tsplit);
    // plans found by tools/plan_search.py on the named configuration (xl, 5 s, batch 1: profiles/r02_plan_search_a.log): the
    // triple-block fc2 GEMM (K = 5632) takes 4 K-splits instead of the model's 6 (4.007 -> 3.980 ms per step: a third less
    // fp32 partial traffic for the combine kernel behind it)
    plan_override[std::make_tuple(250, 2, 1408, 88)] = std::make_pair(256, 4);
    if (const char* e = getenv("FOLEY_PLAN_OVERRIDE")) {
        const char* p = e;
        while (*p) {
            int r, b, n, kb, bn, sp, used = 0;
            if (sscanf(p, "%d:%d:%d:%d=%d:%d%n", &r, &b, &n, &kb, &bn, &sp, &used) == 6 && (bn == 128 || bn == 256) && sp >= 1 && sp <= max_splits)
                plan_override[std::make_tuple(r, b, n, kb)] = std::make_pair(bn, sp);
            else if (used == 0) break;
            p += used;
            while (*p == ';' || *p == ',' || *p == ' ') ++p;
        }
    }
    FOLEY_CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    FOLEY_CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    FOLEY_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&host_step), sizeof(int), cudaHostAllocMapped));
    *host_step = 0;
    FOLEY_CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&host_step_dev), host_step, 0));
    {
        std::string err;
        if (!gemm_init_attributes(&err)) return fail(FOLEY_ERR_CUDA, err);
        FOLEY_CUDA_OK(cudaFuncSetAttribute(attention_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<4, 4>::SMEM));
        FOLEY_CUDA_OK(cudaFuncSetAttribute(attention_kernel<8, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<8, 3, 2>::SMEM));
        if (const char* e = getenv("FOLEY_ATT_KVSPLIT")) att_kv_split = atoi(e) != 0;
        if (const char* e = getenv("FOLEY_SIDE_SPLITS")) side_split_cap = atoi(e);
        if (const char* e = getenv("FOLEY_ATT_TC")) att_tc = atoi(e);
        if (const char* e = getenv("FOLEY_ATT_FUSED")) att_fused = atoi(e) != 0;
        FOLEY_CUDA_OK(attention_tc_init());
        FOLEY_CUDA_OK(cudaFuncSetAttribute(attention_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<8, 3>::SMEM));
    }
    triple.resize(NT);
    single.resize(NS);
    return FOLEY_OK;
}

// ------------------------------------------------------------------------------------------------ weights
static bool name_is_used(const std::string& n) {
    if (n.rfind("dac.", 0) == 0) {
        return n.rfind("dac.decoder.", 0) == 0 || n.rfind("dac.post_quant_conv.", 0) == 0;
    }
    if (n.rfind("final_layer.adaLN_modulation", 0) == 0) return false;  // dead on this path (modulate_layers.py:20-22)
    return true;
}

foley_status Engine::load_tensor(const char* name, const void* data, const int64_t* shape, int ndim, int dtype) {
    if (!name || !data || ndim < 0 || ndim > 4) return fail(FOLEY_ERR_INVALID, "load_tensor: bad arguments");
    if (dtype < FOLEY_DT_BF16 || dtype > FOLEY_DT_F8_E5M2)
        return fail(FOLEY_ERR_INVALID, "load_tensor: dtype must be bf16, f32, f16, f8_e4m3fn or f8_e5m2");
    const std::string n(name);
    if (!name_is_used(n)) return FOLEY_OK;
    FOLEY_CUDA_OK(cudaSetDevice(device));
    RawTensor rt;
    rt.dtype = dtype;
    rt.numel = 1;
    for (int i = 0; i < ndim; ++i) { rt.shape.push_back(shape[i]); rt.numel *= shape[i]; }
    rt.fp8_mode = fp8_wraps(n, ndim) ? fp8_storage : 0;   // reference quantization != none (utils.py:410-485)
    const size_t bytes = static_cast<size_t>(rt.numel) * (dtype == FOLEY_DT_F32 ? 4 : dtype >= FOLEY_DT_F8_E4M3FN ? 1 : 2);
    FOLEY_CUDA_OK(cudaMalloc(&rt.dev, std::max<size_t>(bytes, 16)));
    FOLEY_CUDA_OK(cudaMemcpy(rt.dev, data, bytes, cudaMemcpyDefault));  // host or device source
    auto it = raw.find(n);
    if (it != raw.end() && it->second.dev) cudaFree(it->second.dev);
    raw[n] = rt;
    // A (re)loaded tensor invalidates only the half of the engine it belongs to: DAC tensors are packed by
    // dac_finalize, everything else by finalize (which frees the staging copies, so a reload must bring the whole
    // state dict again; finalize says which tensor is missing otherwise).
    if (n.rfind("dac.", 0) == 0) dac_ready = false;
    else finalized = false;
    return FOLEY_OK;
}

// The whole checkpoint in one call: map the file, parse the header, copy every tensor the path uses from the mapping
// to the device (nodes.py:85-104 without the meta-init / to_empty / load_state_dict / .to(dtype) passes over host RAM).
foley_status Engine::load_safetensors(const char* path, const char* prefix, int64_t* n_loaded) {
    if (!path) return fail(FOLEY_ERR_INVALID, "load_safetensors: null path");
    StFile f;
    std::string err;
    if (!f.open_file(path, &err)) return fail(FOLEY_ERR_INVALID, err);
    const std::string pre(prefix ? prefix : "");
    int64_t count = 0;
    for (const StEntry& e : f.entries) {
        const std::string name = pre + e.name;
        if (!name_is_used(name)) continue;
        const int dt = st_dtype_to_foley(e.dtype);
        if (dt < 0) {
            // integer / f64 tensors: only acceptable for entries the path does not read (e.g. step counters)
            if (name.find("num_batches_tracked") != std::string::npos) continue;
            return fail(FOLEY_ERR_UNSUPPORTED, "load_safetensors: dtype " + e.dtype + " of " + e.name + " is not supported");
        }
        if (e.shape.size() > 4) return fail(FOLEY_ERR_INVALID, "load_safetensors: rank > 4: " + e.name);
        ST_OK(load_tensor(name.c_str(), f.data + e.begin, e.shape.data(), static_cast<int>(e.shape.size()), dt));
        ++count;
    }
    if (n_loaded) *n_loaded = count;
    return FOLEY_OK;
}

foley_status Engine::raw_as_bf16(const std::string& name, bf16** out, RawTensor** rtp) {
    auto it = raw.find(name);
    if (it == raw.end()) return fail(FOLEY_ERR_MISSING, "missing tensor: " + name);
    RawTensor& rt = it->second;
    bf16* d = nullptr;
    FOLEY_CUDA_OK(cudaMalloc(&d, std::max<size_t>(rt.numel * 2, 16)));
    convert_to_bf16_kernel<<<blocks_for(rt.numel, 256), 256>>>(rt.dev, rt.dtype, rt.numel, d, rt.fp8_mode);
    FOLEY_CUDA_OK(cudaGetLastError());
    *out = d;
    if (rtp) *rtp = &rt;
    return FOLEY_OK;
}

foley_status Engine::take_vec(const std::string& name, bf16** out, int64_t n_expected) {
    RawTensor* rt = nullptr;
    ST_OK(raw_as_bf16(name, out, &rt));
    if (rt->numel != n_expected) return fail(FOLEY_ERR_INVALID, "unexpected size for " + name);
    return FOLEY_OK;
}

// Linear [N,K] / conv [N,K,taps] -> K-major GEMM weight [N, taps*K] (tap-major), bias [N].
foley_status Engine::take_linear(const std::string& name, LinearW* out, bool bias, int taps_expected) {
    RawTensor* rt = nullptr;
    bf16* w = nullptr;
    ST_OK(raw_as_bf16(name + ".weight", &w, &rt));
    const int nd = static_cast<int>(rt->shape.size());
    if (nd < 2) return fail(FOLEY_ERR_INVALID, "weight must be >= 2-D: " + name);
    out->n = static_cast<int>(rt->shape[0]);
    out->k = static_cast<int>(rt->shape[1]);
    out->taps = nd == 3 ? static_cast<int>(rt->shape[2]) : 1;
    if (out->taps != taps_expected) return fail(FOLEY_ERR_INVALID, "unexpected kernel size for " + name);
    if (out->taps > 1) {
        bf16* r = nullptr;
        FOLEY_CUDA_OK(cudaMalloc(&r, rt->numel * 2));
        repack_conv_weight_kernel<<<blocks_for(rt->numel, 256), 256>>>(w, out->n, out->k, out->taps, 1, 0, r);
        FOLEY_CUDA_OK(cudaGetLastError());
        FOLEY_CUDA_OK(cudaDeviceSynchronize());
        cudaFree(w);
        w = r;
    }
    out->w = w;
    out->b = nullptr;
    if (bias) ST_OK(take_vec(name + ".bias", &out->b, out->n));
    return FOLEY_OK;
}

// Two [N,K(,taps)] matrices interleaved row-wise (w1 -> even rows, w3 -> odd rows) for the SwiGLU epilogue.
static foley_status take_pair(Engine* e, const std::string& n1, const std::string& n3, LinearW* out, int taps_expected) {
    auto get = [&](const std::string& nm, bf16** w, int* N, int* K, int* taps) -> foley_status {
        auto it = e->raw.find(nm + ".weight");
        if (it == e->raw.end()) return fail(FOLEY_ERR_MISSING, "missing tensor: " + nm + ".weight");
        RawTensor& rt = it->second;
        *N = static_cast<int>(rt.shape[0]);
        *K = static_cast<int>(rt.shape[1]);
        *taps = rt.shape.size() == 3 ? static_cast<int>(rt.shape[2]) : 1;
        FOLEY_CUDA_OK(cudaMalloc(w, rt.numel * 2));
        convert_to_bf16_kernel<<<blocks_for(rt.numel, 256), 256>>>(rt.dev, rt.dtype, rt.numel, *w, rt.fp8_mode);
        FOLEY_CUDA_OK(cudaGetLastError());
        return FOLEY_OK;
    };
    bf16 *w1 = nullptr, *w3 = nullptr;
    int N1, K1, T1, N3, K3, T3;
    ST_OK(get(n1, &w1, &N1, &K1, &T1));
    ST_OK(get(n3, &w3, &N3, &K3, &T3));
    if (N1 != N3 || K1 != K3 || T1 != T3 || T1 != taps_expected) return fail(FOLEY_ERR_INVALID, "SwiGLU pair mismatch: " + n1);
    const long long numel = static_cast<long long>(N1) * K1 * T1;
    bf16* dst = nullptr;
    FOLEY_CUDA_OK(cudaMalloc(&dst, numel * 2 * 2));
    repack_conv_weight_kernel<<<blocks_for(numel, 256), 256>>>(w1, N1, K1, T1, 2, 0, dst);
    repack_conv_weight_kernel<<<blocks_for(numel, 256), 256>>>(w3, N1, K1, T1, 2, 1, dst);
    FOLEY_CUDA_OK(cudaGetLastError());
    FOLEY_CUDA_OK(cudaDeviceSynchronize());
    cudaFree(w1);
    cudaFree(w3);
    out->w = dst; out->b = nullptr; out->n = 2 * N1; out->k = K1; out->taps = T1;
    return FOLEY_OK;
}

foley_status Engine::finalize() {
    if (finalized) {
        if (cfg.with_dac && !dac_ready) ST_OK(dac_finalize());
        return FOLEY_OK;
    }
    FOLEY_CUDA_OK(cudaSetDevice(device));
    if (packed) {   // weights were reloaded after an earlier finalize: drop the old packed copies (and what was built on them)
        FOLEY_CUDA_OK(cudaDeviceSynchronize());
        free_plan();
        if (step_graph) { cudaGraphExecDestroy(step_graph); step_graph = nullptr; }
        free_packed();
    }
    packed = true;
    ST_OK(take_linear("audio_embedder.proj", &audio_embed, true, 1));
    ST_OK(take_pair(this, "visual_proj.w1", "visual_proj.w3", &vis_w13, 1));
    ST_OK(take_linear("visual_proj.w2", &vis_w2, false, 1));
    ST_OK(take_linear("cond_in.linear_1", &cond1, true, 1));
    ST_OK(take_linear("cond_in.linear_2", &cond2, true, 1));
    ST_OK(take_linear("time_in.mlp.0", &time1, true, 1));
    ST_OK(take_linear("time_in.mlp.2", &time2, true, 1));
    ST_OK(take_linear("sync_in.0", &sync0, true, 1));
    ST_OK(take_pair(this, "sync_in.2.w1", "sync_in.2.w3", &sync_w13, 1));
    ST_OK(take_linear("sync_in.2.w2", &sync_w2, false, 1));
    ST_OK(take_linear("final_layer.linear", &final_lin, true, 1));
    ST_OK(take_vec("sync_pos_emb", &sync_pos_emb, 8LL * cfg.sync_dim));
    ST_OK(take_vec("empty_clip_feat", &empty_clip, cfg.clip_dim));
    ST_OK(take_vec("empty_sync_feat", &empty_sync, cfg.sync_dim));
    if (audio_embed.n != C || audio_embed.k != LAT) return fail(FOLEY_ERR_INVALID, "audio_embedder shape");

    // stacked matrices: all triple-block modulations, all single-block modulations, all text K/V projections
    auto alloc_stack = [&](LinearW* w, long long n, int k) -> foley_status {
        w->n = static_cast<int>(n); w->k = k; w->taps = 1;
        FOLEY_CUDA_OK(cudaMalloc(&w->w, n * k * 2));
        FOLEY_CUDA_OK(cudaMalloc(&w->b, n * 2));
        return FOLEY_OK;
    };
    ST_OK(alloc_stack(&mod_triple_all, static_cast<long long>(NT) * 18 * C, C));
    ST_OK(alloc_stack(&mod_single_all, static_cast<long long>(NS) * 6 * C, C));
    ST_OK(alloc_stack(&text_kv_all, static_cast<long long>(NT) * 2 * C, C));
    auto place = [&](const std::string& name, LinearW& dst, long long row_off, int n_rows) -> foley_status {
        LinearW tmp;
        ST_OK(take_linear(name, &tmp, true, 1));
        if (tmp.n != n_rows || tmp.k != dst.k) return fail(FOLEY_ERR_INVALID, "unexpected shape for " + name);
        FOLEY_CUDA_OK(cudaMemcpy(dst.w + row_off * dst.k, tmp.w, static_cast<size_t>(n_rows) * dst.k * 2, cudaMemcpyDeviceToDevice));
        FOLEY_CUDA_OK(cudaMemcpy(dst.b + row_off, tmp.b, static_cast<size_t>(n_rows) * 2, cudaMemcpyDeviceToDevice));
        cudaFree(tmp.w);
        cudaFree(tmp.b);
        return FOLEY_OK;
    };
    const char* st_name[2] = {"audio", "v_cond"};
    for (int i = 0; i < NT; ++i) {
        const std::string p = "triple_blocks." + std::to_string(i) + ".";
        TripleW& t = triple[i];
        for (int s = 0; s < 2; ++s) {
            const std::string sn(st_name[s]);
            ST_OK(place(p + sn + "_mod.linear", mod_triple_all, (static_cast<long long>(i) * 2 + s) * 9 * C, 9 * C));
            ST_OK(take_linear(p + (s == 0 ? "audio_self_attn_qkv" : "v_cond_attn_qkv"), &t.qkv[s], true, 1));
            ST_OK(take_vec(p + (s == 0 ? "audio_self_q_norm.weight" : "v_cond_attn_q_norm.weight"), &t.self_q_norm[s], 128));
            ST_OK(take_vec(p + (s == 0 ? "audio_self_k_norm.weight" : "v_cond_attn_k_norm.weight"), &t.self_k_norm[s], 128));
            ST_OK(take_linear(p + sn + "_self_proj", &t.self_proj[s], true, 1));
            ST_OK(take_linear(p + sn + "_cross_q", &t.cross_q[s], true, 1));
            ST_OK(take_vec(p + sn + "_cross_q_norm.weight", &t.cross_q_norm[s], 128));
            ST_OK(take_linear(p + sn + "_cross_proj", &t.cross_proj[s], true, 1));
            ST_OK(take_linear(p + sn + "_mlp.fc1", &t.fc1[s], true, 1));
            ST_OK(take_linear(p + sn + "_mlp.fc2", &t.fc2[s], true, 1));
            if (t.qkv[s].n != 3 * C || t.fc1[s].n != F || t.fc2[s].k != F) return fail(FOLEY_ERR_INVALID, "triple block shape: " + p);
        }
        ST_OK(place(p + "text_cross_kv", text_kv_all, static_cast<long long>(i) * 2 * C, 2 * C));
        ST_OK(take_vec(p + "text_cross_k_norm.weight", &t.text_k_norm, 128));
    }
    for (int i = 0; i < NS; ++i) {
        const std::string p = "single_blocks." + std::to_string(i) + ".";
        SingleW& s = single[i];
        ST_OK(place(p + "modulation.linear", mod_single_all, static_cast<long long>(i) * 6 * C, 6 * C));
        // QKV rows (H D K) -> (K H D)
        LinearW q;
        ST_OK(take_linear(p + "linear_qkv", &q, true, 1));
        if (q.n != 3 * C || q.k != C) return fail(FOLEY_ERR_INVALID, "single block qkv shape: " + p);
        s.qkv.n = q.n; s.qkv.k = q.k; s.qkv.taps = 1;
        FOLEY_CUDA_OK(cudaMalloc(&s.qkv.w, static_cast<size_t>(q.n) * q.k * 2));
        FOLEY_CUDA_OK(cudaMalloc(&s.qkv.b, static_cast<size_t>(q.n) * 2));
        permute_qkv_rows_kernel<<<blocks_for(static_cast<long long>(q.n) * q.k, 256), 256>>>(q.w, H, 128, q.k, s.qkv.w);
        permute_qkv_rows_kernel<<<blocks_for(q.n, 256), 256>>>(q.b, H, 128, 1, s.qkv.b);
        FOLEY_CUDA_OK(cudaGetLastError());
        FOLEY_CUDA_OK(cudaDeviceSynchronize());
        cudaFree(q.w);
        cudaFree(q.b);
        ST_OK(take_linear(p + "linear1", &s.linear1, true, 3));
        ST_OK(take_pair(this, p + "linear2.w1", p + "linear2.w3", &s.w13, 3));
        ST_OK(take_linear(p + "linear2.w2", &s.w2, false, 3));
        ST_OK(take_vec(p + "q_norm.weight", &s.q_norm, 128));
        ST_OK(take_vec(p + "k_norm.weight", &s.k_norm, 128));
        if (s.w13.n != 2 * Hs || s.w2.k != Hs) return fail(FOLEY_ERR_INVALID, "single block ConvMLP shape: " + p);
    }
    FOLEY_CUDA_OK(cudaDeviceSynchronize());
    // release the DiT staging copies; DAC tensors stay until dac_finalize
    for (auto it = raw.begin(); it != raw.end();) {
        if (it->first.rfind("dac.", 0) != 0) {
            cudaFree(it->second.dev);
            it = raw.erase(it);
        } else {
            ++it;
        }
    }
    if (cfg.with_dac) ST_OK(dac_finalize());
    finalized = true;
    return FOLEY_OK;
}

// ------------------------------------------------------------------------------------------------ plan
template <typename T>
foley_status Engine::palloc(T** p, size_t count) {
    void* d = nullptr;
    FOLEY_CUDA_OK(cudaMalloc(&d, std::max<size_t>(count * sizeof(T), 256)));
    FOLEY_CUDA_OK(cudaMemset(d, 0, std::max<size_t>(count * sizeof(T), 256)));
    plan_allocs.push_back(d);
    *p = static_cast<T*>(d);
    return FOLEY_OK;
}

// Frees one plan-owned buffer (before it is re-allocated larger).
template <typename T>
void Engine::pfree(T*& p) {
    if (!p) return;
    auto it = std::find(plan_allocs.begin(), plan_allocs.end(), static_cast<void*>(p));
    if (it != plan_allocs.end()) {   // (a pointer that outlived its plan is stale: free_plan released it already)
        plan_allocs.erase(it);
        cudaFree(p);
    }
    p = nullptr;
}

void Engine::free_plan() {
    for (void* p : plan_allocs) cudaFree(p);
    plan_allocs.clear();
    plan = Plan();
    graph_valid = false;
    tables_ready = false;
}

foley_status Engine::alloc_plan(int B, int U, int L, int Lv, int S, int T) {
    free_plan();
    Plan p;
    p.B = B; p.U = U; p.B2 = B * U; p.L = L; p.Lv = Lv; p.S = S; p.T = T; p.G = U;
    const size_t B2 = p.B2, Sj = static_cast<size_t>(L) + Lv;
    const int Pmax = std::max(L, std::max(Lv, T));
    ST_OK(palloc(&a_sync, static_cast<size_t>(U) * L * C));
    ST_OK(palloc(&vcond0, static_cast<size_t>(U) * Lv * C));
    ST_OK(palloc(&text_k, static_cast<size_t>(NT) * U * T * C));
    ST_OK(palloc(&text_v, static_cast<size_t>(NT) * U * T * C));
    ST_OK(palloc(&rope_av_a_cos, static_cast<size_t>(L) * 128));
    ST_OK(palloc(&rope_av_a_sin, static_cast<size_t>(L) * 128));
    ST_OK(palloc(&rope_av_v_cos, static_cast<size_t>(Lv) * 128));
    ST_OK(palloc(&rope_av_v_sin, static_cast<size_t>(Lv) * 128));
    ST_OK(palloc(&rope_plain_cos, static_cast<size_t>(Pmax) * 128));
    ST_OK(palloc(&rope_plain_sin, static_cast<size_t>(Pmax) * 128));
    ST_OK(palloc(&rope2_av_a, static_cast<size_t>(L) * 64));
    ST_OK(palloc(&rope2_av_v, static_cast<size_t>(Lv) * 64));
    ST_OK(palloc(&rope2_plain, static_cast<size_t>(Pmax) * 64));
    ST_OK(palloc(&qkv_j, B2 * Sj * 3 * C));
    ST_OK(palloc(&grp_of_sample, B2));
    ST_OK(palloc(&trow_of_grp, B2));
    ST_OK(palloc(&cond_of_grp, B2));
    ST_OK(palloc(&step_dev, 4));
    ST_OK(palloc(&x_in, B2 * L * LAT));
    ST_OK(palloc(&h_a, B2 * L * C));
    ST_OK(palloc(&h_v, B2 * Lv * C));
    ST_OK(palloc(&qkv_a, B2 * L * 3 * C));
    ST_OK(palloc(&qkv_v, B2 * Lv * 3 * C));
    ST_OK(palloc(&Qj, B2 * Sj * C));
    ST_OK(palloc(&Kj, B2 * Sj * C));
    ST_OK(palloc(&Vj, B2 * Sj * C));
    ST_OK(palloc(&attn_out, B2 * Sj * C));
    ST_OK(palloc(&mlp_a, B2 * L * static_cast<size_t>(std::max(F, Hs))));
    ST_OK(palloc(&mlp_v, B2 * Lv * static_cast<size_t>(F)));
    ST_OK(palloc(&y_out, B2 * L * LAT));
    ST_OK(palloc(&audio, B2 * L * C));
    ST_OK(palloc(&vcond, B2 * Lv * C));
    ST_OK(palloc(&part_a, static_cast<size_t>(max_splits) * B2 * L * C));
    ST_OK(palloc(&part_v, static_cast<size_t>(max_splits) * B2 * Lv * 3 * C));   // also the visual QKV partials (3C wide)
    ST_OK(palloc(&lat_dev, static_cast<size_t>(B) * LAT * L));
    {
        const size_t US = static_cast<size_t>(U) * S, UT = static_cast<size_t>(U) * T, ULv = static_cast<size_t>(U) * Lv;
        ST_OK(palloc(&sc_clip, ULv * cfg.clip_dim));
        ST_OK(palloc(&sc_sync, US * cfg.sync_dim));
        ST_OK(palloc(&sc_text, UT * cfg.text_dim));
        ST_OK(palloc(&sc_s0, US * cfg.sync_dim));
        ST_OK(palloc(&sc_s1, US * C));
        ST_OK(palloc(&sc_s2, US * Hy));
        ST_OK(palloc(&sc_s3, US * C));
        ST_OK(palloc(&sc_c1, UT * C));
        ST_OK(palloc(&sc_c2, UT * C));
        ST_OK(palloc(&sc_kv, UT * NT * 2 * C));
        ST_OK(palloc(&sc_v1, ULv * C));
        ST_OK(palloc(&sc_idx, L));
        ST_OK(palloc(&uq_first, US));
        ST_OK(palloc(&uq_src, US));
        ST_OK(palloc(&uq_tok_row, static_cast<size_t>(U) * L));
    }
    p.valid = true;
    plan = p;
    // group-dependent buffers are (re)sized by prepare_timesteps
    vectok_act = nullptr; mod_single = nullptr; vec_all = nullptr; mod_triple = nullptr; sigmas_dev = nullptr; t_dev = nullptr;
    sc_e = sc_h1 = sc_vs = nullptr;
    sol_d[0] = sol_d[1] = sol_d[2] = sol_samp = nullptr; sol_table = nullptr; sol_table_cap = 0;
    mod_rows_cap = 0;
    plan.n_t = 0;
    return FOLEY_OK;
}

// Tile-shape / K-split planner.  Measured (profiles/r01_experiments.md): ~0.35 us per 64-deep k-block for 64-, 128- and
// 256-wide tiles alike (the MMA issue rate, not TMA, sets it), ~5 us of launch + prologue + epilogue per CTA wave, so
// the cost model is  waves x (k-blocks per CTA x t_kb + t_fixed)  and the planner picks the (tile width, split) pair
// that minimises it.
void Engine::plan_gemm(int rows, int batch, int n, int kblocks, bool can_split, int* bn_out, int* splits_out, int split_cap) const {
    {
        auto it = plan_override.find(std::make_tuple(rows, batch, n, kblocks));
        if (it != plan_override.end()) {
            int s = can_split ? it->second.second : 1;
            if (split_cap > 0) s = std::min(s, split_cap);
            *bn_out = it->second.first;
            *splits_out = std::max(1, std::min(s, std::min(max_splits, max_splits_used)));
            plan_seen[std::make_tuple(rows, batch, n, kblocks)] = std::make_pair(*bn_out, *splits_out);
            return;
        }
    }
    const long long m_tiles = static_cast<long long>((rows + 127) / 128) * batch;
    double best = 1e30;
    int best_bn = 128, best_s = 1;
    int smax = can_split ? std::min(max_splits, max_splits_used) : 1;
    if (split_cap > 0) smax = std::min(smax, split_cap);
    for (int bn : {128, 256}) {
        const long long n_tiles = (n + bn - 1) / bn;
        const double t_kb = bn == 128 ? plan_tkb128 : plan_tkb256;
        for (int s = 1; s <= smax; ++s) {
            if (s > 1 && kblocks / s < 4) break;
            const long long ctas = m_tiles * n_tiles * s;
            const long long waves = (ctas + num_sms - 1) / num_sms;
            const int kb_per = (kblocks + s - 1) / s;
            const double t = waves * (kb_per * t_kb + plan_tfix) + (s - 1) * plan_tsplit;   // extra partials cost a little downstream
            if (t < best - 1e-9) { best = t; best_bn = bn; best_s = s; }
        }
    }
    *bn_out = best_bn;
    *splits_out = best_s;
    plan_seen[std::make_tuple(rows, batch, n, kblocks)] = std::make_pair(best_bn, best_s);
}

int Engine::pick_bn(int rows, int batch, int n, int kblocks) const {
    int bn, s;
    plan_gemm(rows, batch, n, kblocks, false, &bn, &s);
    return bn;
}

int Engine::pick_splits(int rows, int batch, int n, int kblocks, int bn) const {
    (void)bn;
    int b2, s;
    plan_gemm(rows, batch, n, kblocks, true, &b2, &s);
    return s;
}

foley_status Engine::gemm(cudaStream_t st, const bf16* A, int rows, int batch, long long lda, long long a_bs,
                          const LinearW& W, int n_off, int n_cnt, GemmEpi epi, int splits, int bn, int max_ctas) {
    if (skip_gemm_once) { skip_gemm_once = false; return FOLEY_OK; }          // tools/ablate_step.py only
    if (debug_skip & (1 << 4)) { if (st == side_stream) return FOLEY_OK; }     // visual-branch GEMMs
    GemmLaunch Lc;
    Lc.a.ptr = A; Lc.a.dtype = DT_BF16; Lc.a.k = W.k; Lc.a.rows = rows; Lc.a.batch = batch; Lc.a.ld = lda;
    Lc.a.batch_stride = a_bs;
    Lc.w = W.w + static_cast<long long>(n_off) * W.k * W.taps;
    Lc.n = n_cnt;
    Lc.taps = W.taps;
    Lc.tap_off0 = -(W.taps / 2);
    Lc.tap_stride = 1;
    Lc.splits = splits;
    Lc.bn = bn;
    Lc.max_ctas = max_ctas;
    if (epi.bias) epi.bias = reinterpret_cast<const bf16*>(epi.bias) + n_off;
    if (epi.out_batch_stride == 0) epi.out_batch_stride = static_cast<long long>(rows) * epi.ldo;
    Lc.epi = epi;
    std::string err;
    if (!launch_gemm(Lc, st, &err)) return fail(FOLEY_ERR_CUDA, err);
    ++launches;
    return FOLEY_OK;
}

// GEMM with fp32 K-split partials followed by the fused reduce + bias + gate + residual + LayerNorm/modulate.
foley_status Engine::proj_combine(cudaStream_t st, const bf16* A, int rows, int batch, long long lda, long long a_bs,
                                  const LinearW& W, float* partials, CombineArgs ca) {
    const int kblocks = W.k * W.taps / 64;
    int bn = 128, splits = 1;
    plan_gemm(rows, batch, W.n, kblocks, true, &bn, &splits, st == side_stream ? side_split_cap : 0);
    GemmEpi e;
    e.mode = EPI_F32;
    e.out = partials;
    e.ldo = W.n;
    e.out_batch_stride = static_cast<long long>(rows) * W.n;
    e.split_stride = static_cast<long long>(rows) * batch * W.n;
    ST_OK(gemm(st, A, rows, batch, lda, a_bs, W, 0, W.n, e, splits, bn));
    ca.partials = partials;
    ca.splits = splits;
    ca.split_stride = e.split_stride;
    ca.C = W.n;
    ca.rows_total = rows * batch;
    if (debug_skip & (1 << 2)) return FOLEY_OK;
    FOLEY_CUDA_OK(launch_k(combine_ln_mod_kernel, dim3(ca.rows_total), dim3(ca.C / 4), 0, st, ca));
    ++launches;
    return FOLEY_OK;
}

// ------------------------------------------------------------------------------------------------ conditions
static void host_rope_table(const std::vector<int>& pos, float theta, std::vector<float>* cos_t, std::vector<float>* sin_t) {
    // get_1d_rotary_pos_embed(use_real=True), posemb_layers.py:160-168: fp32 pow, fp32 outer product
    cos_t->resize(pos.size() * 128);
    sin_t->resize(pos.size() * 128);
    for (size_t p = 0; p < pos.size(); ++p)
        for (int k = 0; k < 64; ++k) {
            const float freq = powf(theta, -(static_cast<float>(2 * k) / 128.0f));
            const float ang = static_cast<float>(pos[p]) * freq;
            const float c = cosf(ang), s = sinf(ang);
            (*cos_t)[p * 128 + 2 * k] = c; (*cos_t)[p * 128 + 2 * k + 1] = c;
            (*sin_t)[p * 128 + 2 * k] = s; (*sin_t)[p * 128 + 2 * k + 1] = s;
        }
}
static std::vector<int> nearest_exact_index(int n_in, int n_out) {
    // ATen: floorf((dst + 0.5f) * scale), scale = (float)n_in / n_out   (UpSample.h nearest_exact_idx)
    std::vector<int> idx(n_out);
    const float scale = static_cast<float>(n_in) / static_cast<float>(n_out);
    for (int i = 0; i < n_out; ++i) {
        volatile float f = (static_cast<float>(i) + 0.5f) * scale;
        idx[i] = std::min(static_cast<int>(floorf(f)), n_in - 1);
    }
    return idx;
}

foley_status Engine::set_conditions(const void* clip, const void* sync, const void* text, int dtype, int U, int Lv,
                                    int S, int T, int L, int B, cudaStream_t st) {
    if (!finalized) return fail(FOLEY_ERR_STATE, "set_conditions before finalize");
    if (U < 1 || U > 2 || B < 1 || L < 1 || Lv < 1 || S < 8 || S % 8 != 0 || T < 1)
        return fail(FOLEY_ERR_INVALID, "set_conditions: bad shape (n_cond in {1,2}, S multiple of 8)");
    if (L < Lv) return fail(FOLEY_ERR_UNSUPPORTED, "set_conditions: L < Lv is not supported");
    FOLEY_CUDA_OK(cudaSetDevice(device));
    if (!plan.valid || plan.B != B || plan.U != U || plan.L != L || plan.Lv != Lv || plan.S != S || plan.T != T)
        ST_OK(alloc_plan(B, U, L, Lv, S, T));
    const Plan& p = plan;
    // scratch (freed with the plan): bf16 copies of the inputs and the embedder intermediates
    bf16 *clip_b = sc_clip, *sync_b = sc_sync, *text_b = sc_text, *s0 = sc_s0, *s1 = sc_s1, *s2 = sc_s2, *s3 = sc_s3;
    bf16 *c1 = sc_c1, *c2 = sc_c2, *kv_all = sc_kv, *v1 = sc_v1;
    int* idx_dev = sc_idx;
    const size_t US = static_cast<size_t>(U) * S, UT = static_cast<size_t>(U) * T, ULv = static_cast<size_t>(U) * Lv;
    auto to_bf16 = [&](const void* src, bf16* dst, long long n) -> foley_status {
        convert_to_bf16_kernel<<<blocks_for(n, 256), 256, 0, st>>>(src, dtype, n, dst, 0);
        FOLEY_CUDA_OK(cudaGetLastError());
        ++launches;
        return FOLEY_OK;
    };
    ST_OK(to_bf16(clip, clip_b, ULv * cfg.clip_dim));
    ST_OK(to_bf16(sync, sync_b, US * cfg.sync_dim));
    ST_OK(to_bf16(text, text_b, UT * cfg.text_dim));

    auto bf = [&](bf16* out, int ldo, const bf16* bias, int act, int mode = EPI_BF16) {
        GemmEpi e;
        e.mode = mode; e.act = act; e.out = out; e.ldo = ldo; e.bias = bias;
        return e;
    };
    // ---- sync branch (hifi_foley.py:755-762)
    sync_add_pos_kernel<<<blocks_for(US * cfg.sync_dim, 256), 256, 0, st>>>(sync_b, sync_pos_emb, US, cfg.sync_dim, S, s0);
    ++launches;
    ST_OK(gemm(st, s0, US, 1, cfg.sync_dim, 0, sync0, 0, C, bf(s1, C, sync0.b, ACT_SILU), 1, 128));
    ST_OK(gemm(st, s1, US, 1, C, 0, sync_w13, 0, 2 * Hy, bf(s2, Hy, nullptr, 0, EPI_SWIGLU), 1, 128));
    ST_OK(gemm(st, s2, US, 1, Hy, 0, sync_w2, 0, C, bf(s3, C, nullptr, 0), 1, 128));
    {
        if (!tables_ready) {   // index / RoPE tables depend only on the plan's shapes: built once per plan
            std::vector<int> idx = nearest_exact_index(S, L);
            FOLEY_CUDA_OK(cudaMemcpyAsync(idx_dev, idx.data(), L * sizeof(int), cudaMemcpyHostToDevice, st));
            FOLEY_CUDA_OK(cudaStreamSynchronize(st));
        }
        gather_rows_kernel<<<blocks_for(static_cast<long long>(U) * L * C / 8, 256), 256, 0, st>>>(s3, idx_dev, U, S, L, C, a_sync);
        ++launches;
    }
    {
        // Distinct rows of the sync-token table: the single-block modulation vectors (13 % of the reference's step) only
        // need to be computed once per distinct row (step()).  Exact row compare on the device, grouping on the host.
        row_first_equal_kernel<<<static_cast<unsigned>(US), 128, 0, st>>>(s3, static_cast<int>(US), C, uq_first);
        ++launches;
        std::vector<int> first(US), src, id(US), idx(L), tok(static_cast<size_t>(U) * L);
        FOLEY_CUDA_OK(cudaMemcpyAsync(first.data(), uq_first, US * sizeof(int), cudaMemcpyDeviceToHost, st));
        FOLEY_CUDA_OK(cudaMemcpyAsync(idx.data(), idx_dev, L * sizeof(int), cudaMemcpyDeviceToHost, st));
        FOLEY_CUDA_OK(cudaStreamSynchronize(st));
        for (size_t r = 0; r < US; ++r) {
            if (first[r] == static_cast<int>(r)) { id[r] = static_cast<int>(src.size()); src.push_back(static_cast<int>(r)); }
            else id[r] = id[first[r]];
        }
        for (int u = 0; u < U; ++u)
            for (int l = 0; l < L; ++l) tok[static_cast<size_t>(u) * L + l] = id[static_cast<size_t>(u) * S + idx[l]];
        FOLEY_CUDA_OK(cudaMemcpyAsync(uq_src, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        FOLEY_CUDA_OK(cudaMemcpyAsync(uq_tok_row, tok.data(), tok.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        FOLEY_CUDA_OK(cudaStreamSynchronize(st));   // host vectors go out of scope
        if (static_cast<int>(src.size()) != n_uq) graph_valid = false;
        n_uq = static_cast<int>(src.size());
    }
    // ---- RoPE tables (hifi_foley.py:797-803, 151-166, 865; positions per oracle.interleaved_positions)
    if (!tables_ready) {
        std::vector<int> pa(L), pv(Lv), pp(std::max(L, std::max(Lv, T)));
        std::vector<int> pick = nearest_exact_index(L, Lv);
        for (int i = 0; i < L; ++i) pa[i] = 2 * i;
        for (int j = 0; j < Lv; ++j) pv[j] = (L == Lv) ? 2 * j + 1 : 2 * pick[j] + 1;
        for (size_t i = 0; i < pp.size(); ++i) pp[i] = static_cast<int>(i);
        std::vector<float> c, s;
        std::vector<float2> cs2;
        auto up = [&](const std::vector<int>& pos, float* dc, float* ds, float2* d2) -> foley_status {
            host_rope_table(pos, cfg.rope_theta, &c, &s);
            cs2.resize(pos.size() * 64);    // the same values, one (cos, sin) per rotation pair
            for (size_t p = 0; p < pos.size(); ++p)
                for (int k = 0; k < 64; ++k) cs2[p * 64 + k] = make_float2(c[p * 128 + 2 * k], s[p * 128 + 2 * k]);
            FOLEY_CUDA_OK(cudaMemcpyAsync(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice, st));
            FOLEY_CUDA_OK(cudaMemcpyAsync(ds, s.data(), s.size() * 4, cudaMemcpyHostToDevice, st));
            FOLEY_CUDA_OK(cudaMemcpyAsync(d2, cs2.data(), cs2.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
            FOLEY_CUDA_OK(cudaStreamSynchronize(st));
            return FOLEY_OK;
        };
        ST_OK(up(pa, rope_av_a_cos, rope_av_a_sin, rope2_av_a));
        ST_OK(up(pv, rope_av_v_cos, rope_av_v_sin, rope2_av_v));
        ST_OK(up(pp, rope_plain_cos, rope_plain_sin, rope2_plain));
        tables_ready = true;
    }
    // ---- text branch: cond_in, then K/V of every triple block at once (step-invariant, hifi_foley.py:289-308)
    ST_OK(gemm(st, text_b, UT, 1, cfg.text_dim, 0, cond1, 0, C, bf(c1, C, cond1.b, ACT_SILU), 1, 128));
    ST_OK(gemm(st, c1, UT, 1, C, 0, cond2, 0, C, bf(c2, C, cond2.b, 0), 1, 128));
    ST_OK(gemm(st, c2, UT, 1, C, 0, text_kv_all, 0, NT * 2 * C, bf(kv_all, NT * 2 * C, text_kv_all.b, 0), 1, 128));
    for (int i = 0; i < NT; ++i) {
        QkvArgs q;
        q.src = kv_all; q.src_ld = NT * 2 * C; q.n_parts = 2; q.H = H; q.L = T; q.rows_total = U * T;
        q.norm_kind = 0; q.eps = 1e-6f; q.cos = rope_plain_cos; q.sin = rope_plain_sin; q.pos = nullptr;
        const long long blk = static_cast<long long>(U) * T * C;
        q.part[0].dst = text_k + i * blk; q.part[0].dst_batch_stride = static_cast<long long>(T) * C;
        q.part[0].dst_head_stride = static_cast<long long>(T) * 128; q.part[0].seq_offset = 0;
        q.part[0].norm_w = triple[i].text_k_norm; q.part[0].src_col = i * 2 * C;
        q.part[1] = q.part[0];
        q.part[1].dst = text_v + i * blk; q.part[1].norm_w = nullptr; q.part[1].src_col = i * 2 * C + C;
        FOLEY_CUDA_OK(launch_k(qk_norm_rope_kernel, dim3(blocks_for(static_cast<long long>(q.rows_total) * 2 * H, 4)), dim3(128), 0, st, q));
        ++launches;
    }
    // ---- visual branch (hifi_foley.py:770)
    ST_OK(gemm(st, clip_b, ULv, 1, cfg.clip_dim, 0, vis_w13, 0, 2 * C, bf(v1, C, nullptr, 0, EPI_SWIGLU), 1, 128));
    ST_OK(gemm(st, v1, ULv, 1, C, 0, vis_w2, 0, C, bf(vcond0, C, nullptr, 0), 1, 128));
    FOLEY_CUDA_OK(cudaGetLastError());
    graph_valid = false;
    (void)p;
    return FOLEY_OK;
}

// time embedding + every triple-block modulation for the given timesteps (hifi_foley.py:744, 191-213)
foley_status Engine::prepare_timesteps(const float* t_host, int n_t, bool per_sample, cudaStream_t st) {
    if (!plan.valid) return fail(FOLEY_ERR_STATE, "conditions not set");
    Plan& p = plan;
    const int G = per_sample ? p.B2 : p.U;
    const size_t need_rows = static_cast<size_t>(per_sample ? G : 1) * p.U * p.S;
    if (n_t > p.n_t || G > p.G || !vec_all || !mod_single || need_rows > mod_rows_cap) {
        // (re)allocate timestep- and group-sized buffers
        p.n_t = std::max(n_t, p.n_t);
        p.G = std::max(G, p.G);
        FOLEY_CUDA_OK(cudaStreamSynchronize(st));   // the old buffers may still be read by work queued on this stream
        pfree(vec_all); pfree(mod_triple); pfree(t_dev); pfree(sigmas_dev); pfree(vectok_act); pfree(mod_single);
        pfree(sc_e); pfree(sc_h1); pfree(sc_vs);
        ST_OK(palloc(&vec_all, static_cast<size_t>(p.n_t) * C));
        ST_OK(palloc(&mod_triple, static_cast<size_t>(p.n_t) * NT * 18 * C));
        ST_OK(palloc(&t_dev, p.n_t));
        ST_OK(palloc(&sigmas_dev, p.n_t + 1));
        // rows: the distinct sync-token rows (<= U*S), once per time vector (one, or one per group when timesteps are per sample)
        const size_t mod_rows = std::max(need_rows, mod_rows_cap);
        ST_OK(palloc(&vectok_act, mod_rows * C));
        ST_OK(palloc(&mod_single, mod_rows * NS * 6 * C));
        mod_rows_cap = mod_rows;
        ST_OK(palloc(&sc_e, static_cast<size_t>(p.n_t) * cfg.freq_dim));
        ST_OK(palloc(&sc_h1, static_cast<size_t>(p.n_t) * C));
        ST_OK(palloc(&sc_vs, static_cast<size_t>(p.n_t) * C));
        graph_valid = false;
    }
    bf16 *e = sc_e, *h1 = sc_h1, *vs = sc_vs;
    FOLEY_CUDA_OK(cudaMemcpyAsync(t_dev, t_host, n_t * sizeof(float), cudaMemcpyHostToDevice, st));
    timestep_embed_kernel<<<blocks_for(static_cast<long long>(n_t) * cfg.freq_dim / 2, 128), 128, 0, st>>>(t_dev, n_t, cfg.freq_dim, e);
    ++launches;
    GemmEpi e1; e1.mode = EPI_BF16; e1.act = ACT_SILU; e1.out = h1; e1.ldo = C; e1.bias = time1.b;
    ST_OK(gemm(st, e, n_t, 1, cfg.freq_dim, 0, time1, 0, C, e1, 1, 128));
    GemmEpi e2; e2.mode = EPI_BF16; e2.out = vec_all; e2.ldo = C; e2.bias = time2.b;
    ST_OK(gemm(st, h1, n_t, 1, C, 0, time2, 0, C, e2, 1, 128));
    silu_bf16_kernel<<<blocks_for(static_cast<long long>(n_t) * C, 256), 256, 0, st>>>(vec_all, vs, static_cast<long long>(n_t) * C);
    ++launches;
    GemmEpi e3; e3.mode = EPI_BF16; e3.out = mod_triple; e3.ldo = static_cast<long long>(NT) * 18 * C; e3.bias = mod_triple_all.b;
    ST_OK(gemm(st, vs, n_t, 1, C, 0, mod_triple_all, 0, NT * 18 * C, e3, 1, 128));
    // group maps
    std::vector<int> gos(p.B2), tog(p.B2, 0), cog(p.B2, 0);
    for (int b = 0; b < p.B2; ++b) gos[b] = per_sample ? b : b / p.B;
    for (int g = 0; g < G; ++g) {
        cog[g] = per_sample ? g / p.B : g;
        tog[g] = (per_sample && n_t > 1) ? g : 0;
    }
    FOLEY_CUDA_OK(cudaMemcpyAsync(grp_of_sample, gos.data(), p.B2 * sizeof(int), cudaMemcpyHostToDevice, st));
    FOLEY_CUDA_OK(cudaMemcpyAsync(trow_of_grp, tog.data(), p.B2 * sizeof(int), cudaMemcpyHostToDevice, st));
    FOLEY_CUDA_OK(cudaMemcpyAsync(cond_of_grp, cog.data(), p.B2 * sizeof(int), cudaMemcpyHostToDevice, st));
    FOLEY_CUDA_OK(cudaMemsetAsync(step_dev, 0, 4 * sizeof(int), st));
    FOLEY_CUDA_OK(cudaStreamSynchronize(st));  // host vectors go out of scope
    cur_G = G;
    if (cur_per_sample != per_sample) graph_valid = false;
    cur_per_sample = per_sample;
    return FOLEY_OK;
}

// ------------------------------------------------------------------------------------------------ one DiT step
foley_status Engine::step(cudaStream_t st) {
    const Plan& p = plan;
    const int B2 = p.B2, L = p.L, Lv = p.Lv, T = p.T, Sj = L + Lv, G = cur_G;
    RowMap rm_a{grp_of_sample, trow_of_grp, cond_of_grp, L};
    RowMap rm_v{grp_of_sample, trow_of_grp, cond_of_grp, Lv};
    const long long mt_stride = static_cast<long long>(NT) * 18 * C;
    const long long ms_tok = static_cast<long long>(NS) * 6 * C;
    auto tmod = [&](int blk, int stream) {
        ModRef m; m.base = mod_triple + (static_cast<long long>(blk) * 2 + stream) * 9 * C; m.sample_stride = mt_stride;
        m.tok_stride = 0; m.by_trow = 1; return m;
    };
    // Single-block modulation vectors are per token, but the per-token condition is the nearest-exact up-sampling of
    // the S sync tokens (hifi_foley.py:755-762): audio token l carries a copy of sync token idx[l].  They are computed
    // once per SYNC token (S ~ L/2 rows) and looked up through the index table — less than half of the reference's
    // modulation FLOPs (13 % of the step) for bit-identical values.
    // (set_conditions finds the n_uq distinct rows of the sync-token table and the row of every (condition, token).)
    const int G_eff = cur_per_sample ? G : 1;   // groups only differ in their time vector when timesteps are per sample
    auto smod = [&](int blk) {
        ModRef m; m.base = mod_single + static_cast<long long>(blk) * 6 * C; m.sample_stride = cur_per_sample ? ms_tok * n_uq : 0;
        m.tok_stride = ms_tok; m.by_trow = 0; m.tok_map = uq_tok_row; return m;
    };
    auto bf = [&](bf16* out, long long ldo, const bf16* bias, int act, int mode = EPI_BF16) {
        GemmEpi e; e.mode = mode; e.act = act; e.out = out; e.ldo = ldo; e.bias = bias; return e;
    };
    auto attn = [&](const bf16* q, const bf16* k, const bf16* v, int Sq, int Sk, long long kv_bs, long long kv_hs,
                    bool cross) -> foley_status {
        if (debug_skip & 1) return FOLEY_OK;
        AttnArgs a;
        a.q = q; a.k = k; a.v = v; a.o = attn_out; a.H = H; a.Sq = Sq; a.Sk = Sk;
        a.q_batch_stride = static_cast<long long>(Sq) * C; a.q_head_stride = static_cast<long long>(Sq) * 128;
        a.kv_batch_stride = kv_bs; a.kv_head_stride = kv_hs;
        a.o_batch_stride = static_cast<long long>(Sq) * C;
        a.kv_batch_map = cross ? cond_of_grp : nullptr;
        a.grp_of_sample = cross ? grp_of_sample : nullptr;
        a.scale_log2 = 1.4426950408889634f / sqrtf(128.0f);
        if (att_tc > 0 || (att_tc < 0 && Sk > ATC_CK)) {   // prepared [B,H,S,128] operands on the tcgen05 kernel
            AttOperand oq, ok_, ov;
            oq.ptr = q; oq.batch_stride = a.q_batch_stride; oq.head_stride = a.q_head_stride; oq.row_stride = 128;
            oq.rows = Sq; oq.heads = H; oq.batch = B2;
            ok_.ptr = k; ok_.batch_stride = kv_bs; ok_.head_stride = kv_hs; ok_.row_stride = 128;
            ok_.rows = Sk; ok_.heads = H; ok_.batch = cross ? p.U : B2;
            ov = ok_; ov.ptr = v;
            AttTcArgs t;
            t.o = attn_out; t.o_batch_stride = a.o_batch_stride; t.H = H; t.Sq = Sq; t.Sk = Sk;
            t.kv_batch_map = a.kv_batch_map; t.grp_of_sample = a.grp_of_sample; t.scale_log2 = a.scale_log2;
            std::string err;
            if (!launch_attention_tc(oq, ok_, ov, t, B2, st, &err)) return fail(FOLEY_ERR_CUDA, err);
            ++launches;
            return FOLEY_OK;
        }
        const long long ctas64 = static_cast<long long>((Sq + 63) / 64) * H * B2;
        if (ctas64 > 2LL * num_sms) {
            dim3 grid((Sq + 127) / 128, H, B2);
            FOLEY_CUDA_OK(launch_k(attention_kernel<8, 3>, grid, dim3(256), AttCfg<8, 3>::SMEM, st, a));
        } else if (att_kv_split) {   // small grid: 8 warps = two key-tile groups over the same 64 queries (in-CTA split-KV)
            dim3 grid((Sq + 63) / 64, H, B2);
            FOLEY_CUDA_OK(launch_k(attention_kernel<8, 3, 2>, grid, dim3(256), AttCfg<8, 3, 2>::SMEM, st, a));
        } else {
            dim3 grid((Sq + 63) / 64, H, B2);
            FOLEY_CUDA_OK(launch_k(attention_kernel<4, 4>, grid, dim3(128), AttCfg<4, 4>::SMEM, st, a));
        }
        ++launches;
        return FOLEY_OK;
    };

    // ---- single-block modulations for this step: x-independent, one GEMM for all NS blocks.  It only has to be ready
    // before the first single block, so it runs on its own graph branch underneath the triple-stream phase.
    const bool mod_branch = mod_on_branch && NT > 0;
    bool mod_joined = false;
    cudaStream_t sm_ = mod_branch ? mod_stream : st;
    {
        if (mod_branch) {
            FOLEY_CUDA_OK(cudaEventRecord(ev_fork, st));
            FOLEY_CUDA_OK(cudaStreamWaitEvent(sm_, ev_fork, 0));
        }
        const long long n4 = static_cast<long long>(G_eff) * n_uq * C / 4;
        FOLEY_CUDA_OK(launch_k(vectok_silu_kernel, dim3(blocks_for(n4, 256)), dim3(256), 0, sm_, sc_s3, vec_all, uq_src,
                               trow_of_grp, G_eff, n_uq, C, vectok_act));
        ++launches;
        skip_gemm_once = (debug_skip >> 3) & 1;
        ST_OK(gemm(sm_, vectok_act, G_eff * n_uq, 1, C, 0, mod_single_all, 0, NS * 6 * C,
                   bf(mod_single, ms_tok, mod_single_all.b, 0), 1, 256, mod_branch ? mod_ctas : 0));
        if (mod_branch) FOLEY_CUDA_OK(cudaEventRecord(ev_mod, sm_));
    }
    // The visual stream (B2*Lv = 80 rows at 5 s) is pure launch/latency overhead next to the audio stream: run it on a
    // side stream (a parallel branch of the captured graph) and meet only at the two attention calls of a block.
    cudaStream_t sv = side_stream;
    auto fork = [&]() -> foley_status {
        FOLEY_CUDA_OK(cudaEventRecord(ev_fork, st));
        FOLEY_CUDA_OK(cudaStreamWaitEvent(sv, ev_fork, 0));
        return FOLEY_OK;
    };
    auto join = [&]() -> foley_status {
        FOLEY_CUDA_OK(cudaEventRecord(ev_join, sv));
        FOLEY_CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0));
        return FOLEY_OK;
    };
    auto combine_on = [&](cudaStream_t s_, const CombineArgs& ca) -> foley_status {
        if (debug_skip & (1 << 2)) return FOLEY_OK;
        FOLEY_CUDA_OK(launch_k(combine_ln_mod_kernel, dim3(ca.rows_total), dim3(ca.C / 4), 0, s_, ca));
        ++launches;
        return FOLEY_OK;
    };
    auto qknorm_on = [&](cudaStream_t s_, const QkvArgs& q) -> foley_status {
        if (debug_skip & (1 << 1)) return FOLEY_OK;
        FOLEY_CUDA_OK(launch_k(qk_norm_rope_kernel, dim3(blocks_for(static_cast<long long>(q.rows_total) * q.n_parts * H, 4)),
                               dim3(128), 0, s_, q));
        ++launches;
        return FOLEY_OK;
    };
    // Audio-stream QKV / cross-Q projections: per-k-block cost does not depend on the tile width (§3.1), so the fastest
    // shape is the widest tile with K split until the SMs are full; the q/k-norm + RoPE kernel that consumes the result
    // sums the fp32 partials (+ bias, bf16 rounding) itself.  Returns the args the consumer needs in `q`.
    auto qkv_gemm = [&](cudaStream_t s_, const bf16* A, int rows, int batch, const LinearW& W, int n_cols, float* partials,
                        int part_cols, bf16* out_bf16, QkvArgs* q) -> foley_status {
        int bn = 128, splits = 1;
        int cap_ = max_splits * part_cols / n_cols;   // workspace cap
        if (s_ == side_stream && side_split_cap > 0) cap_ = std::min(cap_, side_split_cap);
        plan_gemm(rows, batch, n_cols, W.k / 64, qkv_split, &bn, &splits, cap_);
        const long long a_bs = static_cast<long long>(rows) * C;
        if (splits > 1) {
            GemmEpi e;
            e.mode = EPI_F32; e.out = partials; e.ldo = n_cols;
            e.out_batch_stride = static_cast<long long>(rows) * n_cols;
            e.split_stride = static_cast<long long>(rows) * batch * n_cols;
            ST_OK(gemm(s_, A, rows, batch, C, a_bs, W, 0, n_cols, e, splits, bn));
            q->partials = partials; q->splits = splits; q->split_stride = e.split_stride; q->bias = W.b; q->src = nullptr;
        } else {
            ST_OK(gemm(s_, A, rows, batch, C, a_bs, W, 0, n_cols, bf(out_bf16, n_cols, W.b, 0), 1, bn));
            q->src = out_bf16;
        }
        q->src_ld = n_cols;
        return FOLEY_OK;
    };
    const int RV = B2 * Lv;   // visual rows, flattened (no token conv on this stream, so samples need no halo)
    // Fused attention (attention_tc.cuh): the projection GEMMs of both streams write bf16 rows into ONE [B2][Lv + L][n]
    // buffer (visual rows first) and the attention kernel normalises / rotates them while loading its operands — no
    // qk_norm_rope_kernel launch, no [B,H,S,128] round trip.  Needs all keys resident (<= 320), else the prepared path.
    const bool fused = att_tc != 0 && att_fused && Sj <= ATC_CK && !(debug_skip & 3);
    auto fused_proj = [&](cudaStream_t s_, const bf16* A, int rows, const LinearW& W, int n_cols, int row_off) -> foley_status {
        GemmEpi e = bf(qkv_j + static_cast<long long>(row_off) * n_cols, n_cols, W.b, 0);
        e.out_batch_stride = static_cast<long long>(Sj) * n_cols;
        return gemm(s_, A, rows, B2, C, static_cast<long long>(rows) * C, W, 0, n_cols, e, 1, pick_bn(rows, B2, n_cols, C / 64));
    };
    auto att_operand = [&](const bf16* ptr, long long row_stride, int rows, int batch, long long batch_stride, long long head_stride) {
        AttOperand o; o.ptr = ptr; o.row_stride = row_stride; o.rows = rows; o.batch = batch; o.batch_stride = batch_stride;
        o.head_stride = head_stride; o.heads = H; return o;
    };
    auto fused_attn = [&](const AttOperand& oq, const AttOperand& ok_, const AttOperand& ov, AttTcArgs t) -> foley_status {
        t.o = attn_out; t.o_batch_stride = static_cast<long long>(oq.rows) * C; t.H = H; t.Sq = oq.rows; t.Sk = ok_.rows;
        t.scale_log2 = 1.4426950408889634f / sqrtf(128.0f);
        std::string err;
        if (!launch_attention_tc(oq, ok_, ov, t, B2, st, &err)) return fail(FOLEY_ERR_CUDA, err);
        ++launches;
        return FOLEY_OK;
    };
    // ---- embed: audio0 = audio_embedder(x) + a_sync (fp32), v_cond0; LN+modulate for block 0
    {
        CombineArgs ca;
        ca.bias = audio_embed.b; ca.x = audio; ca.x_init = a_sync; ca.h = h_a; ca.eps = 1e-6f;
        ca.mod = tmod(0, 0); ca.shift_chunk = 0; ca.scale_chunk = 1; ca.rm = rm_a;
        ST_OK(proj_combine(st, x_in, L, B2, LAT, static_cast<long long>(L) * LAT, audio_embed, part_a, ca));
        ST_OK(fork());
        CombineArgs cv;
        cv.x = vcond; cv.x_init = vcond0; cv.h = h_v; cv.eps = 1e-6f; cv.mod = tmod(0, 1); cv.rm = rm_v;
        cv.C = C; cv.rows_total = RV;
        ST_OK(combine_on(sv, cv));
    }
    const long long jb = static_cast<long long>(Sj) * C, jh = static_cast<long long>(Sj) * 128;
    for (int i = 0; i < ((debug_skip >> 9) & 1 ? 0 : NT); ++i) {
        const TripleW& w = triple[i];
        // -- joint self attention
        if (fused) {
            ST_OK(fused_proj(st, h_a, L, w.qkv[0], 3 * C, Lv));
            ST_OK(fused_proj(sv, h_v, Lv, w.qkv[1], 3 * C, 0));
            ST_OK(join());
            const long long rs = 3LL * C, bs = static_cast<long long>(Sj) * rs;
            AttTcArgs t;
            t.norm_kind = 0; t.eps = 1e-6f;
            t.qn.rows0 = t.kn.rows0 = Lv;
            t.qn.w[0] = w.self_q_norm[1]; t.qn.w[1] = w.self_q_norm[0];
            t.kn.w[0] = w.self_k_norm[1]; t.kn.w[1] = w.self_k_norm[0];
            t.qn.rope[0] = t.kn.rope[0] = rope2_av_v; t.qn.rope[1] = t.kn.rope[1] = rope2_av_a;
            ST_OK(fused_attn(att_operand(qkv_j, rs, Sj, B2, bs, 128), att_operand(qkv_j + C, rs, Sj, B2, bs, 128),
                             att_operand(qkv_j + 2 * C, rs, Sj, B2, bs, 128), t));
        } else {
        QkvArgs qa_joint, qv_joint;
        ST_OK(qkv_gemm(st, h_a, L, B2, w.qkv[0], 3 * C, part_a, C, qkv_a, &qa_joint));
        ST_OK(qkv_gemm(sv, h_v, RV, 1, w.qkv[1], 3 * C, part_v, 3 * C, qkv_v, &qv_joint));
        for (int s = 0; s < 2; ++s) {
            QkvArgs q = s == 0 ? qa_joint : qv_joint;
            q.n_parts = 3; q.H = H; q.L = s == 0 ? L : Lv;
            q.rows_total = B2 * q.L; q.norm_kind = 0; q.eps = 1e-6f;
            q.cos = s == 0 ? rope_av_a_cos : rope_av_v_cos; q.sin = s == 0 ? rope_av_a_sin : rope_av_v_sin;
            bf16* dsts[3] = {Qj, Kj, Vj};
            const bf16* norms[3] = {w.self_q_norm[s], w.self_k_norm[s], nullptr};
            for (int pz = 0; pz < 3; ++pz) {
                q.part[pz].dst = dsts[pz]; q.part[pz].dst_batch_stride = jb; q.part[pz].dst_head_stride = jh;
                q.part[pz].seq_offset = s == 0 ? Lv : 0; q.part[pz].norm_w = norms[pz]; q.part[pz].src_col = pz * C;
            }
            ST_OK(qknorm_on(s == 0 ? st : sv, q));
        }
        ST_OK(join());
        ST_OK(attn(Qj, Kj, Vj, Sj, Sj, jb, jh, false));
        }
        ST_OK(fork());
        {
            CombineArgs ca;
            ca.bias = w.self_proj[0].b; ca.gate = tmod(i, 0); ca.gate_chunk = 2; ca.x = audio; ca.h = h_a; ca.eps = 1e-6f;
            ca.mod = tmod(i, 0); ca.shift_chunk = 3; ca.scale_chunk = 4; ca.rm = rm_a;
            ST_OK(proj_combine(st, attn_out + static_cast<long long>(Lv) * C, L, B2, C, jb, w.self_proj[0], part_a, ca));
            CombineArgs cv = ca;
            cv.bias = w.self_proj[1].b; cv.gate = tmod(i, 1); cv.mod = tmod(i, 1); cv.x = vcond; cv.h = h_v; cv.rm = rm_v;
            cv.round_x = 1;
            ST_OK(proj_combine(sv, attn_out, Lv, B2, C, jb, w.self_proj[1], part_v, cv));
        }
        // -- cross attention to text
        const long long tblk = static_cast<long long>(p.U) * T * C;
        if (fused && T <= ATC_CK) {
            ST_OK(fused_proj(st, h_a, L, w.cross_q[0], C, Lv));
            ST_OK(fused_proj(sv, h_v, Lv, w.cross_q[1], C, 0));
            ST_OK(join());
            AttTcArgs t;
            t.norm_kind = 0; t.eps = 1e-6f;
            t.qn.rows0 = Lv;
            t.qn.w[0] = w.cross_q_norm[1]; t.qn.w[1] = w.cross_q_norm[0];
            t.qn.rope[0] = t.qn.rope[1] = rope2_plain;
            t.kv_batch_map = cond_of_grp; t.grp_of_sample = grp_of_sample;
            const long long tb = static_cast<long long>(T) * C, th = static_cast<long long>(T) * 128;
            ST_OK(fused_attn(att_operand(qkv_j, C, Sj, B2, static_cast<long long>(Sj) * C, 128),
                             att_operand(text_k + i * tblk, 128, T, p.U, tb, th), att_operand(text_v + i * tblk, 128, T, p.U, tb, th), t));
        } else {
        QkvArgs qa_cross, qv_cross;
        ST_OK(qkv_gemm(st, h_a, L, B2, w.cross_q[0], C, part_a, C, qkv_a, &qa_cross));
        ST_OK(qkv_gemm(sv, h_v, RV, 1, w.cross_q[1], C, part_v, 3 * C, qkv_v, &qv_cross));
        for (int s = 0; s < 2; ++s) {
            QkvArgs q = s == 0 ? qa_cross : qv_cross;
            q.n_parts = 1; q.H = H; q.L = s == 0 ? L : Lv;
            q.rows_total = B2 * q.L; q.norm_kind = 0; q.eps = 1e-6f; q.cos = rope_plain_cos; q.sin = rope_plain_sin;
            q.part[0].dst = Qj; q.part[0].dst_batch_stride = jb; q.part[0].dst_head_stride = jh;
            q.part[0].seq_offset = s == 0 ? Lv : 0; q.part[0].norm_w = w.cross_q_norm[s]; q.part[0].src_col = 0;
            ST_OK(qknorm_on(s == 0 ? st : sv, q));
        }
        ST_OK(join());
        ST_OK(attn(Qj, text_k + i * tblk, text_v + i * tblk, Sj, T, static_cast<long long>(T) * C,
                   static_cast<long long>(T) * 128, true));
        }
        {
            ST_OK(fork());
            CombineArgs ca;
            ca.bias = w.cross_proj[0].b; ca.gate = tmod(i, 0); ca.gate_chunk = 5; ca.x = audio; ca.h = h_a; ca.eps = 1e-6f;
            ca.mod = tmod(i, 0); ca.shift_chunk = 6; ca.scale_chunk = 7; ca.rm = rm_a;
            ST_OK(proj_combine(st, attn_out + static_cast<long long>(Lv) * C, L, B2, C, jb, w.cross_proj[0], part_a, ca));
            CombineArgs cv = ca;
            cv.bias = w.cross_proj[1].b; cv.gate = tmod(i, 1); cv.mod = tmod(i, 1); cv.x = vcond; cv.h = h_v; cv.rm = rm_v;
            cv.round_x = 1;
            ST_OK(proj_combine(sv, attn_out, Lv, B2, C, jb, w.cross_proj[1], part_v, cv));
        }
        // -- MLPs
        if (mod_branch && i == NT - 1) { FOLEY_CUDA_OK(cudaStreamWaitEvent(st, ev_mod, 0)); mod_joined = true; }   // next: LN-modulate with single-block params
        ST_OK(gemm(st, h_a, L, B2, C, static_cast<long long>(L) * C, w.fc1[0], 0, F, bf(mlp_a, F, w.fc1[0].b, ACT_GELU_TANH), 1, pick_bn(L, B2, F, C / 64)));
        ST_OK(gemm(sv, h_v, RV, 1, C, 0, w.fc1[1], 0, F, bf(mlp_v, F, w.fc1[1].b, ACT_GELU_TANH), 1, 64));
        {
            const bool last = i == NT - 1;
            CombineArgs ca;
            ca.bias = w.fc2[0].b; ca.gate = tmod(i, 0); ca.gate_chunk = 8; ca.x = audio; ca.h = h_a; ca.rm = rm_a;
            if (!last) { ca.eps = 1e-6f; ca.mod = tmod(i + 1, 0); ca.shift_chunk = 0; ca.scale_chunk = 1; }
            else if (NS > 0) { ca.eps = 1e-5f; ca.mod = smod(0); ca.shift_chunk = 0; ca.scale_chunk = 1; }
            else { ca.eps = 1e-6f; ca.mod = ModRef(); }
            ST_OK(proj_combine(st, mlp_a, L, B2, F, static_cast<long long>(L) * F, w.fc2[0], part_a, ca));
            CombineArgs cv;
            cv.bias = w.fc2[1].b; cv.gate = tmod(i, 1); cv.gate_chunk = 8; cv.x = vcond; cv.rm = rm_v; cv.round_x = 1;
            if (!last) { cv.h = h_v; cv.eps = 1e-6f; cv.mod = tmod(i + 1, 1); cv.shift_chunk = 0; cv.scale_chunk = 1; }
            ST_OK(proj_combine(sv, mlp_v, RV, 1, F, 0, w.fc2[1], part_v, cv));
        }
    }
    ST_OK(join());   // the side branch must be complete before the step (graph) ends
    if (mod_branch && !mod_joined) FOLEY_CUDA_OK(cudaStreamWaitEvent(st, ev_mod, 0));   // (triple loop skipped by an ablation mask)
    // ---- single-stream blocks (hifi_foley.py:364-390)
    const long long sb = static_cast<long long>(L) * C, sh = static_cast<long long>(L) * 128;
    for (int j = 0; j < ((debug_skip >> 10) & 1 ? 0 : NS); ++j) {
        const SingleW& w = single[j];
        skip_gemm_once = (debug_skip >> 5) & 1;
        if (fused) {
            GemmEpi e = bf(qkv_j, 3 * C, w.qkv.b, 0);
            ST_OK(gemm(st, h_a, L, B2, C, sb, w.qkv, 0, 3 * C, e, 1, pick_bn(L, B2, 3 * C, C / 64)));
            const long long rs = 3LL * C, bs = static_cast<long long>(L) * rs;
            AttTcArgs t;
            t.norm_kind = 1; t.eps = cfg.single_rms_eps;
            t.qn.rows0 = t.kn.rows0 = L;
            t.qn.w[0] = w.q_norm; t.kn.w[0] = w.k_norm;
            t.qn.rope[0] = t.kn.rope[0] = rope2_plain;
            ST_OK(fused_attn(att_operand(qkv_j, rs, L, B2, bs, 128), att_operand(qkv_j + C, rs, L, B2, bs, 128),
                             att_operand(qkv_j + 2 * C, rs, L, B2, bs, 128), t));
        } else {
            QkvArgs q;
            ST_OK(qkv_gemm(st, h_a, L, B2, w.qkv, 3 * C, part_a, C, qkv_a, &q));
            q.n_parts = 3; q.H = H; q.L = L; q.rows_total = B2 * L;
            q.norm_kind = 1; q.eps = cfg.single_rms_eps; q.cos = rope_plain_cos; q.sin = rope_plain_sin;
            bf16* dsts[3] = {Qj, Kj, Vj};
            const bf16* norms[3] = {w.q_norm, w.k_norm, nullptr};
            for (int pz = 0; pz < 3; ++pz) {
                q.part[pz].dst = dsts[pz]; q.part[pz].dst_batch_stride = sb; q.part[pz].dst_head_stride = sh;
                q.part[pz].seq_offset = 0; q.part[pz].norm_w = norms[pz]; q.part[pz].src_col = pz * C;
            }
            ST_OK(qknorm_on(st, q));
            ST_OK(attn(Qj, Kj, Vj, L, L, sb, sh, false));
        }
        {
            CombineArgs ca;
            ca.bias = w.linear1.b; ca.gate = smod(j); ca.gate_chunk = 2; ca.x = audio; ca.h = h_a; ca.eps = 1e-5f;
            ca.mod = smod(j); ca.shift_chunk = 3; ca.scale_chunk = 4; ca.rm = rm_a;
            skip_gemm_once = (debug_skip >> 8) & 1;
            ST_OK(proj_combine(st, attn_out, L, B2, C, sb, w.linear1, part_a, ca));
        }
        skip_gemm_once = (debug_skip >> 6) & 1;
        ST_OK(gemm(st, h_a, L, B2, C, sb, w.w13, 0, 2 * Hs, bf(mlp_a, Hs, nullptr, 0, EPI_SWIGLU), 1, pick_bn(L, B2, 2 * Hs, 3 * C / 64)));
        {
            const bool last = j == NS - 1;
            CombineArgs ca;
            ca.gate = smod(j); ca.gate_chunk = 5; ca.x = audio; ca.h = h_a; ca.rm = rm_a;
            if (!last) { ca.eps = 1e-5f; ca.mod = smod(j + 1); ca.shift_chunk = 0; ca.scale_chunk = 1; }
            else { ca.eps = 1e-6f; ca.mod = ModRef(); }   // final LayerNorm, adaLN is a no-op (mlp_layers.py:97-101)
            skip_gemm_once = (debug_skip >> 7) & 1;
            ST_OK(proj_combine(st, mlp_a, L, B2, Hs, static_cast<long long>(L) * Hs, w.w2, part_a, ca));
        }
    }
    // ---- final linear -> y_out bf16 [B2, L, latent]
    ST_OK(gemm(st, h_a, L, B2, C, sb, final_lin, 0, LAT, bf(y_out, LAT, final_lin.b, 0), 1, 64));
    return FOLEY_OK;
}

// ------------------------------------------------------------------------------------------------ public ops
foley_status Engine::forward(const float* x, const float* t, int n_t, float* out, cudaStream_t st) {
    if (!plan.valid) return fail(FOLEY_ERR_STATE, "foley_dit_forward before foley_set_conditions");
    const Plan& p = plan;
    if (n_t != 1 && n_t != p.B2) return fail(FOLEY_ERR_INVALID, "n_t must be 1 or batch*n_cond");
    FOLEY_CUDA_OK(cudaSetDevice(device));
    ST_OK(prepare_timesteps(t, n_t, n_t > 1, st));
    dim3 blk(32, 8), grid((p.L + 31) / 32, (LAT + 31) / 32, p.B2);
    latents_to_tokens_kernel<<<grid, blk, 0, st>>>(x, p.B2, 1, LAT, p.L, x_in);
    ++launches;
    ST_OK(step(st));
    tokens_to_channels_kernel<<<grid, blk, 0, st>>>(y_out, p.B2, LAT, p.L, out);
    ++launches;
    FOLEY_CUDA_OK(cudaGetLastError());
    return FOLEY_OK;
}

// Per-call stage table of the reference scheduler's multi-stage solvers (scheduling_flow_match_discrete.py:299-373).
// Call i feeds timesteps[i] to the model (utils.py:215-246) while sigma / sigma_next come from `step_index`, which
// only advances after a solver's last stage.
std::vector<SolverCall> solver_table(int solver, const float* sigmas, int n_calls) {
    const int n_stages = solver == FOLEY_SOLVER_KUTTA4 ? 4 : 2;
    std::vector<SolverCall> tab(n_calls);
    int step_index = 0, stage = 0;
    float dt_full = 0.f;
    for (int i = 0; i < n_calls; ++i) {
        SolverCall c{};
        c.store_slot = -1;
        if (stage == 0) {
            dt_full = sigmas[step_index + 1] - sigmas[step_index];   // fp32, like the reference's tensors
            c.save_sample = 1;
            c.store_slot = 0;
            c.dt = solver == FOLEY_SOLVER_HEUN2 ? dt_full : dt_full / 2;
        } else if (stage < n_stages - 1) {   // kutta-4 stages 1, 2
            c.store_slot = stage;
            c.dt = stage == 1 ? dt_full / 2 : dt_full;
        } else {
            c.base_saved = 1;
            c.dt = dt_full;
            if (solver == FOLEY_SOLVER_HEUN2) c.kind = 1;
            else if (solver == FOLEY_SOLVER_KUTTA4) {
                c.kind = 2;
                c.c0 = c.cm = static_cast<float>(1.0 / 6.0);
                c.c1 = c.c2 = static_cast<float>(1.0 / 3.0);
            }
        }
        tab[i] = c;
        if (++stage == n_stages) { stage = 0; ++step_index; }
    }
    return tab;
}

foley_status Engine::denoise(float* latents, const float* sigmas, int n_steps, float guidance, int solver,
                             foley_progress_fn progress, void* user, cudaStream_t st) {
    if (!plan.valid) return fail(FOLEY_ERR_STATE, "foley_denoise before foley_set_conditions");
    if (n_steps < 1) return fail(FOLEY_ERR_INVALID, "n_steps < 1");
    if (solver < FOLEY_SOLVER_EULER || solver > FOLEY_SOLVER_KUTTA4) return fail(FOLEY_ERR_INVALID, "unknown solver id");
    Plan& p = plan;
    FOLEY_CUDA_OK(cudaSetDevice(device));
    if (solver != FOLEY_SOLVER_EULER) {
        const size_t n = static_cast<size_t>(p.B) * LAT * p.L;
        if (!sol_samp) {
            for (int i = 0; i < 3; ++i) ST_OK(palloc(&sol_d[i], n));
            ST_OK(palloc(&sol_samp, n));
            graph_valid = false;
        }
        if (n_steps > sol_table_cap) {
            FOLEY_CUDA_OK(cudaStreamSynchronize(st));
            pfree(sol_table);
            ST_OK(palloc(&sol_table, static_cast<size_t>(n_steps)));
            sol_table_cap = n_steps;
            graph_valid = false;
        }
        const std::vector<SolverCall> tab = solver_table(solver, sigmas, n_steps);
        FOLEY_CUDA_OK(cudaMemcpyAsync(sol_table, tab.data(), tab.size() * sizeof(SolverCall), cudaMemcpyHostToDevice, st));
        FOLEY_CUDA_OK(cudaStreamSynchronize(st));   // `tab` goes out of scope
    }
    std::vector<float> ts(n_steps);
    for (int i = 0; i < n_steps; ++i) ts[i] = sigmas[i] * 1000.0f;   // scheduler.timesteps (scheduling_...py:151)
    ST_OK(prepare_timesteps(ts.data(), n_steps, false, st));
    FOLEY_CUDA_OK(cudaMemcpyAsync(sigmas_dev, sigmas, (n_steps + 1) * sizeof(float), cudaMemcpyHostToDevice, st));
    const size_t lat_bytes = static_cast<size_t>(p.B) * LAT * p.L * sizeof(float);
    FOLEY_CUDA_OK(cudaMemcpyAsync(lat_dev, latents, lat_bytes, cudaMemcpyDeviceToDevice, st));
    dim3 blk(32, 8), grid((p.L + 31) / 32, (LAT + 31) / 32, p.B);
    latents_to_tokens_kernel<<<grid, blk, 0, st>>>(lat_dev, p.B, p.U, LAT, p.L, x_in);
    ++launches;
    FOLEY_CUDA_OK(cudaStreamSynchronize(st));

    auto body = [&]() -> foley_status {
        ST_OK(step(st));
        if (solver == FOLEY_SOLVER_EULER)
            FOLEY_CUDA_OK(launch_k(cfg_euler_kernel, grid, blk, 0, st, y_out, lat_dev, x_in, p.B, p.U, LAT, p.L, guidance,
                                   sigmas_dev, step_dev));
        else   // the table entry decides the stage; the kernel itself is the same for every call and solver
            FOLEY_CUDA_OK(launch_k(cfg_solver_kernel, grid, blk, 0, st, y_out, lat_dev, x_in, sol_d[0], sol_d[1], sol_d[2],
                                   sol_samp, p.B, p.U, LAT, p.L, guidance, sol_table, step_dev));
        FOLEY_CUDA_OK(launch_k(advance_step_kernel, dim3(1), dim3(32), 0, st, step_dev, trow_of_grp, cur_G,
                               static_cast<volatile int*>(host_step_dev)));
        launches += 2;
        FOLEY_CUDA_OK(cudaGetLastError());
        return FOLEY_OK;
    };
    const bool use_graph = use_cuda_graph;
    int64_t per_step = 0;
    const bool solver_kind = solver != FOLEY_SOLVER_EULER;
    if (use_graph && (!graph_valid || graph_guidance != guidance || graph_solver_kind != solver_kind)) {
        if (step_graph) { cudaGraphExecDestroy(step_graph); step_graph = nullptr; }
        const int64_t before = launches;
        cudaGraph_t g = nullptr;
        FOLEY_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        foley_status s = body();
        cudaError_t ce = cudaStreamEndCapture(st, &g);
        if (s != FOLEY_OK) { if (g) cudaGraphDestroy(g); return s; }
        if (ce != cudaSuccess) return fail(FOLEY_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&step_graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return fail(FOLEY_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ce));
        graph_launches_per_step = launches - before;
        launches = before;   // capture does not execute
        graph_valid = true;
        graph_guidance = guidance;
        graph_solver_kind = solver_kind;
    }
    per_step = graph_launches_per_step;
    *static_cast<volatile int*>(host_step) = 0;   // (the stream is idle: prepare_timesteps synchronized it)
    for (int i = 0; i < n_steps; ++i) {           // the whole loop is enqueued without a host round trip
        if (use_graph) {
            FOLEY_CUDA_OK(cudaGraphLaunch(step_graph, st));
            launches += per_step;
        } else {
            ST_OK(body());
        }
    }
    FOLEY_CUDA_OK(cudaMemcpyAsync(latents, lat_dev, lat_bytes, cudaMemcpyDeviceToDevice, st));
    if (progress) {
        // utils.py:247 `pbar.update(1)` per step, on the CALLER's thread: the device publishes the number of completed
        // steps in mapped host memory and this thread polls it while the GPU keeps running (no per-step synchronize).
        int reported = 0;
        for (;;) {
            const cudaError_t q = cudaStreamQuery(st);
            const int done = q == cudaSuccess ? n_steps : std::min(n_steps, static_cast<int>(*static_cast<volatile int*>(host_step)));
            while (reported < done) progress(++reported, user);
            if (q == cudaSuccess || reported >= n_steps) break;
            if (q != cudaErrorNotReady) return fail(FOLEY_ERR_CUDA, std::string("denoise: ") + cudaGetErrorString(q));
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
    }
    return FOLEY_OK;
}

foley_status Engine::debug_read(const char* what, float* dst, int64_t cap, int64_t* n_out) {
    if (!plan.valid) return fail(FOLEY_ERR_STATE, "no plan");
    const Plan& p = plan;
    const std::string w(what ? what : "");
    if (w == "plans") {       // every GEMM shape the planner has been asked about: 6 floats each (rows, batch, n, k-blocks, tile width, splits)
        std::vector<float> v;
        for (const auto& kv : plan_seen)
            for (float f : {static_cast<float>(std::get<0>(kv.first)), static_cast<float>(std::get<1>(kv.first)), static_cast<float>(std::get<2>(kv.first)),
                            static_cast<float>(std::get<3>(kv.first)), static_cast<float>(kv.second.first), static_cast<float>(kv.second.second)})
                v.push_back(f);
        if (n_out) *n_out = static_cast<int64_t>(v.size());
        if (!dst) return FOLEY_OK;
        if (cap < static_cast<int64_t>(v.size())) return fail(FOLEY_ERR_INVALID, "debug_read: destination too small");
        FOLEY_CUDA_OK(cudaMemcpy(dst, v.data(), v.size() * 4, cudaMemcpyDefault));
        return FOLEY_OK;
    }
    const void* src = nullptr;
    int64_t n = 0;
    bool is_bf16 = false;
    if (w == "audio") { src = audio; n = 1LL * p.B2 * p.L * C; }
    else if (w == "v_cond") { src = vcond; n = 1LL * p.B2 * p.Lv * C; }
    else if (w == "a_sync") { src = a_sync; n = 1LL * p.U * p.L * C; is_bf16 = true; }
    else if (w == "vcond0") { src = vcond0; n = 1LL * p.U * p.Lv * C; is_bf16 = true; }
    else if (w == "vec") { src = vec_all; n = 1LL * p.n_t * C; is_bf16 = true; }
    else if (w == "h_a") { src = h_a; n = 1LL * p.B2 * p.L * C; is_bf16 = true; }
    else if (w == "y") { src = y_out; n = 1LL * p.B2 * p.L * LAT; is_bf16 = true; }
    else if (w == "text_k") { src = text_k; n = 1LL * NT * p.U * p.T * C; is_bf16 = true; }
    else if (w == "attn_out") { src = attn_out; n = 1LL * p.B2 * (p.L + p.Lv) * C; is_bf16 = true; }
    else if (w == "mod_triple") { src = mod_triple; n = 1LL * p.n_t * NT * 18 * C; is_bf16 = true; }
    else return fail(FOLEY_ERR_INVALID, "debug_read: unknown buffer " + w);
    if (!src) return fail(FOLEY_ERR_STATE, "debug_read: buffer not allocated: " + w);
    if (n_out) *n_out = n;
    if (!dst) return FOLEY_OK;
    if (cap < n) return fail(FOLEY_ERR_INVALID, "debug_read: destination too small");
    FOLEY_CUDA_OK(cudaDeviceSynchronize());
    if (is_bf16) {
        float* tmp = nullptr;
        FOLEY_CUDA_OK(cudaMalloc(&tmp, n * 4));
        convert_to_f32_kernel<<<blocks_for(n, 256), 256>>>(src, 0, n, tmp);
        cudaError_t e = cudaMemcpy(dst, tmp, n * 4, cudaMemcpyDefault);
        cudaFree(tmp);
        FOLEY_CUDA_OK(e);
    } else {
        FOLEY_CUDA_OK(cudaMemcpy(dst, src, n * 4, cudaMemcpyDefault));
    }
    return FOLEY_OK;
}

}  // namespace foley
