// Condition encoders on the engine's kernels: SigLIP2 vision tower + pooling head
// (reference feature_utils.py:64-79 encode_video_with_siglip2 -> HF SiglipVisionTransformer) and the CLAP text tower
// (feature_utils.py:132-138 encode_text_feat -> HF ClapTextModel), weights taken under their HF state-dict names.
// The reference moves both modules to the device in the DiT's dtype before use (nodes.py:283-284), so this is the
// bf16 module: bf16 weights, fp32 accumulation, one bf16 rounding per op.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "attention_tc64.cuh"
#include "encoders.cuh"
#include "engine.cuh"
#include "gemm_host.cuh"
#include "safetensors.cuh"

namespace foley {

#ifndef ST_OK
#define ST_OK(expr)                         \
    do {                                    \
        foley_status _s = (expr);           \
        if (_s != FOLEY_OK) return _s;      \
    } while (0)
#endif

struct EncLayerW {
    LinearW qkv, o, fc1, fc2;
    bf16 *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
};

class Encoder {
  public:
    foley_encoder_config cfg{};
    int device = 0, num_sms = 148;
    int C = 0, H = 0, NL = 0, F = 0;
    bool finalized = false;
    int att_tc = 1;                        // self-attention of the vision tower: 1 = tcgen05 / TMEM kernel (attention_tc64.cuh) with two threads
                                           // per query row, 2 = the same with one thread per row, 0 = the mma.sync flash kernel
                                           // (FOLEY_ENC_ATT_TC; option "att_tc")
    int layers_run = -1;                   // option "layers_run": stop after this many layers (per-layer parity taps); -1 = all
    int64_t launches = 0;
    std::unordered_map<std::string, RawTensor> raw;
    std::vector<void*> packed;             // every device allocation finalize made
    std::vector<EncLayerW> layers;
    // SigLIP
    LinearW patch, head_q, head_kv, head_o, head_fc1, head_fc2;
    bf16 *pos_emb = nullptr, *post_ln_w = nullptr, *post_ln_b = nullptr, *head_ln_w = nullptr, *head_ln_b = nullptr;
    bf16 *probe = nullptr, *probe_q = nullptr;
    // CLAP
    bf16 *word = nullptr, *pos_tab = nullptr, *type_tab = nullptr, *emb_ln_w = nullptr, *emb_ln_b = nullptr;
    // Synchformer (MotionFormer visual extractor): Linear weights hold IEEE fp16 bits (the module runs under fp16 autocast)
    struct SyncBlockW {
        LinearW qkv_t, proj_t, qkv_s, proj_s, fc1, fc2;
        bf16 *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr, *n3w = nullptr, *n3b = nullptr;
    };
    std::vector<SyncBlockW> sblocks;
    LinearW s_patch, s_in_proj, s_out_proj, s_lin1, s_lin2;
    bf16 *s_norm_w = nullptr, *s_norm_b = nullptr, *s_an1w = nullptr, *s_an1b = nullptr, *s_an2w = nullptr, *s_an2b = nullptr;
    float *s_tok_tab = nullptr, *s_agg_cls = nullptr;
    float *xf = nullptr, *xn = nullptr;    // fp32 residual stream / normalised tokens (Synchformer only)
    // activations (grown on demand)
    long long cap_rows = 0;
    bf16 *x = nullptr, *h = nullptr, *qkv = nullptr, *att = nullptr, *y = nullptr, *mlp = nullptr;
    int *ids_dev = nullptr, *pos_dev = nullptr, *mask_dev = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_null = nullptr;

    ~Encoder();
    foley_status create(const foley_encoder_config* c, int dev);
    foley_status load_tensor(const char* name, const void* data, const int64_t* shape, int ndim, int dtype);
    foley_status load_safetensors(const char* path, const char* prefix, int64_t* n_loaded);
    foley_status finalize();
    foley_status siglip_encode(const float* pixels, int n_frames, void* out, cudaStream_t st);
    foley_status clap_encode(const int32_t* ids, const int32_t* mask, int B, int T, void* out, cudaStream_t st);
    foley_status synchformer_encode(const float* frames, int n_frames, float* out, cudaStream_t st);
    foley_status debug_read(const char* what, float* dst, int64_t cap, int64_t* n_out);
    cudaStream_t pick_stream(void* st) { return st ? static_cast<cudaStream_t>(st) : own_stream; }
    foley_status order_after(void* caller_stream) {   // legacy-stream callers: see Engine::order_after
        if (caller_stream) return FOLEY_OK;
        FOLEY_CUDA_OK(cudaEventRecord(ev_null, own_stream));
        FOLEY_CUDA_OK(cudaStreamWaitEvent(cudaStreamLegacy, ev_null, 0));
        return FOLEY_OK;
    }

  private:
    bool name_is_used(const std::string& n) const;
    foley_status to_bf16(const std::string& name, bf16** out, std::vector<int64_t>* shape);
    foley_status take_vec(const std::string& name, bf16** out, int64_t n_expected);
    foley_status take_linear(const std::string& name, LinearW* out, int n_expected, int k_expected);
    foley_status take_rows(const std::string& wname, const std::string& bname, int row0, int rows, int k, LinearW* out);
    foley_status take_qkv(const std::string& q, const std::string& k, const std::string& v, LinearW* out);
    foley_status ensure_rows(long long rows);
    foley_status gemm(cudaStream_t st, const bf16* A, long long rows, const LinearW& W, bf16* out, int act, int f16 = 0);
    foley_status take_linear_f16(const std::string& name, LinearW* out, int n_expected, int k_expected);
    foley_status take_rows_f16(const std::string& wname, const std::string& bname, int n, int k, LinearW* out);
    foley_status sync_ln(cudaStream_t st, SyncLnArgs a);
    foley_status add_ln(cudaStream_t st, EncLnArgs a);
    foley_status attention(cudaStream_t st, const EncAttnArgs& a, bool small);
    void free_all();
};

static inline unsigned enc_blocks(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

void Encoder::free_all() {
    for (auto& kv : raw)
        if (kv.second.dev) cudaFree(kv.second.dev);
    raw.clear();
    for (void* p : packed) cudaFree(p);
    packed.clear();
    for (bf16** p : {&x, &h, &qkv, &att, &y, &mlp})
        if (*p) { cudaFree(*p); *p = nullptr; }
    for (int** p : {&ids_dev, &pos_dev, &mask_dev})
        if (*p) { cudaFree(*p); *p = nullptr; }
    for (float** p : {&xf, &xn})
        if (*p) { cudaFree(*p); *p = nullptr; }
    cap_rows = 0;
}

Encoder::~Encoder() {
    cudaSetDevice(device);
    free_all();
    if (own_stream) cudaStreamDestroy(own_stream);
    if (ev_null) cudaEventDestroy(ev_null);
}

foley_status Encoder::create(const foley_encoder_config* c, int dev) {
    cfg = *c;
    device = dev;
    C = cfg.hidden_size; H = cfg.num_heads; NL = cfg.num_layers; F = cfg.intermediate_size;
    if (cfg.kind != FOLEY_ENC_SIGLIP_VISION && cfg.kind != FOLEY_ENC_CLAP_TEXT && cfg.kind != FOLEY_ENC_SYNCHFORMER)
        return fail(FOLEY_ERR_INVALID, "encoder kind");
    if (C != 768) return fail(FOLEY_ERR_UNSUPPORTED, "encoder hidden_size must be 768 (SigLIP2-base / CLAP text)");
    if (H <= 0 || C / H != 64 || C % H != 0) return fail(FOLEY_ERR_UNSUPPORTED, "encoder head_dim must be 64");
    if (NL < 1 || F < 64 || F % 64 != 0) return fail(FOLEY_ERR_INVALID, "encoder depth / intermediate size");
    if (cfg.kind == FOLEY_ENC_SIGLIP_VISION) {
        if (cfg.patch_size < 8 || cfg.patch_size % 8 != 0 || cfg.image_size % cfg.patch_size != 0 ||
            (3 * cfg.patch_size * cfg.patch_size) % 64 != 0)
            return fail(FOLEY_ERR_UNSUPPORTED, "SigLIP: patch size must be a multiple of 8 dividing the image size");
    } else if (cfg.kind == FOLEY_ENC_SYNCHFORMER) {
        if (cfg.patch_size != 16 || cfg.image_size != 224)
            return fail(FOLEY_ERR_UNSUPPORTED, "Synchformer: MotionFormer divided_224_16x4 only (224 px, 16 x 16 x 2 tubelets)");
    } else {
        if (cfg.vocab_size < 1 || cfg.max_positions < 4) return fail(FOLEY_ERR_INVALID, "CLAP: vocabulary / positions");
    }
    FOLEY_CUDA_OK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    FOLEY_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(FOLEY_ERR_UNSUPPORTED, "foley_b200 requires an sm_100 (Blackwell B200) device");
    num_sms = prop.multiProcessorCount;
    FOLEY_CUDA_OK(cudaStreamCreate(&own_stream));
    FOLEY_CUDA_OK(cudaEventCreateWithFlags(&ev_null, cudaEventDisableTiming));
    std::string err;
    if (!gemm_init_attributes(&err)) return fail(FOLEY_ERR_CUDA, err);
    FOLEY_CUDA_OK(cudaFuncSetAttribute(enc_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EA_SMEM));
    FOLEY_CUDA_OK(cudaFuncSetAttribute(enc_small_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    FOLEY_CUDA_OK(attention_tc64_init());
    FOLEY_CUDA_OK(cudaFuncSetAttribute(enc_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EA_SMEM));
    FOLEY_CUDA_OK(cudaFuncSetAttribute(enc_small_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    if (const char* e = getenv("FOLEY_ENC_ATT_TC")) att_tc = atoi(e);
    layers.resize(NL);
    if (cfg.kind == FOLEY_ENC_SYNCHFORMER) sblocks.resize(NL);
    return FOLEY_OK;
}

// SigLIP checkpoints carry the text tower too, CLAP ones the pooler / projection: neither is on this path.
static const char* kSyncPrefix = "vfeat_extractor.";
bool Encoder::name_is_used(const std::string& n) const {
    if (cfg.kind == FOLEY_ENC_SIGLIP_VISION) return n.rfind("vision_model.", 0) == 0;
    if (cfg.kind == FOLEY_ENC_SYNCHFORMER) {       // Synchformer state dict: the MotionFormer lives under vfeat_extractor.
        const std::string m = n.rfind(kSyncPrefix, 0) == 0 ? n.substr(strlen(kSyncPrefix)) : n;
        if (m.rfind("patch_embed.", 0) == 0) return false;           // 2-D patch embedding: unused by the video model
        for (const char* p : {"cls_token", "pos_embed", "temp_embed", "patch_embed_3d.", "blocks.", "norm.", "spatial_attn_agg."})
            if (m.rfind(p, 0) == 0) return true;
        return false;
    }
    if (n.rfind("text_model.", 0) != 0) return false;
    if (n.find(".pooler.") != std::string::npos) return false;
    if (n.find("position_ids") != std::string::npos || n.find("token_type_ids") != std::string::npos) return false;
    return true;
}

foley_status Encoder::load_tensor(const char* name, const void* data, const int64_t* shape, int ndim, int dtype) {
    if (!name || !data || ndim < 0 || ndim > 5) return fail(FOLEY_ERR_INVALID, "encoder load_tensor: bad arguments");
    if (dtype != FOLEY_DT_BF16 && dtype != FOLEY_DT_F32 && dtype != FOLEY_DT_F16)
        return fail(FOLEY_ERR_INVALID, "encoder load_tensor: dtype must be bf16, f32 or f16");
    std::string n(name);
    if (!name_is_used(n)) return FOLEY_OK;
    if (cfg.kind == FOLEY_ENC_SYNCHFORMER && n.rfind(kSyncPrefix, 0) == 0) n = n.substr(strlen(kSyncPrefix));
    if (finalized) return fail(FOLEY_ERR_STATE, "encoder weights cannot be reloaded after finalize (create a new encoder)");
    FOLEY_CUDA_OK(cudaSetDevice(device));
    RawTensor rt;
    rt.dtype = dtype;
    rt.numel = 1;
    for (int i = 0; i < ndim; ++i) { rt.shape.push_back(shape[i]); rt.numel *= shape[i]; }
    const size_t bytes = static_cast<size_t>(rt.numel) * (dtype == FOLEY_DT_F32 ? 4 : 2);
    FOLEY_CUDA_OK(cudaMalloc(&rt.dev, std::max<size_t>(bytes, 16)));
    FOLEY_CUDA_OK(cudaMemcpy(rt.dev, data, bytes, cudaMemcpyDefault));
    auto it = raw.find(n);
    if (it != raw.end() && it->second.dev) cudaFree(it->second.dev);
    raw[n] = rt;
    return FOLEY_OK;
}

foley_status Encoder::load_safetensors(const char* path, const char* prefix, int64_t* n_loaded) {
    if (!path) return fail(FOLEY_ERR_INVALID, "encoder load_safetensors: null path");
    StFile f;
    std::string err;
    if (!f.open_file(path, &err)) return fail(FOLEY_ERR_INVALID, err);
    const std::string pre(prefix ? prefix : "");
    int64_t count = 0;
    for (const StEntry& e : f.entries) {
        const std::string name = pre + e.name;
        if (!name_is_used(name)) continue;
        const int dt = st_dtype_to_foley(e.dtype);
        if (dt < 0 || dt > FOLEY_DT_F16) return fail(FOLEY_ERR_UNSUPPORTED, "encoder load_safetensors: dtype " + e.dtype + " of " + e.name);
        if (e.shape.size() > 5) return fail(FOLEY_ERR_INVALID, "encoder load_safetensors: rank > 5: " + e.name);
        ST_OK(load_tensor(name.c_str(), f.data + e.begin, e.shape.data(), static_cast<int>(e.shape.size()), dt));
        ++count;
    }
    if (n_loaded) *n_loaded = count;
    return FOLEY_OK;
}

foley_status Encoder::to_bf16(const std::string& name, bf16** out, std::vector<int64_t>* shape) {
    auto it = raw.find(name);
    if (it == raw.end()) return fail(FOLEY_ERR_MISSING, "missing tensor: " + name);
    RawTensor& rt = it->second;
    bf16* d = nullptr;
    FOLEY_CUDA_OK(cudaMalloc(&d, std::max<size_t>(rt.numel * 2, 16)));
    packed.push_back(d);
    convert_to_bf16_kernel<<<enc_blocks(rt.numel, 256), 256>>>(rt.dev, rt.dtype, rt.numel, d, 0);
    FOLEY_CUDA_OK(cudaGetLastError());
    *out = d;
    if (shape) *shape = rt.shape;
    return FOLEY_OK;
}

foley_status Encoder::take_vec(const std::string& name, bf16** out, int64_t n_expected) {
    std::vector<int64_t> shape;
    ST_OK(to_bf16(name, out, &shape));
    int64_t n = 1;
    for (int64_t s : shape) n *= s;
    if (n != n_expected) return fail(FOLEY_ERR_INVALID, "unexpected size for " + name);
    return FOLEY_OK;
}

foley_status Encoder::take_linear(const std::string& name, LinearW* out, int n_expected, int k_expected) {
    std::vector<int64_t> shape;
    ST_OK(to_bf16(name + ".weight", &out->w, &shape));
    if (shape.size() < 2) return fail(FOLEY_ERR_INVALID, "weight must be >= 2-D: " + name);
    int64_t k = 1;
    for (size_t i = 1; i < shape.size(); ++i) k *= shape[i];      // Conv2d [N, 3, P, P] is already the K-major GEMM matrix
    if (shape[0] != n_expected || k != k_expected) return fail(FOLEY_ERR_INVALID, "unexpected shape for " + name + ".weight");
    out->n = n_expected; out->k = k_expected; out->taps = 1;
    return take_vec(name + ".bias", &out->b, n_expected);
}

// Rows [row0, row0 + rows) of a packed projection (nn.MultiheadAttention.in_proj_weight / in_proj_bias).
foley_status Encoder::take_rows(const std::string& wname, const std::string& bname, int row0, int rows, int k, LinearW* out) {
    bf16 *w = nullptr, *b = nullptr;
    std::vector<int64_t> shape;
    ST_OK(to_bf16(wname, &w, &shape));
    if (shape.size() != 2 || shape[1] != k || shape[0] < row0 + rows) return fail(FOLEY_ERR_INVALID, "unexpected shape for " + wname);
    ST_OK(to_bf16(bname, &b, &shape));
    out->w = w + static_cast<long long>(row0) * k;
    out->b = b + row0;
    out->n = rows; out->k = k; out->taps = 1;
    return FOLEY_OK;
}

// q / k / v Linear layers stacked into one [3C, C] matrix: one GEMM, and the attention kernel reads the three column
// blocks of its output in place.
foley_status Encoder::take_qkv(const std::string& qn, const std::string& kn, const std::string& vn, LinearW* out) {
    bf16 *w = nullptr, *b = nullptr;
    FOLEY_CUDA_OK(cudaMalloc(&w, static_cast<size_t>(3) * C * C * 2));
    packed.push_back(w);
    FOLEY_CUDA_OK(cudaMalloc(&b, static_cast<size_t>(3) * C * 2));
    packed.push_back(b);
    const std::string* names[3] = {&qn, &kn, &vn};
    for (int i = 0; i < 3; ++i) {
        LinearW part;
        ST_OK(take_linear(*names[i], &part, C, C));
        FOLEY_CUDA_OK(cudaMemcpy(w + static_cast<long long>(i) * C * C, part.w, static_cast<size_t>(C) * C * 2, cudaMemcpyDeviceToDevice));
        FOLEY_CUDA_OK(cudaMemcpy(b + i * C, part.b, static_cast<size_t>(C) * 2, cudaMemcpyDeviceToDevice));
    }
    out->w = w; out->b = b; out->n = 3 * C; out->k = C; out->taps = 1;
    return FOLEY_OK;
}

// Linear of a module that runs under fp16 autocast with bf16 parameters: weight and bias reach the GEMM as fp16(bf16(w)).
foley_status Encoder::take_linear_f16(const std::string& name, LinearW* out, int n_expected, int k_expected) {
    ST_OK(take_linear(name, out, n_expected, k_expected));
    const long long nw = static_cast<long long>(n_expected) * k_expected;
    bf16_to_f16_kernel<<<enc_blocks(nw, 256), 256>>>(out->w, nw, reinterpret_cast<__half*>(out->w));     // in place: same element size
    bf16_to_f16_kernel<<<enc_blocks(n_expected, 256), 256>>>(out->b, n_expected, reinterpret_cast<__half*>(out->b));
    FOLEY_CUDA_OK(cudaGetLastError());
    return FOLEY_OK;
}
foley_status Encoder::take_rows_f16(const std::string& wname, const std::string& bname, int n, int k, LinearW* out) {
    ST_OK(take_rows(wname, bname, 0, n, k, out));
    const long long nw = static_cast<long long>(n) * k;
    bf16_to_f16_kernel<<<enc_blocks(nw, 256), 256>>>(out->w, nw, reinterpret_cast<__half*>(out->w));
    bf16_to_f16_kernel<<<enc_blocks(n, 256), 256>>>(out->b, n, reinterpret_cast<__half*>(out->b));
    FOLEY_CUDA_OK(cudaGetLastError());
    return FOLEY_OK;
}

foley_status Encoder::finalize() {
    if (finalized) return FOLEY_OK;
    FOLEY_CUDA_OK(cudaSetDevice(device));
    if (cfg.kind == FOLEY_ENC_SYNCHFORMER) {
        const int P = cfg.patch_size, G = cfg.image_size / P, NSP = G * G;          // 196 spatial locations, 8 temporal positions
        ST_OK(take_linear_f16("patch_embed_3d.proj", &s_patch, C, 3 * 2 * P * P));
        bf16 *cls = nullptr, *pos = nullptr, *temp = nullptr, *acls = nullptr;
        ST_OK(take_vec("cls_token", &cls, C));
        ST_OK(take_vec("pos_embed", &pos, static_cast<int64_t>(1 + NSP) * C));
        ST_OK(take_vec("temp_embed", &temp, 8LL * C));
        const long long ntab = (1LL + 8LL * NSP) * C;
        FOLEY_CUDA_OK(cudaMalloc(&s_tok_tab, ntab * sizeof(float)));
        packed.push_back(s_tok_tab);
        sync_token_table_kernel<<<enc_blocks(ntab, 256), 256>>>(cls, pos, temp, NSP, 8, C, s_tok_tab);
        FOLEY_CUDA_OK(cudaGetLastError());
        for (int i = 0; i < NL; ++i) {
            const std::string l = "blocks." + std::to_string(i) + ".";
            SyncBlockW& w = sblocks[i];
            ST_OK(take_vec(l + "norm1.weight", &w.n1w, C)); ST_OK(take_vec(l + "norm1.bias", &w.n1b, C));
            ST_OK(take_vec(l + "norm2.weight", &w.n2w, C)); ST_OK(take_vec(l + "norm2.bias", &w.n2b, C));
            ST_OK(take_vec(l + "norm3.weight", &w.n3w, C)); ST_OK(take_vec(l + "norm3.bias", &w.n3b, C));
            ST_OK(take_linear_f16(l + "timeattn.qkv", &w.qkv_t, 3 * C, C));
            ST_OK(take_linear_f16(l + "timeattn.proj", &w.proj_t, C, C));
            ST_OK(take_linear_f16(l + "attn.qkv", &w.qkv_s, 3 * C, C));
            ST_OK(take_linear_f16(l + "attn.proj", &w.proj_s, C, C));
            ST_OK(take_linear_f16(l + "mlp.fc1", &w.fc1, F, C));
            ST_OK(take_linear_f16(l + "mlp.fc2", &w.fc2, C, F));
        }
        ST_OK(take_vec("norm.weight", &s_norm_w, C)); ST_OK(take_vec("norm.bias", &s_norm_b, C));
        const std::string ag = "spatial_attn_agg.";
        ST_OK(take_vec(ag + "cls_token", &acls, C));
        FOLEY_CUDA_OK(cudaMalloc(&s_agg_cls, C * sizeof(float)));
        packed.push_back(s_agg_cls);
        bf16_to_f32_kernel<<<enc_blocks(C, 256), 256>>>(acls, C, s_agg_cls);
        FOLEY_CUDA_OK(cudaGetLastError());
        ST_OK(take_rows_f16(ag + "self_attn.in_proj_weight", ag + "self_attn.in_proj_bias", 3 * C, C, &s_in_proj));
        ST_OK(take_linear_f16(ag + "self_attn.out_proj", &s_out_proj, C, C));
        ST_OK(take_linear_f16(ag + "linear1", &s_lin1, F, C));
        ST_OK(take_linear_f16(ag + "linear2", &s_lin2, C, F));
        ST_OK(take_vec(ag + "norm1.weight", &s_an1w, C)); ST_OK(take_vec(ag + "norm1.bias", &s_an1b, C));
        ST_OK(take_vec(ag + "norm2.weight", &s_an2w, C)); ST_OK(take_vec(ag + "norm2.bias", &s_an2b, C));
    } else if (cfg.kind == FOLEY_ENC_SIGLIP_VISION) {
        const std::string vm = "vision_model.";
        const int P = cfg.patch_size, G = cfg.image_size / P;
        ST_OK(take_linear(vm + "embeddings.patch_embedding", &patch, C, 3 * P * P));
        ST_OK(take_vec(vm + "embeddings.position_embedding.weight", &pos_emb, static_cast<int64_t>(G) * G * C));
        for (int i = 0; i < NL; ++i) {
            const std::string l = vm + "encoder.layers." + std::to_string(i) + ".";
            EncLayerW& w = layers[i];
            ST_OK(take_vec(l + "layer_norm1.weight", &w.ln1_w, C));
            ST_OK(take_vec(l + "layer_norm1.bias", &w.ln1_b, C));
            ST_OK(take_qkv(l + "self_attn.q_proj", l + "self_attn.k_proj", l + "self_attn.v_proj", &w.qkv));
            ST_OK(take_linear(l + "self_attn.out_proj", &w.o, C, C));
            ST_OK(take_vec(l + "layer_norm2.weight", &w.ln2_w, C));
            ST_OK(take_vec(l + "layer_norm2.bias", &w.ln2_b, C));
            ST_OK(take_linear(l + "mlp.fc1", &w.fc1, F, C));
            ST_OK(take_linear(l + "mlp.fc2", &w.fc2, C, F));
        }
        ST_OK(take_vec(vm + "post_layernorm.weight", &post_ln_w, C));
        ST_OK(take_vec(vm + "post_layernorm.bias", &post_ln_b, C));
        ST_OK(take_vec(vm + "head.probe", &probe, C));
        ST_OK(take_rows(vm + "head.attention.in_proj_weight", vm + "head.attention.in_proj_bias", 0, C, C, &head_q));
        ST_OK(take_rows(vm + "head.attention.in_proj_weight", vm + "head.attention.in_proj_bias", C, 2 * C, C, &head_kv));
        ST_OK(take_linear(vm + "head.attention.out_proj", &head_o, C, C));
        ST_OK(take_vec(vm + "head.layernorm.weight", &head_ln_w, C));
        ST_OK(take_vec(vm + "head.layernorm.bias", &head_ln_b, C));
        ST_OK(take_linear(vm + "head.mlp.fc1", &head_fc1, F, C));
        ST_OK(take_linear(vm + "head.mlp.fc2", &head_fc2, C, F));
        // the probe is a parameter: its query projection is the same for every frame (SiglipMultiheadAttentionPoolingHead)
        FOLEY_CUDA_OK(cudaMalloc(&probe_q, static_cast<size_t>(C) * 2));
        packed.push_back(probe_q);
        ST_OK(gemm(own_stream, probe, 1, head_q, probe_q, ACT_NONE));
    } else {
        const std::string tm = "text_model.";
        ST_OK(take_vec(tm + "embeddings.word_embeddings.weight", &word, static_cast<int64_t>(cfg.vocab_size) * C));
        ST_OK(take_vec(tm + "embeddings.position_embeddings.weight", &pos_tab, static_cast<int64_t>(cfg.max_positions) * C));
        {
            std::vector<int64_t> shape;
            ST_OK(to_bf16(tm + "embeddings.token_type_embeddings.weight", &type_tab, &shape));
            if (shape.size() != 2 || shape[1] != C) return fail(FOLEY_ERR_INVALID, "unexpected shape for token_type_embeddings");
        }
        ST_OK(take_vec(tm + "embeddings.LayerNorm.weight", &emb_ln_w, C));
        ST_OK(take_vec(tm + "embeddings.LayerNorm.bias", &emb_ln_b, C));
        for (int i = 0; i < NL; ++i) {
            const std::string l = tm + "encoder.layer." + std::to_string(i) + ".";
            EncLayerW& w = layers[i];
            ST_OK(take_qkv(l + "attention.self.query", l + "attention.self.key", l + "attention.self.value", &w.qkv));
            ST_OK(take_linear(l + "attention.output.dense", &w.o, C, C));
            ST_OK(take_vec(l + "attention.output.LayerNorm.weight", &w.ln1_w, C));
            ST_OK(take_vec(l + "attention.output.LayerNorm.bias", &w.ln1_b, C));
            ST_OK(take_linear(l + "intermediate.dense", &w.fc1, F, C));
            ST_OK(take_linear(l + "output.dense", &w.fc2, C, F));
            ST_OK(take_vec(l + "output.LayerNorm.weight", &w.ln2_w, C));
            ST_OK(take_vec(l + "output.LayerNorm.bias", &w.ln2_b, C));
        }
    }
    FOLEY_CUDA_OK(cudaDeviceSynchronize());
    for (auto& kv : raw)
        if (kv.second.dev) cudaFree(kv.second.dev);
    raw.clear();
    finalized = true;
    return FOLEY_OK;
}

foley_status Encoder::ensure_rows(long long rows) {
    if (rows <= cap_rows) return FOLEY_OK;
    FOLEY_CUDA_OK(cudaDeviceSynchronize());
    for (bf16** p : {&x, &h, &qkv, &att, &y, &mlp})
        if (*p) { cudaFree(*p); *p = nullptr; }
    for (int** p : {&ids_dev, &pos_dev, &mask_dev})
        if (*p) { cudaFree(*p); *p = nullptr; }
    for (float** p : {&xf, &xn})
        if (*p) { cudaFree(*p); *p = nullptr; }
    cap_rows = 0;
    const size_t r = static_cast<size_t>(rows);
    if (cfg.kind == FOLEY_ENC_SYNCHFORMER) {
        FOLEY_CUDA_OK(cudaMalloc(&xf, r * C * sizeof(float)));
        FOLEY_CUDA_OK(cudaMalloc(&xn, r * C * sizeof(float)));
    }
    FOLEY_CUDA_OK(cudaMalloc(&x, r * C * 2));
    FOLEY_CUDA_OK(cudaMalloc(&h, r * C * 2));
    FOLEY_CUDA_OK(cudaMalloc(&qkv, r * 3 * C * 2));
    FOLEY_CUDA_OK(cudaMalloc(&att, r * C * 2));
    FOLEY_CUDA_OK(cudaMalloc(&y, r * C * 2));
    FOLEY_CUDA_OK(cudaMalloc(&mlp, r * F * 2));
    FOLEY_CUDA_OK(cudaMalloc(&ids_dev, r * sizeof(int)));
    FOLEY_CUDA_OK(cudaMalloc(&pos_dev, r * sizeof(int)));
    FOLEY_CUDA_OK(cudaMalloc(&mask_dev, r * sizeof(int)));
    cap_rows = rows;
    return FOLEY_OK;
}

// out[rows, N] = act(A[rows, K] W^T + b), bf16.  Tile width: the widest that still fills the SMs; big grids (the 40960-row
// GEMMs of a 5 s clip) go to the persistent kernel (gemm_host.cuh).
foley_status Encoder::gemm(cudaStream_t st, const bf16* A, long long rows, const LinearW& W, bf16* out, int act, int f16) {
    GemmLaunch L;
    L.a.ptr = A; L.a.dtype = DT_BF16; L.a.k = W.k; L.a.rows = rows; L.a.batch = 1; L.a.ld = W.k; L.a.batch_stride = rows * W.k;
    L.w = W.w; L.n = W.n; L.taps = 1; L.tap_off0 = 0; L.tap_stride = 1; L.splits = 1;
    const long long mt = (rows + 127) / 128;
    int bn = 256;
    while (bn > 64 && mt * ((W.n + bn - 1) / bn) < num_sms) bn >>= 1;
    L.bn = bn;
    L.epi.mode = EPI_BF16; L.epi.act = act; L.epi.out = out; L.epi.ldo = W.n; L.epi.bias = W.b; L.f16 = f16;
    L.epi.out_batch_stride = rows * W.n;
    std::string err;
    if (!launch_gemm(L, st, &err)) return fail(FOLEY_ERR_CUDA, err);
    ++launches;
    return FOLEY_OK;
}

foley_status Encoder::add_ln(cudaStream_t st, EncLnArgs a) {
    FOLEY_CUDA_OK(launch_k(enc_add_ln_kernel<3>, dim3(enc_blocks(a.rows, 8)), dim3(256), 0, st, a));
    ++launches;
    return FOLEY_OK;
}

foley_status Encoder::sync_ln(cudaStream_t st, SyncLnArgs a) {
    a.eps = cfg.layer_norm_eps;
    FOLEY_CUDA_OK(launch_k(sync_add_ln_kernel<3>, dim3(enc_blocks(a.rows, 8)), dim3(256), 0, st, a));
    ++launches;
    return FOLEY_OK;
}

// One query per (sample, head): the block-per-unit kernel (the one-warp kernel walked 1569 keys serially: 263 us per call).
template <bool kHalf>
static foley_status launch_cls_attention(const EncAttnArgs& a, cudaStream_t st) {
    const size_t smem = (static_cast<size_t>((a.Sk + 3) & ~3) + ECA_WARPS * 64) * sizeof(float);
    if (smem > 48 * 1024) return fail(FOLEY_ERR_UNSUPPORTED, "class-token attention: too many keys");
    FOLEY_CUDA_OK(launch_k(enc_cls_attention_kernel<kHalf>, dim3(static_cast<unsigned>(a.B) * a.H), dim3(32 * ECA_WARPS), smem, st, a));
    return FOLEY_OK;
}

foley_status Encoder::attention(cudaStream_t st, const EncAttnArgs& a, bool small) {
    if (small && a.Sq == 1) {
        ST_OK(launch_cls_attention<false>(a, st));
    } else if (small) {
        const size_t smem = static_cast<size_t>(ESA_WARPS) * a.Sk * sizeof(float);
        if (smem > 64 * 1024) return fail(FOLEY_ERR_UNSUPPORTED, "small attention: too many keys");
        const long long units = static_cast<long long>(a.B) * a.H * a.Sq;
        FOLEY_CUDA_OK(launch_k(enc_small_attention_kernel<false>, dim3(enc_blocks(units, ESA_WARPS)), dim3(32 * ESA_WARPS), smem, st, a));
    } else if (att_tc) {
        AttTc64Args t;
        t.o = a.o; t.o_batch_stride = a.o_batch_stride; t.o_row_stride = a.o_row_stride; t.H = a.H; t.Sq = a.Sq; t.Sk = a.Sk;
        t.scale_log2 = a.scale * 1.4426950408889634f;
        std::string err;
        if (!launch_attention_tc64(a.q, a.k, a.v, a.q_row_stride, a.q_batch_stride, a.kv_row_stride, a.kv_batch_stride, a.B, t, st, &err,
                                   att_tc == 1 ? 2 : 1))
            return fail(FOLEY_ERR_CUDA, err);
    } else {
        dim3 grid((a.Sq + EA_BM - 1) / EA_BM, a.H, a.B);
        FOLEY_CUDA_OK(launch_k(enc_attention_kernel<false>, grid, dim3(32 * EA_NW), EA_SMEM, st, a));
    }
    ++launches;
    return FOLEY_OK;
}

// pixels: fp32 [n_frames, 3, IMG, IMG] (the output of foley_preprocess_frames) -> out: bf16 [n_frames, C] pooler_output.
foley_status Encoder::siglip_encode(const float* pixels, int n_frames, void* out, cudaStream_t st) {
    if (cfg.kind != FOLEY_ENC_SIGLIP_VISION) return fail(FOLEY_ERR_STATE, "not a SigLIP vision encoder");
    if (!finalized) return fail(FOLEY_ERR_STATE, "encoder used before finalize");
    if (!pixels || !out || n_frames < 1) return fail(FOLEY_ERR_INVALID, "siglip_encode: bad argument");
    FOLEY_CUDA_OK(cudaSetDevice(device));
    const int P = cfg.patch_size, IMG = cfg.image_size, G = IMG / P, NP = G * G;
    const int per_pass = cfg.max_frames_per_pass > 0 ? cfg.max_frames_per_pass : 48;
    ST_OK(ensure_rows(static_cast<long long>(std::min(per_pass, n_frames)) * NP));
    const int nl = layers_run >= 0 ? std::min(layers_run, NL) : NL;
    for (int f0 = 0; f0 < n_frames; f0 += per_pass) {
        const int Tc = std::min(per_pass, n_frames - f0);
        const long long rows = static_cast<long long>(Tc) * NP;
        const float* px = pixels + static_cast<long long>(f0) * 3 * IMG * IMG;
        // ---- embeddings: Conv2d(k = stride = P) as im2col + GEMM, + position embeddings; LayerNorm of layer 0
        const long long n8 = static_cast<long long>(Tc) * 3 * IMG * (IMG / 8);
        FOLEY_CUDA_OK(launch_k(enc_patchify_kernel, dim3(enc_blocks(n8, 256)), dim3(256), 0, st, px, Tc, IMG, P, att));
        ++launches;
        ST_OK(gemm(st, att, rows, patch, y, ACT_NONE));
        {
            EncLnArgs a;
            a.y = y; a.res = pos_emb; a.res_mod = NP; a.x_out = x; a.rows = rows; a.eps = cfg.layer_norm_eps; a.h_out = h;
            a.ln_w = nl > 0 ? layers[0].ln1_w : post_ln_w; a.ln_b = nl > 0 ? layers[0].ln1_b : post_ln_b;
            ST_OK(add_ln(st, a));
        }
        for (int i = 0; i < nl; ++i) {
            const EncLayerW& w = layers[i];
            ST_OK(gemm(st, h, rows, w.qkv, qkv, ACT_NONE));
            EncAttnArgs aa;
            aa.q = qkv; aa.k = qkv + C; aa.v = qkv + 2 * C; aa.o = att;
            aa.B = Tc; aa.H = H; aa.Sq = NP; aa.Sk = NP;
            aa.q_row_stride = aa.kv_row_stride = 3LL * C; aa.q_batch_stride = aa.kv_batch_stride = 3LL * C * NP;
            aa.o_row_stride = C; aa.o_batch_stride = static_cast<long long>(C) * NP;
            aa.scale = 0.125f;
            ST_OK(attention(st, aa, false));
            ST_OK(gemm(st, att, rows, w.o, y, ACT_NONE));
            EncLnArgs a1;
            a1.y = y; a1.res = x; a1.x_out = x; a1.ln_w = w.ln2_w; a1.ln_b = w.ln2_b; a1.h_out = h; a1.rows = rows; a1.eps = cfg.layer_norm_eps;
            ST_OK(add_ln(st, a1));
            ST_OK(gemm(st, h, rows, w.fc1, mlp, ACT_GELU_TANH));
            ST_OK(gemm(st, mlp, rows, w.fc2, y, ACT_NONE));
            EncLnArgs a2 = a1;
            const bool last = i + 1 == nl;
            a2.ln_w = last ? post_ln_w : layers[i + 1].ln1_w;
            a2.ln_b = last ? post_ln_b : layers[i + 1].ln1_b;
            ST_OK(add_ln(st, a2));
        }
        // ---- attention-pooling head (HF SiglipMultiheadAttentionPoolingHead; nn.MultiheadAttention with need_weights:
        // bmm scores and softmax both rounded to bf16)
        ST_OK(gemm(st, h, rows, head_kv, qkv, ACT_NONE));        // [rows, 2C]: k | v
        EncAttnArgs pa;
        pa.q = probe_q; pa.k = qkv; pa.v = qkv + C; pa.o = att;
        pa.B = Tc; pa.H = H; pa.Sq = 1; pa.Sk = NP;
        pa.q_row_stride = C; pa.q_batch_stride = 0;
        pa.kv_row_stride = 2LL * C; pa.kv_batch_stride = 2LL * C * NP;
        pa.o_row_stride = C; pa.o_batch_stride = C;
        pa.scale = 0.125f; pa.round_scores = 1;
        ST_OK(attention(st, pa, true));
        ST_OK(gemm(st, att, Tc, head_o, y, ACT_NONE));           // hidden_state = attention(probe, x, x)[0]
        bf16* resid = qkv;                                        // the k | v projection is dead by now: rows [0, Tc) hold the residual
        EncLnArgs hl;
        hl.y = y; hl.x_out = resid; hl.ln_w = head_ln_w; hl.ln_b = head_ln_b; hl.h_out = h; hl.rows = Tc; hl.eps = cfg.layer_norm_eps;
        ST_OK(add_ln(st, hl));
        ST_OK(gemm(st, h, Tc, head_fc1, mlp, ACT_GELU_TANH));
        ST_OK(gemm(st, mlp, Tc, head_fc2, att, ACT_NONE));
        EncLnArgs fin;
        fin.y = att; fin.res = resid; fin.x_out = static_cast<bf16*>(out) + static_cast<long long>(f0) * C; fin.rows = Tc;
        ST_OK(add_ln(st, fin));
    }
    return FOLEY_OK;
}

// ids / mask: HOST int32 [B, T] (tokenizer output: input_ids, attention_mask) -> out: bf16 [B, T, C] last_hidden_state.
foley_status Encoder::clap_encode(const int32_t* ids, const int32_t* mask, int B, int T, void* out, cudaStream_t st) {
    if (cfg.kind != FOLEY_ENC_CLAP_TEXT) return fail(FOLEY_ERR_STATE, "not a CLAP text encoder");
    if (!finalized) return fail(FOLEY_ERR_STATE, "encoder used before finalize");
    if (!ids || !out || B < 1 || T < 1) return fail(FOLEY_ERR_INVALID, "clap_encode: bad argument");
    FOLEY_CUDA_OK(cudaSetDevice(device));
    const long long rows = static_cast<long long>(B) * T;
    // RoBERTa position ids (ClapTextEmbeddings.create_position_ids_from_input_ids): cumulative count of non-pad tokens
    std::vector<int> pos(rows), msk(rows);
    for (int b = 0; b < B; ++b) {
        int run = 0;
        for (int t = 0; t < T; ++t) {
            const int id = ids[b * T + t];
            if (id < 0 || id >= cfg.vocab_size) return fail(FOLEY_ERR_INVALID, "clap_encode: token id outside the vocabulary");
            const int keep = id != cfg.pad_token_id;
            run += keep;
            const int p = keep ? run + cfg.pad_token_id : cfg.pad_token_id;
            if (p >= cfg.max_positions) return fail(FOLEY_ERR_INVALID, "clap_encode: sequence longer than max_position_embeddings");
            pos[b * T + t] = p;
            msk[b * T + t] = mask ? (mask[b * T + t] != 0) : 1;
        }
    }
    ST_OK(ensure_rows(rows));
    FOLEY_CUDA_OK(cudaMemcpyAsync(ids_dev, ids, rows * sizeof(int), cudaMemcpyHostToDevice, st));
    FOLEY_CUDA_OK(cudaMemcpyAsync(pos_dev, pos.data(), rows * sizeof(int), cudaMemcpyHostToDevice, st));
    FOLEY_CUDA_OK(cudaMemcpyAsync(mask_dev, msk.data(), rows * sizeof(int), cudaMemcpyHostToDevice, st));
    FOLEY_CUDA_OK(cudaStreamSynchronize(st));   // pos / msk are stack-owned host buffers
    const int nl = layers_run >= 0 ? std::min(layers_run, NL) : NL;
    bf16* xo = nl == 0 ? static_cast<bf16*>(out) : x;
    FOLEY_CUDA_OK(launch_k(enc_text_embed_kernel<3>, dim3(enc_blocks(rows, 8)), dim3(256), 0, st, ids_dev, pos_dev, word, type_tab,
                           pos_tab, emb_ln_w, emb_ln_b, cfg.layer_norm_eps, rows, xo));
    ++launches;
    for (int i = 0; i < nl; ++i) {
        const EncLayerW& w = layers[i];
        ST_OK(gemm(st, x, rows, w.qkv, qkv, ACT_NONE));
        EncAttnArgs aa;
        aa.q = qkv; aa.k = qkv + C; aa.v = qkv + 2 * C; aa.o = att;
        aa.B = B; aa.H = H; aa.Sq = T; aa.Sk = T;
        aa.q_row_stride = aa.kv_row_stride = 3LL * C; aa.q_batch_stride = aa.kv_batch_stride = 3LL * C * T;
        aa.o_row_stride = C; aa.o_batch_stride = static_cast<long long>(C) * T;
        aa.scale = 0.125f; aa.key_mask = mask_dev;
        ST_OK(attention(st, aa, true));
        ST_OK(gemm(st, att, rows, w.o, y, ACT_NONE));
        EncLnArgs a1;                       // post-LN: x = LayerNorm(dense(attn) + x)
        a1.y = y; a1.res = x; a1.ln_w = w.ln1_w; a1.ln_b = w.ln1_b; a1.h_out = x; a1.rows = rows; a1.eps = cfg.layer_norm_eps;
        ST_OK(add_ln(st, a1));
        ST_OK(gemm(st, x, rows, w.fc1, mlp, ACT_NONE));
        FOLEY_CUDA_OK(launch_k(enc_gelu_erf_kernel, dim3(enc_blocks(rows * F / 8, 256)), dim3(256), 0, st, mlp, rows * F / 8));
        ++launches;
        ST_OK(gemm(st, mlp, rows, w.fc2, y, ACT_NONE));
        EncLnArgs a2 = a1;
        a2.ln_w = w.ln2_w; a2.ln_b = w.ln2_b;
        a2.h_out = i + 1 == nl ? static_cast<bf16*>(out) : x;
        ST_OK(add_ln(st, a2));
    }
    return FOLEY_OK;
}

// encode_video_with_sync (feature_utils.py:81-106) + Synchformer.forward -> MotionFormer.forward (motionformer.py:178-213):
// frames fp32 [n_frames, 3, 224, 224] (25 fps, preprocessed) -> out fp32 [segments * 8, C], segments = (n_frames - 16) / 8 + 1
// windows of 16 frames every 8 frames; per window 8 x 14 x 14 tubelet tokens + a class token through 12 divided space-time
// blocks, the final norm, and one spatial aggregation layer (class token per temporal position).
foley_status Encoder::synchformer_encode(const float* frames, int n_frames, float* out, cudaStream_t st) {
    if (cfg.kind != FOLEY_ENC_SYNCHFORMER) return fail(FOLEY_ERR_STATE, "not a Synchformer encoder");
    if (!finalized) return fail(FOLEY_ERR_STATE, "encoder used before finalize");
    if (!frames || !out || n_frames < 16) return fail(FOLEY_ERR_INVALID, "synchformer_encode: needs at least 16 frames");
    FOLEY_CUDA_OK(cudaSetDevice(device));
    const int P = cfg.patch_size, IMG = cfg.image_size, G = IMG / P, NSP = G * G, NTOK = 1 + 8 * NSP, SEQ = 1 + NSP;
    const int S = (n_frames - 16) / 8 + 1;
    const int per_pass = cfg.max_frames_per_pass > 0 ? cfg.max_frames_per_pass : 16;     // segments per pass
    ST_OK(ensure_rows(static_cast<long long>(std::min(per_pass, S)) * 8 * SEQ));         // 8 * 197 > 1569 rows per segment
    const int nl = layers_run >= 0 ? std::min(layers_run, NL) : NL;
    __half* hh = reinterpret_cast<__half*>(h);
    __half* hy = reinterpret_cast<__half*>(y);
    __half* hqkv = reinterpret_cast<__half*>(qkv);
    __half* hatt = reinterpret_cast<__half*>(att);
    for (int s0 = 0; s0 < S; s0 += per_pass) {
        const int Sc = std::min(per_pass, S - s0);
        const long long rows = static_cast<long long>(Sc) * NTOK;
        const float* fr = frames + static_cast<long long>(s0) * 8 * 3 * IMG * IMG;
        // ---- tubelet embedding: Conv3d as im2col + GEMM (fp16), class token + position / temporal tables (fp32)
        const long long n8 = static_cast<long long>(Sc) * 8 * 2 * 3 * IMG * (IMG / 8);
        FOLEY_CUDA_OK(launch_k(sync_patchify_kernel, dim3(enc_blocks(n8, 256)), dim3(256), 0, st, fr, Sc, IMG, P, reinterpret_cast<__half*>(mlp)));
        ++launches;
        ST_OK(gemm(st, mlp, static_cast<long long>(Sc) * 8 * NSP, s_patch, y, ACT_NONE, 1));
        {
            SyncLnArgs a;
            a.patch = hy; a.tok_tab = s_tok_tab; a.tokens = NTOK; a.x_dst = xf; a.rows = rows; a.h_out = hh;
            a.ln_w = nl > 0 ? sblocks[0].n3w : s_norm_w; a.ln_b = nl > 0 ? sblocks[0].n3b : s_norm_b;
            if (nl == 0) { a.h_out = nullptr; a.n_out = xn; a.n_out_skip = NTOK; }
            ST_OK(sync_ln(st, a));
        }
        auto cls_attention = [&]() -> foley_status {       // the class token attends to every token of its segment
            EncAttnArgs ca;
            ca.q = qkv; ca.k = qkv + C; ca.v = qkv + 2 * C; ca.o = att;
            ca.B = Sc; ca.H = H; ca.Sq = 1; ca.Sk = NTOK;
            ca.q_row_stride = ca.kv_row_stride = 3LL * C; ca.q_batch_stride = ca.kv_batch_stride = 3LL * C * NTOK;
            ca.o_row_stride = C; ca.o_batch_stride = static_cast<long long>(C) * NTOK;
            ca.scale = 0.125f; ca.round_scores = 1;
            ST_OK(launch_cls_attention<true>(ca, st));
            ++launches;
            return FOLEY_OK;
        };
        for (int i = 0; i < nl; ++i) {
            const SyncBlockW& w = sblocks[i];
            // -- divided TIME attention (norm3 -> timeattn), vit_helper.py:155-158
            ST_OK(gemm(st, h, rows, w.qkv_t, qkv, ACT_NONE, 1));
            ST_OK(cls_attention());
            FOLEY_CUDA_OK(launch_k(sync_time_attention_kernel, dim3(enc_blocks(static_cast<long long>(Sc) * NSP * H, 4)), dim3(128), 0, st,
                                   static_cast<const __half*>(hqkv), hatt, Sc, NTOK, NSP, H, 0.125f));
            ++launches;
            ST_OK(gemm(st, att, rows, w.proj_t, y, ACT_NONE, 1));
            {
                SyncLnArgs a;
                a.x = xf; a.y = hy; a.x_dst = xf; a.ln_w = w.n1w; a.ln_b = w.n1b; a.h_out = hh; a.rows = rows;
                ST_OK(sync_ln(st, a));
            }
            // -- divided SPACE attention (norm1 -> attn), vit_helper.py:160-163
            ST_OK(gemm(st, h, rows, w.qkv_s, qkv, ACT_NONE, 1));
            ST_OK(cls_attention());
            {
                EncAttnArgs sa;       // frame b2 of segment b1: 196 queries, keys = [class token; the 196 tokens of the frame]
                sa.q = qkv + 3LL * C; sa.k = qkv + C; sa.v = qkv + 2 * C; sa.o = att + C;
                sa.B = Sc * 8; sa.batch2 = 8; sa.H = H; sa.Sq = NSP; sa.Sk = SEQ;
                sa.q_row_stride = sa.kv_row_stride = 3LL * C;
                sa.q_batch_stride = sa.kv_batch_stride = 3LL * C * NTOK; sa.q_batch_stride2 = 3LL * C * NSP;
                sa.kv_batch_stride2 = NSP;                                   // rows
                sa.o_row_stride = C; sa.o_batch_stride = static_cast<long long>(C) * NTOK; sa.o_batch_stride2 = static_cast<long long>(C) * NSP;
                sa.scale = 0.125f; sa.round_scores = 1;
                dim3 grid((sa.Sq + EA_BM - 1) / EA_BM, sa.H, sa.B);
                FOLEY_CUDA_OK(launch_k(enc_attention_kernel<true>, grid, dim3(32 * EA_NW), EA_SMEM, st, sa));
                ++launches;
            }
            ST_OK(gemm(st, att, rows, w.proj_s, y, ACT_NONE, 1));
            {
                SyncLnArgs a;
                a.x = xf; a.y = hy; a.x_dst = xf; a.ln_w = w.n2w; a.ln_b = w.n2b; a.h_out = hh; a.rows = rows;
                ST_OK(sync_ln(st, a));
            }
            // -- MLP (norm2 -> fc1 -> GELU -> fc2)
            ST_OK(gemm(st, h, rows, w.fc1, mlp, ACT_GELU_ERF, 1));
            ST_OK(gemm(st, mlp, rows, w.fc2, y, ACT_NONE, 1));
            {
                SyncLnArgs a;
                const bool last = i + 1 == nl;
                a.x = xf; a.y = hy; a.x_dst = xf; a.rows = rows;
                if (!last) { a.ln_w = sblocks[i + 1].n3w; a.ln_b = sblocks[i + 1].n3b; a.h_out = hh; }
                else { a.ln_w = s_norm_w; a.ln_b = s_norm_b; a.n_out = xn; a.n_out_skip = NTOK; }   // x[:, 1:] -> norm (motionformer.py:186-187)
                ST_OK(sync_ln(st, a));
            }
        }
        // ---- spatial aggregation layer (SpatialTransformerEncoderLayer, motionformer.py:330-355: nn.TransformerEncoderLayer,
        // norm_first, a class token per (segment, temporal position); only the class-token row leaves the layer)
        const long long rows_a = static_cast<long long>(Sc) * 8 * SEQ, nseq = static_cast<long long>(Sc) * 8;
        {
            SyncLnArgs a;
            a.seq_src = xn; a.seq_cls = s_agg_cls; a.seq_len = SEQ; a.ln_w = s_an1w; a.ln_b = s_an1b; a.h_out = hh; a.rows = rows_a;
            ST_OK(sync_ln(st, a));
        }
        ST_OK(gemm(st, h, rows_a, s_in_proj, qkv, ACT_NONE, 1));
        {
            EncAttnArgs ca;
            ca.q = qkv; ca.k = qkv + C; ca.v = qkv + 2 * C; ca.o = att;
            ca.B = static_cast<int>(nseq); ca.H = H; ca.Sq = 1; ca.Sk = SEQ;
            ca.q_row_stride = ca.kv_row_stride = 3LL * C; ca.q_batch_stride = ca.kv_batch_stride = 3LL * C * SEQ;
            ca.o_row_stride = C; ca.o_batch_stride = C;
            ca.scale = 0.125f;
            ST_OK(launch_cls_attention<true>(ca, st));
            ++launches;
        }
        ST_OK(gemm(st, att, nseq, s_out_proj, y, ACT_NONE, 1));
        {
            SyncLnArgs a;       // x1 = cls + sa(...): the class-token rows of the layer's residual stream
            a.seq_cls = s_agg_cls; a.seq_len = 0; a.y = hy; a.x_dst = xf; a.ln_w = s_an2w; a.ln_b = s_an2b; a.h_out = hh; a.rows = nseq;
            ST_OK(sync_ln(st, a));
        }
        ST_OK(gemm(st, h, nseq, s_lin1, mlp, ACT_GELU_ERF, 1));
        ST_OK(gemm(st, mlp, nseq, s_lin2, y, ACT_NONE, 1));
        {
            SyncLnArgs a;
            a.x = xf; a.y = hy; a.x_dst = out + static_cast<long long>(s0) * 8 * C; a.rows = nseq;
            ST_OK(sync_ln(st, a));
        }
    }
    return FOLEY_OK;
}

__global__ void enc_bf16_to_f32_kernel(const __nv_bfloat16* src, long long n, float* dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __bfloat162float(src[i]);
}

// Activation buffers of the last call, for per-stage parity tests: "x" (residual stream), "h" (last LayerNorm output),
// "qkv", "att", "y", "mlp".  dst: HOST fp32.
foley_status Encoder::debug_read(const char* what, float* dst, int64_t cap, int64_t* n_out) {
    if (!what || !dst) return fail(FOLEY_ERR_INVALID, "debug_read: null argument");
    const std::string w(what);
    const bf16* src = nullptr;
    long long width = C;
    if (w == "x") src = x; else if (w == "h") src = h; else if (w == "att") src = att; else if (w == "y") src = y;
    else if (w == "qkv") { src = qkv; width = 3LL * C; } else if (w == "mlp") { src = mlp; width = F; }
    else return fail(FOLEY_ERR_INVALID, "debug_read: unknown buffer " + w);
    if (!src) return fail(FOLEY_ERR_STATE, "debug_read: no call has been made yet");
    const long long n = std::min<long long>(cap, cap_rows * width);
    FOLEY_CUDA_OK(cudaSetDevice(device));
    float* tmp = nullptr;
    FOLEY_CUDA_OK(cudaMalloc(&tmp, std::max<size_t>(n * 4, 16)));
    FOLEY_CUDA_OK(cudaDeviceSynchronize());
    enc_bf16_to_f32_kernel<<<enc_blocks(n, 256), 256>>>(src, n, tmp);
    cudaError_t e = cudaMemcpy(dst, tmp, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    FOLEY_CUDA_OK(e);
    if (n_out) *n_out = n;
    return FOLEY_OK;
}

}  // namespace foley

// ------------------------------------------------------------------------------------------------ C ABI
using namespace foley;

struct foley_encoder {
    Encoder impl;
};

#define ENC_GUARD_BEGIN try {
#define ENC_GUARD_END                                                          \
    } catch (const std::bad_alloc&) {                                          \
        return fail(FOLEY_ERR_CUDA, "out of host memory");                     \
    } catch (const std::exception& ex) {                                       \
        return fail(FOLEY_ERR_CUDA, std::string("internal error: ") + ex.what()); \
    }

extern "C" foley_status foley_encoder_create(const foley_encoder_config* cfg, int device, foley_encoder** out) {
    if (!cfg || !out) return fail(FOLEY_ERR_INVALID, "foley_encoder_create: null argument");
    ENC_GUARD_BEGIN
    foley_encoder* e = new foley_encoder();
    foley_status s = e->impl.create(cfg, device);
    if (s != FOLEY_OK) { delete e; return s; }
    *out = e;
    return FOLEY_OK;
    ENC_GUARD_END
}
extern "C" void foley_encoder_destroy(foley_encoder* e) { delete e; }
extern "C" foley_status foley_encoder_load_tensor(foley_encoder* e, const char* name, const void* data, const int64_t* shape,
                                                  int32_t ndim, int32_t dtype) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null encoder");
    ENC_GUARD_BEGIN
    return e->impl.load_tensor(name, data, shape, ndim, dtype);
    ENC_GUARD_END
}
extern "C" foley_status foley_encoder_load_safetensors(foley_encoder* e, const char* path, const char* prefix, int64_t* n_loaded) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null encoder");
    ENC_GUARD_BEGIN
    return e->impl.load_safetensors(path, prefix, n_loaded);
    ENC_GUARD_END
}
extern "C" foley_status foley_encoder_finalize(foley_encoder* e) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null encoder");
    ENC_GUARD_BEGIN
    return e->impl.finalize();
    ENC_GUARD_END
}
extern "C" foley_status foley_siglip_encode(foley_encoder* e, const float* pixels, int32_t n_frames, void* out, void* stream) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null encoder");
    ENC_GUARD_BEGIN
    foley_status s = e->impl.siglip_encode(pixels, n_frames, out, e->impl.pick_stream(stream));
    if (s != FOLEY_OK) return s;
    return e->impl.order_after(stream);
    ENC_GUARD_END
}
extern "C" foley_status foley_clap_text_encode(foley_encoder* e, const int32_t* ids, const int32_t* mask, int32_t batch, int32_t T,
                                               void* out, void* stream) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null encoder");
    ENC_GUARD_BEGIN
    foley_status s = e->impl.clap_encode(ids, mask, batch, T, out, e->impl.pick_stream(stream));
    if (s != FOLEY_OK) return s;
    return e->impl.order_after(stream);
    ENC_GUARD_END
}
extern "C" foley_status foley_synchformer_encode(foley_encoder* e, const float* frames, int32_t n_frames, float* out, void* stream) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null encoder");
    ENC_GUARD_BEGIN
    foley_status s = e->impl.synchformer_encode(frames, n_frames, out, e->impl.pick_stream(stream));
    if (s != FOLEY_OK) return s;
    return e->impl.order_after(stream);
    ENC_GUARD_END
}
extern "C" foley_status foley_encoder_set_option(foley_encoder* e, const char* key, int64_t value) {
    if (!e || !key) return fail(FOLEY_ERR_INVALID, "foley_encoder_set_option: null argument");
    const std::string k(key);
    if (k == "layers_run") { e->impl.layers_run = static_cast<int>(value); return FOLEY_OK; }
    if (k == "att_tc") { e->impl.att_tc = static_cast<int>(value); return FOLEY_OK; }
    return fail(FOLEY_ERR_INVALID, "unknown encoder option " + k);
}
extern "C" int64_t foley_encoder_launch_count(const foley_encoder* e) { return e ? e->impl.launches : 0; }
extern "C" foley_status foley_encoder_debug_read(foley_encoder* e, const char* what, float* dst, int64_t cap, int64_t* n_out) {
    if (!e) return fail(FOLEY_ERR_INVALID, "null encoder");
    ENC_GUARD_BEGIN
    return e->impl.debug_read(what, dst, cap, n_out);
    ENC_GUARD_END
}

// softmax(Q K^T * scale) V for head_dim 64, exported for unit tests.  impl 0: flash kernel (no mask); impl 1: one warp per
// query row (key mask, optional bf16 rounding of scores); impl 2: tcgen05 / TMEM kernel (no mask); impl 3: one CTA per
// (sample, head) for a single query over many keys (class-token / pooling-probe queries).
extern "C" foley_status foley_attention_d64(const void* q, const void* k, const void* v, void* out, int32_t batch, int32_t heads,
                                            int32_t Sq, int32_t Sk, int64_t q_batch_stride, int64_t q_row_stride,
                                            int64_t kv_batch_stride, int64_t kv_row_stride, int64_t o_batch_stride,
                                            int64_t o_row_stride, float scale, const int32_t* key_mask, int32_t round_scores,
                                            int32_t impl, void* stream) {
    if (!q || !k || !v || !out || batch < 1 || heads < 1 || Sq < 1 || Sk < 1) return fail(FOLEY_ERR_INVALID, "foley_attention_d64: bad argument");
    if (impl != 1 && impl != 3 && (key_mask || round_scores)) return fail(FOLEY_ERR_UNSUPPORTED, "foley_attention_d64: the flash kernels take no mask");
    EncAttnArgs a;
    a.q = static_cast<const bf16*>(q); a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v); a.o = static_cast<bf16*>(out);
    a.B = batch; a.H = heads; a.Sq = Sq; a.Sk = Sk;
    a.q_batch_stride = q_batch_stride; a.q_row_stride = q_row_stride; a.kv_batch_stride = kv_batch_stride; a.kv_row_stride = kv_row_stride;
    a.o_batch_stride = o_batch_stride; a.o_row_stride = o_row_stride; a.scale = scale; a.key_mask = key_mask; a.round_scores = round_scores;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static bool attr = false;
    if (!attr) {
        FOLEY_CUDA_OK(cudaFuncSetAttribute(enc_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EA_SMEM));
        FOLEY_CUDA_OK(cudaFuncSetAttribute(enc_small_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        FOLEY_CUDA_OK(attention_tc64_init());
        attr = true;
    }
    if (impl == 2 || impl == 4) {         // 2: two threads per query row (default), 4: one thread per row
        AttTc64Args t;
        t.o = a.o; t.o_batch_stride = o_batch_stride; t.o_row_stride = o_row_stride; t.H = heads; t.Sq = Sq; t.Sk = Sk;
        t.scale_log2 = scale * 1.4426950408889634f;
        std::string err;
        if (!launch_attention_tc64(a.q, a.k, a.v, q_row_stride, q_batch_stride, kv_row_stride, kv_batch_stride, batch, t, st, &err,
                                   impl == 2 ? 2 : 1))
            return fail(FOLEY_ERR_CUDA, err);
        return FOLEY_OK;
    }
    if (impl == 3) {
        if (Sq != 1) return fail(FOLEY_ERR_INVALID, "foley_attention_d64: impl 3 is the one-query kernel");
        return launch_cls_attention<false>(a, st);
    }
    if (impl == 1) {
        const size_t smem = static_cast<size_t>(ESA_WARPS) * Sk * sizeof(float);
        if (smem > 64 * 1024) return fail(FOLEY_ERR_UNSUPPORTED, "foley_attention_d64: too many keys for the small kernel");
        const long long units = static_cast<long long>(batch) * heads * Sq;
        FOLEY_CUDA_OK(launch_k(enc_small_attention_kernel<false>, dim3(enc_blocks(units, ESA_WARPS)), dim3(32 * ESA_WARPS), smem, st, a));
    } else {
        dim3 grid((Sq + EA_BM - 1) / EA_BM, heads, batch);
        FOLEY_CUDA_OK(launch_k(enc_attention_kernel<false>, grid, dim3(32 * EA_NW), EA_SMEM, st, a));
    }
    return FOLEY_OK;
}
