// DAC-VAE decoder (dac.py:98-149, 280-303; dac_vae/nn/layers.py) on the tcgen05 GEMM kernel, kind::tf32:
// every Conv1d / ConvTranspose1d is a multi-tap GEMM over channel-last fp32 activations [B, T, C], with
// per-sample zero padding supplied by TMA out-of-bounds fill.  The reference runs this stack in fp32 through
// cuDNN with torch.backends.cudnn.allow_tf32 = True (torch default; nodes.py:398), i.e. also on TF32 tensor
// cores (profiles/r01_torch_probe.json: 2.9e-4 relative error vs fp64 for conv1d and conv_transpose1d).
//
//  * weight-norm is folded once at load (w = g * v / ||v||, per dim-0 slice; for ConvTranspose1d that is per
//    INPUT channel) and weights are rounded to tf32.
//  * Snake1d of the *next* layer is applied in the producing GEMM's epilogue (out2), the residual add of
//    ResidualUnit in the 1x1 conv's epilogue, so each conv reads/writes its activations once.
//  * ConvTranspose1d(k=2s, stride s, pad ceil(s/2), out_pad s%2) is the polyphase 2-tap GEMM
//      out[i, r*C_out + co] = sum_ci x[i, ci] w[ci, co, r] + x[i-1, ci] w[ci, co, r+s],   y[t = i*s + r - pad]
//    whose [T_in+1, s*C_out] row-major output *is* the channel-last y shifted by pad rows.
#include <algorithm>
#include <cmath>

#include "engine.cuh"

namespace foley {

#ifndef ST_OK
#define ST_OK(expr)                         \
    do {                                    \
        foley_status _s = (expr);           \
        if (_s != FOLEY_OK) return _s;      \
    } while (0)
#endif


struct DacLayer {
    float* w = nullptr;      // [N, taps*K] tf32-rounded
    float* bias = nullptr;   // [C_out]
    float* alpha = nullptr;  // snake alpha applied to this layer's OUTPUT for the next conv (may be null)
    int n = 0, k = 0, taps = 1, off0 = 0, tstride = 1;
    int c_out = 0;           // channels of the output (n = stride*c_out for the transposed conv)
    int up = 1;              // upsampling factor (transposed conv) or 1
    int pad = 0;             // transposed conv padding
    int kind = 0;            // 0 conv (out2 only), 1 conv + residual (out and out2), 2 transposed conv, 3 plain (out only)
};

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// weight-norm fold: one block per dim-0 slice; writes w = g * v / ||v|| in place into `out` (same shape as v)
__global__ void fold_weight_norm_kernel(const float* g, const float* v, long long slice, float* out) {
    __shared__ float red[32];
    const long long base = static_cast<long long>(blockIdx.x) * slice;
    float s = 0.f;
    for (long long i = threadIdx.x; i < slice; i += blockDim.x) { const float x = v[base + i]; s += x * x; }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    const float scale = g[blockIdx.x] / sqrtf(red[0]);
    for (long long i = threadIdx.x; i < slice; i += blockDim.x) out[base + i] = v[base + i] * scale;
}

// Conv1d weight [N, K, taps] (fp32) -> [N, taps*K] tap-major, tf32-rounded
__global__ void dac_pack_conv_kernel(const float* w, int N, int K, int taps, float* dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(N) * K * taps;
    if (i >= total) return;
    const int k = static_cast<int>(i % K);
    const int tap = static_cast<int>((i / K) % taps);
    const int n = static_cast<int>(i / (static_cast<long long>(K) * taps));
    dst[i] = to_tf32(w[(static_cast<long long>(n) * K + k) * taps + tap]);
}
// ConvTranspose1d weight [C_in, C_out, 2s] -> [s*C_out, 2*C_in]: dst[r*C_out+co, tau*C_in+ci] = w[ci, co, r + tau*s]
__global__ void dac_pack_convT_kernel(const float* w, int Cin, int Cout, int s, float* dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(s) * Cout * 2 * Cin;
    if (i >= total) return;
    const int ci = static_cast<int>(i % Cin);
    const int tau = static_cast<int>((i / Cin) % 2);
    const long long row = i / (2LL * Cin);
    const int co = static_cast<int>(row % Cout), r = static_cast<int>(row / Cout);
    dst[i] = to_tf32(w[(static_cast<long long>(ci) * Cout + co) * (2 * s) + r + tau * s]);
}

// z [B, ch, L] -> [B, L, ch]
__global__ void dac_transpose_in_kernel(const float* z, int ch, int L, float* out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, l = l0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < ch && l < L) ? z[(static_cast<long long>(b) * ch + c) * L + l] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        if (l < L && c < ch) out[(static_cast<long long>(b) * L + l) * ch + c] = tile[threadIdx.x][i];
    }
}

// Final Conv1d(C -> 1, k=7, pad 3) + tanh on CUDA cores (N = 1 has no tensor-core shape): one thread per output
// sample, the 7*C weights in shared memory, activation rows read as float4 (channel-last, L1/L2 resident).
__global__ void dac_final_conv_tanh_kernel(const float* s, const float* w /*[7][C]*/, const float* bias, int C,
                                           long long T, float* wav) {
    extern __shared__ float wsm[];
    for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) wsm[i] = w[i];
    __syncthreads();
    const int b = blockIdx.y;
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float* sb = s + static_cast<long long>(b) * T * C;
    float acc = bias[0];
    for (int tap = 0; tap < 7; ++tap) {
        const long long tt = t + tap - 3;
        if (tt < 0 || tt >= T) continue;
        const float4* row = reinterpret_cast<const float4*>(sb + tt * C);
        const float* wr = wsm + tap * C;
        for (int c = 0; c < C / 4; ++c) {
            const float4 v = row[c];
            acc += v.x * wr[4 * c] + v.y * wr[4 * c + 1] + v.z * wr[4 * c + 2] + v.w * wr[4 * c + 3];
        }
    }
    wav[static_cast<long long>(b) * T + t] = tanhf(acc);
}

static foley_status raw_f32(Engine* e, const std::string& name, float** out, RawTensor** rtp) {
    auto it = e->raw.find(name);
    if (it == e->raw.end()) return fail(FOLEY_ERR_MISSING, "missing tensor: " + name);
    RawTensor& rt = it->second;
    float* d = nullptr;
    FOLEY_CUDA_OK(cudaMalloc(&d, std::max<size_t>(rt.numel * 4, 16)));
    convert_to_f32_kernel<<<blocks_for(rt.numel, 256), 256>>>(rt.dev, rt.dtype, rt.numel, d);
    FOLEY_CUDA_OK(cudaGetLastError());
    *out = d;
    if (rtp) *rtp = &rt;
    return FOLEY_OK;
}

// Loads a (weight-normed or plain) conv weight as fp32 in its original layout.
static foley_status folded_weight(Engine* e, const std::string& base, bool weight_norm, float** out, std::vector<int64_t>* shape) {
    RawTensor* rt = nullptr;
    if (!weight_norm) {
        ST_OK(raw_f32(e, base + ".weight", out, &rt));
        *shape = rt->shape;
        return FOLEY_OK;
    }
    float *g = nullptr, *v = nullptr;
    RawTensor* rg = nullptr;
    ST_OK(raw_f32(e, base + ".parametrizations.weight.original0", &g, &rg));
    ST_OK(raw_f32(e, base + ".parametrizations.weight.original1", &v, &rt));
    *shape = rt->shape;
    if (rg->numel != rt->shape[0]) return fail(FOLEY_ERR_INVALID, "weight-norm g/v mismatch: " + base);
    const long long slice = rt->numel / rt->shape[0];
    fold_weight_norm_kernel<<<static_cast<unsigned>(rt->shape[0]), 256>>>(g, v, slice, v);
    FOLEY_CUDA_OK(cudaGetLastError());
    FOLEY_CUDA_OK(cudaDeviceSynchronize());
    cudaFree(g);
    *out = v;
    return FOLEY_OK;
}

void delete_dac_layer(DacLayer* l) { delete l; }

foley_status Engine::dac_finalize() {
    if (dac_ready) return FOLEY_OK;
    if (!dac_layers.empty() || !dac_allocs.empty()) {   // DAC weights reloaded: drop the previous packed decoder
        FOLEY_CUDA_OK(cudaDeviceSynchronize());
        free_dac();
    }
    const std::string P = "dac.";
    auto keep = [&](void* p) { dac_allocs.push_back(p); };
    auto vec = [&](const std::string& name, float** out) -> foley_status {
        ST_OK(raw_f32(this, P + name, out, nullptr));
        keep(*out);
        return FOLEY_OK;
    };
    auto conv = [&](const std::string& base, bool wn, int dil, DacLayer* L) -> foley_status {
        float* w = nullptr;
        std::vector<int64_t> sh;
        ST_OK(folded_weight(this, P + base, wn, &w, &sh));
        if (sh.size() != 3) return fail(FOLEY_ERR_INVALID, "conv weight must be 3-D: " + base);
        L->n = static_cast<int>(sh[0]); L->k = static_cast<int>(sh[1]); L->taps = static_cast<int>(sh[2]);
        L->c_out = L->n; L->up = 1;
        L->off0 = -((L->taps - 1) / 2) * dil; L->tstride = dil;
        const long long numel = static_cast<long long>(L->n) * L->k * L->taps;
        FOLEY_CUDA_OK(cudaMalloc(&L->w, numel * 4));
        dac_pack_conv_kernel<<<blocks_for(numel, 256), 256>>>(w, L->n, L->k, L->taps, L->w);
        FOLEY_CUDA_OK(cudaGetLastError());
        FOLEY_CUDA_OK(cudaDeviceSynchronize());
        cudaFree(w);
        keep(L->w);
        ST_OK(vec(base + ".bias", &L->bias));
        return FOLEY_OK;
    };
    // discover depth: decoder.model.{1..n}.block.* are DecoderBlocks
    int n_blocks = 0;
    while (raw.count(P + "decoder.model." + std::to_string(n_blocks + 1) + ".block.0.alpha")) ++n_blocks;
    if (n_blocks == 0) return fail(FOLEY_ERR_MISSING, "DAC decoder weights not loaded (dac.decoder.model.*)");
    dac_layers.clear();
    auto add = [&](DacLayer L) { dac_layers.push_back(new DacLayer(L)); };

    DacLayer pq; pq.kind = 3;
    ST_OK(conv("post_quant_conv", false, 1, &pq));
    add(pq);
    DacLayer c0; c0.kind = 0;
    ST_OK(conv("decoder.model.0", true, 1, &c0));
    ST_OK(vec("decoder.model.1.block.0.alpha", &c0.alpha));
    add(c0);
    for (int i = 0; i < n_blocks; ++i) {
        const std::string b = "decoder.model." + std::to_string(i + 1) + ".block.";
        // transposed conv
        DacLayer t; t.kind = 2;
        float* w = nullptr;
        std::vector<int64_t> sh;
        ST_OK(folded_weight(this, P + b + "1", true, &w, &sh));
        const int Cin = static_cast<int>(sh[0]), Cout = static_cast<int>(sh[1]), k2 = static_cast<int>(sh[2]);
        const int s = k2 / 2;
        t.n = s * Cout; t.k = Cin; t.taps = 2; t.off0 = 0; t.tstride = -1; t.c_out = Cout; t.up = s; t.pad = (s + 1) / 2;
        const long long numel = static_cast<long long>(t.n) * 2 * Cin;
        FOLEY_CUDA_OK(cudaMalloc(&t.w, numel * 4));
        dac_pack_convT_kernel<<<blocks_for(numel, 256), 256>>>(w, Cin, Cout, s, t.w);
        FOLEY_CUDA_OK(cudaGetLastError());
        FOLEY_CUDA_OK(cudaDeviceSynchronize());
        cudaFree(w);
        keep(t.w);
        ST_OK(vec(b + "1.bias", &t.bias));
        ST_OK(vec(b + "2.block.0.alpha", &t.alpha));
        add(t);
        const int dils[3] = {1, 3, 9};
        for (int j = 0; j < 3; ++j) {
            const std::string r = b + std::to_string(j + 2) + ".block.";
            DacLayer c7; c7.kind = 0;
            ST_OK(conv(r + "1", true, dils[j], &c7));
            ST_OK(vec(r + "2.alpha", &c7.alpha));
            add(c7);
            DacLayer c1; c1.kind = 1;
            ST_OK(conv(r + "3", true, 1, &c1));
            std::string next_alpha;
            if (j < 2) next_alpha = b + std::to_string(j + 3) + ".block.0.alpha";
            else if (i + 1 < n_blocks) next_alpha = "decoder.model." + std::to_string(i + 2) + ".block.0.alpha";
            else next_alpha = "decoder.model." + std::to_string(n_blocks + 1) + ".alpha";
            ST_OK(vec(next_alpha, &c1.alpha));
            add(c1);
        }
    }
    // final conv C -> 1: weights as [7][C] fp32 (CUDA-core kernel, exact fp32)
    {
        DacLayer f; f.kind = 4;
        float* w = nullptr;
        std::vector<int64_t> sh;
        ST_OK(folded_weight(this, P + "decoder.model." + std::to_string(n_blocks + 2), true, &w, &sh));
        if (sh[0] != 1 || sh[2] != 7) return fail(FOLEY_ERR_INVALID, "final DAC conv must be [1, C, 7]");
        f.n = 1; f.k = static_cast<int>(sh[1]); f.taps = 7;
        std::vector<float> hw(static_cast<size_t>(f.k) * 7), hp(static_cast<size_t>(f.k) * 7);
        FOLEY_CUDA_OK(cudaMemcpy(hw.data(), w, hw.size() * 4, cudaMemcpyDeviceToHost));
        for (int c = 0; c < f.k; ++c)
            for (int tap = 0; tap < 7; ++tap) hp[static_cast<size_t>(tap) * f.k + c] = hw[static_cast<size_t>(c) * 7 + tap];
        FOLEY_CUDA_OK(cudaMalloc(&f.w, hp.size() * 4));
        FOLEY_CUDA_OK(cudaMemcpy(f.w, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
        cudaFree(w);
        keep(f.w);
        ST_OK(vec("decoder.model." + std::to_string(n_blocks + 2) + ".bias", &f.bias));
        add(f);
    }
    for (auto it = raw.begin(); it != raw.end();) {
        if (it->first.rfind("dac.", 0) == 0) { cudaFree(it->second.dev); it = raw.erase(it); }
        else ++it;
    }
    dac_ready = true;
    return FOLEY_OK;
}

foley_status Engine::dac_decode(const float* z, int batch, int L, float* wav, cudaStream_t st) {
    if (!dac_ready) return fail(FOLEY_ERR_STATE, "DAC decoder weights not loaded");
    if (batch < 1 || L < 1) return fail(FOLEY_ERR_INVALID, "dac_decode: bad shape");
    FOLEY_CUDA_OK(cudaSetDevice(device));
    // buffer capacity: max over stages of T*C per sample
    long long T = L, need = static_cast<long long>(L) * LAT;
    for (const DacLayer* l : dac_layers) {
        if (l->kind == 4) continue;
        T *= l->up;
        need = std::max(need, (T + 1) * l->c_out);
    }
    need *= batch;
    if (need > dac_buf_elems) {
        FOLEY_CUDA_OK(cudaStreamSynchronize(st));
        for (int i = 0; i < 4; ++i) {
            if (dac_buf[i]) {
                cudaFree(dac_buf[i]);
                dac_allocs.erase(std::remove(dac_allocs.begin(), dac_allocs.end(), static_cast<void*>(dac_buf[i])), dac_allocs.end());
            }
            FOLEY_CUDA_OK(cudaMalloc(&dac_buf[i], need * 4 + 4096));
            dac_allocs.push_back(dac_buf[i]);
        }
        dac_buf_elems = need;
    }
    float *X = dac_buf[0], *S = dac_buf[1], *M = dac_buf[2], *X2 = dac_buf[3];
    dim3 blk(32, 8), grid((L + 31) / 32, (LAT + 31) / 32, batch);
    dac_transpose_in_kernel<<<grid, blk, 0, st>>>(z, LAT, L, M);
    ++launches;
    T = L;
    const float* in = M;      // current conv input
    auto run = [&](const DacLayer& l, const float* A, long long T_in, float* out, float* out2, const float* resid,
                   long long* T_out) -> foley_status {
        GemmLaunch g;
        g.a.ptr = A; g.a.dtype = DT_F32; g.a.k = l.k; g.a.rows = T_in; g.a.batch = batch; g.a.ld = l.k;
        g.a.batch_stride = T_in * l.k;
        g.w = l.w; g.n = l.n; g.taps = l.taps; g.tap_off0 = l.off0; g.tap_stride = l.tstride; g.splits = 1;
        g.bn = l.n >= 256 ? 256 : (l.n >= 128 ? 128 : 64);
        g.epi.mode = EPI_DAC;
        g.epi.bias = l.bias;
        g.epi.alpha = out2 ? l.alpha : nullptr;
        g.epi.ch_mod = l.c_out;
        g.epi.ldo = l.n;
        if (l.kind == 2) {
            const long long To = T_in * l.up;
            g.out_rows = T_in + 1;
            g.epi.out_batch_stride = To * l.c_out;
            g.epi.flat_lo = static_cast<long long>(l.pad) * l.c_out;
            g.epi.flat_hi = (static_cast<long long>(l.pad) + To) * l.c_out;
            const long long shift = static_cast<long long>(l.pad) * l.c_out;
            g.epi.out = out ? out - shift : nullptr;
            g.epi.out2 = out2 ? out2 - shift : nullptr;
            *T_out = To;
        } else {
            g.epi.out_batch_stride = T_in * l.n;
            g.epi.out = out;
            g.epi.out2 = out2;
            g.epi.resid = resid;
            *T_out = T_in;
        }
        std::string err;
        if (!launch_gemm(g, st, &err)) return fail(FOLEY_ERR_CUDA, "DAC " + err);
        ++launches;
        return FOLEY_OK;
    };
    size_t li = 0;
    long long To = 0;
    // post_quant_conv: M -> X (raw)
    ST_OK(run(*dac_layers[li++], in, T, X, nullptr, nullptr, &To));
    // conv0: X -> S = snake(conv0(X))
    ST_OK(run(*dac_layers[li++], X, T, nullptr, S, nullptr, &To));
    while (li < dac_layers.size() && dac_layers[li]->kind == 2) {
        // transposed conv: S -> X (raw), M (snake'd)   [T *= stride]
        ST_OK(run(*dac_layers[li++], S, T, X, M, nullptr, &To));
        T = To;
        float *x = X, *s = M, *mid = S, *x2 = X2;
        for (int j = 0; j < 3; ++j) {
            ST_OK(run(*dac_layers[li++], s, T, nullptr, mid, nullptr, &To));      // conv7(snake(x)) -> snake -> mid
            // conv1 + residual: x2 = x + conv1(mid); s' = snake_next(x2).  s is free now -> reuse as s'
            ST_OK(run(*dac_layers[li++], mid, T, x2, s, x, &To));
            std::swap(x, x2);
        }
        // after 3 units the current (raw, snake'd) pair is (x, s); make S hold the snake'd tensor for the next stage
        if (s != S) {
            // s is M here; next transposed conv reads from S: swap roles by pointer
            std::swap(S, M);
        }
        X = x; X2 = x2;
    }
    const DacLayer& f = *dac_layers[li];
    if (f.kind != 4) return fail(FOLEY_ERR_STATE, "DAC layer list corrupted");
    {
        dim3 g2(static_cast<unsigned>((T + 255) / 256), batch);
        dac_final_conv_tanh_kernel<<<g2, 256, 7 * f.k * sizeof(float), st>>>(S, f.w, f.bias, f.k, T, wav);
        ++launches;
    }
    FOLEY_CUDA_OK(cudaGetLastError());
    return FOLEY_OK;
}

}  // namespace foley
