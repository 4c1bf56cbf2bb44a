// Row-wise (HBM/L2-bound) kernels of the DiT step: everything between the tensor-core GEMMs.
// One CTA owns one token row (C/4 threads, one float4 each), LayerNorm statistics are shuffle + one
// shared-memory hop, and every global access is a coalesced 8/16-byte vector.
//
// Rounding points follow the reference's CUDA-autocast dataflow (oracle/foley_oracle.py "cuda_bf16",
// SURVEY.md Appendix B): fp32 LayerNorm, (1+scale) rounded to bf16, result rounded to bf16 for the next
// GEMM; bf16 gate multiply; fp32 audio residual stream, bf16-valued visual stream.
#pragma once
#include "ptx.cuh"

namespace foley {

// Where a row finds its modulation vectors (shift / scale / gate), all bf16:
//   ptr = base + grp_or_trow * sample_stride + l * tok_stride + chunk * C
// by_trow = 1: triple blocks, indexed by the timestep row of the sample's group (per-sample vectors);
// by_trow = 0: single blocks, indexed by the sample's group, per-token vectors.
struct ModRef {
    const __nv_bfloat16* base = nullptr;
    long long sample_stride = 0;
    long long tok_stride = 0;
    int by_trow = 0;
    const int* tok_map = nullptr;   // per-token vectors stored once per DISTINCT source row: row of token l of condition
                                    // c is tok_map[c * L + l] (rows of all conditions share one table)
};

struct RowMap {
    const int* grp_of_sample;  // [B2]
    const int* trow_of_grp;    // [G]
    const int* cond_of_grp;    // [G]
    int L;                     // rows (tokens) per sample
};

__device__ __forceinline__ const __nv_bfloat16* mod_ptr(const ModRef& m, const RowMap& rm, int b, int l, int chunk, int C) {
    const int g = rm.grp_of_sample[b];
    const long long s = m.by_trow ? rm.trow_of_grp[g] : g;
    const int lt = m.tok_map ? m.tok_map[rm.cond_of_grp[g] * rm.L + l] : l;
    return m.base + s * m.sample_stride + static_cast<long long>(lt) * m.tok_stride + static_cast<long long>(chunk) * C;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void load_bf16x4(const __nv_bfloat16* p, float (&o)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
}

struct CombineArgs {
    // ---- residual update (skipped when partials == nullptr):
    //   y = bf16r(sum_s partials[s][row] + bias);  if gate: y = bf16r(y * gate);  x = x_src + y
    const float* partials = nullptr;     // [splits][rows_total][C]
    int splits = 0;
    long long split_stride = 0;
    const __nv_bfloat16* bias = nullptr; // [C] or null
    ModRef gate;                         // gate.base == nullptr -> no gate
    int gate_chunk = 0;
    float* x = nullptr;                  // residual stream fp32 [B2*L][C] (in/out)
    const __nv_bfloat16* x_init = nullptr;  // if set: x_src = x_init[cond_of_grp, l] (bf16, first layer) instead of x
    int round_x = 0;                     // 1: keep the stream bf16-valued (visual stream)
    // ---- LayerNorm + modulate (skipped when h == nullptr):
    //   h = bf16r(LN(x) * bf16r(1 + scale) + shift)
    __nv_bfloat16* h = nullptr;          // [B2*L][C]
    float eps = 1e-6f;
    ModRef mod;                          // mod.base == nullptr -> plain LN
    int shift_chunk = 0, scale_chunk = 1;
    int C = 0;
    int rows_total = 0;
    RowMap rm;
};

// Block-wide sum for blocks of up to 16 warps: one __syncthreads per call (distinct scratch per call site).
__device__ __forceinline__ float block_sum(float v, float* scratch /*[16]*/) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (i < nw) t += scratch[i];
    return t;
}

// grid: one CTA per token row, C/4 threads, one float4 per thread: every access is a fully coalesced 16-byte
// vector and a 500-row call still puts ~40 warps on every SM (the first version used one warp per row and was
// latency-bound at 27 us per call, 30 % of the step; profiles/r01_launches_xl_v1.txt).
__global__ void __launch_bounds__(512) combine_ln_mod_kernel(const CombineArgs a) {
    __shared__ float red0[16], red1[16];
    const int row = blockIdx.x;
    const int b = row / a.rm.L, l = row - b * a.rm.L;
    const int C = a.C;
    const int c = threadIdx.x * 4;
    // Everything that does not depend on the predecessor kernel (bias, gate, shift, scale: weights or modulation
    // vectors produced much earlier in the step) is fetched before the programmatic-dependency wait.
    float bv[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {1.f, 1.f, 1.f, 1.f}, sv[4], cv[4];
    const bool has_part = a.partials != nullptr;
    if (has_part && a.bias) load_bf16x4(a.bias + c, bv);
    if (has_part && a.gate.base) load_bf16x4(mod_ptr(a.gate, a.rm, b, l, a.gate_chunk, C) + c, gv);
    const bool has_mod = a.h != nullptr && a.mod.base != nullptr;
    if (has_mod) {
        load_bf16x4(mod_ptr(a.mod, a.rm, b, l, a.shift_chunk, C) + c, sv);
        load_bf16x4(mod_ptr(a.mod, a.rm, b, l, a.scale_chunk, C) + c, cv);
    }
    pdl_wait();
    pdl_trigger();
    float x[4];
    float* xrow = a.x + static_cast<long long>(row) * C;

    if (a.x_init) {
        const int g = a.rm.grp_of_sample[b];
        load_bf16x4(a.x_init + (static_cast<long long>(a.rm.cond_of_grp[g]) * a.rm.L + l) * C + c, x);
    } else {
        const float4 v = *reinterpret_cast<const float4*>(xrow + c);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }

    if (has_part) {
        const float* pr = a.partials + static_cast<long long>(row) * C + c;
        float4 acc = *reinterpret_cast<const float4*>(pr);
#pragma unroll 4
        for (int s = 1; s < a.splits; ++s) {
            const float4 p = *reinterpret_cast<const float4*>(pr + s * a.split_stride);
            acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
        float y[4] = {acc.x + bv[0], acc.y + bv[1], acc.z + bv[2], acc.w + bv[3]};
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = bf16_round(y[j]);
        if (a.gate.base) {
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = bf16_round(y[j] * gv[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j] += y[j];
            if (a.round_x) x[j] = bf16_round(x[j]);
        }
    }
    if (has_part || a.x_init) *reinterpret_cast<float4*>(xrow + c) = make_float4(x[0], x[1], x[2], x[3]);

    if (a.h) {
        // two-pass LayerNorm in fp32 (mean, then centred second moment), like ATen's CUDA kernel
        const float mean = block_sum((x[0] + x[1]) + (x[2] + x[3]), red0) / C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { x[j] -= mean; q += x[j] * x[j]; }
        const float rstd = rsqrtf(block_sum(q, red1) / C + a.eps);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = x[j] * rstd;
        if (has_mod) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = __fadd_rn(__fmul_rn(o[j], bf16_round(1.0f + cv[j])), sv[j]);
        }
        store_bf16x4(a.h + static_cast<long long>(row) * C + c, o);
    }
}

// ------------------------------------------------------------------------------------------------
// q/k RMSNorm + RoPE + scatter into the attention layout [B, H, S_total, D] (D = 128; one warp per
// (row, head), 4 consecutive channels per lane so each rotation pair lives in one thread).
//   norm_kind 0: reference RMSNorm (norm_layers.py:49-51): bf16r(bf16r(x * rstd(eps=1e-6)) * w)
//   norm_kind 1: torch.nn.RMSNorm(eps=None) on bf16: bf16r(x * rstd(eps=fp32 eps) * w)
struct QkvPart {
    __nv_bfloat16* dst = nullptr;        // [B][H][S_total][128]
    long long dst_batch_stride = 0;      // elements
    long long dst_head_stride = 0;
    int seq_offset = 0;                  // row l lands at seq_offset + l
    const __nv_bfloat16* norm_w = nullptr;  // [128]; nullptr -> plain copy (v)
    int src_col = 0;                     // column offset of this part inside the source row
};

struct QkvArgs {
    const __nv_bfloat16* src = nullptr;  // [B*L][src_ld], parts laid out (part, H, D)
    int src_ld = 0;
    int n_parts = 0;
    QkvPart part[3];
    int H = 0;
    int L = 0;                           // rows per sample
    int rows_total = 0;
    int norm_kind = 0;
    float eps = 1e-6f;
    // split-K producer: when `partials` is set the source row is bf16r(sum_s partials[s][row][col] + bias[col]) — what
    // the bf16 GEMM epilogue would have written — and `src` is not read
    const float* partials = nullptr;     // [splits][rows_total][src_ld] fp32
    int splits = 0;
    long long split_stride = 0;
    const __nv_bfloat16* bias = nullptr; // [src_ld] or null
    const float* cos = nullptr;          // [P][128] fp32 tables
    const float* sin = nullptr;
    const int* pos = nullptr;            // [L] table row of token l; nullptr -> l
};

__global__ void __launch_bounds__(128) qk_norm_rope_kernel(const QkvArgs a) {
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int per_row = a.n_parts * a.H;
    const bool active = w < a.rows_total * per_row;
    const int row = active ? w / per_row : 0;
    const int ph = active ? w - row * per_row : 0;
    const int part = ph / a.H, head = ph - part * a.H;
    const int b = row / a.L, l = row - b * a.L;
    const QkvPart& p = a.part[part];
    // norm weights and RoPE tables are static: fetch them before the programmatic-dependency wait
    float wv[4] = {1.f, 1.f, 1.f, 1.f};
    float4 c = make_float4(1.f, 1.f, 1.f, 1.f), s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && p.norm_w) {
        load_bf16x4(p.norm_w + lane * 4, wv);
        const int prow = a.pos ? a.pos[l] : l;
        c = *reinterpret_cast<const float4*>(a.cos + static_cast<long long>(prow) * 128 + lane * 4);
        s = *reinterpret_cast<const float4*>(a.sin + static_cast<long long>(prow) * 128 + lane * 4);
    }
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};
    const int col = p.src_col + head * 128 + lane * 4;
    if (active && a.partials && a.bias) load_bf16x4(a.bias + col, bias4);   // static too
    pdl_wait();
    pdl_trigger();
    if (!active) return;
    float v[4];
    if (a.partials) {
        const float* pr = a.partials + static_cast<long long>(row) * a.src_ld + col;
        float4 acc = *reinterpret_cast<const float4*>(pr);
        for (int sp = 1; sp < a.splits; ++sp) {
            const float4 q = *reinterpret_cast<const float4*>(pr + sp * a.split_stride);
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        }
        v[0] = bf16_round(acc.x + bias4[0]); v[1] = bf16_round(acc.y + bias4[1]);
        v[2] = bf16_round(acc.z + bias4[2]); v[3] = bf16_round(acc.w + bias4[3]);
    } else {
        load_bf16x4(a.src + static_cast<long long>(row) * a.src_ld + col, v);
    }
    if (p.norm_w) {
        const float ss = warp_sum(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
        const float rstd = rsqrtf(ss * (1.0f / 128.0f) + a.eps);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (a.norm_kind == 0) v[j] = bf16_round(bf16_round(v[j] * rstd) * wv[j]);
            else v[j] = bf16_round(__fmul_rn(__fmul_rn(v[j], rstd), wv[j]));
        }
        // (x0, x1) -> (x0*cos - x1*sin, x1*cos + x0*sin)   (attn_layers.py:112-148)
        const float o0 = __fadd_rn(__fmul_rn(v[0], c.x), __fmul_rn(-v[1], s.x));
        const float o1 = __fadd_rn(__fmul_rn(v[1], c.y), __fmul_rn(v[0], s.y));
        const float o2 = __fadd_rn(__fmul_rn(v[2], c.z), __fmul_rn(-v[3], s.z));
        const float o3 = __fadd_rn(__fmul_rn(v[3], c.w), __fmul_rn(v[2], s.w));
        v[0] = o0; v[1] = o1; v[2] = o2; v[3] = o3;
    }
    __nv_bfloat16* dst = p.dst + static_cast<long long>(b) * p.dst_batch_stride +
                         static_cast<long long>(head) * p.dst_head_stride +
                         static_cast<long long>(p.seq_offset + l) * 128 + lane * 4;
    store_bf16x4(dst, v);
}

// ------------------------------------------------------------------------------------------------ small ops
// e[r, :] = bf16r([cos(t*f), sin(t*f)]), f_k = exp(-ln(1e4) * k / half)   (embed_layers.py:76-103)
__global__ void timestep_embed_kernel(const float* t, int n_t, int dim, __nv_bfloat16* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim / 2;
    if (i >= n_t * half) return;
    const int r = i / half, k = i - r * half;
    const float f = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
    const float arg = t[r] * f;
    out[static_cast<long long>(r) * dim + k] = __float2bfloat16_rn(cosf(arg));
    out[static_cast<long long>(r) * dim + half + k] = __float2bfloat16_rn(sinf(arg));
}

// out = bf16r(silu(in)) elementwise over bf16 (ModulateDiT's activation on the bf16 time vector).
__global__ void silu_bf16_kernel(const __nv_bfloat16* in, __nv_bfloat16* out, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __bfloat162float(in[i]);
    out[i] = __float2bfloat16_rn(x / (1.0f + expf(-x)));
}

// out[ge, k, :] = bf16r(silu(float(src[row_src[k], :]) + float(vec[trow(ge), :])))   (hifi_foley.py:866-867 followed by
// ModulateDiT's SiLU on the fp32 per-token condition), for the n_rows DISTINCT rows of the sync-token table and
// G_eff time vectors.
__global__ void vectok_silu_kernel(const __nv_bfloat16* src, const __nv_bfloat16* vec, const int* row_src,
                                   const int* trow_of_grp, int G_eff, int n_rows, int C, __nv_bfloat16* out) {
    pdl_wait();
    pdl_trigger();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long n4 = static_cast<long long>(G_eff) * n_rows * C / 4;
    if (i >= n4) return;
    const long long e = i * 4;
    const int c = static_cast<int>(e % C);
    const long long gk = e / C;
    const int k = static_cast<int>(gk % n_rows), ge = static_cast<int>(gk / n_rows);
    float a[4], v[4], o[4];
    load_bf16x4(src + static_cast<long long>(row_src[k]) * C + c, a);
    load_bf16x4(vec + static_cast<long long>(trow_of_grp[ge]) * C + c, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float s = a[j] + v[j];
        o[j] = s / (1.0f + expf(-s));
    }
    store_bf16x4(out + e, o);
}

// first[r] = smallest r' <= r whose row equals row r (exact, 16-byte compares): finds the distinct rows of the sync-token
// table (the unconditional half repeats with period 8, text-to-audio repeats everywhere).  One CTA per row.
__global__ void row_first_equal_kernel(const __nv_bfloat16* rows, int n_rows, int C, int* first) {
    const int r = blockIdx.x;
    __shared__ int differs;
    const uint4* mine = reinterpret_cast<const uint4*>(rows + static_cast<long long>(r) * C);
    const int n16 = C / 8;
    int found = r;
    for (int cand = 0; cand < r; ++cand) {
        if (threadIdx.x == 0) differs = 0;
        __syncthreads();
        const uint4* other = reinterpret_cast<const uint4*>(rows + static_cast<long long>(cand) * C);
        int d = 0;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) {
            const uint4 a = mine[i], b = other[i];
            d |= (a.x != b.x) | (a.y != b.y) | (a.z != b.z) | (a.w != b.w);
        }
        if (d) differs = 1;
        __syncthreads();
        if (!differs) { found = cand; break; }
        __syncthreads();
    }
    if (threadIdx.x == 0) first[r] = found;
}

// sync features: out[u, s, :] = bf16r(sync[u, s, :] + pos_emb[s % 8, :])   (hifi_foley.py:757-758)
__global__ void sync_add_pos_kernel(const __nv_bfloat16* sync, const __nv_bfloat16* pos_emb, long long rows, int dim,
                                    int S, __nv_bfloat16* out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * dim) return;
    const int c = static_cast<int>(i % dim);
    const long long r = i / dim;
    const int s = static_cast<int>(r % S);
    out[i] = __float2bfloat16_rn(__bfloat162float(sync[i]) + __bfloat162float(pos_emb[(s & 7) * dim + c]));
}

// nearest-exact gather along the sequence: out[u, l, :] = in[u, idx[l], :]
__global__ void gather_rows_kernel(const __nv_bfloat16* in, const int* idx, int U, int S, int L, int C,
                                   __nv_bfloat16* out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long n8 = static_cast<long long>(U) * L * C / 8;
    if (i >= n8) return;
    const long long e = i * 8;
    const int c = static_cast<int>(e % C);
    const long long ul = e / C;
    const int l = static_cast<int>(ul % L), u = static_cast<int>(ul / L);
    *reinterpret_cast<uint4*>(out + e) =
        *reinterpret_cast<const uint4*>(in + (static_cast<long long>(u) * S + idx[l]) * C + c);
}

// latents fp32 [B, ch, L] -> model input bf16 [B2, L, ch] for every group copy (utils.py:205,223)
__global__ void latents_to_tokens_kernel(const float* lat, int B, int n_rep, int ch, int L, __nv_bfloat16* out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, l = l0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < ch && l < L) ? lat[(static_cast<long long>(b) * ch + c) * L + l] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        if (l < L && c < ch) {
            const __nv_bfloat16 v = __float2bfloat16_rn(tile[threadIdx.x][i]);
            for (int r = 0; r < n_rep; ++r)
                out[((static_cast<long long>(r) * B + b) * L + l) * ch + c] = v;
        }
    }
}

// model output bf16 [B2, L, ch] -> fp32 [B2, ch, L]   (unpatchify1d, hifi_foley.py:926-936)
__global__ void tokens_to_channels_kernel(const __nv_bfloat16* y, int B2, int ch, int L, float* out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < ch && l < L) ? __bfloat162float(y[(static_cast<long long>(b) * L + l) * ch + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, l = l0 + threadIdx.x;
        if (c < ch && l < L) out[(static_cast<long long>(b) * ch + c) * L + l] = tile[threadIdx.x][i];
    }
}

// CFG combine (bf16, utils.py:241-243) + Euler update (fp32, scheduling_flow_match_discrete.py:262-297) +
// next step's bf16 model input.  y: [n_cond*B, L, ch] bf16 (uncond rows first); latents [B, ch, L] fp32.
// step_state: [0] = step index (incremented by thread 0 of block 0 after use), sigmas on device.
__global__ void cfg_euler_kernel(const __nv_bfloat16* y, float* lat, __nv_bfloat16* x_next, int B, int n_cond,
                                 int ch, int L, float guidance, const float* sigmas, const int* step_ptr) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
    const int step = *step_ptr;
    const float dt = sigmas[step + 1] - sigmas[step];
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        float v = 0.f;
        if (l < L && c < ch) {
            if (n_cond == 2) {
                const float u = __bfloat162float(y[(static_cast<long long>(b) * L + l) * ch + c]);
                const float t = __bfloat162float(y[(static_cast<long long>(B + b) * L + l) * ch + c]);
                v = bf16_round(u + bf16_round(guidance * bf16_round(t - u)));
            } else {
                v = __bfloat162float(y[(static_cast<long long>(b) * L + l) * ch + c]);
            }
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, l = l0 + threadIdx.x;
        if (c < ch && l < L) {
            const long long o = (static_cast<long long>(b) * ch + c) * L + l;
            const float nv = __fadd_rn(lat[o], __fmul_rn(tile[threadIdx.x][i], dt));
            lat[o] = nv;
            tile[threadIdx.x][i] = nv;
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        if (l < L && c < ch) {
            const __nv_bfloat16 v = __float2bfloat16_rn(tile[i][threadIdx.x]);
            for (int r = 0; r < n_cond; ++r)
                x_next[((static_cast<long long>(r) * B + b) * L + l) * ch + c] = v;
        }
    }
}

// Multi-stage flow-matching solvers (heun-2 / midpoint-2 / kutta-4; scheduling_flow_match_discrete.py:299-373) inside the
// engine loop: same CFG combine as cfg_euler_kernel, then one stage of the reference scheduler's state machine.  The
// host precomputes one SolverCall per model call (the reference feeds consecutive `timesteps` entries to the model while
// `step_index` only advances after the last stage — mirrored by the table, not "fixed"); the kernel reads the entry of
// the current call through the device step counter, so the captured graph is the same for every call.
struct SolverCall {
    float dt;            // step size of this stage (dt, or dt/2)
    float c0, c1, c2, cm; // kind 2: derivative = ((c0*d0 + c1*d1) + c2*d2) + cm*mo, every product and sum rounded (torch eager)
    int kind;            // 0: derivative = model output; 1: 0.5*(d0 + mo) (heun last stage); 2: kutta last stage
    int store_slot;      // >= 0: keep the model output as derivative d[store_slot]
    int save_sample;     // 1: keep the incoming latents as the step's base sample
    int base_saved;      // 1: update from the saved base sample instead of the incoming latents
};
__global__ void cfg_solver_kernel(const __nv_bfloat16* y, float* lat, __nv_bfloat16* x_next, float* d0, float* d1,
                                  float* d2, float* samp0, int B, int n_cond, int ch, int L, float guidance,
                                  const SolverCall* table, const int* step_ptr) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
    const SolverCall sc = table[*step_ptr];
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        float v = 0.f;
        if (l < L && c < ch) {
            if (n_cond == 2) {
                const float u = __bfloat162float(y[(static_cast<long long>(b) * L + l) * ch + c]);
                const float t = __bfloat162float(y[(static_cast<long long>(B + b) * L + l) * ch + c]);
                v = bf16_round(u + bf16_round(guidance * bf16_round(t - u)));
            } else {
                v = __bfloat162float(y[(static_cast<long long>(b) * L + l) * ch + c]);
            }
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, l = l0 + threadIdx.x;
        if (c < ch && l < L) {
            const long long o = (static_cast<long long>(b) * ch + c) * L + l;
            const float mo = tile[threadIdx.x][i];
            const float cur = lat[o];
            if (sc.save_sample) samp0[o] = cur;
            if (sc.store_slot == 0) d0[o] = mo;
            else if (sc.store_slot == 1) d1[o] = mo;
            else if (sc.store_slot == 2) d2[o] = mo;
            float der = mo;
            if (sc.kind == 1) {
                der = __fmul_rn(0.5f, __fadd_rn(d0[o], mo));
            } else if (sc.kind == 2) {
                float a = __fmul_rn(sc.c0, d0[o]);
                a = __fadd_rn(a, __fmul_rn(sc.c1, d1[o]));
                a = __fadd_rn(a, __fmul_rn(sc.c2, d2[o]));
                der = __fadd_rn(a, __fmul_rn(sc.cm, mo));
            }
            const float base = sc.base_saved ? samp0[o] : cur;
            const float nv = __fadd_rn(base, __fmul_rn(der, sc.dt));
            lat[o] = nv;
            tile[threadIdx.x][i] = nv;
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        if (l < L && c < ch) {
            const __nv_bfloat16 v = __float2bfloat16_rn(tile[i][threadIdx.x]);
            for (int r = 0; r < n_cond; ++r)
                x_next[((static_cast<long long>(r) * B + b) * L + l) * ch + c] = v;
        }
    }
}

// Advances the device-side step counter and the per-group timestep rows (one tiny launch per step so the
// whole step can be replayed as one CUDA graph).
// `host_step` (mapped pinned host memory, may be null) receives the number of completed steps: the host polls it to
// drive the progress callback (utils.py:247 pbar.update) without synchronizing the stream after every step.
__global__ void advance_step_kernel(int* step_ptr, int* trow_of_grp, int G, volatile int* host_step) {
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int s = *step_ptr + 1;
        *step_ptr = s;
        for (int g = 0; g < G; ++g) trow_of_grp[g] = s;
        if (host_step) { *host_step = s; __threadfence_system(); }
    }
}

// ------------------------------------------------------------------------------------------------ weight repack
// ---- source dtypes of checkpoints: bf16 / f32 / f16 and the two FP8 storage formats (FOLEY_DT_*)
__device__ __forceinline__ float decode_f8_e4m3fn(uint8_t b) {   // 1-4-3, bias 7, no inf, NaN = S.1111.111
    const uint32_t sign = (b & 0x80u) << 24, e = (b >> 3) & 0xFu, m = b & 0x7u;
    if (e == 0xFu && m == 0x7u) return __uint_as_float(sign | 0x7FC00000u);
    if (e == 0) return __uint_as_float(sign | __float_as_uint(static_cast<float>(m) * 0.001953125f));   // m * 2^-9
    return __uint_as_float(sign | ((e + 120u) << 23) | (m << 20));
}
__device__ __forceinline__ float decode_f8_e5m2(uint8_t b) {     // the top byte of an fp16
    return __half2float(__ushort_as_half(static_cast<unsigned short>(b) << 8));
}
__device__ __forceinline__ float load_any_as_float(const void* src, int src_dtype, long long i) {
    switch (src_dtype) {
        case 0: return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
        case 1: return reinterpret_cast<const float*>(src)[i];
        case 2: return __half2float(reinterpret_cast<const __half*>(src)[i]);
        case 3: return decode_f8_e4m3fn(reinterpret_cast<const uint8_t*>(src)[i]);
        default: return decode_f8_e5m2(reinterpret_cast<const uint8_t*>(src)[i]);
    }
}
// Round-to-nearest-even through an FP8 storage format (what `weight.to(torch.float8_*)` followed by the per-forward
// upcast does in the reference's FP8WeightWrapper, utils.py:316-356).  mode: 1 = e4m3fn, 2 = e5m2.  Overflow -> NaN
// (e4m3fn) / inf (e5m2), like torch's casts; weights never get there.
__device__ __forceinline__ float round_through_fp8(float v, int mode) {
    if (mode == 1) return decode_f8_e4m3fn(static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_NOSAT, __NV_E4M3)));
    if (mode == 2) return decode_f8_e5m2(static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_NOSAT, __NV_E5M2)));
    return v;
}
// fp8_mode != 0: the value is first rounded to bf16 (load_state_dict into bf16 parameters), then stored in FP8; a
// checkpoint tensor that already is in that FP8 format passes through unchanged.
__global__ void convert_to_bf16_kernel(const void* src, int src_dtype, long long n, __nv_bfloat16* dst, int fp8_mode) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = bf16_round(load_any_as_float(src, src_dtype, i));
    if (fp8_mode != 0 && src_dtype != fp8_mode + 2) v = round_through_fp8(v, fp8_mode);
    dst[i] = __float2bfloat16_rn(v);
}
__global__ void convert_to_f32_kernel(const void* src, int src_dtype, long long n, float* dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = load_any_as_float(src, src_dtype, i);
}

// Conv weight [N, K, taps] -> tap-major GEMM rows dst[(n*row_mul + row_add) , tap*K + k]; also used for plain
// [N, K] matrices (taps = 1).  row_mul/row_add interleave two matrices (SwiGLU w1/w3 pairs).
__global__ void repack_conv_weight_kernel(const __nv_bfloat16* src, int N, int K, int taps, int row_mul, int row_add,
                                          __nv_bfloat16* dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(N) * K * taps;
    if (i >= total) return;
    const int k = static_cast<int>(i % K);
    const int tap = static_cast<int>((i / K) % taps);
    const int n = static_cast<int>(i / (static_cast<long long>(K) * taps));
    dst[(static_cast<long long>(n) * row_mul + row_add) * (static_cast<long long>(K) * taps) + static_cast<long long>(tap) * K + k] =
        src[(static_cast<long long>(n) * K + k) * taps + tap];
}

// Single-block QKV rows (H D K) -> (K H D)   (hifi_foley.py:362): dst row (k*H*D + h*D + d) = src row (h*D*3 + d*3 + k)
__global__ void permute_qkv_rows_kernel(const __nv_bfloat16* src, int H, int D, long long row_len, __nv_bfloat16* dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(3) * H * D * row_len;
    if (i >= total) return;
    const long long col = i % row_len;
    const long long drow = i / row_len;
    const int k = static_cast<int>(drow / (H * D));
    const int hd = static_cast<int>(drow % (H * D));
    const int h = hd / D, d = hd % D;
    dst[i] = src[(static_cast<long long>(h) * D * 3 + d * 3 + k) * row_len + col];
}

}  // namespace foley
