// Kernels of the condition encoders: SigLIP2 vision tower + attention-pooling head and the
// CLAP (RoBERTa) text tower, as the reference runs them — the whole module cast to the DiT's dtype
// (`hunyuan_deps[key].to(device, dtype=target_dtype)`, reference nodes.py:283-284), i.e. bf16 weights, bf16 activations,
// one bf16 rounding after every op, no autocast (feature_utils.py:64-79, 132-138).  The Linear layers run on the
// engine's tcgen05 GEMM (gemm.cuh); this header holds what sits between them:
//   enc_patchify_kernel     Conv2d(3, C, k = p, stride = p) as im2col rows for the GEMM (HF SiglipVisionEmbeddings)
//   enc_add_ln_kernel       residual add (+ position table) and LayerNorm in one pass over a row (HBM-bound)
//   enc_text_embed_kernel   word + token-type + position embeddings and LayerNorm (HF ClapTextEmbeddings)
//   enc_attention_kernel    softmax(QK^T/8)V for head_dim 64, flash-style on mma.sync, operands read in place from the
//                           fused QKV projection output (row stride 3C, head stride 64)
//   enc_small_attention_kernel  one warp per query row: key-padding mask (CLAP), single-probe pooling (SigLIP head)
#pragma once
#include "attention.cuh"
#include "ptx.cuh"
#include "rowwise.cuh"

namespace foley {

// ------------------------------------------------------------------------------------------------ patchify
// pixels fp32 [T, 3, IMG, IMG] -> bf16 rows [(t, py, px)][(c, ky, kx)]: exactly the flattening of the Conv2d weight
// [C_out, 3, P, P], so the patch embedding is one GEMM.  A thread converts 8 consecutive pixels of an image row
// (32 bytes in, 16 bytes out); consecutive threads walk along the image row (coalesced reads).
__global__ void enc_patchify_kernel(const float* __restrict__ px, int T, int IMG, int P, __nv_bfloat16* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int xg = IMG / 8;
    const long long total = static_cast<long long>(T) * 3 * IMG * xg;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x0 = static_cast<int>(i % xg) * 8;
    const int y = static_cast<int>((i / xg) % IMG);
    const int c = static_cast<int>((i / (static_cast<long long>(xg) * IMG)) % 3);
    const int t = static_cast<int>(i / (static_cast<long long>(xg) * IMG * 3));
    const float4* src = reinterpret_cast<const float4*>(px + ((static_cast<long long>(t) * 3 + c) * IMG + y) * IMG + x0);
    const float4 a = src[0], b = src[1];
    const int G = IMG / P;
    const int pr = y / P, ky = y % P, pc = x0 / P, kx = x0 % P;
    const long long row = (static_cast<long long>(t) * G + pr) * G + pc;
    const int col = (c * P + ky) * P + kx;
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(out + row * (3LL * P * P) + col) = o;
}

// ------------------------------------------------------------------------------------------------ add + LayerNorm
// One warp per row of C = 256 * NCH channels (NCH = 3: 768); the row stays in registers (a lane owns 8 channels of every
// 256-wide chunk), statistics in fp32 (mean, then the centred second moment: two passes over registers), the output
// rounded to bf16 once — torch's LayerNorm on bf16 tensors.
//   x = res ? bf16(y + res[row % res_mod]) : y      (y == nullptr: x = res row)
//   x_out (optional) <- x;  h_out (optional) <- bf16((x - mean) * rstd * w + b)
// Algorithmic bytes per row: 2C per tensor touched (y, res, x_out, h_out).
struct EncLnArgs {
    const __nv_bfloat16* y = nullptr;
    const __nv_bfloat16* res = nullptr;
    long long res_mod = 0;                 // > 0: res is a table of res_mod rows indexed by row % res_mod (position embeddings)
    __nv_bfloat16* x_out = nullptr;
    const __nv_bfloat16* ln_w = nullptr;   // nullptr: no LayerNorm
    const __nv_bfloat16* ln_b = nullptr;
    __nv_bfloat16* h_out = nullptr;
    float eps = 1e-6f;
    long long rows = 0;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = __low2float(p[j]); f[2 * j + 1] = __high2float(p[j]); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int NCH>
__device__ __forceinline__ void enc_ln_row(float (&x)[NCH][8], const __nv_bfloat16* w, const __nv_bfloat16* b, float eps,
                                           __nv_bfloat16* dst, int lane) {
    constexpr int C = NCH * 256;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += x[c][j];
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = x[c][j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int off = c * 256 + lane * 8;
        float wf[8], bf[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(w + off), wf);
        unpack8(*reinterpret_cast<const uint4*>(b + off), bf);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((x[c][j] - mean) * rstd, wf[j], bf[j]);
        *reinterpret_cast<uint4*>(dst + off) = pack8(o);
    }
}

template <int NCH>
__global__ void __launch_bounds__(256) enc_add_ln_kernel(const EncLnArgs a) {
    pdl_wait();
    pdl_trigger();
    constexpr int C = NCH * 256;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.rows) return;
    float x[NCH][8];
    const long long rrow = a.res_mod > 0 ? row % a.res_mod : row;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int off = c * 256 + lane * 8;
        if (a.y) {
            unpack8(*reinterpret_cast<const uint4*>(a.y + row * C + off), x[c]);
            if (a.res) {
                float r[8];
                unpack8(*reinterpret_cast<const uint4*>(a.res + rrow * C + off), r);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[c][j] = bf16_round(x[c][j] + r[j]);
            }
        } else {
            unpack8(*reinterpret_cast<const uint4*>(a.res + rrow * C + off), x[c]);
        }
        if (a.x_out) *reinterpret_cast<uint4*>(a.x_out + row * C + off) = pack8(x[c]);
    }
    if (a.ln_w && a.h_out) enc_ln_row<NCH>(x, a.ln_w, a.ln_b, a.eps, a.h_out + row * C, lane);
}

// nn.GELU() (erf) in place on a bf16 tensor: the CLAP text tower's intermediate activation (2 x 77 rows; the bf16 GEMM
// epilogues of the DiT step carry SiLU / GELU-tanh only, and are left exactly as they were).
__global__ void enc_gelu_erf_kernel(__nv_bfloat16* __restrict__ x, long long n8) {
    pdl_wait();
    pdl_trigger();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    uint4* p = reinterpret_cast<uint4*>(x) + i;
    float f[8];
    unpack8(*p, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.5f * f[j] * (1.0f + erff(f[j] * 0.70710678118654752f));
    *p = pack8(f);
}

// ------------------------------------------------------------------------------------------------ CLAP text embeddings
// x = bf16(bf16(word[id] + type[0]) + pos[pos_id]); out = LayerNorm(x)   (HF ClapTextEmbeddings.forward, bf16 module)
template <int NCH>
__global__ void __launch_bounds__(256) enc_text_embed_kernel(const int* __restrict__ ids, const int* __restrict__ pos_ids,
                                                             const __nv_bfloat16* __restrict__ word,
                                                             const __nv_bfloat16* __restrict__ type0,
                                                             const __nv_bfloat16* __restrict__ pos,
                                                             const __nv_bfloat16* ln_w, const __nv_bfloat16* ln_b, float eps,
                                                             long long rows, __nv_bfloat16* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    constexpr int C = NCH * 256;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const long long id = ids[row], pid = pos_ids[row];
    float x[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int off = c * 256 + lane * 8;
        float w[8], t[8], p[8];
        unpack8(*reinterpret_cast<const uint4*>(word + id * C + off), w);
        unpack8(*reinterpret_cast<const uint4*>(type0 + off), t);
        unpack8(*reinterpret_cast<const uint4*>(pos + pid * C + off), p);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[c][j] = bf16_round(bf16_round(w[j] + t[j]) + p[j]);
    }
    enc_ln_row<NCH>(x, ln_w, ln_b, eps, out + row * C, lane);
}

// ------------------------------------------------------------------------------------------------ attention, head_dim 64
// Element (b, h, r, d) of an operand at ptr + b*batch_stride + r*row_stride + h*64 + d: the q / k / v column blocks of the
// fused projection output [rows, 3C] are read in place; the output is token-major [b][r][h*64 + d] with its own strides.
struct EncAttnArgs {
    const __nv_bfloat16* q = nullptr;
    const __nv_bfloat16* k = nullptr;
    const __nv_bfloat16* v = nullptr;
    __nv_bfloat16* o = nullptr;
    int B = 0, H = 0, Sq = 0, Sk = 0;
    long long q_batch_stride = 0, kv_batch_stride = 0, o_batch_stride = 0;   // elements
    long long q_row_stride = 0, kv_row_stride = 0, o_row_stride = 0;
    const int* key_mask = nullptr;      // small kernel only: [B, Sk], 0 = padded key (HF extended attention mask)
    float scale = 0.125f;
    int round_scores = 0;               // scores (and, small kernel, normalised probabilities) rounded to the operand type before /
                                        // after the softmax: the bmm + softmax path of F.multi_head_attention_forward (need_weights)
                                        // and the einsum attention of the Synchformer under fp16 autocast (vit_helper.py:27-35)
    long long kv_gap = 0;               // flash kernel: key j > 0 lives kv_gap rows further on (key 0 = the class token of the segment,
                                        // keys 1.. = the tokens of one frame: divided space attention, vit_helper.py:56-110)
    long long q_batch_stride2 = 0, kv_batch_stride2 = 0, o_batch_stride2 = 0;   // second-level batch: b = b1 * batch2 + b2 adds b2 * stride2
    int batch2 = 1;
};

constexpr int EA_D = 64, EA_BN = 64, EA_NW = 8, EA_BM = 16 * EA_NW, EA_NST = 3;
constexpr int EA_TILE = EA_BN * EA_D * 2;                                 // one K or V tile: 8 KB
constexpr int EA_SMEM = EA_BM * EA_D * 2 + EA_NST * 2 * EA_TILE;          // 16 KB + 48 KB

__device__ __forceinline__ uint32_t swz64(int row, int chunk) {   // 16-byte chunk of a [rows][64] bf16 tile (128-byte rows)
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void ea_load_tile(uint32_t smem_base, const __nv_bfloat16* g, long long row_stride, int r0, int S, int rows) {
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        const bool ok = (r0 + r) < S;
        cp_async16(smem_base + swz64(r, c), g + static_cast<long long>(ok ? r0 + r : 0) * row_stride + c * 8, ok);
    }
}
__device__ __forceinline__ float ea_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// K / V tile with the class-token gap: key 0 is row 0, key j > 0 is row j + gap.
__device__ __forceinline__ void ea_load_tile_gap(uint32_t smem_base, const __nv_bfloat16* g, long long row_stride, int r0, int S, int rows,
                                                 long long gap) {
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        const int key = r0 + r;
        const bool ok = key < S;
        const long long grow = ok ? (key > 0 ? key + gap : 0) : 0;
        cp_async16(smem_base + swz64(r, c), g + grow * row_stride + c * 8, ok);
    }
}

// One CTA = 128 queries of one (sample, head): 8 warps x 16 query rows, 64-key tiles through a 3-deep cp.async ring,
// online softmax in the log2 domain, P rounded to the operand type for the PV product (the row sum taken from the rounded
// values).  kHalf: IEEE fp16 operands (the Synchformer under fp16 autocast) instead of bf16.
template <bool kHalf>
__global__ void __launch_bounds__(32 * EA_NW) enc_attention_kernel(const EncAttnArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t ea_smem[];
    const uint32_t sQ = smem_u32(ea_smem);
    const uint32_t sK0 = sQ + EA_BM * EA_D * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * EA_BM;
    const int b1 = b / a.batch2, b2 = b - b1 * a.batch2;
    const __nv_bfloat16* Q = a.q + b1 * a.q_batch_stride + b2 * a.q_batch_stride2 + h * EA_D;
    const __nv_bfloat16* K = a.k + b1 * a.kv_batch_stride + h * EA_D;
    const __nv_bfloat16* V = a.v + b1 * a.kv_batch_stride + h * EA_D;
    const long long gap = a.kv_gap + b2 * a.kv_batch_stride2;      // (kv_batch_stride2 in ROWS: frame b2 starts b2 * stride2 rows further on)
    const float scale_log2 = a.scale * 1.4426950408889634f;

    const int n_tiles = (a.Sk + EA_BN - 1) / EA_BN;
    auto load_stage = [&](int t) {
        const uint32_t dst = sK0 + (t % EA_NST) * 2 * EA_TILE;
        ea_load_tile_gap(dst, K, a.kv_row_stride, t * EA_BN, a.Sk, EA_BN, gap);
        ea_load_tile_gap(dst + EA_TILE, V, a.kv_row_stride, t * EA_BN, a.Sk, EA_BN, gap);
    };
    ea_load_tile(sQ, Q, a.q_row_stride, q0, a.Sq, EA_BM);
    cp_async_commit();
    for (int t = 0; t < EA_NST - 1; ++t) {
        if (t < n_tiles) load_stage(t);
        cp_async_commit();
    }

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    for (int t = 0; t < n_tiles; ++t) {
        {
            const int tn = t + EA_NST - 1;
            if (tn < n_tiles) load_stage(tn);
            cp_async_commit();
        }
        cp_async_wait_group<EA_NST - 1>();
        __syncthreads();
        const uint32_t sK = sK0 + (t % EA_NST) * 2 * EA_TILE;
        const uint32_t sV = sK + EA_TILE;
        if (t == 0) {
            const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                ldsm_x4(sQ + swz64(row, kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
        }
        const int k0 = t * EA_BN;
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t b0, b1, b2, b3;
                const int krow = jp * 16 + (lane & 7) + (lane >> 4) * 8;
                ldsm_x4(sK + swz64(krow, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
                if constexpr (kHalf) { mma_f16_16816(s[2 * jp], qf[kk], b0, b1); mma_f16_16816(s[2 * jp + 1], qf[kk], b2, b3); }
                else { mma_bf16_16816(s[2 * jp], qf[kk], b0, b1); mma_bf16_16816(s[2 * jp + 1], qf[kk], b2, b3); }
            }
        }
        if (a.round_scores) {      // einsum / bmm output of the reference: the scaled scores exist in the operand type
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[j][e] = kHalf ? f16_round(s[j][e]) : bf16_round(s[j][e]);
        }
        float mx[2] = {-INFINITY, -INFINITY};
        if (k0 + EA_BN <= a.Sk) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { s[j][e] *= scale_log2; mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]); }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = k0 + j * 8 + (lane & 3) * 2;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    s[j][e] = (key + (e & 1)) < a.Sk ? s[j][e] * scale_log2 : -INFINITY;
                    mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
                }
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = ea_ex2(m_run[r] - m_new);
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
        uint32_t pf[4][4];
        float ls[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = ea_ex2(s[j][0] - m_run[0]), p1 = ea_ex2(s[j][1] - m_run[0]);
            const float p2 = ea_ex2(s[j][2] - m_run[1]), p3 = ea_ex2(s[j][3] - m_run[1]);
            if constexpr (kHalf) {
                const __half2 lo = __floats2half2_rn(p0, p1), hi = __floats2half2_rn(p2, p3);
                ls[0] += __low2float(lo) + __high2float(lo);
                ls[1] += __low2float(hi) + __high2float(hi);
                pf[j >> 1][(j & 1) * 2 + 0] = *reinterpret_cast<const uint32_t*>(&lo);
                pf[j >> 1][(j & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
            } else {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
                ls[0] += __low2float(lo) + __high2float(lo);
                ls[1] += __low2float(hi) + __high2float(hi);
                pf[j >> 1][(j & 1) * 2 + 0] = *reinterpret_cast<const uint32_t*>(&lo);
                pf[j >> 1][(j & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
            }
        }
        l_run[0] += ls[0];
        l_run[1] += ls[1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= corr[0]; o[j][1] *= corr[0];
            o[j][2] *= corr[1]; o[j][3] *= corr[1];
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t b0, b1, b2, b3;
                const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm_x4_t(sV + swz64(vrow, jp * 2 + (lane >> 4)), b0, b1, b2, b3);
                if constexpr (kHalf) { mma_f16_16816(o[2 * jp], pf[kk], b0, b1); mma_f16_16816(o[2 * jp + 1], pf[kk], b2, b3); }
                else { mma_bf16_16816(o[2 * jp], pf[kk], b0, b1); mma_bf16_16816(o[2 * jp + 1], pf[kk], b2, b3); }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    const int row0 = q0 + warp * 16 + (lane >> 2), row1 = row0 + 8;
    __nv_bfloat16* O = a.o + b1 * a.o_batch_stride + b2 * a.o_batch_stride2 + h * EA_D + (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (row0 < a.Sq)
            *reinterpret_cast<uint32_t*>(O + row0 * a.o_row_stride + j * 8) =
                kHalf ? pack_f16x2(o[j][0] * inv0, o[j][1] * inv0) : pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
        if (row1 < a.Sq)
            *reinterpret_cast<uint32_t*>(O + row1 * a.o_row_stride + j * 8) =
                kHalf ? pack_f16x2(o[j][2] * inv1, o[j][3] * inv1) : pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
}

// One warp per (sample, head, query row): scores into shared memory (Sk floats per warp), exact softmax, P = bf16(p / sum),
// out = sum_j P_j v_j.  Serves the CLAP text tower (2 x <= 514 tokens, key-padding mask) and the SigLIP pooling head
// (one probe query over the 1024 patch tokens of every frame).
__device__ __forceinline__ void unpack8h(const uint4& u, float (&f)[8]) {
    const __half2* p = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = __low2float(p[j]); f[2 * j + 1] = __high2float(p[j]); }
}
constexpr int ESA_WARPS = 4;
template <bool kHalf>
__global__ void __launch_bounds__(32 * ESA_WARPS) enc_small_attention_kernel(const EncAttnArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float esa_scores[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sc = esa_scores + static_cast<long long>(warp) * a.Sk;
    const long long unit = static_cast<long long>(blockIdx.x) * ESA_WARPS + warp;
    const long long total = static_cast<long long>(a.B) * a.H * a.Sq;
    if (unit >= total) return;
    const int qi = static_cast<int>(unit % a.Sq);
    const int h = static_cast<int>((unit / a.Sq) % a.H);
    const int b = static_cast<int>(unit / (static_cast<long long>(a.Sq) * a.H));
    const __nv_bfloat16* Q = a.q + b * a.q_batch_stride + qi * a.q_row_stride + h * EA_D;
    const __nv_bfloat16* K = a.k + b * a.kv_batch_stride + h * EA_D;
    const __nv_bfloat16* V = a.v + b * a.kv_batch_stride + h * EA_D;
    float qv[EA_D];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float f[8];
        if constexpr (kHalf) unpack8h(*reinterpret_cast<const uint4*>(Q + c * 8), f);
        else unpack8(*reinterpret_cast<const uint4*>(Q + c * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[c * 8 + j] = f[j];
    }
    float mx = -INFINITY;
    for (int key = lane; key < a.Sk; key += 32) {
        const __nv_bfloat16* kr = K + key * a.kv_row_stride;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float f[8];
            if constexpr (kHalf) unpack8h(*reinterpret_cast<const uint4*>(kr + c * 8), f);
            else unpack8(*reinterpret_cast<const uint4*>(kr + c * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) dot = fmaf(qv[c * 8 + j], f[j], dot);
        }
        float s = dot * a.scale;
        if (a.round_scores) s = kHalf ? f16_round(s) : bf16_round(s);
        if (a.key_mask && a.key_mask[static_cast<long long>(b) * a.Sk + key] == 0) s = -INFINITY;
        sc[key] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int key = lane; key < a.Sk; key += 32) {
        const float p = __expf(sc[key] - mx);
        sc[key] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    float acc0 = 0.f, acc1 = 0.f;
    const __nv_bfloat16* vp = V + lane * 2;
    for (int key = 0; key < a.Sk; ++key) {
        if constexpr (kHalf) {
            const float p = f16_round(sc[key] * inv);
            const __half2 v2 = *reinterpret_cast<const __half2*>(vp + key * a.kv_row_stride);
            acc0 = fmaf(p, __low2float(v2), acc0);
            acc1 = fmaf(p, __high2float(v2), acc1);
        } else {
            const float p = bf16_round(sc[key] * inv);
            const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(vp + key * a.kv_row_stride);
            acc0 = fmaf(p, __low2float(v2), acc0);
            acc1 = fmaf(p, __high2float(v2), acc1);
        }
    }
    __nv_bfloat16* O = a.o + b * a.o_batch_stride + qi * a.o_row_stride + h * EA_D + lane * 2;
    *reinterpret_cast<uint32_t*>(O) = kHalf ? pack_f16x2(acc0, acc1) : pack_bf16x2(acc0, acc1);
}

// ONE query per (sample, head) over many keys (the class-token queries of the Synchformer: 1569 keys; the SigLIP pooling probe:
// 1024 keys): one CTA of 8 warps per (sample, head).  Phase 1: a thread owns whole keys (64-channel dot products), scores
// (rounded to the operand type when round_scores) into shared memory, block-wide maximum and sum; phase 2: warp w takes the w-th
// slice of the keys, a lane owns two channels, P = round(p / sum) exactly as the one-warp kernel; the 8 partial outputs are
// summed through shared memory.  Same arithmetic as enc_small_attention_kernel (the summation order over keys differs).
constexpr int ECA_WARPS = 8;
template <bool kHalf>
__global__ void __launch_bounds__(32 * ECA_WARPS) enc_cls_attention_kernel(const EncAttnArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float eca_smem[];
    float* sc = eca_smem;                          // [Sk]
    float* red = eca_smem + ((a.Sk + 3) & ~3);     // [ECA_WARPS][64] partial outputs; first 2 * ECA_WARPS floats double as reduction slots
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
    const __nv_bfloat16* Q = a.q + b * a.q_batch_stride + h * EA_D;
    const __nv_bfloat16* K = a.k + b * a.kv_batch_stride + h * EA_D;
    const __nv_bfloat16* V = a.v + b * a.kv_batch_stride + h * EA_D;
    float qv[EA_D];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float f[8];
        if constexpr (kHalf) unpack8h(*reinterpret_cast<const uint4*>(Q + c * 8), f);
        else unpack8(*reinterpret_cast<const uint4*>(Q + c * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[c * 8 + j] = f[j];
    }
    float mx = -INFINITY;
    for (int key = threadIdx.x; key < a.Sk; key += 32 * ECA_WARPS) {
        const __nv_bfloat16* kr = K + key * a.kv_row_stride;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float f[8];
            if constexpr (kHalf) unpack8h(*reinterpret_cast<const uint4*>(kr + c * 8), f);
            else unpack8(*reinterpret_cast<const uint4*>(kr + c * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) dot = fmaf(qv[c * 8 + j], f[j], dot);
        }
        float sv = dot * a.scale;
        if (a.round_scores) sv = kHalf ? f16_round(sv) : bf16_round(sv);
        if (a.key_mask && a.key_mask[static_cast<long long>(b) * a.Sk + key] == 0) sv = -INFINITY;
        sc[key] = sv;
        mx = fmaxf(mx, sv);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < ECA_WARPS; ++w) mx = fmaxf(mx, red[w]);
    float sum = 0.f;
    for (int key = threadIdx.x; key < a.Sk; key += 32 * ECA_WARPS) {
        const float p = __expf(sc[key] - mx);
        sc[key] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[ECA_WARPS + warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < ECA_WARPS; ++w) sum += red[ECA_WARPS + w];
    const float inv = 1.0f / sum;
    __syncthreads();                                // the reduction slots are about to hold partial outputs
    const int per = (a.Sk + ECA_WARPS - 1) / ECA_WARPS;
    const int k0 = warp * per, k1 = min(a.Sk, k0 + per);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const __nv_bfloat16* vp = V + lane * 2;
    auto term = [&](int key, float& x0, float& x1) {
        if constexpr (kHalf) {
            const float p = f16_round(sc[key] * inv);
            const __half2 v2 = *reinterpret_cast<const __half2*>(vp + key * a.kv_row_stride);
            x0 = fmaf(p, __low2float(v2), x0);
            x1 = fmaf(p, __high2float(v2), x1);
        } else {
            const float p = bf16_round(sc[key] * inv);
            const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(vp + key * a.kv_row_stride);
            x0 = fmaf(p, __low2float(v2), x0);
            x1 = fmaf(p, __high2float(v2), x1);
        }
    };
    int key = k0;
    for (; key + 1 < k1; key += 2) { term(key, acc0, acc1); term(key + 1, acc2, acc3); }
    if (key < k1) term(key, acc0, acc1);
    red[warp * 64 + lane * 2] = acc0 + acc2;
    red[warp * 64 + lane * 2 + 1] = acc1 + acc3;
    __syncthreads();
    if (warp == 0) {
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int w = 0; w < ECA_WARPS; ++w) { o0 += red[w * 64 + lane * 2]; o1 += red[w * 64 + lane * 2 + 1]; }
        __nv_bfloat16* O = a.o + b * a.o_batch_stride + h * EA_D + lane * 2;
        *reinterpret_cast<uint32_t*>(O) = kHalf ? pack_f16x2(o0, o1) : pack_bf16x2(o0, o1);
    }
}

// ================================================================================================ Synchformer (MotionFormer)
// The reference runs its Synchformer visual extractor under torch.autocast(fp16) on a module whose parameters were moved to
// the DiT's dtype (feature_utils.py:100-102, nodes.py:283-284): Linear / Conv3d / einsum in fp16 (fp32 accumulation), LayerNorm
// and softmax in fp32, and a residual stream that is fp32 from the first concatenation on (bf16 class token + fp16 patch
// embeddings promote to fp32).  These kernels keep exactly those types: x fp32, GEMM operands fp16.

// frames fp32 [T, 3, IMG, IMG] -> fp16 im2col rows of Conv3d(3, C, (2, P, P), stride (2, P, P)) over 16-frame windows that
// start every 8 frames: row (s, t', py, px), column (c, kt, ky, kx) — the flattening of the weight [C, 3, 2, P, P].
__global__ void sync_patchify_kernel(const float* __restrict__ frames, int S, int IMG, int P, __half* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int xg = IMG / 8;
    const long long total = static_cast<long long>(S) * 8 * 2 * 3 * IMG * xg;       // (s, t', kt, c, y, x8)
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x0 = static_cast<int>(i % xg) * 8;
    long long r = i / xg;
    const int y = static_cast<int>(r % IMG); r /= IMG;
    const int c = static_cast<int>(r % 3); r /= 3;
    const int kt = static_cast<int>(r % 2); r /= 2;
    const int tp = static_cast<int>(r % 8);
    const int sgm = static_cast<int>(r / 8);
    const int frame = sgm * 8 + tp * 2 + kt;
    const float4* src = reinterpret_cast<const float4*>(frames + ((static_cast<long long>(frame) * 3 + c) * IMG + y) * IMG + x0);
    const float4 a = src[0], b = src[1];
    const int G = IMG / P;
    const long long row = ((static_cast<long long>(sgm) * 8 + tp) * G + y / P) * G + x0 / P;
    const int col = ((c * 2 + kt) * P + y % P) * P + x0 % P;
    uint4 o;
    o.x = pack_f16x2(a.x, a.y); o.y = pack_f16x2(a.z, a.w); o.z = pack_f16x2(b.x, b.y); o.w = pack_f16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(out + row * (6LL * P * P) + col) = o;
}

// Residual stream (fp32) update + LayerNorm, one warp per token row of C = 256 * NCH channels:
//   x = x_init ? x_init row : x;   x += float(y) (fp16) if y;   x_out <- x (fp32);   h_out <- fp16(LayerNorm(x) * w + b);
//   n_out (optional) <- LayerNorm(x) * w + b in fp32 (the final norm, whose output is concatenated with a class token).
// x_init rows come through a row map: token t of segment s reads table row (t == 0 ? 0 : 1 + ...) — see SyncLnArgs.
struct SyncLnArgs {
    const float* x = nullptr;              // [rows, C] fp32 residual stream (read unless one of the build modes below is set)
    float* x_dst = nullptr;                // where the updated stream goes (may alias x); nullptr: not written
    const __half* y = nullptr;             // [rows, C] fp16 branch output to add; nullptr: none
    const __half* patch = nullptr;         // embedding mode: fp16 patch embeddings [segments * (tokens - 1), C]
    const float* tok_tab = nullptr;        // embedding mode: fp32 [tokens, C]: row 0 = cls + pos[0], row t = pos / temporal table
    int tokens = 0;                        // tokens per segment (1569) in embedding mode
    const float* seq_src = nullptr;        // aggregation mode: fp32 normalised tokens [segments * 8 * 196, C] (class token dropped)
    const float* seq_cls = nullptr;        // aggregation mode: fp32 [C] class token of the aggregation layer; with seq_len == 0 every row
                                           // starts from this vector (the class-token rows of the aggregation layer's residual)
    int seq_len = 0;                       // aggregation mode: 197
    const __nv_bfloat16* ln_w = nullptr;   // bf16 parameters of the module (fp32 arithmetic)
    const __nv_bfloat16* ln_b = nullptr;
    __half* h_out = nullptr;
    float* n_out = nullptr;
    long long n_out_skip = 0;              // > 0: rows with (row % n_out_skip == 0) are class tokens and are dropped from n_out
    float eps = 1e-6f;
    long long rows = 0;
};

template <int NCH>
__global__ void __launch_bounds__(256) sync_add_ln_kernel(const SyncLnArgs a) {
    pdl_wait();
    pdl_trigger();
    constexpr int C = NCH * 256;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.rows) return;
    float x[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int off = c * 256 + lane * 8;
        if (a.tok_tab) {                       // embeddings: class token row, or patch + position / temporal table
            const long long sgm = row / a.tokens;
            const int t = static_cast<int>(row - sgm * a.tokens);
            const float4* tp = reinterpret_cast<const float4*>(a.tok_tab + static_cast<long long>(t) * C + off);
            const float4 t0 = tp[0], t1 = tp[1];
            x[c][0] = t0.x; x[c][1] = t0.y; x[c][2] = t0.z; x[c][3] = t0.w; x[c][4] = t1.x; x[c][5] = t1.y; x[c][6] = t1.z; x[c][7] = t1.w;
            if (t > 0) {
                float p[8];
                unpack8h(*reinterpret_cast<const uint4*>(a.patch + (sgm * (a.tokens - 1) + t - 1) * C + off), p);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[c][j] += p[j];
            }
        } else if (a.seq_cls) {                // aggregation sequences: [cls; 196 normalised tokens of one frame] (or cls only)
            const long long sq = a.seq_len > 0 ? row / a.seq_len : 0;
            const int t = a.seq_len > 0 ? static_cast<int>(row - sq * a.seq_len) : 0;
            const float* src = t == 0 ? a.seq_cls + off : a.seq_src + (sq * (a.seq_len - 1) + t - 1) * C + off;
            const float4 t0 = reinterpret_cast<const float4*>(src)[0], t1 = reinterpret_cast<const float4*>(src)[1];
            x[c][0] = t0.x; x[c][1] = t0.y; x[c][2] = t0.z; x[c][3] = t0.w; x[c][4] = t1.x; x[c][5] = t1.y; x[c][6] = t1.z; x[c][7] = t1.w;
        } else {
            const float4* xp = reinterpret_cast<const float4*>(a.x + row * C + off);
            const float4 t0 = xp[0], t1 = xp[1];
            x[c][0] = t0.x; x[c][1] = t0.y; x[c][2] = t0.z; x[c][3] = t0.w; x[c][4] = t1.x; x[c][5] = t1.y; x[c][6] = t1.z; x[c][7] = t1.w;
        }
        if (a.y) {
            float p[8];
            unpack8h(*reinterpret_cast<const uint4*>(a.y + row * C + off), p);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[c][j] += p[j];
        }
        if (a.x_dst) {
            float4* xo = reinterpret_cast<float4*>(a.x_dst + row * C + off);
            xo[0] = make_float4(x[c][0], x[c][1], x[c][2], x[c][3]);
            xo[1] = make_float4(x[c][4], x[c][5], x[c][6], x[c][7]);
        }
    }
    if (!a.ln_w) return;
    float sm = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) sm += x[c][j];
    const float mean = warp_sum(sm) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = x[c][j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + a.eps);
    const bool drop = a.n_out_skip > 0 && (row % a.n_out_skip) == 0;
    const long long nrow = a.n_out_skip > 0 ? row - row / a.n_out_skip - 1 : row;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int off = c * 256 + lane * 8;
        float wf[8], bf[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(a.ln_w + off), wf);
        unpack8(*reinterpret_cast<const uint4*>(a.ln_b + off), bf);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((x[c][j] - mean) * rstd, wf[j], bf[j]);
        if (a.h_out)
            *reinterpret_cast<uint4*>(a.h_out + row * C + off) =
                make_uint4(pack_f16x2(o[0], o[1]), pack_f16x2(o[2], o[3]), pack_f16x2(o[4], o[5]), pack_f16x2(o[6], o[7]));
        if (a.n_out && !drop) {
            float4* no = reinterpret_cast<float4*>(a.n_out + nrow * C + off);
            no[0] = make_float4(o[0], o[1], o[2], o[3]);
            no[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

// Divided TIME attention of one (segment, spatial location, head) per warp (vit_helper.py:56-110 with "b (f n) d -> (b n) f d"):
// the 8 frame tokens of a location attend to [class token; the same 8 tokens].  qkv: fp16 [segments * tokens, 3C] (q | k | v,
// heads 64 apart); out: fp16 [segments * tokens, C].  Lane = (query frame, 16-channel quarter).  Scores are rounded to fp16
// (einsum under autocast), the softmax is fp32, the normalised probabilities are rounded to fp16 for the product with v.
__global__ void __launch_bounds__(128) sync_time_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int segments,
                                                                  int tokens, int n_sp, int H, float scale) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const long long unit = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long total = static_cast<long long>(segments) * n_sp * H;
    if (unit >= total) return;
    const int h = static_cast<int>(unit % H);
    const int n = static_cast<int>((unit / H) % n_sp);
    const long long sgm = unit / (static_cast<long long>(H) * n_sp);
    const int C = H * 64, f = lane >> 2, part = lane & 3;
    const long long base = sgm * tokens;
    const long long qrow = base + 1 + static_cast<long long>(f) * n_sp + n;
    float qv[16];
    {
        const __half* qp = qkv + qrow * 3 * C + h * 64 + part * 16;
        float t[8];
        unpack8h(*reinterpret_cast<const uint4*>(qp), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[j] = t[j];
        unpack8h(*reinterpret_cast<const uint4*>(qp + 8), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[8 + j] = t[j];
    }
    float sc[9];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const long long krow = j == 0 ? base : base + 1 + static_cast<long long>(j - 1) * n_sp + n;
        const __half* kp = qkv + krow * 3 * C + C + h * 64 + part * 16;
        float t[8], d = 0.f;
        unpack8h(*reinterpret_cast<const uint4*>(kp), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) d = fmaf(qv[i], t[i], d);
        unpack8h(*reinterpret_cast<const uint4*>(kp + 8), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) d = fmaf(qv[8 + i], t[i], d);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        sc[j] = f16_round(d * scale);
        mx = fmaxf(mx, sc[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.0f / sum;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const long long krow = j == 0 ? base : base + 1 + static_cast<long long>(j - 1) * n_sp + n;
        const __half* vp = qkv + krow * 3 * C + 2 * C + h * 64 + part * 16;
        const float p = f16_round(sc[j] * inv);
        float t[8];
        unpack8h(*reinterpret_cast<const uint4*>(vp), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, t[i], acc[i]);
        unpack8h(*reinterpret_cast<const uint4*>(vp + 8), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[8 + i] = fmaf(p, t[i], acc[8 + i]);
    }
    __half* op = out + qrow * C + h * 64 + part * 16;
    *reinterpret_cast<uint4*>(op) = make_uint4(pack_f16x2(acc[0], acc[1]), pack_f16x2(acc[2], acc[3]), pack_f16x2(acc[4], acc[5]), pack_f16x2(acc[6], acc[7]));
    *reinterpret_cast<uint4*>(op + 8) = make_uint4(pack_f16x2(acc[8], acc[9]), pack_f16x2(acc[10], acc[11]), pack_f16x2(acc[12], acc[13]), pack_f16x2(acc[14], acc[15]));
}

// Parameter conversions at finalize
__global__ void bf16_to_f16_kernel(const __nv_bfloat16* src, long long n, __half* dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(__bfloat162float(src[i]));
}
// total_pos_embed of MotionFormer.forward_features (video_model_builder.py:185-196, POS_EMBED "separate") with the class token
// folded into row 0: tab[0] = fp32(cls) + fp32(pos[0]); tab[1 + t*n_sp + n] = fp32(bf16(pos[1 + n] + temp[t])).
__global__ void sync_token_table_kernel(const __nv_bfloat16* cls, const __nv_bfloat16* pos, const __nv_bfloat16* temp, int n_sp, int n_t, int C,
                                        float* tab) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = (1LL + static_cast<long long>(n_sp) * n_t) * C;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    const long long tok = i / C;
    if (tok == 0) { tab[i] = __bfloat162float(cls[c]) + __bfloat162float(pos[c]); return; }
    const int n = static_cast<int>((tok - 1) % n_sp), t = static_cast<int>((tok - 1) / n_sp);
    tab[i] = bf16_round(__bfloat162float(pos[static_cast<long long>(1 + n) * C + c]) + __bfloat162float(temp[static_cast<long long>(t) * C + c]));
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* src, long long n, float* dst) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __bfloat162float(src[i]);
}

}  // namespace foley
