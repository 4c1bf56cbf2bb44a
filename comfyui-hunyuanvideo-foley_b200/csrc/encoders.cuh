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
    int round_scores = 0;               // small kernel only: scores and probabilities rounded to bf16 like the bmm + softmax
                                        // path of F.multi_head_attention_forward (need_weights = True)
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

// One CTA = 128 queries of one (sample, head): 8 warps x 16 query rows, 64-key tiles through a 3-deep cp.async ring,
// online softmax in the log2 domain, P rounded to bf16 for the PV product (the row sum taken from the rounded values).
__global__ void __launch_bounds__(32 * EA_NW) enc_attention_kernel(const EncAttnArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t ea_smem[];
    const uint32_t sQ = smem_u32(ea_smem);
    const uint32_t sK0 = sQ + EA_BM * EA_D * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * EA_BM;
    const __nv_bfloat16* Q = a.q + b * a.q_batch_stride + h * EA_D;
    const __nv_bfloat16* K = a.k + b * a.kv_batch_stride + h * EA_D;
    const __nv_bfloat16* V = a.v + b * a.kv_batch_stride + h * EA_D;
    const float scale_log2 = a.scale * 1.4426950408889634f;

    const int n_tiles = (a.Sk + EA_BN - 1) / EA_BN;
    auto load_stage = [&](int t) {
        const uint32_t dst = sK0 + (t % EA_NST) * 2 * EA_TILE;
        ea_load_tile(dst, K, a.kv_row_stride, t * EA_BN, a.Sk, EA_BN);
        ea_load_tile(dst + EA_TILE, V, a.kv_row_stride, t * EA_BN, a.Sk, EA_BN);
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
                mma_bf16_16816(s[2 * jp], qf[kk], b0, b1);
                mma_bf16_16816(s[2 * jp + 1], qf[kk], b2, b3);
            }
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
            const __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
            ls[0] += __low2float(lo) + __high2float(lo);
            ls[1] += __low2float(hi) + __high2float(hi);
            pf[j >> 1][(j & 1) * 2 + 0] = *reinterpret_cast<const uint32_t*>(&lo);
            pf[j >> 1][(j & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
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
                mma_bf16_16816(o[2 * jp], pf[kk], b0, b1);
                mma_bf16_16816(o[2 * jp + 1], pf[kk], b2, b3);
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
    __nv_bfloat16* O = a.o + b * a.o_batch_stride + h * EA_D + (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (row0 < a.Sq)
            *reinterpret_cast<uint32_t*>(O + row0 * a.o_row_stride + j * 8) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
        if (row1 < a.Sq)
            *reinterpret_cast<uint32_t*>(O + row1 * a.o_row_stride + j * 8) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
}

// One warp per (sample, head, query row): scores into shared memory (Sk floats per warp), exact softmax, P = bf16(p / sum),
// out = sum_j P_j v_j.  Serves the CLAP text tower (2 x <= 514 tokens, key-padding mask) and the SigLIP pooling head
// (one probe query over the 1024 patch tokens of every frame).
constexpr int ESA_WARPS = 4;
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
        unpack8(*reinterpret_cast<const uint4*>(Q + c * 8), f);
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
            unpack8(*reinterpret_cast<const uint4*>(kr + c * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) dot = fmaf(qv[c * 8 + j], f[j], dot);
        }
        float s = dot * a.scale;
        if (a.round_scores) s = bf16_round(s);
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
        const float p = bf16_round(sc[key] * inv);
        const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(vp + key * a.kv_row_stride);
        acc0 = fmaf(p, __low2float(v2), acc0);
        acc1 = fmaf(p, __high2float(v2), acc1);
    }
    __nv_bfloat16* O = a.o + b * a.o_batch_stride + qi * a.o_row_stride + h * EA_D + lane * 2;
    *reinterpret_cast<uint32_t*>(O) = pack_bf16x2(acc0, acc1);
}

}  // namespace foley
