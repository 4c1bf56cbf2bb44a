// Fused attention for head_dim 128 (bf16 in, fp32 softmax, bf16 P for the PV product, bf16 out):
// replaces F.scaled_dot_product_attention at attn_layers.py:422 (joint audio+visual self-attention,
// cross-attention to the 77 text tokens) and hifi_foley.py:383 (single-stream self-attention).
// No mask, no dropout, scale 1/sqrt(128).
//
// Sequence lengths here are short (<= ~3500 keys) and attention is < 2 % of the step's FLOPs at 5 s,
// so this is a compact flash-style kernel on mma.sync.m16n8k16 with cp.async-staged, XOR-swizzled
// shared-memory tiles: one CTA = 64 queries of one (batch, head), 4 warps x 16 query rows, 64-key tiles.
// Q, K, V: [B, H, S, 128] bf16 (written by qk_norm_rope_kernel); O: [B, Sq, H*128] bf16 token-major so it is
// directly the A operand of the following projection GEMM.
#pragma once
#include "ptx.cuh"

namespace foley {

struct AttnArgs {
    const __nv_bfloat16* q = nullptr;
    const __nv_bfloat16* k = nullptr;
    const __nv_bfloat16* v = nullptr;
    __nv_bfloat16* o = nullptr;
    int H = 0, Sq = 0, Sk = 0;
    long long q_batch_stride = 0, q_head_stride = 0;    // elements
    long long kv_batch_stride = 0, kv_head_stride = 0;
    long long o_batch_stride = 0;                       // elements; row stride is H*128
    const int* kv_batch_map = nullptr;                  // kv batch of query batch b (cross-attn: cond of group)
    const int* grp_of_sample = nullptr;                 // optional second-level map: kv = map[grp_of_sample[b]]
    float scale_log2 = 0.f;                             // scale * log2(e)
};

constexpr int ATT_BN = 64, ATT_D = 128;
constexpr int ATT_TILE = ATT_BN * ATT_D * 2;                   // 16 KB
// Two shapes: 4 warps x 16 query rows with a 4-deep K/V ring (small grids: more CTAs, deeper prefetch) and 8 warps
// x 16 rows with a 3-deep ring (large batches: twice the warps per SM, K/V tiles shared by twice the queries).
// KVS = 2: the warps form two groups that share the 16*NW/2 query rows and take alternate 64-key tiles (in-CTA split-KV;
// the two partial softmax states are merged through shared memory at the end): halves the serial chain of tile
// iterations, which is what bounds the small grids of one 5 s clip (110 CTAs x 5 tiles, one dependent chain per warp).
template <int NW, int NST, int KVS = 1>
struct AttCfg {
    static constexpr int BM = 16 * NW / KVS;
    static constexpr int SMEM = BM * ATT_D * 2 + NST * KVS * 2 * ATT_TILE;
};

__device__ __forceinline__ uint32_t swz(int row, int chunk) {  // byte offset of a 16-byte chunk in a [rows][128] bf16 tile
    return static_cast<uint32_t>(row * 256 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Loads `rows` x 128 bf16 rows [r0, r0+rows) of a [S,128] matrix into a swizzled tile; rows >= S are zero.
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* g, int r0, int S, int rows) {
    for (int i = threadIdx.x; i < rows * 16; i += blockDim.x) {
        const int r = i >> 4, c = i & 15;
        const bool ok = (r0 + r) < S;
        cp_async16(smem_base + swz(r, c), g + static_cast<long long>(ok ? r0 + r : 0) * ATT_D + c * 8, ok);
    }
}

template <int NW, int NST, int KVS = 1>
__global__ void __launch_bounds__(32 * NW) attention_kernel(const AttnArgs a) {
    constexpr int ATT_BM = AttCfg<NW, NST, KVS>::BM;
    constexpr int ATT_STAGES = NST;
    constexpr int STAGE_BYTES = KVS * 2 * ATT_TILE;   // one ring stage = KVS consecutive key tiles (K and V each)
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t att_smem[];
    const uint32_t sQ = smem_u32(att_smem);
    const uint32_t sK0 = sQ + ATT_BM * ATT_D * 2;
    const int warp_all = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp_all / (NW / KVS);            // which key tile of a stage this warp takes
    const int warp = warp_all - grp * (NW / KVS);     // which 16 query rows
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_BM;
    int kb = b;
    if (a.kv_batch_map) kb = a.kv_batch_map[a.grp_of_sample ? a.grp_of_sample[b] : b];
    const __nv_bfloat16* Q = a.q + b * a.q_batch_stride + h * a.q_head_stride;
    const __nv_bfloat16* K = a.k + kb * a.kv_batch_stride + h * a.kv_head_stride;
    const __nv_bfloat16* V = a.v + kb * a.kv_batch_stride + h * a.kv_head_stride;

    // K/V tiles stream through a 4-deep cp.async ring: one exposed L2 latency per CTA instead of one per tile
    // (the first version loaded, waited and computed tile by tile: 14 us per call at S = 290).
    const int n_tiles = (a.Sk + ATT_BN - 1) / ATT_BN;
    const int n_iters = (n_tiles + KVS - 1) / KVS;
    auto load_stage = [&](int it) {   // key tiles it*KVS .. it*KVS+KVS-1 (tiles past the end are zero-filled and masked)
        const uint32_t dst = sK0 + (it % ATT_STAGES) * STAGE_BYTES;
#pragma unroll
        for (int g = 0; g < KVS; ++g) {
            const int tile = it * KVS + g;
            if (tile < n_tiles) {
                load_tile(dst + g * 2 * ATT_TILE, K, tile * ATT_BN, a.Sk, ATT_BN);
                load_tile(dst + g * 2 * ATT_TILE + ATT_TILE, V, tile * ATT_BN, a.Sk, ATT_BN);
            }
        }
    };
    load_tile(sQ, Q, q0, a.Sq, ATT_BM);
    cp_async_commit();
    for (int t = 0; t < ATT_STAGES - 1; ++t) {
        if (t < n_iters) load_stage(t);
        cp_async_commit();
    }

    uint32_t qf[8][4];
    float o[16][4];
#pragma unroll
    for (int j = 0; j < 16; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    for (int it = 0; it < n_iters; ++it) {
        const int t = it * KVS + grp;        // this warp group's key tile
        const int k0 = t * ATT_BN;
        {
            const int tn = it + ATT_STAGES - 1;   // refill the stage consumed in the previous iteration
            if (tn < n_iters) load_stage(tn);
            cp_async_commit();
        }
        cp_async_wait_group<ATT_STAGES - 1>();   // Q and stage `it` have landed
        __syncthreads();
        const uint32_t sK = sK0 + (it % ATT_STAGES) * STAGE_BYTES + grp * 2 * ATT_TILE;
        const uint32_t sV = sK + ATT_TILE;
        if (it == 0) {
            // Q fragments for this warp's 16 rows: 8 k-steps x 4 regs
            const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
                ldsm_x4(sQ + swz(row, kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
        }

        if (t < n_tiles) {
            // S = Q K^T : 16 x 64 per warp = 8 n-tiles
            float s[8][4];
    #pragma unroll
            for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
    #pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
    #pragma unroll
                for (int jp = 0; jp < 4; ++jp) {  // pairs of key n-tiles
                    uint32_t b0, b1, b2, b3;
                    const int krow = jp * 16 + (lane & 7) + (lane >> 4) * 8;
                    ldsm_x4(sK + swz(krow, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
                    mma_bf16_16816(s[2 * jp], qf[kk], b0, b1);
                    mma_bf16_16816(s[2 * jp + 1], qf[kk], b2, b3);
                }
            }
            // mask keys beyond Sk, scale into log2 domain, online softmax
            float mx[2] = {-INFINITY, -INFINITY};
    #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = k0 + j * 8 + (lane & 3) * 2;
    #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool ok = (key + (e & 1)) < a.Sk;
                    s[j][e] = ok ? s[j][e] * a.scale_log2 : -INFINITY;
                    mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
                }
            }
            float corr[2];
    #pragma unroll
            for (int r = 0; r < 2; ++r) {
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
                const float m_new = fmaxf(m_run[r], mx[r]);
                corr[r] = exp2f(m_run[r] - m_new);
                m_run[r] = m_new;
                l_run[r] *= corr[r];
            }
            uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
            float ls[2] = {0.f, 0.f};
    #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p0 = exp2f(s[j][0] - m_run[0]), p1 = exp2f(s[j][1] - m_run[0]);
                const float p2 = exp2f(s[j][2] - m_run[1]), p3 = exp2f(s[j][3] - m_run[1]);
                // the PV product consumes bf16 P; accumulate the row sum from the same rounded values
                const __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
                ls[0] += __low2float(lo) + __high2float(lo);
                ls[1] += __low2float(hi) + __high2float(hi);
                pf[j >> 1][(j & 1) * 2 + 0] = *reinterpret_cast<const uint32_t*>(&lo);
                pf[j >> 1][(j & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
            }
            l_run[0] += ls[0];
            l_run[1] += ls[1];
    #pragma unroll
            for (int j = 0; j < 16; ++j) {
                o[j][0] *= corr[0]; o[j][1] *= corr[0];
                o[j][2] *= corr[1]; o[j][3] *= corr[1];
            }
            // O += P V : 16 d n-tiles, 4 k-steps over the 64 keys
    #pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
    #pragma unroll
                for (int jp = 0; jp < 8; ++jp) {  // pairs of d n-tiles
                    uint32_t b0, b1, b2, b3;
                    const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    ldsm_x4_t(sV + swz(vrow, jp * 2 + (lane >> 4)), b0, b1, b2, b3);
                    mma_bf16_16816(o[2 * jp], pf[kk], b0, b1);
                    mma_bf16_16816(o[2 * jp + 1], pf[kk], b2, b3);
                }
            }
        }
        __syncthreads();   // every warp is done with this stage before it is refilled
    }

    if constexpr (KVS == 2) {
        // merge the two groups' online-softmax states: group 1 parks (m, l, O) in the idle K/V ring, group 0 folds it in
        float* mg = reinterpret_cast<float*>(att_smem + ATT_BM * ATT_D * 2) + (warp * 32 + lane) * 68;
        if (grp == 1) {
            mg[0] = m_run[0]; mg[1] = m_run[1]; mg[2] = l_run[0]; mg[3] = l_run[1];
#pragma unroll
            for (int j = 0; j < 16; ++j) { mg[4 + 4 * j] = o[j][0]; mg[5 + 4 * j] = o[j][1]; mg[6 + 4 * j] = o[j][2]; mg[7 + 4 * j] = o[j][3]; }
        }
        __syncthreads();
        if (grp == 1) return;
        float sc0[2], sc1[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float m1 = mg[r], m = fmaxf(m_run[r], m1);
            sc0[r] = exp2f(m_run[r] - m);
            sc1[r] = exp2f(m1 - m);          // group 1 without a tile: m1 = -inf -> 0
            l_run[r] = l_run[r] * sc0[r] + mg[2 + r] * sc1[r];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            o[j][0] = o[j][0] * sc0[0] + mg[4 + 4 * j] * sc1[0];
            o[j][1] = o[j][1] * sc0[0] + mg[5 + 4 * j] * sc1[0];
            o[j][2] = o[j][2] * sc0[1] + mg[6 + 4 * j] * sc1[1];
            o[j][3] = o[j][3] * sc0[1] + mg[7 + 4 * j] * sc1[1];
        }
    }
    // finalize: divide by the row sums (reduced over the quad) and store bf16
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    const int row0 = q0 + warp * 16 + (lane >> 2), row1 = row0 + 8;
    __nv_bfloat16* O = a.o + b * a.o_batch_stride + h * ATT_D + (lane & 3) * 2;
    const long long ld = static_cast<long long>(a.H) * ATT_D;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (row0 < a.Sq)
            *reinterpret_cast<uint32_t*>(O + row0 * ld + j * 8) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
        if (row1 < a.Sq)
            *reinterpret_cast<uint32_t*>(O + row1 * ld + j * 8) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
}

}  // namespace foley
