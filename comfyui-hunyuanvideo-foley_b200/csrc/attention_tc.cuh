// tcgen05 / TMEM attention for head_dim 128 with the q/k RMSNorm + RoPE of the producing projection folded into the
// operand load: replaces, per call, qk_norm_rope_kernel + attention_kernel (attn_layers.py:323-456 joint self-attention
// and cross-attention to the text tokens, hifi_foley.py:364-390 single-stream self-attention).
//
// One CTA = 128 query rows of one (sample, head), 12 warps in three roles (softmax, O accumulation, copy / MMA issue).  Per chunk of keys (the whole sequence when it has <= 320 keys
// — every shape of a 5 s clip — else 128 keys at a time, double-buffered, with the online-softmax rescale):
//   load   single threads issue TMA boxes (128-160 rows x 64 channels, 128-byte swizzle = the canonical tcgen05 operand layout)
//          straight from the projection GEMM's bf16 output (or from prepared [B,H,S,128] tensors); rows past the end of
//          the sequence are zero-filled by the copy engine;
//   norm   (only when a norm weight is given) all warps RMS-normalise and rotate the landed Q / K rows in place
//          (16 threads per row, 8 channels each: both reference RMSNorm flavours, interleaved RoPE);
//   S      = Q K^T   tcgen05.mma (M = 128, N = keys <= 256 + rest, K = 128) into TMEM columns [0, 320);
//   P      softmax: a thread owns a row of S (tcgen05.ld 32x32b), four warps per TMEM lane quarter split the columns,
//          exact row max (no rescale inside a chunk), exp2 in the scaled log2 domain, P rounded to bf16 into shared
//          memory (K-major A operand; in the one-chunk mode it aliases the K tile, which is dead by then);
//   O      = P V     tcgen05.mma with V as an MN-major B operand (keys x d rows exactly as loaded: no transpose pass)
//          into TMEM columns [384, 512), read back once per chunk and accumulated in registers;
//   store  the bf16 tile goes through shared memory so that global stores are whole 256-byte row segments.
#pragma once
#include "gemm_host.cuh"
#include "ptx.cuh"

namespace foley {

struct AttNorm {                           // rows [0, rows0) of the sequence use slot 0, the rest slot 1 (joint attention:
    int rows0 = 0;                         // visual tokens, then audio tokens)
    const __nv_bfloat16* w[2] = {nullptr, nullptr};   // [128] RMSNorm weight: normalise + rotate; nullptr -> rows are used as stored
    const float2* rope[2] = {nullptr, nullptr};       // [rows][64] (cos, sin) of every rotation pair, indexed by the row inside its group
};
struct AttOperand {                        // element (b, h, r, d) at ptr + b*batch_stride + h*head_stride + r*row_stride + d
    const __nv_bfloat16* ptr = nullptr;
    long long batch_stride = 0, head_stride = 0, row_stride = 0;
    int rows = 0, heads = 0, batch = 0;
};
struct AttTcArgs {
    AttNorm qn, kn;
    __nv_bfloat16* o = nullptr;            // [B, Sq, H*128] token-major
    long long o_batch_stride = 0;
    int H = 0, Sq = 0, Sk = 0;
    const int* kv_batch_map = nullptr;     // kv sample of query sample b (cross-attention: condition of the group)
    const int* grp_of_sample = nullptr;    // optional second level: kv = map[grp_of_sample[b]]
    float scale_log2 = 0.f;                // softmax scale * log2(e)
    int norm_kind = 0;                     // 0: bf16r(bf16r(x*rstd)*w) (norm_layers.py:49-51); 1: nn.RMSNorm on bf16, single rounding
    float eps = 1e-6f;
    int probe_chunk = -1;                  // >= 0: thread 0 of CTA (0,0,0) stamps clock64 around the phases of this chunk (foley_debug_times)
};

constexpr int ATC_CK = 320;                          // short mode: the whole sequence (<= 320 keys) is one resident chunk
constexpr int ATC_CK_LONG = 128;                     // long mode: 128-key chunks, double-buffered (the next chunk lands under this one's math)
constexpr int ATC_CS = 2;                            // softmax column shares: softmax warps per TMEM lane quarter
constexpr int ATC_SOFTMAX = 128 * ATC_CS;            // warps 0..7: S -> P (a thread owns a row and 1/ATC_CS of its columns)
constexpr int ATC_CORR = 128;                        // warps 8..11: O accumulation (a thread owns a row and all 128 channels);
                                                     // lane 0 of warp 8 also issues every TMA copy and every MMA
constexpr int ATC_WORKERS = ATC_SOFTMAX;             // threads of the in-place norm pass
constexpr int ATC_THREADS = ATC_SOFTMAX + ATC_CORR;  // 384: register files are handed out per 4 warps — 168 registers each
constexpr int ATC_BOX_Q = 128;                       // rows per TMA box: the Q tile is one box per d half,
constexpr int ATC_BOX_KV_SHORT = 160;                // a resident K / V tile at most two (a TMA instruction costs its issuing
constexpr int ATC_BOX_KV_LONG = 128;                 // thread ~60 ns: few big boxes), a 128-key chunk one
constexpr int ATC_Q_BYTES = 2 * 128 * 128;           // two 64-wide d halves of [128 rows x 128 B]
constexpr int ATC_KV_BYTES = 2 * ATC_CK * 128;       // short mode: two d halves of [320 rows x 128 B]; P ([128 x 320] bf16) aliases the K tile
constexpr int ATC_LONG_STAGE = 2 * (2 * ATC_CK_LONG * 128);   // long mode: K + V of one 128-key chunk = 64 KB; P behind the two stages
constexpr int ATC_O_COL = 384;                       // TMEM: S columns [0, 320), O columns [384, 512)
constexpr int ATC_O_PITCH = 272;                     // output staging: 256 B per row + 16 B (conflict-free both ways)
constexpr int ATC_SMEM = 1024 + ATC_Q_BYTES + 2 * ATC_KV_BYTES + 12 * 128 * 4 + 128;   // (barriers: 9 x 8 bytes + the TMEM slot)

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t atc_tile_addr(uint32_t base, int rows_cap, int r, int sub) {
    return base + static_cast<uint32_t>(sub >> 3) * static_cast<uint32_t>(rows_cap * 128) + static_cast<uint32_t>(r) * 128u +
           ((static_cast<uint32_t>(sub & 7) ^ static_cast<uint32_t>(r & 7)) << 4);
}

// One thread: TMA boxes of `box` rows covering rows [row0, row0 + n_rows) (whole boxes are copied, rows past the tensor's
// extent arrive as zeros) of (head h, sample b) into both d halves of a tile.  Returns the bytes that will land.
__device__ __forceinline__ uint32_t atc_tile_bytes(int n_rows, int box) { return static_cast<uint32_t>((n_rows + box - 1) / box) * 2u * box * 128u; }
__device__ __forceinline__ void atc_issue_tile(const CUtensorMap* tm, uint64_t* bar, int b, int h, int row0, int n_rows, int box,
                                               uint32_t smem_base, int rows_cap) {
    for (int r = 0; r < n_rows; r += box) {
#pragma unroll
        for (int half = 0; half < 2; ++half)
            tma_load_4d(smem_base + half * (rows_cap * 128) + r * 128, tm, bar, half * 64, row0 + r, h, b);
    }
}

// In-place RMSNorm + RoPE of the landed rows that carry a norm weight (16 threads per row, 8 channels = 4 rotation pairs
// each; the RMS statistics take four shuffle steps).  The (cos, sin) pairs of up to ATC_NB pieces are fetched before any
// of them is consumed.
constexpr int ATC_NB = 5;
__device__ __forceinline__ void atc_norm_rows(const AttNorm& nm, int row0, int n_rows, int total_rows, uint32_t smem_base,
                                              int rows_cap, int norm_kind, float eps) {
    if (nm.w[0] == nullptr && nm.w[1] == nullptr) return;
    const int sub = threadIdx.x & 15;
    uint4 wr0 = make_uint4(0u, 0u, 0u, 0u), wr1 = wr0;
    if (nm.w[0]) wr0 = *reinterpret_cast<const uint4*>(nm.w[0] + sub * 8);
    if (nm.w[1]) wr1 = *reinterpret_cast<const uint4*>(nm.w[1] + sub * 8);
    constexpr int RSTEP = ATC_WORKERS / 16;
    for (int rb = threadIdx.x >> 4; rb < n_rows; rb += RSTEP * ATC_NB) {   // (trip counts are warp-uniform: n_rows % 16 == 0)
        float4 t0[ATC_NB], t1[ATC_NB];
        bool normed[ATC_NB], seg1[ATC_NB];
#pragma unroll
        for (int u = 0; u < ATC_NB; ++u) {
            const int r = rb + u * RSTEP, g = row0 + r;
            const bool s1 = g >= nm.rows0;
            const int l = s1 ? g - nm.rows0 : g;
            const float2* rope = s1 ? nm.rope[1] : nm.rope[0];
            seg1[u] = s1;
            normed[u] = r < n_rows && g < total_rows && (s1 ? nm.w[1] : nm.w[0]) != nullptr;
            t0[u] = t1[u] = make_float4(1.f, 0.f, 1.f, 0.f);
            if (normed[u]) {
                const float4* tp = reinterpret_cast<const float4*>(rope + static_cast<long long>(l) * 64 + sub * 4);
                t0[u] = tp[0];
                t1[u] = tp[1];
            }
        }
#pragma unroll
        for (int u = 0; u < ATC_NB; ++u) {
            const int r = rb + u * RSTEP;
            if (r >= n_rows) break;                     // warp-uniform
            const uint32_t addr = atc_tile_addr(smem_base, rows_cap, r, sub);
            const uint4 raw = ld_shared_v4(addr);
            float v[8];
            {
                const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int j = 0; j < 4; ++j) { v[2 * j] = __low2float(p2[j]); v[2 * j + 1] = __high2float(p2[j]); }
            }
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
            // every lane takes part (the two rows of a warp may sit in different groups / past the end)
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (normed[u]) {
                const float rstd = rsqrtf(ss * (1.0f / 128.0f) + eps);
                const uint4 wr = seg1[u] ? wr1 : wr0;
                const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&wr);
                const float cs[4] = {t0[u].x, t0[u].z, t1[u].x, t1[u].z};
                const float sn[4] = {t0[u].y, t0[u].w, t1[u].y, t1[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float w0 = __low2float(w2[j]), w1 = __high2float(w2[j]);
                    float x0, x1;
                    if (norm_kind == 0) {
                        x0 = bf16_round(bf16_round(v[2 * j] * rstd) * w0);
                        x1 = bf16_round(bf16_round(v[2 * j + 1] * rstd) * w1);
                    } else {
                        x0 = bf16_round(__fmul_rn(__fmul_rn(v[2 * j], rstd), w0));
                        x1 = bf16_round(__fmul_rn(__fmul_rn(v[2 * j + 1], rstd), w1));
                    }
                    // (x0, x1) -> (x0*cos - x1*sin, x1*cos + x0*sin)   (attn_layers.py:112-148), products and sums rounded separately
                    v[2 * j] = __fadd_rn(__fmul_rn(x0, cs[j]), __fmul_rn(-x1, sn[j]));
                    v[2 * j + 1] = __fadd_rn(__fmul_rn(x1, cs[j]), __fmul_rn(x0, sn[j]));
                }
                st_shared_v4(addr, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
        }
    }
}

// MN-major shared-memory descriptor (B operand = V stored as [keys][64 d] rows of 128 bytes, 128-byte swizzle):
// LBO = distance between the two 64-wide d blocks, SBO = distance between 8-key groups.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// exponentials of one 32-column group of a row (scaled log2 domain), packed to bf16 pairs; returns the fp32 row-sum share
__device__ __forceinline__ float atc_exps(const uint32_t (&v)[32], uint32_t (&pk)[16], int c0, int kn, float scale_log2, float m_new) {
    float l = 0.f;
    if (c0 + 32 <= kn) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2_fast(fmaf(__uint_as_float(v[j]), scale_log2, -m_new));
            const float p1 = ex2_fast(fmaf(__uint_as_float(v[j + 1]), scale_log2, -m_new));
            l += p0 + p1;
            pk[j >> 1] = pack_bf16x2(p0, p1);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const float p0 = (c0 + j) < kn ? ex2_fast(fmaf(__uint_as_float(v[j]), scale_log2, -m_new)) : 0.f;
            const float p1 = (c0 + j + 1) < kn ? ex2_fast(fmaf(__uint_as_float(v[j + 1]), scale_log2, -m_new)) : 0.f;
            l += p0 + p1;
            pk[j >> 1] = pack_bf16x2(p0, p1);
        }
    }
    return l;
}
// o_acc = o_acc * corr + O chunk read back from TMEM (this thread's channels)
template <int N>
__device__ __forceinline__ void atc_take_o(float (&o_acc)[N], uint32_t taddr, float corr) {
#pragma unroll
    for (int g = 0; g < N / 32; ++g) {
        uint32_t ov[32];
        tmem_ld_32x32(taddr + g * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o_acc[g * 32 + j] = fmaf(o_acc[g * 32 + j], corr, __uint_as_float(ov[j]));
    }
}
// 32 bf16 of row `row`, columns [c0, c0 + 32), into the K-major swizzled P tile
__device__ __forceinline__ void atc_put(const uint32_t (&pk)[16], uint32_t sP, int row, int c0) {
    const uint32_t blk = static_cast<uint32_t>(c0 >> 6), ch0 = static_cast<uint32_t>((c0 & 63) >> 3);
    const uint32_t base = sP + blk * 16384u + static_cast<uint32_t>(row) * 128u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        st_shared_v4(base + (((ch0 + j) ^ static_cast<uint32_t>(row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const AttTcArgs a) {
    extern __shared__ uint8_t atc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sR = sQ + ATC_Q_BYTES;           // K / V / P region (160 KB)
    float* red_all = reinterpret_cast<float*>(smem + ATC_Q_BYTES + 2 * ATC_KV_BYTES);   // [2 chunks][2 shares][128] partial row maxima,
    float* red_sum = red_all + 512;                 // [2 shares][128] row-sum shares,
    float* corr_s = red_all + 768;                  // [4 chunks][128] correction factor of chunk c for every row
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATC_Q_BYTES + 2 * ATC_KV_BYTES + 12 * 128 * 4);
    uint64_t* full_k = bars;            // [2]  K (+ Q) of a stage has landed
    uint64_t* full_v = bars + 2;        // [2]  V of a stage has landed
    uint64_t* bar_s = bars + 4;         // [2]  S(c) is complete in TMEM buffer c & 1
    uint64_t* bar_o = bars + 6;         //      PV(c) is complete (O chunk in TMEM; P and the V tile are free)
    uint64_t* p_ready = bars + 7;       //      every softmax thread has written its share of P(c) (and has read S(c))
    uint64_t* o_free = bars + 8;        //      every accumulation thread has read the O chunk
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool softmax_role = warp < ATC_SOFTMAX / 32;
    const bool control = threadIdx.x == ATC_SOFTMAX;      // lane 0 of the first accumulation warp
    const int qq = warp & 3;            // TMEM lane quarter
    const int cs = (warp >> 2) & (ATC_CS - 1);   // softmax: which share of the columns
    const int row = qq * 32 + lane;     // query row of the tile (= TMEM lane)
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;

    const bool probe0 = a.probe_chunk >= 0 && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    if (probe0) { g_foley_times[0] = clock64(); g_foley_times[15] = globaltimer_ns(); }
    const bool long_mode = a.Sk > ATC_CK;
    const int ck = long_mode ? ATC_CK_LONG : a.Sk;
    const int cap = long_mode ? ATC_CK_LONG : ATC_CK;      // rows per d-half block of a K / V tile
    const int n_chunks = (a.Sk + ck - 1) / ck;
    const uint32_t sP = long_mode ? sR + 2 * ATC_LONG_STAGE : sR;   // short mode: P aliases the (dead) K tile
    const uint32_t s_col = long_mode ? 128u : 0u;          // long mode: S(c) lives in TMEM columns [(c & 1) * 128, +128)
    auto k_tile = [&](int c) { return long_mode ? sR + static_cast<uint32_t>(c & 1) * ATC_LONG_STAGE : sR; };
    auto v_tile = [&](int c) { return long_mode ? sR + static_cast<uint32_t>(c & 1) * ATC_LONG_STAGE + 2 * ATC_CK_LONG * 128 : sR + ATC_KV_BYTES; };
    auto chunk_rows = [&](int c) { return (min(ck, a.Sk - c * ck) + 15) & ~15; };
    const bool k_normed = a.kn.w[0] != nullptr || a.kn.w[1] != nullptr;

    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (threadIdx.x == 32) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        for (int i = 0; i < 2; ++i) { mbar_init(&full_k[i], 1); mbar_init(&full_v[i], 1); mbar_init(&bar_s[i], 1); }
        mbar_init(bar_o, 1);
        mbar_init(p_ready, ATC_SOFTMAX);
        mbar_init(o_free, ATC_CORR);
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(qq * 32) << 16);

    int kb = b;
    pdl_wait();
    pdl_trigger();
    if (a.kv_batch_map) kb = a.kv_batch_map[a.grp_of_sample ? a.grp_of_sample[b] : b];
    if (probe0) g_foley_times[1] = clock64();

    if (softmax_role) {
        // ================================================================== 8 softmax warps: (norm,) S -> P
        float m_run = -INFINITY, l_run = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            float* red = red_all + (c & 1) * 256;     // (double-buffered: a fast thread's chunk c+1 never meets a slow reader of chunk c)
            auto row_max = [&]() {   // max over the shares' partial maxima of this row
                float m = red[row];
#pragma unroll
                for (int i = 1; i < ATC_CS; ++i) m = fmaxf(m, red[i * 128 + row]);
                return m;
            };
            const int key0 = c * ck;
            const int kn = min(ck, a.Sk - key0);
            const int kpad = (kn + 15) & ~15;
            const bool probe = probe0 && c == a.probe_chunk;
            if (probe) g_foley_times[2] = clock64();
            if (c == 0 || k_normed) {
                mbar_wait(&full_k[c & 1], (c >> 1) & 1, 0x610 + c);
                if (probe) g_foley_times[3] = clock64();
                if (c == 0) atc_norm_rows(a.qn, q0, 128, a.Sq, sQ, 128, a.norm_kind, a.eps);
                atc_norm_rows(a.kn, key0, kpad, a.Sk, k_tile(c), cap, a.norm_kind, a.eps);
                fence_proxy_async();
                asm volatile("bar.sync 2, %0;" ::"n"(ATC_SOFTMAX + 32) : "memory");   // + the control warp
            }
            if (probe) g_foley_times[4] = clock64();
            mbar_wait(&bar_s[c & 1], (c >> 1) & 1, 0x600 + c);
            tc_fence_after();
            if (probe) g_foley_times[5] = clock64();

            // ---- softmax of this thread's row over its share of the 32-column groups
            const uint32_t s_row = t_row + static_cast<uint32_t>(c & 1) * s_col;
            const int n_grp = (kpad + 31) >> 5;
            const int g_lo = (n_grp * cs) / ATC_CS, g_hi = (n_grp * (cs + 1)) / ATC_CS;
            float l_chunk = 0.f;
            float corr;
            if (long_mode) {
                // at most two groups (128 keys / 32 / ATC_CS): S is read once and stays in registers; the exponentials are
                // computed before PV(c-1) is waited for, P is written after it (one P buffer)
                uint32_t va[32], vb[32], pka[16], pkb[16];
                const bool has_a = g_lo < g_hi, has_b = g_lo + 1 < g_hi;
                float mx = -INFINITY;
                if (has_a) tmem_ld_32x32(s_row + g_lo * 32, va);
                if (has_b) tmem_ld_32x32(s_row + (g_lo + 1) * 32, vb);
                tmem_ld_wait();
                if (has_a) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (g_lo * 32 + j) < kn ? __uint_as_float(va[j]) : -INFINITY);
                }
                if (has_b) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, ((g_lo + 1) * 32 + j) < kn ? __uint_as_float(vb[j]) : -INFINITY);
                }
                red[cs * 128 + row] = mx;
                asm volatile("bar.sync 1, %0;" ::"n"(ATC_SOFTMAX) : "memory");
                if (probe) g_foley_times[6] = clock64();
                const float m_new = fmaxf(m_run, row_max() * a.scale_log2);
                corr = ex2_fast(m_run - m_new);     // first chunk: exp2(-inf) = 0
                m_run = m_new;
                if (cs == 0) corr_s[(c & 3) * 128 + row] = corr;    // for the accumulation warps (read after PV(c) completes)
                if (has_a) l_chunk += atc_exps(va, pka, g_lo * 32, kn, a.scale_log2, m_new);
                if (has_b) l_chunk += atc_exps(vb, pkb, (g_lo + 1) * 32, kn, a.scale_log2, m_new);
                if (c > 0) mbar_wait(bar_o, (c - 1) & 1, 0x680 + c);   // PV(c-1), which ran under the lines above, frees the P buffer
                if (has_a) atc_put(pka, sP, row, g_lo * 32);
                if (has_b) atc_put(pkb, sP, row, (g_lo + 1) * 32);
            } else {
                // one resident chunk of up to 320 keys: two passes over S (max, then exponentials), no rescale at all
                uint32_t v[32], pk[16];
                float mx = -INFINITY;
                for (int g = g_lo; g < g_hi; ++g) {
                    tmem_ld_32x32(s_row + g * 32, v);
                    tmem_ld_wait();
                    const int c0 = g * 32;
                    if (c0 + 32 <= kn) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c0 + j) < kn ? __uint_as_float(v[j]) : -INFINITY);
                    }
                }
                red[cs * 128 + row] = mx;
                asm volatile("bar.sync 1, %0;" ::"n"(ATC_SOFTMAX) : "memory");
                if (probe) g_foley_times[6] = clock64();
                const float m_new = fmaxf(m_run, row_max() * a.scale_log2);
                corr = ex2_fast(m_run - m_new);
                m_run = m_new;
                if (cs == 0) corr_s[(c & 3) * 128 + row] = corr;
                for (int g = g_lo; g < g_hi; ++g) {
                    tmem_ld_32x32(s_row + g * 32, v);
                    tmem_ld_wait();
                    l_chunk += atc_exps(v, pk, g * 32, kn, a.scale_log2, m_new);
                    atc_put(pk, sP, row, g * 32);
                }
            }
            l_run = l_run * corr + l_chunk;
            if (c + 1 == n_chunks) red_sum[cs * 128 + row] = l_run;   // final row-sum shares (read behind barrier 3)
            if (probe) g_foley_times[7] = clock64();
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(p_ready);
            if (probe) g_foley_times[8] = clock64();
        }
    } else {
        // ================================================================== 4 accumulation warps (+ copy / MMA issue by lane 0 of warp 8)
        const int box_kv = long_mode ? ATC_BOX_KV_LONG : ATC_BOX_KV_SHORT;
        auto issue_k = [&](int c) {         // K of chunk c (with Q for the first chunk) -> stage c & 1
            const int st = c & 1, rows = chunk_rows(c);
            mbar_expect_tx(&full_k[st], atc_tile_bytes(rows, box_kv) + (c == 0 ? atc_tile_bytes(128, ATC_BOX_Q) : 0u));
            if (c == 0) atc_issue_tile(&tm_q, &full_k[0], b, h, q0, 128, ATC_BOX_Q, sQ, 128);
            atc_issue_tile(&tm_k, &full_k[st], kb, h, c * ck, rows, box_kv, k_tile(c), cap);
        };
        auto issue_v = [&](int c) {
            const int st = c & 1, rows = chunk_rows(c);
            mbar_expect_tx(&full_v[st], atc_tile_bytes(rows, box_kv));
            atc_issue_tile(&tm_v, &full_v[st], kb, h, c * ck, rows, box_kv, v_tile(c), cap);
        };
        auto issue_s = [&](int c) {         // S(c) = Q K(c)^T; more than 256 keys: two MMAs of about half the columns each
            const int kpad = chunk_rows(c);     // (an MMA costs max(69, N/2) cycles: 160 + 144 columns beat 256 + 48)
            const int n_first = kpad <= 256 ? kpad : ((kpad + 31) >> 5) << 4;
            const uint32_t d0 = tmem_base + static_cast<uint32_t>(c & 1) * s_col;
            for (int n_off = 0; n_off < kpad; n_off += n_first) {
                const int n = min(n_first, kpad - n_off);
                const uint32_t idesc = make_idesc(1, 128, n);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t a_desc = make_smem_desc_sw128(sQ + (ks >> 2) * 16384 + (ks & 3) * 32);
                    const uint64_t b_desc = make_smem_desc_sw128(k_tile(c) + (ks >> 2) * (cap * 128) + n_off * 128 + (ks & 3) * 32);
                    umma_bf16(d0 + n_off, a_desc, b_desc, idesc, ks != 0);
                }
            }
            umma_commit(&bar_s[c & 1]);
        };
        if (control) {
            issue_k(0);
            issue_v(0);
            if (n_chunks > 1) { issue_k(1); issue_v(1); }
        }
        float o_acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) o_acc[j] = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            if (warp == ATC_SOFTMAX / 32) {
                if (c == 0 || k_normed) {
                    // Q / K(c) are normalised in place by the softmax warps first (or this is the first chunk): they meet this warp here
                    asm volatile("bar.sync 2, %0;" ::"n"(ATC_SOFTMAX + 32) : "memory");
                    if (control) {
                        tc_fence_after();
                        mbar_wait(&full_k[c & 1], (c >> 1) & 1, 0x610 + c);
                        issue_s(c);
                        if (c == 0 && !k_normed && n_chunks > 1) {      // K needs no norm pass: S runs two chunks ahead
                            mbar_wait(&full_k[1], 0, 0x611);
                            tc_fence_after();
                            issue_s(1);
                        }
                    }
                }
                if (control) {
                    const int kpad = chunk_rows(c);
                    // refills first (nothing here blocks): V(c+1) into the stage PV(c-1) has released (this warp waited on it in
                    // the previous iteration), K(c+2) into the stage S(c) has released
                    if (c >= 1 && c + 1 < n_chunks) issue_v(c + 1);
                    if (c + 2 < n_chunks) {
                        mbar_wait(&bar_s[c & 1], (c >> 1) & 1, 0x670 + c);
                        issue_k(c + 2);
                    }
                    mbar_wait(p_ready, c & 1, 0x640 + c);               // P(c) written, S(c) consumed by every softmax thread
                    if (c > 0) mbar_wait(o_free, (c - 1) & 1, 0x660 + c);   // O(c-1) read by every accumulation thread
                    mbar_wait(&full_v[c & 1], (c >> 1) & 1, 0x620 + c);
                    tc_fence_after();
                    // ---- O_chunk = P V: K = kpad keys in steps of 16, N = 128 (d), V MN-major
                    const uint32_t idesc = make_idesc(1, 128, 128) | (1u << 16);
                    for (int kk = 0; kk < (kpad >> 4); ++kk) {
                        const uint64_t a_desc = make_smem_desc_sw128(sP + (kk >> 2) * 16384 + (kk & 3) * 32);
                        const uint64_t b_desc = make_smem_desc_mn_sw128(v_tile(c) + kk * 2048, static_cast<uint32_t>(cap * 128), 1024u);
                        umma_bf16(tmem_base + ATC_O_COL, a_desc, b_desc, idesc, kk != 0);
                    }
                    umma_commit(bar_o);
                    if (c + 2 < n_chunks && !k_normed) {                // S runs two chunks ahead: TMEM buffer c & 1 is free (p_ready(c))
                        mbar_wait(&full_k[c & 1], ((c + 2) >> 1) & 1, 0x630 + c);
                        tc_fence_after();
                        issue_s(c + 2);
                    }
                }
                __syncwarp();
            }
            // ---- o_acc = o_acc * corr(c) + O chunk (this thread's row, all 128 channels), 16 columns at a time
            mbar_wait(bar_o, c & 1, 0x6a0 + c);
            tc_fence_after();
            const float corr = corr_s[(c & 3) * 128 + row];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint32_t ov[16];
                tmem_ld_32x16(t_row + ATC_O_COL + g * 16, ov);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) o_acc[g * 16 + j] = fmaf(o_acc[g * 16 + j], corr, __uint_as_float(ov[j]));
            }
            tc_fence_before();
            mbar_arrive(o_free);
        }
        // ---- normalise by the row sum (all softmax shares) and stage the bf16 row in the dead Q / K memory
        if (probe0) g_foley_times[9] = clock64();
        asm volatile("bar.sync 3, %0;" ::"n"(ATC_THREADS) : "memory");   // row-sum shares visible; every MMA has completed
        {
            float lsum = red_sum[row];
#pragma unroll
            for (int i = 1; i < ATC_CS; ++i) lsum += red_sum[i * 128 + row];
            const float inv = __fdividef(1.0f, lsum);
            const uint32_t dst = sQ + static_cast<uint32_t>(row) * ATC_O_PITCH;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                st_shared_v4(dst + j * 16, pack_bf16x2(o_acc[8 * j] * inv, o_acc[8 * j + 1] * inv),
                             pack_bf16x2(o_acc[8 * j + 2] * inv, o_acc[8 * j + 3] * inv),
                             pack_bf16x2(o_acc[8 * j + 4] * inv, o_acc[8 * j + 5] * inv),
                             pack_bf16x2(o_acc[8 * j + 6] * inv, o_acc[8 * j + 7] * inv));
        }
    }
    if (softmax_role) asm volatile("bar.sync 3, %0;" ::"n"(ATC_THREADS) : "memory");   // (the accumulation warps' side is above)
    __syncthreads();                                                                  // staged tile complete
    {   // whole 256-byte row segments to global, all 12 warps
        __nv_bfloat16* O = a.o + b * a.o_batch_stride + h * 128;
        const long long ld = static_cast<long long>(a.H) * 128;
        for (int piece = threadIdx.x; piece < 128 * 16; piece += ATC_THREADS) {
            const int r = piece >> 4, pc = piece & 15;
            if (q0 + r < a.Sq) {
                const uint4 u = ld_shared_v4(sQ + static_cast<uint32_t>(r) * ATC_O_PITCH + static_cast<uint32_t>(pc) * 16u);
                *reinterpret_cast<uint4*>(O + static_cast<long long>(q0 + r) * ld + pc * 8) = u;
            }
        }
    }
    if (probe0) g_foley_times[12] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
        if (probe0) { g_foley_times[13] = clock64(); g_foley_times[14] = globaltimer_ns(); }
    }
}

// ---- host side
// 4-D view [d = 128, rows, heads, batch] of an operand, box = 64 channels x box_rows rows, 128-byte swizzle, zero fill.
inline bool encode_att_map(CUtensorMap* out, const AttOperand& t, int box_rows, std::string* err) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { if (err) *err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
    cuuint64_t dims[4] = {128, static_cast<cuuint64_t>(t.rows), static_cast<cuuint64_t>(t.heads > 0 ? t.heads : 1),
                          static_cast<cuuint64_t>(t.batch > 0 ? t.batch : 1)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(t.row_stride) * 2, static_cast<cuuint64_t>(t.head_stride) * 2,
                             static_cast<cuuint64_t>(t.batch_stride) * 2};
    for (int i = 1; i < 3; ++i)
        if (strides[i] == 0) strides[i] = strides[0];     // extent-1 dimensions
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_rows), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(t.ptr) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
        if (err) *err = "attention operands must be 16-byte aligned (pointer and strides)";
        return false;
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(t.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) {
            char buf[256];
            snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(attention) failed (%d): rows=%d heads=%d batch=%d strides=%lld/%lld/%lld",
                     static_cast<int>(r), t.rows, t.heads, t.batch, t.row_stride, t.head_stride, t.batch_stride);
            *err = buf;
        }
        return false;
    }
    return true;
}

// The dynamic shared-memory opt-in, once, up front (never inside a stream capture).
inline cudaError_t attention_tc_init() {
    return cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM);
}
inline bool launch_attention_tc(const AttOperand& q, const AttOperand& k, const AttOperand& v, const AttTcArgs& a, int batch,
                                cudaStream_t st, std::string* err) {
    CUtensorMap mq, mk, mv;
    const int box_kv = a.Sk > ATC_CK ? ATC_BOX_KV_LONG : ATC_BOX_KV_SHORT;
    if (!encode_att_map(&mq, q, ATC_BOX_Q, err) || !encode_att_map(&mk, k, box_kv, err) || !encode_att_map(&mv, v, box_kv, err)) return false;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>((a.Sq + 127) / 128), static_cast<unsigned>(a.H), static_cast<unsigned>(batch));
    cfg.blockDim = dim3(ATC_THREADS);
    cfg.dynamicSmemBytes = ATC_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, attention_tc_kernel, mq, mk, mv, a);
    if (e != cudaSuccess) {
        if (err) *err = std::string("attention launch failed: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

}  // namespace foley
