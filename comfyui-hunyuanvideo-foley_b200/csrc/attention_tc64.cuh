// tcgen05 / TMEM attention for head_dim 64 — the SigLIP2 vision tower's self-attention (HF SiglipAttention -> SDPA,
// reference feature_utils.py:64-79): 12 heads x 1024 patch tokens per frame, no mask, scale 1/8.
//
// One CTA = 128 query rows of one (frame, head), 160 threads: warps 0..3 are the softmax / accumulation warps (a thread
// owns ONE query row = one TMEM lane: the row maximum needs no exchange between threads), lane 0 of warp 4 issues every
// TMA copy and every MMA.  112 KB of shared memory and 256 TMEM columns per CTA, so TWO CTAs are resident per SM and one
// CTA's MMAs run under the other's exponentials (the kernel is bound by the MUFU unit: 128 x 128 ex2 per chunk).
// Per 128-key chunk c:
//   S(c) = Q K(c)^T      tcgen05.mma M = 128, N = 128, K = 64 into TMEM columns [0, 128)      (issued right after P(c-1) is read)
//   pass 1               row maximum of S(c) (tcgen05.ld 32x32b, four 32-column groups), rescale factor
//   O(c-1)               read back from TMEM columns [128, 192) and folded into the fp32 register accumulator
//   pass 2               p = ex2(s * scale * log2 e - m), rounded to bf16 into shared memory as a K-major A operand
//   O(c) = P(c) V(c)     tcgen05.mma M = 128, N = 64, K = 128 with V as an MN-major B operand (rows as loaded)
// K / V chunks are double-buffered TMA boxes read in place from the fused QKV projection output [rows, 3C]
// (4-D tensor map: channel, token, head, frame); tokens past the end arrive as zeros and are masked in the softmax.
#pragma once
#include "attention_tc.cuh"
#include "ptx.cuh"

namespace foley {

struct AttTc64Args {
    __nv_bfloat16* o = nullptr;            // element (b, r, h, d) at o + b*o_batch_stride + r*o_row_stride + h*64 + d
    long long o_batch_stride = 0, o_row_stride = 0;
    int H = 0, Sq = 0, Sk = 0;
    float scale_log2 = 0.f;                // softmax scale * log2(e)
};

constexpr int A64_THREADS = 160;                          // HALVES = 1: 4 row warps + the control warp (HALVES = 2: 288)
constexpr int A64_CK = 128;                               // keys per chunk
constexpr int A64_TILE = 128 * 128;                       // a [128 rows x 64 channels] bf16 tile: 16 KB
constexpr int A64_OFF_K = A64_TILE;                       // Q | K0 K1 | V0 V1 | P (two 64-key blocks) | barriers
constexpr int A64_OFF_V = A64_OFF_K + 2 * A64_TILE;
constexpr int A64_OFF_P = A64_OFF_V + 2 * A64_TILE;
constexpr int A64_OFF_BAR = A64_OFF_P + 2 * A64_TILE;
constexpr int A64_OFF_MX = A64_OFF_BAR + 128;             // HALVES = 2: [2][128] bf16 row-maximum exchange between the two column halves
constexpr int A64_SMEM = A64_OFF_MX + 512;                // 115328 bytes: two CTAs per SM (limit 115712)
constexpr int A64_O_COL = 128;                            // TMEM: S columns [0, 128), O columns [128, 192)
constexpr int A64_O_PITCH = 144;                          // output staging: 128 B per row + 16 B (conflict-free)

// HALVES = 2: TWO threads per query row (warps w and w + 4 share a TMEM lane quarter and split the 128 columns of an S chunk
// and the 64 channels of O): 8 row warps per CTA, four per scheduler with two CTAs per SM, so that one warp's exponentials
// run under another's tcgen05.ld / maximum / P stores (HALVES = 1 left the MUFU unit ~42 % busy: ncu, profiles/).  The two
// threads of a row agree on the running maximum through a bf16 value rounded UP (any common upper bound is a valid softmax
// shift) exchanged in shared memory behind a 64-thread named barrier.
template <int HALVES>
__global__ void __launch_bounds__(32 * (4 * HALVES + 1), 2)
attention_tc64_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const AttTc64Args a) {
    constexpr int THREADS = 32 * (4 * HALVES + 1), ROW_THREADS = 128 * HALVES, CTRL_WARP = 4 * HALVES;
    constexpr int NG = 4 / HALVES;                            // 32-column groups of an S chunk per thread
    constexpr int OC = 64 / HALVES;                           // O channels per thread
    extern __shared__ __align__(1024) uint8_t a64_smem[];
    const uint32_t sQ = smem_u32(a64_smem);
    const uint32_t sK = sQ + A64_OFF_K, sV = sQ + A64_OFF_V, sP = sQ + A64_OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a64_smem + A64_OFF_BAR);
    uint64_t* full_k = bars;            // [2]  K (+ Q with the first chunk) of a stage has landed
    uint64_t* full_v = bars + 2;        // [2]  V of a stage has landed
    uint64_t* bar_s = bars + 4;         //      S(c) is complete in TMEM
    uint64_t* bar_o = bars + 5;         //      PV(c) is complete (O chunk in TMEM; P and the V stage are free)
    uint64_t* p_ready = bars + 6;       //      every row thread has written P(c), read S(c) and read O(c-1)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int n_chunks = (a.Sk + A64_CK - 1) / A64_CK;
    auto chunk_rows = [&](int c) { return (min(A64_CK, a.Sk - c * A64_CK) + 15) & ~15; };

    if ((sQ & 1023u) != 0) { if (threadIdx.x == 0) atomicCAS(&g_foley_dbg[0], 0u, 0x7a0u); return; }   // layout assumption of the descriptors
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    if (threadIdx.x == ROW_THREADS) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        for (int i = 0; i < 2; ++i) { mbar_init(&full_k[i], 1); mbar_init(&full_v[i], 1); }
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        mbar_init(p_ready, ROW_THREADS);
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == CTRL_WARP) {
        // ================================================================== control lane: TMA copies and MMAs
        if (lane == 0) {
            auto issue_k = [&](int c) {
                const int st = c & 1;
                mbar_expect_tx(&full_k[st], (c == 0 ? 2u : 1u) * A64_TILE);
                if (c == 0) tma_load_4d(sQ, &tm_q, &full_k[0], 0, q0, h, b);
                tma_load_4d(sK + st * A64_TILE, &tm_k, &full_k[st], 0, c * A64_CK, h, b);
            };
            auto issue_v = [&](int c) {
                const int st = c & 1;
                mbar_expect_tx(&full_v[st], A64_TILE);
                tma_load_4d(sV + st * A64_TILE, &tm_v, &full_v[st], 0, c * A64_CK, h, b);
            };
            auto issue_s = [&](int c) {         // S(c) = Q K(c)^T: four K = 16 steps over the 64 channels
                const uint32_t idesc = make_idesc(1, 128, chunk_rows(c));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(tmem_base, make_smem_desc_sw128(sQ + ks * 32), make_smem_desc_sw128(sK + (c & 1) * A64_TILE + ks * 32),
                              idesc, ks != 0);
                umma_commit(bar_s);
            };
            issue_k(0);
            issue_v(0);
            if (n_chunks > 1) { issue_k(1); issue_v(1); }
            mbar_wait(&full_k[0], 0, 0x710);
            tc_fence_after();
            issue_s(0);
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(p_ready, c & 1, 0x740 + (c & 15));      // P(c) written; S(c) and O(c-1) consumed by every row thread
                tc_fence_after();
                if (c + 1 < n_chunks) {                            // S first: pass 1 of the next chunk runs under PV(c)
                    mbar_wait(&full_k[(c + 1) & 1], ((c + 1) >> 1) & 1, 0x720 + (c & 15));
                    tc_fence_after();
                    issue_s(c + 1);
                }
                mbar_wait(&full_v[c & 1], (c >> 1) & 1, 0x730 + (c & 15));
                tc_fence_after();
                {   // O(c) = P(c) V(c): K = keys in steps of 16, N = 64 channels, V MN-major
                    const uint32_t idesc = make_idesc(1, 128, 64) | (1u << 16);
                    const int ksteps = chunk_rows(c) >> 4;
                    for (int kk = 0; kk < ksteps; ++kk)
                        umma_bf16(tmem_base + A64_O_COL, make_smem_desc_sw128(sP + (kk >> 2) * A64_TILE + (kk & 3) * 32),
                                  make_smem_desc_mn_sw128(sV + (c & 1) * A64_TILE + kk * 2048, A64_TILE, 1024u), idesc, kk != 0);
                    umma_commit(bar_o);
                }
                // refills: K(c+2) into the stage S(c) has released; V(c+1) into the stage PV(c-1) has released (c >= 1)
                if (c + 2 < n_chunks) issue_k(c + 2);
                if (c >= 1 && c + 1 < n_chunks) issue_v(c + 1);
            }
        }
        __syncwarp();
    } else {
        // ================================================================== row warps: softmax + O accumulation
        const int q = warp & 3, hf = warp >> 2;                 // TMEM lane quarter; which half of the columns (HALVES = 2)
        const int row = q * 32 + lane;                          // query row of the tile = TMEM lane
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        __nv_bfloat16* mxs = reinterpret_cast<__nv_bfloat16*>(a64_smem + A64_OFF_MX);
        float o_acc[OC];
#pragma unroll
        for (int j = 0; j < OC; ++j) o_acc[j] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, corr_prev = 1.f;
        auto take_o = [&](float corr) {                          // o_acc = o_acc * corr + this thread's channels of the O chunk
#pragma unroll
            for (int g = 0; g < OC / 32; ++g) {
                uint32_t ov[32];
                tmem_ld_32x32(t_row + A64_O_COL + hf * OC + g * 32, ov);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) o_acc[g * 32 + j] = fmaf(o_acc[g * 32 + j], corr, __uint_as_float(ov[j]));
            }
        };
        for (int c = 0; c < n_chunks; ++c) {
            const int kn = min(A64_CK, a.Sk - c * A64_CK);
            const int n_grp = (kn + 31) >> 5;
            mbar_wait(bar_s, c & 1, 0x700 + (c & 15));
            tc_fence_after();
            // ---- pass 1: row maximum over this thread's column groups
            float mx = -INFINITY;
#pragma unroll 1
            for (int gi = 0; gi < NG; ++gi) {
                const int g = hf * NG + gi;
                if (g >= n_grp) break;
                uint32_t v[32];
                tmem_ld_32x32(t_row + g * 32, v);
                tmem_ld_wait();
                if (g * 32 + 32 <= kn) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (g * 32 + j) < kn ? __uint_as_float(v[j]) : -INFINITY);
                }
            }
            float m_new;
            if constexpr (HALVES == 2) {
                // both threads of the row use max(up(own), up(partner)): identical on both sides, >= the true maximum.
                // (single-buffered: a thread writes its next value only after bar_s(c+1), i.e. after every row thread has
                //  arrived on p_ready(c), which comes after this read)
                const __nv_bfloat16 up = __float2bfloat16_ru(mx * a.scale_log2);
                mxs[hf * 128 + row] = up;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                m_new = fmaxf(m_run, fmaxf(__bfloat162float(up), __bfloat162float(mxs[(1 - hf) * 128 + row])));
            } else {
                m_new = fmaxf(m_run, mx * a.scale_log2);
            }
            const float corr = ex2_fast(m_run - m_new);         // first chunk: ex2(-inf) = 0
            m_run = m_new;
            // ---- O(c-1): PV(c-1) ran under pass 1; its completion also frees the P buffer
            if (c > 0) {
                mbar_wait(bar_o, (c - 1) & 1, 0x750 + (c & 15));
                tc_fence_after();
                take_o(corr_prev);
            }
            corr_prev = corr;
            // ---- pass 2: exponentials -> bf16 P (K-major swizzled A operand), row sum from the fp32 values
            float l_chunk = 0.f;
#pragma unroll 1
            for (int gi = 0; gi < NG; ++gi) {
                const int g = hf * NG + gi;
                if (g >= n_grp) break;
                uint32_t v[32], pk[16];
                tmem_ld_32x32(t_row + g * 32, v);
                tmem_ld_wait();
                l_chunk += atc_exps(v, pk, g * 32, kn, a.scale_log2, m_new);
                atc_put(pk, sP, row, g * 32);
            }
            l_run = l_run * corr + l_chunk;
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(p_ready);
        }
        // ---- last chunk's O, normalise, stage the bf16 row (128 bytes) in the dead Q / K memory
        mbar_wait(bar_o, (n_chunks - 1) & 1, 0x760);
        tc_fence_after();
        take_o(corr_prev);
        if constexpr (HALVES == 2) {    // the row sum is the sum of the two halves' sums (same maximum sequence); P is dead by now
            float* ls = reinterpret_cast<float*>(a64_smem + A64_OFF_P);
            ls[hf * 128 + row] = l_run;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            l_run += ls[(1 - hf) * 128 + row];
        }
        const float inv = __fdividef(1.0f, l_run);
        const uint32_t dst = sQ + static_cast<uint32_t>(row) * A64_O_PITCH + static_cast<uint32_t>(hf * OC * 2);
#pragma unroll
        for (int j = 0; j < OC / 8; ++j)
            st_shared_v4(dst + j * 16, pack_bf16x2(o_acc[8 * j] * inv, o_acc[8 * j + 1] * inv),
                         pack_bf16x2(o_acc[8 * j + 2] * inv, o_acc[8 * j + 3] * inv),
                         pack_bf16x2(o_acc[8 * j + 4] * inv, o_acc[8 * j + 5] * inv),
                         pack_bf16x2(o_acc[8 * j + 6] * inv, o_acc[8 * j + 7] * inv));
    }
    tc_fence_before();
    __syncthreads();                                            // staged tile complete; every MMA has completed
    {   // whole 128-byte row segments to global: 8 lanes per row
        __nv_bfloat16* O = a.o + b * a.o_batch_stride + h * 64;
        for (int piece = threadIdx.x; piece < 128 * 8; piece += THREADS) {
            const int r = piece >> 3, pc = piece & 7;
            if (q0 + r < a.Sq) {
                const uint4 u = ld_shared_v4(sQ + static_cast<uint32_t>(r) * A64_O_PITCH + static_cast<uint32_t>(pc) * 16u);
                *reinterpret_cast<uint4*>(O + static_cast<long long>(q0 + r) * a.o_row_stride + pc * 8) = u;
            }
        }
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ---- host side
// 4-D view [d = 64, rows, heads, batch] of an operand stored as (b, r, h, d) at ptr + b*batch_stride + r*row_stride + h*64 + d.
inline bool encode_att64_map(CUtensorMap* out, const __nv_bfloat16* ptr, long long rows, long long heads, long long batch,
                             long long row_stride, long long batch_stride, std::string* err) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { if (err) *err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
    cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(heads), static_cast<cuuint64_t>(batch > 0 ? batch : 1)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(row_stride) * 2, 128, static_cast<cuuint64_t>(batch_stride) * 2};
    if (strides[2] == 0) strides[2] = strides[0];
    cuuint32_t box[4] = {64, 128, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (strides[2] & 15)) {
        if (err) *err = "attention operands must be 16-byte aligned (pointer and strides)";
        return false;
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) {
            char buf[256];
            snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(attention d64) failed (%d): rows=%lld heads=%lld batch=%lld strides=%lld/%lld",
                     static_cast<int>(r), rows, heads, batch, row_stride, batch_stride);
            *err = buf;
        }
        return false;
    }
    return true;
}

inline cudaError_t attention_tc64_init() {
    cudaError_t e = cudaFuncSetAttribute(attention_tc64_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, A64_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(attention_tc64_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, A64_SMEM);
}

// q / k / v: (b, r, h, d) operands with their own row / batch strides (elements); heads are 64 channels apart.
inline bool launch_attention_tc64(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, long long q_row_stride,
                                  long long q_batch_stride, long long kv_row_stride, long long kv_batch_stride, int batch,
                                  const AttTc64Args& a, cudaStream_t st, std::string* err, int halves = 2) {
    CUtensorMap mq, mk, mv;
    if (!encode_att64_map(&mq, q, a.Sq, a.H, batch, q_row_stride, q_batch_stride, err) ||
        !encode_att64_map(&mk, k, a.Sk, a.H, batch, kv_row_stride, kv_batch_stride, err) ||
        !encode_att64_map(&mv, v, a.Sk, a.H, batch, kv_row_stride, kv_batch_stride, err))
        return false;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>((a.Sq + 127) / 128), static_cast<unsigned>(a.H), static_cast<unsigned>(batch));
    cfg.blockDim = dim3(halves == 2 ? 288 : A64_THREADS);
    cfg.dynamicSmemBytes = A64_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    const cudaError_t e = halves == 2 ? cudaLaunchKernelEx(&cfg, attention_tc64_kernel<2>, mq, mk, mv, a)
                                      : cudaLaunchKernelEx(&cfg, attention_tc64_kernel<1>, mq, mk, mv, a);
    if (e != cudaSuccess) {
        if (err) *err = std::string("attention (d64) launch failed: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

}  // namespace foley
