// tcgen05 + TMA GEMM / multi-tap 1-D convolution kernel for sm_100a.
//
//   C[b, r, n] = sum_{tap} sum_{k} A[b, r + off0 + tap*tap_stride, k] * W[n, tap*K + k]
//
// A is a K-major activation tensor [batch, rows, K] (bf16, or fp32 consumed as tf32); rows outside
// [0, rows) read as zero through TMA out-of-bounds fill, which is what gives every sample of a
// batch its own zero halo for the k=3 / k=7 convolutions and the transposed-conv polyphase form.
// W is a K-major weight matrix [N, taps*K].  One CTA computes one 128 x BN output tile (optionally one
// K-split of it): warp 0 = TMA producer of the weight tiles, warp 10 = TMA producer of the activation tiles (two
// single-thread producers: one thread issuing both loads of a k-block was the slowest agent of the ring), warp 1 =
// TMEM allocator + single-thread tcgen05.mma issuer, warps 2..9 = epilogue (tcgen05.ld -> registers -> fused
// epilogue -> global).
//
// Replaces, on the reference path: F.linear / Conv1d(k=1,3) in hifi_foley.py:216-331,364-390,
// mlp_layers.py:104-149, and the DAC decoder convs in dac.py:28-44,98-149.
#pragma once
#include "ptx.cuh"

#ifndef FOLEY_GEMM_TWO_PRODUCERS
#define FOLEY_GEMM_TWO_PRODUCERS 1   // 0: warp 0 issues both loads of a k-block (A/B switch for the round-2 bisect of the w1|w3 shape)
#endif
#ifndef FOLEY_PAIR_DIRECT
#define FOLEY_PAIR_DIRECT 1   // CTA-pair tiles: both CTAs' TMA loads complete on the LEADER's full barrier (cta_group::2 TMA form; the
                              // leader expects the pair's bytes) instead of the peer forwarding "stage landed" with a remote arrive
#endif
// (SwiGLU tiles keep storing 32 bytes per row and chunk straight from registers: sending them through the staging tile +
//  TMA stores measured slower both in round 1 (+0.8 us) and with round 2's arrive-only barriers (w1|w3 29.8 vs 29.0 us), and
//  so did keeping the next chunk's tcgen05.ld in flight (29.2 us, 168 registers): that epilogue is bound by its math — two
//  MUFU ops and three bf16 roundings per output — not by its stores.)
#ifndef FOLEY_PDL_EARLY_TRIGGER
#define FOLEY_PDL_EARLY_TRIGGER 1   // griddepcontrol.launch_dependents right after this kernel's own dependency wait (0: after its mainloop): the next kernel
                                    // becomes resident (prologue, weight prefetch) under this mainloop wherever resources allow: 3.937 -> 3.926 ms per step
#endif
#ifndef FOLEY_EPI_WARPS_BF16
#define FOLEY_EPI_WARPS_BF16 8    // 16 measured: fc1 (GELU) 14.5 -> 13.7 us, but w2 / w1|w3 0.3 us slower; step 4.14 -> 4.17 ms
#endif

namespace foley {

enum EpiMode : int {
    EPI_BF16 = 0,    // out(bf16) = act(bf16r(acc + bias))
    EPI_SWIGLU = 1,  // columns hold (w1_j, w3_j) pairs: out(bf16)[j] = bf16r(bf16r(silu(bf16r(a))) * bf16r(b))
    EPI_F32 = 2,     // out(fp32)[split] = acc     (raw partial sums; consumer applies bias/gate)
    EPI_DAC = 3,     // fp32 conv epilogue: y = acc + bias (+resid); out = y; out2 = snake(y) / tanh(y)
};
enum ActMode : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU_TANH = 2, ACT_TANH = 3, ACT_GELU_ERF = 4 };   // GELU_ERF (nn.GELU()): fp16 instantiations only

struct GemmEpi {
    int mode = EPI_BF16;
    int act = ACT_NONE;
    void* out = nullptr;            // primary output
    long long ldo = 0;              // elements between output rows
    long long out_batch_stride = 0; // elements between samples
    long long split_stride = 0;     // EPI_F32: elements between K-split partials
    const void* bias = nullptr;     // bf16[N] (DiT; fp16[N] in the fp16 instantiations) or fp32[ch_mod] (DAC); may be null
    // --- EPI_DAC only
    void* out2 = nullptr;           // snake(y) with alpha (next conv's input); may be null
    const float* resid = nullptr;   // residual input, same indexing as out; may be null
    const float* alpha = nullptr;   // snake alpha per channel for out2
    int ch_mod = 0;                 // channel = col % ch_mod (bias/alpha index); 0 -> col
    long long flat_lo = 0, flat_hi = 0;  // if flat_hi > flat_lo: store only when flat_lo <= r*ldo+col < flat_hi
};

struct GemmArgs {
    int rows;        // valid rows per sample
    int n;           // output columns (N)
    int kb_per_tap;  // K / BLOCK_K
    int taps;
    int tap_off0;    // A row offset of tap 0
    int tap_stride;  // A row offset increment per tap
    int splits;      // K splits (gridDim.z); the taps*kb_per_tap k-blocks are divided evenly
    int dbg_stop;    // bring-up aid: 0 = full kernel, 1 = setup only, 2 = TMA only, 3 = TMA+MMA, 4 = MMA only (first ring
                     // pass re-used), 8 = full kernel + clock64 stamps of CTA 0 (foley_debug_times)
    int prefetch_b;  // L2 prefetch distance of the weight operand in k-blocks (0 = off)
    int pf_mod;      // CTAs with blockIdx.x % pf_mod == 0 issue the prefetch (m-tiles sharing a weight column)
    int cluster_m;   // > 1: the cluster_m CTAs of a (cluster_m,1,1) cluster (consecutive m-tiles, same weight tile) each load
                     // 1/cluster_m of the B tile and MULTICAST it to all of them (one L2 read instead of cluster_m)
    GemmEpi epi;
};

// kPair: the CTA pair of a (2,1,1) cluster computes a 256 x BN tile with tcgen05.mma.cta_group::2 — each CTA holds
// its own 128 rows of A and only HALF of the B tile, so the per-SM TMA ingress per MMA cycle (what bounds this
// kernel) drops by a third and the same shared memory holds twice the MMA time.
template <int BN, bool kTF32, bool kPair = false>
struct GemmCfg {
    static constexpr int BM = 128;
    static constexpr int BK_BYTES = 128;                      // one swizzle-128B row
    static constexpr int BK = kTF32 ? 32 : 64;                // elements per k-block
    static constexpr int UMMA_K = kTF32 ? 8 : 16;             // 32 bytes per instruction
    static constexpr int A_BYTES = BM * BK_BYTES;             // 16 KB
    static constexpr int B_ROWS = kPair ? BN / 2 : BN;        // rows of B this CTA loads
    static constexpr int B_BYTES = B_ROWS * BK_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 8 ? 8 : (200 * 1024 / STAGE_BYTES);
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 3 * 256 * 4 /*epilogue params*/;
    // weight-TMA warp, MMA warp, epilogue warps, activation-TMA warp.  8 epilogue warps (two per TMEM lane quarter);
    // the bf16 (DiT) kernels can be built with 16 (-DFOLEY_EPI_WARPS_BF16=16: no net gain, see above), the tf32 (DAC)
    // kernels always keep 8: their epilogue needs > 100 registers.
    static constexpr int EPI_WARPS = FOLEY_EPI_WARPS_BF16 == 16 && !kTF32 ? 16 : 8;
    static constexpr int EPI_GROUPS = EPI_WARPS / 4;          // warps per TMEM lane quarter = interleaved column-chunk sets
    static constexpr int A_WARP = FOLEY_GEMM_TWO_PRODUCERS ? 2 + EPI_WARPS : -1;   // warp id of the activation producer
    static constexpr int THREADS = 32 * (2 + EPI_WARPS + (FOLEY_GEMM_TWO_PRODUCERS ? 1 : 0));
};

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case ACT_SILU: return x / (1.0f + expf(-x));
        case ACT_GELU_TANH: return gelu_tanh_f(x);
        case ACT_TANH: return tanhf(x);
        default: return x;
    }
}

// kF16: the 16-bit operands, the bias and the output are IEEE fp16 instead of bf16 (EPI_BF16 mode only; a separate
// instantiation so that the bf16 kernels of the DiT step compile to exactly the code they had before the mode existed:
// as a run-time flag it cost the step 1.3 %).
template <int BN, bool kTF32, bool kPair = false, bool kF16 = false>
__global__ void __launch_bounds__((GemmCfg<BN, kTF32, kPair>::THREADS), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ CUtensorMap tm_c, const GemmArgs g) {
    using Cfg = GemmCfg<BN, kTF32, kPair>;
    extern __shared__ uint8_t smem_raw[];
    // swizzle-128B tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::STAGES;
    uint64_t* tmem_full_bar = bars + 2 * Cfg::STAGES;
    uint64_t* peer_ready = bars + 2 * Cfg::STAGES + 1;   // pair mode, leader only: the peer CTA's stage has landed
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3 * Cfg::STAGES + 1);
    float* epi_f = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);   // barriers use <= 26*8 bytes   // [3][256]: bias, alpha, 1/alpha

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool tprobe = g.dbg_stop == 8 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    if (tprobe && threadIdx.x == 0) { g_foley_times[0] = clock64(); g_foley_times[15] = globaltimer_ns(); }
    const int n0 = blockIdx.y * BN;
    const int m_tiles = (g.rows + Cfg::BM - 1) / Cfg::BM;
    const int batch = blockIdx.x / m_tiles;          // x carries (batch, m-tile): it may exceed 65535 (DAC)
    const int m0 = (blockIdx.x % m_tiles) * Cfg::BM;
    const int split = blockIdx.z;

    const int kb_total = g.taps * g.kb_per_tap;
    const int kb_begin = static_cast<int>((static_cast<long long>(kb_total) * split) / g.splits);
    const int kb_end = static_cast<int>((static_cast<long long>(kb_total) * (split + 1)) / g.splits);
    const int num_kb = kb_end - kb_begin;

    // CTA pair: rank 0 (even m-tile) is the leader — it owns the full barriers and issues the MMAs for both.
    const uint32_t pair_rank = kPair ? cluster_ctarank() : 0;
    const bool leader = pair_rank == 0;

    // Each CTA's loads complete on its OWN full barrier (two arrivals per phase: the weight producer's and the
    // activation producer's expect_tx); in pair mode the peer forwards "stage landed" to the leader with one remote
    // arrive per stage (remote complete_tx from TMA, the cta_group::2 TMA form, measured slower here).
    // weight-tile multicast across the m-tiles of a cluster (single-CTA tiles only)
    const uint32_t cm = kPair ? 1u : static_cast<uint32_t>(g.cluster_m > 1 ? g.cluster_m : 1);
    const uint32_t crank = cm > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = static_cast<uint16_t>((1u << cm) - 1u);
    constexpr bool kDirect = kPair && FOLEY_PAIR_DIRECT && FOLEY_GEMM_TWO_PRODUCERS;
    const uint32_t lead_full = kDirect ? mapa_u32(smem_u32(full_bar), 0) : 0;   // the leader's full barriers (shared::cluster)
    auto load_b = [&](int s, int kb) {
        if constexpr (kDirect) {
            if (leader) mbar_expect_tx(&full_bar[s], 2 * Cfg::B_BYTES);   // both halves of the B tile land on this barrier
            tma_load_3d_pair(smem_b + s * Cfg::B_BYTES, &tm_b, lead_full + s * 8, kb * Cfg::BK,
                             n0 + static_cast<int>(pair_rank) * Cfg::B_ROWS, 0);
        } else if (cm > 1) {
            // this CTA's slice of the B tile goes to every CTA of the cluster; its own barrier collects all cm slices
            const uint32_t slice_rows = static_cast<uint32_t>(BN) / cm;
            mbar_expect_tx(&full_bar[s], Cfg::B_BYTES);
            tma_load_3d_mc(smem_b + s * Cfg::B_BYTES + crank * slice_rows * Cfg::BK_BYTES, &tm_b, &full_bar[s], kb * Cfg::BK,
                           n0 + static_cast<int>(crank * slice_rows), 0, cmask);
        } else {
            mbar_expect_tx(&full_bar[s], Cfg::B_BYTES);
            tma_load_3d(smem_b + s * Cfg::B_BYTES, &tm_b, &full_bar[s], kb * Cfg::BK,
                        n0 + static_cast<int>(pair_rank) * Cfg::B_ROWS, 0);   // rank-3 map; pair: this CTA's half
        }
    };
    auto load_a = [&](int s, int kcol, int arow) {
        if constexpr (kDirect) {
            if (leader) mbar_expect_tx(&full_bar[s], 2 * Cfg::A_BYTES);
            tma_load_3d_pair(smem_a + s * Cfg::A_BYTES, &tm_a, lead_full + s * 8, kcol, arow, batch);
        } else {
            mbar_expect_tx(&full_bar[s], Cfg::A_BYTES);
            tma_load_3d(smem_a + s * Cfg::A_BYTES, &tm_a, &full_bar[s], kcol, arow, batch);
        }
    };
    const int pre = g.dbg_stop == 1 ? 0 : (num_kb < Cfg::STAGES ? num_kb : Cfg::STAGES);
    // direct pair mode: the peer's loads signal the LEADER's barriers, which exist only after the cluster sync below
    const int pre_early = ((kDirect && !leader) || cm > 1) ? 0 : pre;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        tma_prefetch_desc(&tm_c);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 2);
            mbar_init(&empty_bar[s], cm);   // multicast: a stage is refilled (in every CTA) once all cm CTAs have consumed it
            mbar_init(&peer_ready[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
        // Weights do not depend on the previous kernel (programmatic dependent launch) nor on the TMEM allocation
        // going on in warp 1: fill the ring with B tiles right away so the weight stream's latency hides under the
        // rest of the prologue and under the predecessor kernel's tail.
        for (int i = 0; i < pre_early; ++i) load_b(i, kb_begin + i);
    } else if (warp == 1) {
        if constexpr (kPair) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_ptr_smem);
        else tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
    }
    tc_fence_before();
    __syncthreads();
    if (kPair || cm > 1) cluster_sync_all();   // the peers' barriers must exist before anything is signalled to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (tprobe && threadIdx.x == 0) g_foley_times[1] = clock64();

    // epilogue geometry (also needed by warp 0, which issues the TMA stores of the staged epilogues)
    constexpr int NG = Cfg::EPI_GROUPS;          // column-chunk sets: chunk c0 = half*32 + it*32*NG belongs to set `half`
    constexpr int EPI_THREADS = 32 * Cfg::EPI_WARPS;
    constexpr int NIT = (BN + 32 * NG - 1) / (32 * NG);
    static_assert(NIT <= 8, "one named barrier (ids 2..9) per column group");

    if (g.dbg_stop == 1) {
        // setup / teardown only
    } else if (warp == 0) {
        // ------------------------------------------------------------- TMA producer: weight tiles past the first ring pass
#if FOLEY_GEMM_TWO_PRODUCERS
        if (lane == 0) {
            int s = pre_early == Cfg::STAGES ? 0 : pre_early;
            uint32_t ph = pre_early == Cfg::STAGES ? 1 : 0;   // round r >= 1 waits for the consumer's release of round r-1: parity (r & 1) ^ 1
            for (int i = pre_early; i < (g.dbg_stop == 4 ? 0 : num_kb); ++i) {
                if (i >= Cfg::STAGES) { if (!mbar_wait(&empty_bar[s], ph ^ 1, 0x100 + i)) break; }
                load_b(s, kb_begin + i);
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
            }
        }
#else
        if (lane == 0) {   // one thread issues both operands (same barrier protocol: two expect_tx arrivals per stage)
            int tap = kb_begin / g.kb_per_tap;
            int kk = kb_begin - tap * g.kb_per_tap;
            int arow = m0 + g.tap_off0 + tap * g.tap_stride;
            int s = 0;
            uint32_t ph = 0;
            pdl_wait();
            for (int i = 0; i < (g.dbg_stop == 4 ? pre : num_kb); ++i) {
                if (i >= pre) {
                    if (!mbar_wait(&empty_bar[s], ph ^ 1, 0x100 + i)) break;
                    load_b(s, kb_begin + i);
                }
                mbar_expect_tx(&full_bar[s], Cfg::A_BYTES);
                tma_load_3d(smem_a + s * Cfg::A_BYTES, &tm_a, &full_bar[s], kk * Cfg::BK, arow, batch);
                if (++kk == g.kb_per_tap) { kk = 0; arow += g.tap_stride; }
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
            }
        }
#endif
        // ---- after the mainloop this warp issues the TMA stores of the staged (bf16 / fp32) epilogues: the epilogue warps
        // arrive on one named barrier per column group as soon as they have written it to the staging tile and move on
        __syncwarp();
        if (g.epi.mode != EPI_DAC && g.epi.mode != EPI_SWIGLU) {
            const int bpc = g.epi.mode == EPI_F32 ? 4 : 2;      // output bytes per accumulator column
            const int esz = bpc;
            const bool live_w = (g.dbg_stop & 7) == 0 && num_kb > 0;
            int boxes_issued = 0;
            for (int it = 0; it < NIT; ++it) {
                asm volatile("bar.sync %0, %1;" ::"r"(2 + it), "n"(EPI_THREADS + 32) : "memory");
                if (lane == 0 && live_w) {
                    const int cols_done = (it + 1) * 32 * NG < BN ? (it + 1) * 32 * NG : BN;
                    const int boxes_done = (cols_done * bpc) >> 7;
                    for (; boxes_issued < boxes_done; ++boxes_issued) {
                        const int col_elem = (n0 * bpc + boxes_issued * 128) / esz;
                        if (col_elem < g.n) tma_store_4d(&tm_c, smem + boxes_issued * 16384, col_elem, m0, batch, split);
                    }
                    tma_store_commit();
                }
            }
            if (lane == 0) tma_store_wait_read();               // the staging memory must outlive the reads
        }
    } else if (warp == Cfg::A_WARP) {
        // ------------------------------------------------------------- TMA producer: activation tiles
        if (lane == 0 && FOLEY_GEMM_TWO_PRODUCERS) {
            int tap = kb_begin / g.kb_per_tap;
            int kk = kb_begin - tap * g.kb_per_tap;
            int arow = m0 + g.tap_off0 + tap * g.tap_stride;
            int s = 0;
            uint32_t ph = 0;
            pdl_wait();   // activations (A) are the predecessor's output
            if (tprobe) g_foley_times[2] = clock64();
            for (int i = 0; i < (g.dbg_stop == 4 ? pre : num_kb); ++i) {
                if (i >= Cfg::STAGES) { if (!mbar_wait(&empty_bar[s], ph ^ 1, 0x180 + i)) break; }
                load_a(s, kk * Cfg::BK, arow);
                if (++kk == g.kb_per_tap) { kk = 0; arow += g.tap_stride; }
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------- MMA issuer (one thread; pair: leader CTA only)
        if (lane == 0 && !leader && !kDirect) {
            // peer CTA of a pair: forward each landed stage to the leader
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < num_kb; ++i) {
                if (!mbar_wait(&full_bar[s], ph, 0x400 + i)) break;
                mbar_arrive_cluster(mapa_u32(smem_u32(&peer_ready[s]), 0));
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
            }
        }
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc(kTF32 ? 2 : (kF16 ? 0 : 1), kPair ? 2 * Cfg::BM : Cfg::BM, BN);
            int s = -1;
            uint32_t ph = 1;
            for (int i = 0; i < num_kb; ++i) {
                if (++s == Cfg::STAGES) s = 0;
                if (s == 0) ph ^= 1;
                if (g.dbg_stop != 4 || i < pre) {   // dbg 4: MMA-only rate, the first ring pass is re-used without reloading
                if (!mbar_wait(&full_bar[s], ph, 0x200 + i)) break;
                if constexpr (kPair && !kDirect) { if (!mbar_wait(&peer_ready[s], ph, 0x500 + i)) break; }
                }
                tc_fence_after();
                if (tprobe && i == 0) g_foley_times[3] = clock64();
                if (g.dbg_stop == 2) {
                    if constexpr (kPair) { mbar_arrive(&empty_bar[s]); mbar_arrive_cluster(mapa_u32(smem_u32(&empty_bar[s]), 1)); }
                    else if (cm > 1) { for (uint32_t rk = 0; rk < cm; ++rk) mbar_arrive_cluster(mapa_u32(smem_u32(&empty_bar[s]), rk)); }
                    else mbar_arrive(&empty_bar[s]);
                    continue;
                }
                const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + s * Cfg::A_BYTES));
                const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + s * Cfg::B_BYTES));
#pragma unroll
                for (int k = 0; k < Cfg::BK / Cfg::UMMA_K; ++k) {
                    // advance 32 bytes (= 2 x 16 B units) along K inside the swizzle atom
                    // (Alternating the K-slices of a k-block between two TMEM accumulators changes nothing: the ~150-cycle
                    // floor per tcgen05.mma, whatever N <= 256, is an issue-rate property, not an accumulator dependency.)
                    const uint32_t acc = (i | k) != 0;
                    if constexpr (kPair) {
                        if constexpr (kTF32) umma_tf32_pair(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
                        else umma_bf16_pair(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
                    } else {
                        if constexpr (kTF32) umma_tf32(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
                        else umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
                    }
                }
                // free the stage (in both CTAs of a pair) once these MMAs retire
                if constexpr (kPair) umma_commit_pair(&empty_bar[s]);
                else if (cm > 1) umma_commit_mc(&empty_bar[s], cmask);
                else umma_commit(&empty_bar[s]);
            }
            if (g.dbg_stop == 2) {
                mbar_arrive(tmem_full_bar);
                if constexpr (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_full_bar), 1));
            } else {
                if constexpr (kPair) umma_commit_pair(tmem_full_bar);   // accumulators complete in both CTAs
                else umma_commit(tmem_full_bar);
            }
            if (tprobe) g_foley_times[4] = clock64();
        }
    } else {
        // ------------------------------------------------------------- epilogue warps (2..9)
        // Two warps per TMEM lane quarter; each takes every other 32-column chunk.  Per-column parameters (bias,
        // snake alpha and 1/alpha) are staged in shared memory while the mainloop runs.
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;    // which interleaved set of column chunks
        const int r = m0 + q * 32 + lane;    // output row within the sample
        const GemmEpi& e = g.epi;
        {
            const int et = threadIdx.x - 64;
            for (int c = et; c < BN; c += EPI_THREADS) {
                const int col = n0 + c;
                float bv = 0.f, al = 0.f, ial = 0.f;
                if (col < g.n) {
                    if (e.mode == EPI_DAC) {
                        const int ch = e.ch_mod ? col % e.ch_mod : col;
                        if (e.bias) bv = reinterpret_cast<const float*>(e.bias)[ch];
                        if (e.alpha) { al = e.alpha[ch]; ial = 1.0f / (al + 1e-9f); }
                    } else if (e.bias) {
                        if constexpr (kF16) bv = __half2float(reinterpret_cast<const __half*>(e.bias)[col]);
                        else bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.bias)[col]);
                    }
                }
                epi_f[c] = bv; epi_f[256 + c] = al; epi_f[512 + c] = ial;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        }
        pdl_wait();                          // outputs may alias buffers the predecessor still reads
#if FOLEY_PDL_EARLY_TRIGGER
        pdl_trigger();                       // as soon as this kernel runs: the next kernel becomes resident wherever resources allow
#endif
        // (Parking these warps in a hardware barrier until warp 1 has seen the accumulator barrier, instead of letting
        // all of them poll it through the mainloop, measured no difference: try_wait suspends in hardware.)
        const bool acc_ok = mbar_wait(tmem_full_bar, 0, 0x300);
#if !FOLEY_PDL_EARLY_TRIGGER
        pdl_trigger();                       // mainloop done: the next kernel may start its prologue
#endif
        tc_fence_after();
        if (tprobe && threadIdx.x == 64) g_foley_times[5] = clock64();
        const bool row_ok = r < g.rows;
        const long long row_off = static_cast<long long>(batch) * e.out_batch_stride +
                                  static_cast<long long>(r) * e.ldo;
        // SwiGLU tiles write only 32 bytes per row and chunk: their direct stores were never the bottleneck, and the
        // staged path costs them ~0.7 us of barriers (measured), so they keep storing from registers.
        if (e.mode != EPI_DAC && e.mode != EPI_SWIGLU) {
            // ---- DiT epilogues (bf16 + bias + activation, fp32 K-split partials): the tile goes through shared memory (the
            // idle operand ring) and leaves as TMA stores.  A thread owns one accumulator ROW, so direct stores made every
            // warp instruction touch 32 different lines (measured 2.5-6.5 us per tile, LSU-transaction bound); staged, the
            // tile is written to global as whole 128-byte lines by the copy engine.  Staging layout = TMA boxes of
            // 128 rows x 128 B with the 128-byte swizzle (16-byte piece index ^ (row & 7)): conflict-free for one row
            // per thread.  Round 2 (tools/gemm_micro.py --dbg 8 showed 3.9 us per bf16 tile, 2.9 us per fp32 tile):
            //   * the activation switch is hoisted out of the element loop (it cost ~60 cycles per element),
            //   * the tcgen05.ld of the next chunk is in flight while this one is converted,
            //   * the epilogue warps only ARRIVE on a per-chunk named barrier; warp 0 (idle after the mainloop) waits on it
            //     and issues the TMA stores, so no epilogue warp stalls behind the store issue (~0.35 us per chunk).
            const int bpc = e.mode == EPI_F32 ? 4 : 2;
            const uint32_t r_local = static_cast<uint32_t>(q * 32 + lane);
            const uint32_t stage_base = smem_u32(smem);
            const uint32_t epi_s = smem_u32(epi_f);
            const bool live = acc_ok && (g.dbg_stop & 7) == 0 && num_kb > 0;
            const uint32_t sw = r_local & 7u;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 32;
            auto chunk_ok = [&](int it) {
                const int c0 = half * 32 + it * 32 * NG;
                return live && it < NIT && c0 < BN && n0 + c0 < g.n;
            };
            // One accumulator chunk (32 columns of this thread's row) -> fused math -> swizzled staging row.
            auto emit = [&](uint32_t (&v)[32], int c0) {
                const uint32_t byte_off = static_cast<uint32_t>(c0 * bpc);
                const uint32_t row_addr = stage_base + (byte_off >> 7) * 16384u + r_local * 128u;
                const uint32_t piece0 = (byte_off & 127u) >> 4;
                if (e.mode == EPI_F32) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(row_addr + (((piece0 + j) ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    return;
                }
                float a[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {     // bias of these 32 columns (staged as fp32)
                    uint32_t b0, b1, b2, b3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(epi_s + (c0 + 4 * j) * 4));
                    a[4 * j] = __uint_as_float(v[4 * j]) + __uint_as_float(b0);
                    a[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + __uint_as_float(b1);
                    a[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + __uint_as_float(b2);
                    a[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + __uint_as_float(b3);
                }
                if constexpr (kF16) {      // fp16 module (Synchformer under autocast): nn.GELU() or no activation
                    if (e.act == ACT_GELU_ERF) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { const float x = f16_round(a[j]); a[j] = 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
                    } else if (e.act != ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) a[j] = apply_act(f16_round(a[j]), e.act);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_shared_v4(row_addr + (((piece0 + j) ^ sw) << 4), pack_f16x2(a[8 * j], a[8 * j + 1]), pack_f16x2(a[8 * j + 2], a[8 * j + 3]),
                                     pack_f16x2(a[8 * j + 4], a[8 * j + 5]), pack_f16x2(a[8 * j + 6], a[8 * j + 7]));
                    return;
                }
                if (e.act == ACT_SILU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float x = bf16_round(a[j]); a[j] = __fdividef(x, 1.0f + __expf(-x)); }
                } else if (e.act == ACT_GELU_TANH) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = gelu_tanh_f(bf16_round(a[j]));
                } else if (e.act != ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = apply_act(bf16_round(a[j]), e.act);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    st_shared_v4(row_addr + (((piece0 + j) ^ sw) << 4), pack_bf16x2(a[8 * j], a[8 * j + 1]), pack_bf16x2(a[8 * j + 2], a[8 * j + 3]),
                                 pack_bf16x2(a[8 * j + 4], a[8 * j + 5]), pack_bf16x2(a[8 * j + 6], a[8 * j + 7]));
            };
            // registers written by an in-flight tcgen05.ld must not be touched before tcgen05.wait::ld: an empty asm with
            // "+r" operands pins every later use behind it
            auto pin = [](uint32_t (&v)[32]) {
#pragma unroll
                for (int j = 0; j < 32; ++j) asm volatile("" : "+r"(v[j]));
            };
            auto step = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int it) {
                tmem_ld_wait();
                if (chunk_ok(it + 1)) tmem_ld_32x32(t_row + (it + 1) * 32 * NG, nxt);
                if (chunk_ok(it)) {
                    pin(cur);
                    if (tprobe && threadIdx.x == 64 && it == 0) g_foley_times[9] = clock64();
                    emit(cur, half * 32 + it * 32 * NG);
                    if (tprobe && threadIdx.x == 64 && it == 0) g_foley_times[10] = clock64();
                }
                fence_proxy_async();                          // generic-proxy smem writes -> visible to the TMA store
                asm volatile("bar.arrive %0, %1;" ::"r"(2 + it), "n"(EPI_THREADS + 32) : "memory");   // warp 0 stores this column group
            };
            uint32_t va[32], vb[32];
            if (chunk_ok(0)) tmem_ld_32x32(t_row, va);
#pragma unroll 1
            for (int it = 0; it < NIT; it += 2) {
                step(va, vb, it);
                if (it + 1 < NIT) step(vb, va, it + 1);
            }
            if (tprobe && threadIdx.x == 64) g_foley_times[13] = clock64();
        } else
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN && acc_ok && (g.dbg_stop & 7) == 0; c0 += 32 * NG) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
            tmem_ld_wait();
            const int col = n0 + c0;
            if (!row_ok || col >= g.n || num_kb <= 0) continue;
            const float* pb = epi_f + c0;
            if (e.mode == EPI_SWIGLU) {
                __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out) + row_off + (col >> 1);
                uint32_t packed[8];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float g0 = bf16_round(__uint_as_float(v[j])), u0 = bf16_round(__uint_as_float(v[j + 1]));
                    const float g1 = bf16_round(__uint_as_float(v[j + 2])), u1 = bf16_round(__uint_as_float(v[j + 3]));
                    const float s0 = bf16_round(__fdividef(g0, 1.0f + __expf(-g0))) * u0;
                    const float s1 = bf16_round(__fdividef(g1, 1.0f + __expf(-g1))) * u1;
                    packed[j >> 2] = pack_bf16x2(s0, s1);
                }
                uint4* dst = reinterpret_cast<uint4*>(out);
                dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            } else if constexpr (kTF32) {  // EPI_DAC (fp32 / tf32 decoder convolutions only)
                const long long flat0 = static_cast<long long>(r) * e.ldo + col;
                const bool windowed = e.flat_hi > e.flat_lo;
                float* out = e.out ? reinterpret_cast<float*>(e.out) + row_off + col : nullptr;
                float* out2 = e.out2 ? reinterpret_cast<float*>(e.out2) + row_off + col : nullptr;
                const float* res = e.resid ? e.resid + row_off + col : nullptr;
                const bool full_ok = !windowed || (flat0 >= e.flat_lo && flat0 + 32 <= e.flat_hi);
                const bool any_ok = !windowed || (flat0 + 32 > e.flat_lo && flat0 < e.flat_hi);
                if (!any_ok) continue;
                const bool snake = e.alpha != nullptr && e.act != ACT_TANH;
                float y[32];
                if (res && full_ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 rv = reinterpret_cast<const float4*>(res)[j];
                        y[4 * j] = rv.x; y[4 * j + 1] = rv.y; y[4 * j + 2] = rv.z; y[4 * j + 3] = rv.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        y[j] = (res && flat0 + j >= e.flat_lo && flat0 + j < e.flat_hi) ? res[j] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(v[j]) + pb[j];
                if (full_ok) {
                    if (out) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<float4*>(out)[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
                    }
                    if (out2) {
                        if (e.act == ACT_TANH) {          // (the switch outside the element loop: see the staged epilogue)
#pragma unroll
                            for (int j = 0; j < 32; ++j) y[j] = tanhf(y[j]);
                        } else if (snake) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) { const float sn = __sinf(pb[256 + j] * y[j]); y[j] = fmaf(pb[512 + j] * sn, sn, y[j]); }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<float4*>(out2)[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (flat0 + j >= e.flat_lo && flat0 + j < e.flat_hi) {
                            if (out) out[j] = y[j];
                            if (out2) {
                                float z = y[j];
                                if (e.act == ACT_TANH) z = tanhf(z);
                                else if (snake) { const float sn = __sinf(pb[256 + j] * z); z = fmaf(pb[512 + j] * sn, sn, z); }
                                out2[j] = z;
                            }
                        }
                    }
                }
            }
        }
    }

    if (tprobe && threadIdx.x == 64) g_foley_times[6] = clock64();
    tc_fence_before();
    __syncthreads();
    if (tprobe && threadIdx.x == 0) g_foley_times[7] = clock64();
    if (kPair || cm > 1) cluster_sync_all();   // the peers may still read this CTA's B half / signal its barriers
    if (warp == 1) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
        else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
        if (tprobe && lane == 0) { g_foley_times[8] = clock64(); g_foley_times[14] = globaltimer_ns(); }
    }
}

}  // namespace foley
