// Persistent variant of the tcgen05 + TMA GEMM for grids of more than one wave (every GEMM of a batch >= 4 step, the
// stacked modulation GEMM at batch 1): one CTA per SM walks over its tiles, the accumulator is double-buffered in
// TMEM (2 x 256 of the 512 columns), so the epilogue of tile j (tcgen05.ld, fused math, stores) runs while the TMA / MMA
// warps are already in the mainloop of tile j+1, and the per-CTA prologue (barrier init, TMEM allocation, descriptor
// prefetch) is paid once.  In the one-tile-per-CTA kernel those two cost ~5 us per tile next to 8-24 us of mainloop
// (profiles/r01_experiments.md).  Same arithmetic and the same epilogue math as gemm_tcgen05_kernel<256, false>:
// results are bit-identical.  bf16 operands, 128 x 256 tiles, modes EPI_BF16 / EPI_SWIGLU / EPI_F32.
//
// Roles (352 threads): warp 0 = TMA producer of weight tiles, warp 10 = TMA producer of activation tiles, warp 1 = MMA
// issuer, warps 2..9 = epilogue.  Barriers: full / empty per ring stage (3 stages), acc_full / acc_empty per TMEM buffer.
#pragma once
#include "gemm.cuh"

namespace foley {

struct PersistCfg {
    static constexpr int BN = 256, BM = 128, BK = 64, UMMA_K = 16;
    static constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = 3;                                  // deeper rings measured no faster
    static constexpr int STAGING_BYTES = 4 * 16384;                   // bf16 tile = four [128 x 128 B] store boxes
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int THREADS = 352;
};

struct PersistTiles {
    int m_tiles;        // 128-row tiles per sample
    int mb_total;       // m_tiles * batch
    int n_tiles;        // 256-column tiles
    int num_tiles;      // mb_total * n_tiles * splits
};

template <bool kF16 = false>      // fp16 operands / bias / output (see gemm.cuh): a separate instantiation
__global__ void __launch_bounds__(PersistCfg::THREADS, 1)
gemm_persistent_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                       const __grid_constant__ CUtensorMap tm_c, const GemmArgs g, const PersistTiles pt) {
    using Cfg = PersistCfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::STAGES * Cfg::A_BYTES;
    uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + Cfg::STAGES;     // [STAGES]
    uint64_t* acc_full = bars + 2 * Cfg::STAGES;  // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_total = g.taps * g.kb_per_tap;

    // tile t -> (split, n-tile, batch, m-tile); m fastest so that neighbouring CTAs share a weight tile in L2
    auto decode = [&](int t, int* split, int* n0, int* batch, int* m0) {
        const int per_split = pt.mb_total * pt.n_tiles;
        *split = t / per_split;
        const int r = t - *split * per_split;
        const int ni = r / pt.mb_total, mb = r - ni * pt.mb_total;
        *n0 = ni * Cfg::BN;
        *batch = mb / pt.m_tiles;
        *m0 = (mb - *batch * pt.m_tiles) * Cfg::BM;
    };
    auto kb_range = [&](int split, int* kb_begin, int* kb_end) {
        *kb_begin = static_cast<int>((static_cast<long long>(kb_total) * split) / g.splits);
        *kb_end = static_cast<int>((static_cast<long long>(kb_total) * (split + 1)) / g.splits);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        tma_prefetch_desc(&tm_c);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 1); }
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc<512>(tmem_ptr_smem);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ------------------------------------------------------------- weight tiles (static data: no dependency wait)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            long long n_issued = 0;
            for (int t = blockIdx.x; t < pt.num_tiles; t += gridDim.x) {
                int split, n0, batch, m0, kb0, kb1;
                decode(t, &split, &n0, &batch, &m0);
                kb_range(split, &kb0, &kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++n_issued) {
                    if (n_issued >= Cfg::STAGES) { if (!mbar_wait(&empty_bar[s], ph ^ 1, 0x900)) return; }
                    mbar_expect_tx(&full_bar[s], Cfg::B_BYTES);
                    tma_load_3d(smem_b + s * Cfg::B_BYTES, &tm_b, &full_bar[s], kb * Cfg::BK, n0, 0);
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 10) {
        // ------------------------------------------------------------- activation tiles
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            long long n_issued = 0;
            pdl_wait();   // activations are the predecessor's output
            for (int t = blockIdx.x; t < pt.num_tiles; t += gridDim.x) {
                int split, n0, batch, m0, kb0, kb1;
                decode(t, &split, &n0, &batch, &m0);
                kb_range(split, &kb0, &kb1);
                int tap = kb0 / g.kb_per_tap;
                int kk = kb0 - tap * g.kb_per_tap;
                int arow = m0 + g.tap_off0 + tap * g.tap_stride;
                for (int kb = kb0; kb < kb1; ++kb, ++n_issued) {
                    if (n_issued >= Cfg::STAGES) { if (!mbar_wait(&empty_bar[s], ph ^ 1, 0x980)) return; }
                    mbar_expect_tx(&full_bar[s], Cfg::A_BYTES);
                    tma_load_3d(smem_a + s * Cfg::A_BYTES, &tm_a, &full_bar[s], kk * Cfg::BK, arow, batch);
                    if (++kk == g.kb_per_tap) { kk = 0; arow += g.tap_stride; }
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(kF16 ? 0 : 1, Cfg::BM, Cfg::BN);
            int s = 0;
            uint32_t ph = 0;
            int j = 0;   // local tile counter
            for (int t = blockIdx.x; t < pt.num_tiles; t += gridDim.x, ++j) {
                int split, n0, batch, m0, kb0, kb1;
                decode(t, &split, &n0, &batch, &m0);
                kb_range(split, &kb0, &kb1);
                const int buf = j & 1, use = j >> 1;
                if (use > 0) { if (!mbar_wait(&acc_empty[buf], (use - 1) & 1, 0xA00)) return; }   // epilogue released it
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * Cfg::BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!mbar_wait(&full_bar[s], ph, 0xA80)) return;
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + s * Cfg::A_BYTES));
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + s * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < Cfg::BK / Cfg::UMMA_K; ++k)
                        umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ------------------------------------------------------------- epilogue warps 2..9
        const int q = warp & 3, half = (warp - 2) >> 2;
        const GemmEpi& e = g.epi;
        const int bpc = e.mode == EPI_F32 ? 4 : (e.mode == EPI_BF16 ? 2 : 1);
        const int n_out = e.mode == EPI_SWIGLU ? g.n >> 1 : g.n;
        const uint32_t r_local = static_cast<uint32_t>(q * 32 + lane);
        const uint32_t stage_base = smem_u32(staging);
        const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(e.bias);
        pdl_wait();   // outputs may alias buffers the predecessor still reads
        int j = 0;
        for (int t = blockIdx.x; t < pt.num_tiles; t += gridDim.x, ++j) {
            int split, n0, batch, m0;
            decode(t, &split, &n0, &batch, &m0);
            const int buf = j & 1, use = j >> 1;
            const bool ok = mbar_wait(&acc_full[buf], use & 1, 0xB00) && g.dbg_stop == 0;   // dbg 3: hand-off only
            if (t + static_cast<int>(gridDim.x) >= pt.num_tiles) pdl_trigger();   // last tile of this CTA: mainloops are done
            tc_fence_after();
            const int r = m0 + q * 32 + lane;
            const bool row_ok = r < g.rows;
            const long long row_off = static_cast<long long>(batch) * e.out_batch_stride + static_cast<long long>(r) * e.ldo;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg::BN + half * 32;
            if (e.mode != EPI_SWIGLU) {
                // staged epilogue (see gemm.cuh): 64 KB of staging = four [128 x 128 B] store boxes = the whole bf16 tile
                // or half an fp32 tile; a box is rewritten only after the bulk store that read it last has drained
                const int boxes_per_it = e.mode == EPI_F32 ? 2 : 1;          // 64 columns = 256 B (fp32) / 128 B (bf16) per row
                if (threadIdx.x == 64) tma_store_wait_read();                // previous tile's stores
                asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll 1
                for (int it = 0; it < 4; ++it) {
                    const int c0 = half * 32 + it * 64;
                    if (e.mode == EPI_F32 && it == 2) {                      // second half of an fp32 tile reuses the boxes
                        if (threadIdx.x == 64) tma_store_wait_read();
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                    }
                    const int box0 = (it * boxes_per_it) & 3;               // first staging box of this column group
                    if (ok && n0 + c0 < g.n) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_row + it * 64, v);
                        tmem_ld_wait();
                        const uint32_t sw = r_local & 7u;
                        if (e.mode == EPI_F32) {
                            const uint32_t row_addr = stage_base + static_cast<uint32_t>(box0 + half) * 16384u + r_local * 128u;
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj)
                                st_shared_v4(row_addr + ((static_cast<uint32_t>(jj) ^ sw) << 4), v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                        } else {
                            const uint32_t row_addr = stage_base + static_cast<uint32_t>(box0) * 16384u + r_local * 128u;
                            const uint32_t piece0 = static_cast<uint32_t>(half) * 4u;
                            // (the activation switch sits OUTSIDE the element loop: inside, it cost ~60 cycles per element;
                            //  the bias of the 32 columns comes as four 16-byte loads)
                            float a[32];
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj) a[jj] = __uint_as_float(v[jj]);
                            if (bias && kF16) {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const uint4 bw = __ldg(reinterpret_cast<const uint4*>(bias + n0 + c0) + jj);
                                    const __half2* b2 = reinterpret_cast<const __half2*>(&bw);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        a[8 * jj + 2 * k] += __low2float(b2[k]);
                                        a[8 * jj + 2 * k + 1] += __high2float(b2[k]);
                                    }
                                }
                            } else if (bias) {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const uint4 bw = __ldg(reinterpret_cast<const uint4*>(bias + n0 + c0) + jj);
                                    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bw);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        a[8 * jj + 2 * k] += __low2float(b2[k]);
                                        a[8 * jj + 2 * k + 1] += __high2float(b2[k]);
                                    }
                                }
                            }
                            if constexpr (kF16) {
                                if (e.act == ACT_GELU_ERF) {
#pragma unroll
                                    for (int jj = 0; jj < 32; ++jj) { const float x = f16_round(a[jj]); a[jj] = 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
                                } else if (e.act != ACT_NONE) {
#pragma unroll
                                    for (int jj = 0; jj < 32; ++jj) a[jj] = apply_act(f16_round(a[jj]), e.act);
                                }
                            } else if (e.act == ACT_SILU) {
#pragma unroll
                                for (int jj = 0; jj < 32; ++jj) { const float x = bf16_round(a[jj]); a[jj] = __fdividef(x, 1.0f + __expf(-x)); }
                            } else if (e.act == ACT_GELU_TANH) {
#pragma unroll
                                for (int jj = 0; jj < 32; ++jj) a[jj] = gelu_tanh_f(bf16_round(a[jj]));
                            } else if (e.act != ACT_NONE) {
#pragma unroll
                                for (int jj = 0; jj < 32; ++jj) a[jj] = apply_act(bf16_round(a[jj]), e.act);
                            }
                            uint32_t packed[16];
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) packed[jj] = kF16 ? pack_f16x2(a[2 * jj], a[2 * jj + 1]) : pack_bf16x2(a[2 * jj], a[2 * jj + 1]);
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                st_shared_v4(row_addr + (((piece0 + jj) ^ sw) << 4), packed[4 * jj], packed[4 * jj + 1],
                                             packed[4 * jj + 2], packed[4 * jj + 3]);
                        }
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (threadIdx.x == 64 && ok) {
                        for (int b = 0; b < boxes_per_it; ++b) {
                            const int col_elem = n0 + it * 64 + b * 32;      // fp32: 32 columns per box; bf16: one 64-column box
                            if (col_elem < n_out) tma_store_4d(&tm_c, staging + (box0 + b) * 16384, col_elem, m0, batch, split);
                        }
                        tma_store_commit();
                    }
                }
            } else {
#pragma unroll 1
                for (int c0 = half * 32; c0 < Cfg::BN && ok; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + (c0 - half * 32), v);
                    tmem_ld_wait();
                    const int col = n0 + c0;
                    if (!ok || !row_ok || col >= g.n) continue;
                    if (e.mode == EPI_SWIGLU) {
                        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out) + row_off + (col >> 1);
                        uint32_t packed[8];
#pragma unroll
                        for (int jj = 0; jj < 32; jj += 4) {
                            const float g0 = bf16_round(__uint_as_float(v[jj])), u0 = bf16_round(__uint_as_float(v[jj + 1]));
                            const float g1 = bf16_round(__uint_as_float(v[jj + 2])), u1 = bf16_round(__uint_as_float(v[jj + 3]));
                            const float s0 = bf16_round(__fdividef(g0, 1.0f + __expf(-g0))) * u0;
                            const float s1 = bf16_round(__fdividef(g1, 1.0f + __expf(-g1))) * u1;
                            packed[jj >> 2] = pack_bf16x2(s0, s1);
                        }
                        uint4* dst = reinterpret_cast<uint4*>(out);
                        dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                    }
                }
            }
            // hand the TMEM buffer back to the MMA warp once all eight warps have read it
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) mbar_arrive(&acc_empty[buf]);
            (void)bpc;
        }
        if (threadIdx.x == 64) tma_store_wait_read();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace foley
