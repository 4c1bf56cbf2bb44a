// Host side of the tcgen05 GEMM: TMA descriptor encoding (driver entry point fetched through the
// runtime so the library has no link-time dependency on libcuda) and the launcher.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <string>

#include "gemm.cuh"
#include "gemm_persistent.cuh"

namespace foley {

using PFN_encodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

enum DType : int { DT_BF16 = 0, DT_F32 = 1 };

inline CUtensorMapL2promotion tma_l2_promotion() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FOLEY_TMA_L2PROMO");
        v = e ? atoi(e) : 3;
    }
    switch (v) {
        case 0: return CU_TENSOR_MAP_L2_PROMOTION_NONE;
        case 1: return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
        case 2: return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        default: return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }
}

// A K-major operand view: element (b, r, k) lives at ptr + b*batch_stride + r*ld + k (elements).
struct Operand {
    const void* ptr = nullptr;
    int dtype = DT_BF16;
    long long k = 0;             // contiguous extent
    long long rows = 0;
    long long batch = 1;
    long long ld = 0;            // elements between rows
    long long batch_stride = 0;  // elements between samples
};

inline bool encode_operand_map(CUtensorMap* out, const Operand& t, int box_rows, std::string* err) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { if (err) *err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
    const int esz = t.dtype == DT_BF16 ? 2 : 4;
    const cuuint32_t box_k = 128 / esz;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(t.k), static_cast<cuuint64_t>(t.rows),
                          static_cast<cuuint64_t>(t.batch > 0 ? t.batch : 1)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(t.ld) * esz,
                             static_cast<cuuint64_t>(t.batch > 1 ? t.batch_stride : t.ld * t.rows) * esz};
    if (strides[1] == 0) strides[1] = strides[0];
    cuuint32_t box[3] = {box_k, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(t.ptr) & 15) || (strides[0] & 15) || (strides[1] & 15)) {
        if (err) *err = "TMA operand must be 16-byte aligned (pointer and strides)";
        return false;
    }
    CUresult r = enc(out, t.dtype == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32,
                     3, const_cast<void*>(t.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, tma_l2_promotion(),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) {
            char buf[256];
            snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): k=%lld rows=%lld batch=%lld ld=%lld bs=%lld",
                     static_cast<int>(r), t.k, t.rows, t.batch, t.ld, t.batch_stride);
            *err = buf;
        }
        return false;
    }
    return true;
}

// Output view of the DiT epilogues for the TMA-store path: element (split, b, r, c) at out + split*split_stride +
// b*batch_stride + r*ld + c.  Box = 128 rows x 128 bytes, 128-byte swizzle (the staging layout of the epilogue).
inline bool encode_out_map(CUtensorMap* out, void* ptr, bool f32, long long n_out, long long rows, long long batch,
                           long long splits, long long ld, long long batch_stride, long long split_stride,
                           std::string* err) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { if (err) *err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
    const int esz = f32 ? 4 : 2;
    if (batch < 1) batch = 1;
    if (splits < 1) splits = 1;
    if (batch_stride <= 0) batch_stride = rows * ld;
    if (split_stride <= 0) split_stride = batch * batch_stride;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(n_out), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch),
                          static_cast<cuuint64_t>(splits)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * esz, static_cast<cuuint64_t>(batch_stride) * esz,
                             static_cast<cuuint64_t>(split_stride) * esz};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(128 / esz), 128, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
        if (err) *err = "GEMM output must be 16-byte aligned (pointer and strides)";
        return false;
    }
    CUresult r = enc(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) {
            char buf[256];
            snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(out) failed (%d): n=%lld rows=%lld batch=%lld splits=%lld ld=%lld",
                     static_cast<int>(r), n_out, rows, batch, splits, ld);
            *err = buf;
        }
        return false;
    }
    return true;
}

// Programmatic dependent launch for every engine kernel (FOLEY_PDL=0 disables).
inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FOLEY_PDL");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

// Launch helper for the non-GEMM kernels of the step: same stream order, plus the programmatic-serialization
// attribute so a kernel's launch latency and prologue overlap its predecessor's tail.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int default_prefetch_distance() {
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("FOLEY_GEMM_PREFETCH");
        v = e ? atoi(e) : 0;
    }
    return v;
}

// Weight-tile multicast across the m-tiles of a cluster (FOLEY_GEMM_MCAST = largest cluster size tried: 1 off, 2, 4).
// Off by default: correct (bit-identical, tools/pair_check.py) but 2-8 % slower than plain tiles on every shape of the step in
// both rounds (profiles/r02_mcast_check.log) although cuBLAS's kernels for these shapes all use 1x4 / 4x1 / 2x4 clusters.
inline int default_mcast() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FOLEY_GEMM_MCAST");
        v = e ? atoi(e) : 1;
        if (v != 2 && v != 4) v = 1;
    }
    return v;
}

// CTA-pair tiles on by default (FOLEY_GEMM_PAIR=0 disables).
inline int default_pair_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FOLEY_GEMM_PAIR");
        v = e ? atoi(e) : 0;   // measured 10-15 % slower than single-CTA tiles at every shape of this model
    }
    return v;
}

struct GemmLaunch {
    Operand a;                 // activations [batch, rows, K]
    const void* w = nullptr;   // weights [N, taps*K], K-major, same dtype as a
    long long n = 0;
    int taps = 1, tap_off0 = 0, tap_stride = 1;
    int splits = 1;
    int bn = 128;              // tile width: 64, 128 or 256
    int dbg_stop = 0;
    int prefetch_b = -1;       // -1: default distance
    int pair = -1;             // CTA-pair (cta_group::2) tiles: -1 default policy, 0 off, 1 on when the shape allows
    int cluster_m = -1;        // weight-tile multicast across this many consecutive m-tiles: -1 default policy, 1 off, 2 / 4
    int max_ctas = 0;          // persistent kernel only: cap on the grid (0 = one CTA per SM).  A GEMM that runs on a side branch of the
                               // step graph under latency-critical kernels should not take every SM
    int f16 = 0;               // EPI_BF16 only: operands, bias and output are IEEE fp16 instead of bf16 (same 16-bit layouts and tensor
                               // maps; kind::f16 instruction-descriptor format 0): the Synchformer runs under fp16 autocast
    long long out_rows = 0;    // output rows per sample (0 -> a.rows); may exceed a.rows (halo rows read as zero)
    GemmEpi epi;
};

template <int BN, bool kTF32, bool kPair = false, bool kF16 = false>
inline cudaError_t launch_gemm_inst(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                                    const GemmArgs& args, dim3 grid, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, kTF32, kPair>;
    auto kern = gemm_tcgen05_kernel<BN, kTF32, kPair, kF16>;
    const int cx = kPair ? 2 : (args.cluster_m > 1 ? args.cluster_m : 1), cy = 1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cs);
        if (cs == cudaStreamCaptureStatusNone) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cx * cy > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cx;
        attr[na].val.clusterDim.y = cy;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, args);
}

// Opts every instantiation into its dynamic shared-memory size up front (must not happen lazily inside a
// stream capture).
inline bool gemm_init_attributes(std::string* err) {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kern, int bytes) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    set(gemm_persistent_kernel<false>, PersistCfg::SMEM_BYTES);
    set(gemm_persistent_kernel<true>, PersistCfg::SMEM_BYTES);
    set(gemm_tcgen05_kernel<64, false, false, true>, GemmCfg<64, false>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<128, false, false, true>, GemmCfg<128, false>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<256, false, false, true>, GemmCfg<256, false>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<64, false>, GemmCfg<64, false>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<128, false>, GemmCfg<128, false>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<256, false>, GemmCfg<256, false>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<128, false, true>, GemmCfg<128, false, true>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<256, false, true>, GemmCfg<256, false, true>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<256, true, true>, GemmCfg<256, true, true>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<64, true>, GemmCfg<64, true>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<128, true>, GemmCfg<128, true>::SMEM_BYTES);
    set(gemm_tcgen05_kernel<256, true>, GemmCfg<256, true>::SMEM_BYTES);
    if (e != cudaSuccess) {
        if (err) *err = std::string("cudaFuncSetAttribute(gemm): ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

inline bool launch_gemm(const GemmLaunch& L, cudaStream_t stream, std::string* err) {
    const bool tf32 = L.a.dtype == DT_F32;
    const int bk = tf32 ? 32 : 64;
    if (L.a.k % bk != 0) { if (err) *err = "GEMM K must be a multiple of the 128-byte k-block"; return false; }
    if (L.n % 16 != 0) { if (err) *err = "GEMM N must be a multiple of 16"; return false; }
    if (L.epi.mode == EPI_DAC && !tf32) { if (err) *err = "the DAC epilogue exists for fp32 / tf32 operands only"; return false; }
    if (L.f16 && (tf32 || L.epi.mode != EPI_BF16)) { if (err) *err = "fp16 operands exist for the 16-bit output mode only"; return false; }
    const long long out_rows_c = L.out_rows > 0 ? L.out_rows : L.a.rows;
    const long long mt = ((out_rows_c + 127) / 128) * (L.a.batch > 0 ? L.a.batch : 1);
    const long long nt = (L.n + L.bn - 1) / L.bn;
    // CTA-pair (cta_group::2) tiles: need an even number of m-tiles and a 128/256-wide tile
    int pair = L.pair >= 0 ? L.pair : default_pair_mode();
    if (mt % 2 != 0 || (L.bn != 256 && L.bn != 128) || (tf32 && L.bn != 256) || L.f16) pair = 0;
    // multicast clusters: cm consecutive m-tiles (grid.x) share every weight tile; bf16 single-CTA tiles only
    int cm = L.cluster_m > 0 ? L.cluster_m : default_mcast();
    if (pair || tf32 || L.bn < 128 || L.f16) cm = 1;
    while (cm > 1 && mt % cm != 0) cm >>= 1;
    {   // grids that go to the persistent kernel (below) keep whole-tile loads
        static int persist_env = -1;
        if (persist_env < 0) { const char* ev = getenv("FOLEY_GEMM_PERSIST"); persist_env = ev ? atoi(ev) : 1; }
        int sms_ = 148;
        const long long tiles_ = mt * nt * (L.splits < 1 ? 1 : L.splits);
        const bool mode_ok_ = L.epi.mode == EPI_BF16 || L.epi.mode == EPI_SWIGLU || (L.epi.mode == EPI_F32 && persist_env >= 2);
        if (persist_env && mode_ok_ && !tf32 && L.bn == 256 && !pair && tiles_ * 2 >= static_cast<long long>(sms_) * 3) cm = 1;
    }
    const int cy = 1;
    CUtensorMap ma, mb;
    if (!encode_operand_map(&ma, L.a, 128 / cy, err)) return false;
    Operand wb;
    wb.ptr = L.w; wb.dtype = L.a.dtype; wb.k = L.a.k * L.taps; wb.rows = L.n; wb.batch = 1;
    wb.ld = wb.k; wb.batch_stride = wb.k * wb.rows;
    if (!encode_operand_map(&mb, wb, pair ? L.bn / 2 : L.bn / cm, err)) return false;

    const long long out_rows = L.out_rows > 0 ? L.out_rows : L.a.rows;
    CUtensorMap mc = ma;   // EPI_DAC stores directly; the DiT epilogues leave through TMA stores
    if (L.epi.mode != EPI_DAC) {
        const bool f32 = L.epi.mode == EPI_F32;
        if (!encode_out_map(&mc, L.epi.out, f32, L.epi.mode == EPI_SWIGLU ? L.n / 2 : L.n, out_rows,
                            L.a.batch > 0 ? L.a.batch : 1, L.splits < 1 ? 1 : L.splits, L.epi.ldo, L.epi.out_batch_stride,
                            L.epi.split_stride, err))
            return false;
    }
    GemmArgs args;
    args.rows = static_cast<int>(out_rows);
    args.n = static_cast<int>(L.n);
    args.kb_per_tap = static_cast<int>(L.a.k / bk);
    args.taps = L.taps;
    args.tap_off0 = L.tap_off0;
    args.tap_stride = L.tap_stride;
    args.splits = L.splits < 1 ? 1 : L.splits;
    args.epi = L.epi;
    args.dbg_stop = L.dbg_stop;
    args.prefetch_b = L.prefetch_b < 0 ? default_prefetch_distance() : L.prefetch_b;
    args.pf_mod = 1;
    args.cluster_m = cm;
    const int m_tiles = static_cast<int>((out_rows + 127) / 128);
    dim3 grid(static_cast<unsigned>(m_tiles * (L.a.batch > 0 ? L.a.batch : 1)),
              static_cast<unsigned>(((nt + cy - 1) / cy) * cy), static_cast<unsigned>(args.splits));
    cudaError_t e = cudaSuccess;
    // Grids of more than ~1.5 waves of 256-wide bf16 tiles go to the persistent kernel (epilogue of tile j under the
    // mainloop of tile j+1, one prologue per SM).  FOLEY_GEMM_PERSIST=0 disables.
    {
        static int persist = -1, sms = 0;
        if (persist < 0) {
            const char* ev = getenv("FOLEY_GEMM_PERSIST");
            persist = ev ? atoi(ev) : 1;
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        const long long tiles = static_cast<long long>(grid.x) * grid.y * grid.z;
        // (fp32 K-split partials stay on the one-tile kernel unless FOLEY_GEMM_PERSIST=2: their 128 KB tile needs two
        // staging rounds here and measured 3 % slower; bf16 / SwiGLU outputs measured 5-25 % faster)
        const bool mode_ok = L.epi.mode == EPI_BF16 || L.epi.mode == EPI_SWIGLU || (L.epi.mode == EPI_F32 && persist >= 2);
        if (persist && mode_ok && !tf32 && L.bn == 256 && !pair && (L.dbg_stop == 0 || L.dbg_stop == 3) && sms > 0 &&
            tiles * 2 >= static_cast<long long>(sms) * 3 && tiles < (1LL << 30)) {
            static bool p_attr = false;
            if (!p_attr) {   // (Engine::create sets it up front; this covers stand-alone foley_gemm calls)
                cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
                cudaStreamIsCapturing(stream, &cs);
                if (cs == cudaStreamCaptureStatusNone) {
                    cudaFuncSetAttribute(gemm_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PersistCfg::SMEM_BYTES);
                    cudaFuncSetAttribute(gemm_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PersistCfg::SMEM_BYTES);
                    p_attr = true;
                }
            }
            PersistTiles pt;
            pt.m_tiles = m_tiles;
            pt.mb_total = static_cast<int>(grid.x);
            pt.n_tiles = static_cast<int>(grid.y);
            pt.num_tiles = static_cast<int>(tiles);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(static_cast<unsigned>(std::min<long long>(tiles, L.max_ctas > 0 ? std::min(L.max_ctas, sms) : sms)));
            cfg.blockDim = dim3(PersistCfg::THREADS);
            cfg.dynamicSmemBytes = PersistCfg::SMEM_BYTES;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pdl_enabled() ? 1 : 0;
            e = L.f16 ? cudaLaunchKernelEx(&cfg, gemm_persistent_kernel<true>, ma, mb, mc, args, pt)
                          : cudaLaunchKernelEx(&cfg, gemm_persistent_kernel<false>, ma, mb, mc, args, pt);
            if (e != cudaSuccess) {
                if (err) *err = std::string("persistent GEMM launch failed: ") + cudaGetErrorString(e);
                return false;
            }
            return true;
        }
    }
    if (!tf32 && L.f16) {          // fp16 operands (Synchformer): single-CTA tiles
        if (L.bn == 64) e = launch_gemm_inst<64, false, false, true>(ma, mb, mc, args, grid, stream);
        else if (L.bn == 128) e = launch_gemm_inst<128, false, false, true>(ma, mb, mc, args, grid, stream);
        else if (L.bn == 256) e = launch_gemm_inst<256, false, false, true>(ma, mb, mc, args, grid, stream);
        else { if (err) *err = "unsupported BN"; return false; }
    } else if (!tf32) {
        if (L.bn == 64) e = launch_gemm_inst<64, false>(ma, mb, mc, args, grid, stream);
        else if (L.bn == 128) e = pair ? launch_gemm_inst<128, false, true>(ma, mb, mc, args, grid, stream)
                                       : launch_gemm_inst<128, false>(ma, mb, mc, args, grid, stream);
        else if (L.bn == 256) e = pair ? launch_gemm_inst<256, false, true>(ma, mb, mc, args, grid, stream)
                                       : launch_gemm_inst<256, false>(ma, mb, mc, args, grid, stream);
        // (in-between widths that put one tile on every SM — 208 columns x 148 CTAs for the w1|w3 shape, 192 for fc1 — were
        //  measured in round 2: no gain, 29.5 vs 28.9 us; the mainloop is bound by the aggregate L2 -> shared-memory copy
        //  rate of all CTAs (~17.5 TB/s), not by the MMA width: profiles/r02_experiments.md)
        else { if (err) *err = "unsupported BN"; return false; }
    } else {
        if (L.bn == 64) e = launch_gemm_inst<64, true>(ma, mb, mc, args, grid, stream);
        else if (L.bn == 128) e = launch_gemm_inst<128, true>(ma, mb, mc, args, grid, stream);
        else if (L.bn == 256) e = pair ? launch_gemm_inst<256, true, true>(ma, mb, mc, args, grid, stream)
                                       : launch_gemm_inst<256, true>(ma, mb, mc, args, grid, stream);
        else { if (err) *err = "unsupported BN"; return false; }
    }
    if (e != cudaSuccess) {
        if (err) *err = std::string("GEMM launch failed: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

}  // namespace foley
