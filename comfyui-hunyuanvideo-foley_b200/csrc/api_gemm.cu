// C-ABI: error/version plumbing and the low-level GEMM entry point (include/foley_b200.h).
#include "common.cuh"
#include "gemm_host.cuh"

namespace foley {
std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
}  // namespace foley

using namespace foley;

extern "C" const char* foley_last_error(void) { return last_error_ref().c_str(); }
extern "C" const char* foley_version(void) { return "foley_b200 0.1 sm_100a (tcgen05+TMA)"; }

extern "C" foley_status foley_gemm(const void* a, int32_t dtype, int64_t batch, int64_t rows, int64_t k,
                                   int64_t lda, int64_t a_batch_stride, const void* w, int64_t n,
                                   int32_t taps, int32_t tap_off0, int32_t tap_stride, int32_t splits,
                                   int32_t bn, int32_t mode, int32_t act, const void* bias, void* out,
                                   int64_t ldo, int64_t out_batch_stride, int64_t split_stride,
                                   void* stream) {
    if (!a || !w || !out) return fail(FOLEY_ERR_INVALID, "foley_gemm: null pointer");
    if (dtype != FOLEY_DT_BF16 && dtype != FOLEY_DT_F32) return fail(FOLEY_ERR_INVALID, "foley_gemm: dtype");
    if (mode < 0 || mode > 2) return fail(FOLEY_ERR_INVALID, "foley_gemm: mode must be 0,1,2");
    GemmLaunch L;
    L.a.ptr = a;
    L.a.dtype = dtype == FOLEY_DT_BF16 ? DT_BF16 : DT_F32;
    L.a.k = k; L.a.rows = rows; L.a.batch = batch; L.a.ld = lda; L.a.batch_stride = a_batch_stride;
    L.w = w; L.n = n;
    L.taps = taps; L.tap_off0 = tap_off0; L.tap_stride = tap_stride;
    L.splits = splits; L.bn = bn & 0xFFFF;
    L.dbg_stop = (bn >> 16) & 0xF;  // bring-up aid: upper bits of bn select a partial pipeline
    L.epi.mode = mode; L.epi.act = act; L.epi.bias = bias; L.epi.out = out; L.epi.ldo = ldo;
    L.epi.out_batch_stride = out_batch_stride; L.epi.split_stride = split_stride;
    std::string err;
    if (!launch_gemm(L, static_cast<cudaStream_t>(stream), &err)) return fail(FOLEY_ERR_CUDA, err);
    return FOLEY_OK;
}

// Reads and clears the device debug words ([0] = first timed-out mbarrier wait code).
extern "C" foley_status foley_debug_flags(uint32_t* out4) {
    unsigned int h[4] = {0, 0, 0, 0};
    FOLEY_CUDA_OK(cudaMemcpyFromSymbol(h, g_foley_dbg, sizeof h));
    unsigned int z[4] = {0, 0, 0, 0};
    FOLEY_CUDA_OK(cudaMemcpyToSymbol(g_foley_dbg, z, sizeof z));
    for (int i = 0; i < 4; ++i) out4[i] = h[i];
    return FOLEY_OK;
}
