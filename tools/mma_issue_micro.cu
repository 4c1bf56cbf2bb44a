// Issue-rate micro-benchmark for tcgen05.mma (kind::f16, bf16 operands from shared memory, cta_group::1, M = 128):
// how many SM cycles does one MMA of K = 16 cost as a function of N?  One thread per CTA issues `iters` k-blocks of
// 4 MMAs (one 64-deep swizzle-128B k-block, as in csrc/gemm.cuh) back to back on a 4-stage ring of shared-memory tiles
// that is never reloaded: no TMA, no barrier waits, one tcgen05.commit per k-block (mode 0), or only one commit at the
// very end (mode 1).  Reported: clock64 cycles per MMA for one CTA alone and for one CTA on each of the 148 SMs.
//
// Build + run (GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I comfyui-hunyuanvideo-foley_b200/csrc \
//        tools/mma_issue_micro.cu -o tools/_bin/mma_issue_micro && tools/_bin/mma_issue_micro
//
// Round-1's planner assumed "one tcgen05.mma costs ~150 cycles whatever N <= 256" (profiles/r01_experiments.md); the
// VERDICT asked for this proof before more design is hung on it.
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace foley;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int mode, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int STAGES = 4;
    constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    __shared__ uint64_t bars[STAGES + 1];
    __shared__ uint32_t tmem_slot;
    // small finite bf16 values (0x3c00 = 0.0078125 patterns mixed by the index) so the datapath toggles
    for (int i = threadIdx.x; i < (STAGES * (A_BYTES + B_BYTES)) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ ((i * 2654435761u) & 0x00ff00ffu);
    if (threadIdx.x == 0) {
        for (int s = 0; s <= STAGES; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc<256>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc(1, 128, N);
        const unsigned long long t0 = clock64();
        int s = 0;
        for (int i = 0; i < iters; ++i) {
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + s * A_BYTES));
            const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (i | k) != 0);
            if (mode == 0) umma_commit(&bars[s]);
            if (++s == STAGES) s = 0;
        }
        umma_commit(&bars[STAGES]);
        mbar_wait(&bars[STAGES], 0, 0x700);
        const unsigned long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

template <int N>
static void run(int grid, int iters, int mode, unsigned long long* d_cyc) {
    const int smem = 4 * (128 * 128 + N * 128) + 1024;
    cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<unsigned long long> h(grid);
    float best_ms = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        mma_rate_kernel<N><<<grid, 128, smem>>>(iters, mode, d_cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best_ms = ms < best_ms ? ms : best_ms;
    }
    cudaMemcpy(h.data(), d_cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0, mn = ~0ull;
    for (auto v : h) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
    const double n_mma = 4.0 * iters;
    const double tflops = 2.0 * 128 * N * 16 * n_mma * grid / (best_ms * 1e-3) / 1e12;
    printf("{\"N\": %d, \"grid\": %d, \"commit_per_kblock\": %d, \"cycles_per_mma_min\": %.1f, \"cycles_per_mma_max\": %.1f, "
           "\"us_per_kblock\": %.4f, \"tflops_chip\": %.1f}\n",
           N, grid, mode == 0, mn / n_mma, mx / n_mma, best_ms * 1e3 / iters, tflops);
}

int main() {
    unsigned long long* d_cyc = nullptr;
    cudaMalloc(&d_cyc, 1024 * sizeof(unsigned long long));
    const int iters = 4000;
    for (int mode = 0; mode < 2; ++mode)
        for (int grid : {1, 148}) {
            run<32>(grid, iters, mode, d_cyc);
            run<64>(grid, iters, mode, d_cyc);
            run<96>(grid, iters, mode, d_cyc);
            run<128>(grid, iters, mode, d_cyc);
            run<192>(grid, iters, mode, d_cyc);
            run<208>(grid, iters, mode, d_cyc);
            run<256>(grid, iters, mode, d_cyc);
        }
    return 0;
}
