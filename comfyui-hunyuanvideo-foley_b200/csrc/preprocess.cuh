// Frame preprocessing for the condition encoders as two fused HBM-bound kernels.
//
// Replaces, bit for bit, what the reference does on the CPU frame by frame (nodes.py:293-317, 184-196; utils.py:270-273):
//   IMAGE fp32 [N,H,W,3] --(x*255).byte()--> uint8 --index_select(8 / 25 fps picks)--> torchvision v2
//   Resize(bicubic, antialias=True) [-> CenterCrop] -> ToDtype(float32, scale=True) -> Normalize(0.5, 0.5).
// The resize is ATen's uint8 antialiased bicubic (UpSampleKernel.cpp: _compute_indices_int16_weights_aa +
// upsample_avx_bilinear_bicubic_uint8): separable, horizontal pass first, int16 fixed-point weights per axis with a
// shared precision, (acc + 2^(p-1)) >> p, saturation to uint8 between and after the passes.  The filter banks are
// built on the host in double precision exactly as ATen does; the kernels do the integer arithmetic.
//
//   pass 1  resize_rows_kernel   one CTA per (input row, frame): the fp32 HWC row is read once, coalesced, quantised to
//                                uint8 into shared memory, and every output column of the three channels is produced
//                                from there -> uint8 planar [T,3,H,out_w]
//   pass 2  resize_cols_normalize_kernel   one thread per output pixel of the crop window, coalesced along x over the
//                                uint8 intermediate; writes the normalised fp32 [T,3,crop_h,crop_w] the encoders take
//
// Algorithmic bytes per output frame: 12*H*W read (fp32 RGB) + 3*H*out_w written and ~taps/stride times re-read from L2
// + 12*crop_h*crop_w written.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace foley {

struct ResizeBank {            // one axis of the separable filter
    int in_size = 0, out_size = 0, max_interp = 0, precision = 0;
    std::vector<int32_t> xmin, xsize;
    std::vector<int16_t> w;    // [out_size][max_interp]
};

inline double cubic_aa_filter(double x) {   // Keys cubic, a = -0.5 (HelperInterpCubic::aa_filter)
    const double a = -0.5;
    x = std::fabs(x);
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
    if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
    return 0.0;
}

// HelperInterpBase::_compute_index_ranges_weights<double> + _compute_index_ranges_int16_weights, align_corners = false,
// scale = in / out (torchvision passes sizes, not scale factors).
inline ResizeBank make_resize_bank(int in_size, int out_size) {
    ResizeBank b;
    b.in_size = in_size; b.out_size = out_size;
    const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
    const int interp_size = 4;
    const double support = scale >= 1.0 ? (interp_size * 0.5) * scale : interp_size * 0.5;
    b.max_interp = static_cast<int>(std::ceil(support)) * 2 + 1;
    const double invscale = scale >= 1.0 ? 1.0 / scale : 1.0;
    b.xmin.resize(out_size); b.xsize.resize(out_size);
    std::vector<double> wd(static_cast<size_t>(out_size) * b.max_interp, 0.0);
    double wt_max = 0.0;
    for (int i = 0; i < out_size; ++i) {
        const double center = scale * (i + 0.5);
        const int64_t lo = std::max<int64_t>(static_cast<int64_t>(center - support + 0.5), 0);
        int64_t size = std::min<int64_t>(static_cast<int64_t>(center + support + 0.5), in_size) - lo;
        size = std::min<int64_t>(std::max<int64_t>(size, 0), b.max_interp);
        double* wp = &wd[static_cast<size_t>(i) * b.max_interp];
        double total = 0.0;
        for (int64_t j = 0; j < size; ++j) {
            wp[j] = cubic_aa_filter((j + lo - center + 0.5) * invscale);
            total += wp[j];
        }
        if (total != 0.0)
            for (int64_t j = 0; j < size; ++j) { wp[j] /= total; wt_max = std::max(wt_max, wp[j]); }
        b.xmin[i] = static_cast<int32_t>(lo);
        b.xsize[i] = static_cast<int32_t>(size);
    }
    int precision = 0;
    for (; precision < 22; ++precision)
        if (static_cast<int>(0.5 + wt_max * (1 << (precision + 1))) >= (1 << 15)) break;
    b.precision = precision;
    b.w.resize(wd.size());
    for (size_t k = 0; k < wd.size(); ++k) {
        const double v = wd[k] * (1 << precision);
        b.w[k] = static_cast<int16_t>(v < 0 ? v - 0.5 : v + 0.5);
    }
    return b;
}

__device__ __forceinline__ uint8_t quantize_u8(float v) {   // (image * 255.0).byte(): truncation, C-style wrap
    return static_cast<uint8_t>(static_cast<int>(truncf(__fmul_rn(v, 255.0f))));
}
__device__ __forceinline__ uint8_t clip_u8(int v) { return static_cast<uint8_t>(min(max(v, 0), 255)); }

// pass 1 — grid (H, T), any block size.  image: fp32 [n,H,W,3]; tmp: uint8 [T,3,H,out_w].
// The row is staged as one packed 0x00BBGGRR word per pixel: one shared-memory read per filter tap serves the three
// channels (byte-planar staging made this kernel L1/LSU-bound at 93 % L1TEX throughput, 37 % of HBM: profiles/).
__global__ void resize_rows_kernel(const float* __restrict__ image, const int* __restrict__ frame_idx, int H, int W,
                                   int out_w, const int* __restrict__ xmin, const int* __restrict__ xsize,
                                   const int16_t* __restrict__ wts, int max_interp, int precision, int col_lo, int col_hi,
                                   uint8_t* __restrict__ tmp) {
    extern __shared__ uint32_t row_px[];   // [W]
    const int y = blockIdx.x, t = blockIdx.y;
    const float* src = image + (static_cast<long long>(frame_idx[t]) * H + y) * W * 3;
    if ((W & 3) == 0) {   // 4 pixels = 12 floats = three 16-byte loads per thread and step (rows are 16-byte aligned)
        const float4* src4 = reinterpret_cast<const float4*>(src);
        for (int p4 = threadIdx.x; p4 < (W >> 2); p4 += blockDim.x) {
            const float4 a = __ldg(src4 + p4 * 3), b = __ldg(src4 + p4 * 3 + 1), c = __ldg(src4 + p4 * 3 + 2);
            uint4 o;
            o.x = quantize_u8(a.x) | (quantize_u8(a.y) << 8) | (quantize_u8(a.z) << 16);
            o.y = quantize_u8(a.w) | (quantize_u8(b.x) << 8) | (quantize_u8(b.y) << 16);
            o.z = quantize_u8(b.z) | (quantize_u8(b.w) << 8) | (quantize_u8(c.x) << 16);
            o.w = quantize_u8(c.y) | (quantize_u8(c.z) << 8) | (quantize_u8(c.w) << 16);
            reinterpret_cast<uint4*>(row_px)[p4] = o;
        }
    } else {
        for (int x = threadIdx.x; x < W; x += blockDim.x)
            row_px[x] = quantize_u8(__ldg(src + 3 * x)) | (quantize_u8(__ldg(src + 3 * x + 1)) << 8) |
                        (quantize_u8(__ldg(src + 3 * x + 2)) << 16);
    }
    __syncthreads();
    const int round_add = 1 << (precision - 1);
    for (int xo = col_lo + threadIdx.x; xo < col_hi; xo += blockDim.x) {
        const int x0 = xmin[xo], n = xsize[xo];
        const int16_t* w = wts + static_cast<long long>(xo) * max_interp;
        int a0 = round_add, a1 = round_add, a2 = round_add;
        for (int j = 0; j < n; ++j) {
            const int wj = w[j];
            const uint32_t px = row_px[x0 + j];
            a0 += wj * static_cast<int>(px & 255u);
            a1 += wj * static_cast<int>((px >> 8) & 255u);
            a2 += wj * static_cast<int>(px >> 16);
        }
        const long long o = ((static_cast<long long>(t) * 3) * H + y) * out_w + xo;
        tmp[o] = clip_u8(a0 >> precision);
        tmp[o + static_cast<long long>(H) * out_w] = clip_u8(a1 >> precision);
        tmp[o + 2LL * H * out_w] = clip_u8(a2 >> precision);
    }
}

__device__ __forceinline__ float normalize_px(int u) {
    // ToDtype(float32, scale=True): x.float() * (1/255); Normalize(0.5, 0.5): (x - 0.5) / 0.5   (all fp32, separately rounded)
    const float f = __fmul_rn(static_cast<float>(u), 0.003921568859368563f);
    return __fdiv_rn(__fsub_rn(f, 0.5f), 0.5f);
}

// pass 2 — one thread per FOUR horizontally adjacent output pixels (one 32-bit read of the uint8 intermediate per
// tap and one 16-byte store).  tmp: uint8 [T,3,H,out_w] -> out fp32 [T,3,crop_h,crop_w] normalised.  quad = 1 needs
// out_w, crop_left and crop_w to be multiples of 4 (the host picks); quad = 0 is the one-pixel-per-thread form.
__global__ void resize_cols_normalize_kernel(const uint8_t* __restrict__ tmp, int T, int H, int out_w, int crop_top,
                                             int crop_left, int crop_h, int crop_w, const int* __restrict__ ymin,
                                             const int* __restrict__ ysize, const int16_t* __restrict__ wts,
                                             int max_interp, int precision, int quad, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int px_per = quad ? 4 : 1;
    const int wq = crop_w / px_per;
    const long long total = static_cast<long long>(T) * 3 * crop_h * wq;
    if (i >= total) return;
    const int x = static_cast<int>(i % wq) * px_per;
    const int y = static_cast<int>((i / wq) % crop_h);
    const long long tc = i / (static_cast<long long>(wq) * crop_h);
    const uint8_t* plane = tmp + tc * H * out_w + (x + crop_left);
    const int yo = y + crop_top;
    // (equal sizes: ATen skips the pass; the bank then is the exact identity 2^p * delta, same bytes)
    const int y0 = ymin[yo], n = ysize[yo];
    const int16_t* w = wts + static_cast<long long>(yo) * max_interp;
    const int round_add = 1 << (precision - 1);
    float* dst = out + (tc * crop_h + y) * crop_w + x;
    if (quad) {
        int a0 = round_add, a1 = round_add, a2 = round_add, a3 = round_add;
        const uint8_t* col = plane + static_cast<long long>(y0) * out_w;
        for (int j = 0; j < n; ++j, col += out_w) {
            const int wj = w[j];
            const uint32_t p = *reinterpret_cast<const uint32_t*>(col);
            a0 += wj * static_cast<int>(p & 255u);
            a1 += wj * static_cast<int>((p >> 8) & 255u);
            a2 += wj * static_cast<int>((p >> 16) & 255u);
            a3 += wj * static_cast<int>(p >> 24);
        }
        *reinterpret_cast<float4*>(dst) = make_float4(normalize_px(clip_u8(a0 >> precision)), normalize_px(clip_u8(a1 >> precision)),
                                                      normalize_px(clip_u8(a2 >> precision)), normalize_px(clip_u8(a3 >> precision)));
    } else {
        int acc = round_add;
        for (int j = 0; j < n; ++j) acc += static_cast<int>(w[j]) * plane[static_cast<long long>(y0 + j) * out_w];
        *dst = normalize_px(clip_u8(acc >> precision));
    }
}

// Host driver.  frame_idx: T host indices into the image batch.  The resize target is (resize_h, resize_w); the
// output is the crop window [crop_top, +out_h) x [crop_left, +out_w) of it.
inline foley_status preprocess_frames(const float* image, int n_frames, int H, int W, const int32_t* frame_idx, int T,
                                      int resize_h, int resize_w, int crop_top, int crop_left, int out_h, int out_w,
                                      float* out, cudaStream_t st) {
    if (!image || !out || !frame_idx) return fail(FOLEY_ERR_INVALID, "preprocess_frames: null argument");
    if (T < 1 || n_frames < 1 || H < 1 || W < 1 || resize_h < 1 || resize_w < 1 || out_h < 1 || out_w < 1)
        return fail(FOLEY_ERR_INVALID, "preprocess_frames: empty shape");
    if (crop_top < 0 || crop_left < 0 || crop_top + out_h > resize_h || crop_left + out_w > resize_w)
        return fail(FOLEY_ERR_INVALID, "preprocess_frames: crop window outside the resized frame");
    if (W > 12 * 1024) return fail(FOLEY_ERR_UNSUPPORTED, "preprocess_frames: frame wider than 12288 pixels");
    for (int t = 0; t < T; ++t)
        if (frame_idx[t] < 0 || frame_idx[t] >= n_frames) return fail(FOLEY_ERR_INVALID, "preprocess_frames: frame index out of range");
    const ResizeBank bx = make_resize_bank(W, resize_w), by = make_resize_bank(H, resize_h);
    // One stream-ordered scratch block: [tmp u8 | idx | xmin | xsize | ymin | ysize | wx | wy], 16-byte aligned parts.
    auto al = [](size_t n) { return (n + 15) & ~size_t(15); };
    const size_t n_tmp = al(static_cast<size_t>(T) * 3 * H * resize_w);
    const size_t o_idx = n_tmp, o_xmin = o_idx + al(sizeof(int) * T), o_xsize = o_xmin + al(sizeof(int) * resize_w);
    const size_t o_ymin = o_xsize + al(sizeof(int) * resize_w), o_ysize = o_ymin + al(sizeof(int) * resize_h);
    const size_t o_wx = o_ysize + al(sizeof(int) * resize_h), o_wy = o_wx + al(sizeof(int16_t) * bx.w.size());
    const size_t total = o_wy + al(sizeof(int16_t) * by.w.size());
    uint8_t* blk = nullptr;
    FOLEY_CUDA_OK(cudaMallocAsync(&blk, total, st));
    auto bail = [&](cudaError_t e, const char* what) {
        cudaFreeAsync(blk, st);
        return fail(FOLEY_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
    {   // all tables in ONE pageable upload (each such copy synchronises with the stream: seven of them cost more than the kernels)
        std::vector<uint8_t> host(total - n_tmp);
        auto put = [&](size_t off, const void* src, size_t bytes) { std::memcpy(host.data() + (off - n_tmp), src, bytes); };
        put(o_idx, frame_idx, sizeof(int) * T);
        put(o_xmin, bx.xmin.data(), sizeof(int) * resize_w);
        put(o_xsize, bx.xsize.data(), sizeof(int) * resize_w);
        put(o_ymin, by.xmin.data(), sizeof(int) * resize_h);
        put(o_ysize, by.xsize.data(), sizeof(int) * resize_h);
        put(o_wx, bx.w.data(), sizeof(int16_t) * bx.w.size());
        put(o_wy, by.w.data(), sizeof(int16_t) * by.w.size());
        cudaError_t e = cudaMemcpyAsync(blk + n_tmp, host.data(), host.size(), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return bail(e, "preprocess_frames: table upload");   // (staged before the call returns)
    }
    // pass 1 only for the columns the crop window keeps
    resize_rows_kernel<<<dim3(H, T), 256, static_cast<size_t>(W) * 4, st>>>(
        image, reinterpret_cast<const int*>(blk + o_idx), H, W, resize_w, reinterpret_cast<const int*>(blk + o_xmin),
        reinterpret_cast<const int*>(blk + o_xsize), reinterpret_cast<const int16_t*>(blk + o_wx), bx.max_interp, bx.precision,
        crop_left, crop_left + out_w, blk);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bail(e, "resize_rows_kernel");
    // four pixels per thread when every 4-pixel group is 4-byte aligned in the intermediate and 16-byte aligned in the output
    const int quad = (resize_w % 4 == 0 && crop_left % 4 == 0 && out_w % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    const long long n_thr = static_cast<long long>(T) * 3 * out_h * (out_w / (quad ? 4 : 1));
    resize_cols_normalize_kernel<<<static_cast<unsigned>((n_thr + 255) / 256), 256, 0, st>>>(
        blk, T, H, resize_w, crop_top, crop_left, out_h, out_w, reinterpret_cast<const int*>(blk + o_ymin),
        reinterpret_cast<const int*>(blk + o_ysize), reinterpret_cast<const int16_t*>(blk + o_wy), by.max_interp, by.precision,
        quad, out);
    e = cudaGetLastError();
    if (e != cudaSuccess) return bail(e, "resize_cols_normalize_kernel");
    FOLEY_CUDA_OK(cudaFreeAsync(blk, st));
    return FOLEY_OK;
}

}  // namespace foley
