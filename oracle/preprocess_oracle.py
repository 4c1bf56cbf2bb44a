"""CPU oracle of the frame-preprocessing path (TEST INFRASTRUCTURE ONLY — tests/, __graft_entry__.smoke() and the
bench's cpu_baseline may import it; the product path is csrc/preprocess.cuh and never calls this).

Restates, in numpy integer arithmetic, what the reference does to video frames before the condition encoders:

  nodes.py:293-317   hold-last-frame padding, IMAGE float [N,H,W,3] -> uint8 via (x*255).byte(), frame picks at 8 / 25 fps
                     with torch.linspace(0, n-1, int(duration*fps)).long()
  nodes.py:184-196   torchvision v2 pipelines: Resize(bicubic, antialias=True) [-> CenterCrop(224)] -> ToDtype(float32,
                     scale=True) -> Normalize(0.5, 0.5)

The arithmetic lives in third-party code that is not under /root/reference: torchvision 0.26 `resize_image` (uint8 on CPU
is resized natively) -> ATen `_upsample_bicubic2d_aa` uint8 kernel (aten/src/ATen/native/cpu/UpSampleKernel.cpp,
`_compute_indices_int16_weights_aa` + `upsample_avx_bilinear_bicubic_uint8`; torch 2.11): separable, horizontal pass first,
per-pass int16 fixed-point weights, uint8 intermediate.  Parity is pinned by running exactly those pipelines here
(tools/make_golden.py preprocess -> tests/golden/preprocess_*.pt) and by tests/test_preprocess_oracle.py, which compares
this restatement with torchvision bit for bit on many shapes.
"""
import math

import numpy as np


# ---------------------------------------------------------------------------------------------------------------------
# frame selection (nodes.py:293-317)
# ---------------------------------------------------------------------------------------------------------------------
def frame_indices(num_frames_to_process, duration, fps):
    """torch.linspace(0, n-1, int(duration*fps)).long(): fp32 linspace (symmetric two-sided formula of ATen), truncation."""
    steps = int(duration * fps)
    if steps <= 0:
        return np.zeros((0,), np.int64)
    if steps == 1:
        return np.zeros((1,), np.int64)
    start, end = np.float32(0.0), np.float32(num_frames_to_process - 1)
    step = (end - start) / np.float32(steps - 1)
    i = np.arange(steps)
    half = steps // 2
    lo = (start + step * i.astype(np.float32)).astype(np.float32)
    hi = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, lo, hi).astype(np.int64)


def source_frame(idx, total_input_frames):
    """Frame `idx` of the padded slice -> index into the input batch (the last frame is held, nodes.py:298-303)."""
    return np.minimum(idx, total_input_frames - 1)


def to_uint8(image_f32):
    """(image * 255.0).byte(): fp32 multiply, truncation toward zero, wrap-around like a C cast for out-of-range values."""
    v = (image_f32.astype(np.float32) * np.float32(255.0)).astype(np.float32)
    return np.trunc(v).astype(np.int64).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# antialiased bicubic resize, uint8 (ATen)
# ---------------------------------------------------------------------------------------------------------------------
def _cubic_aa(x):
    """Keys cubic, a = -0.5 (PIL's choice; UpSampleKernel.cpp HelperInterpCubic::aa_filter)."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def aa_weights_int16(in_size, out_size):
    """Per output index: (xmin, xsize, int16 weights[max_interp]) and the shared fixed-point precision.
    ATen: _compute_index_ranges_weights<double> + _compute_index_ranges_int16_weights (align_corners=False, no scale)."""
    scale = float(in_size) / float(out_size)
    interp_size = 4
    support = (interp_size * 0.5) * scale if scale >= 1.0 else interp_size * 0.5
    max_interp = int(math.ceil(support)) * 2 + 1
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    xmin = np.zeros(out_size, np.int64)
    xsize = np.zeros(out_size, np.int64)
    w = np.zeros((out_size, max_interp), np.float64)
    wt_max = 0.0
    for i in range(out_size):
        center = scale * (i + 0.5)
        lo = max(int(center - support + 0.5), 0)
        size = min(int(center + support + 0.5), in_size) - lo
        size = min(max(size, 0), max_interp)
        total = 0.0
        for j in range(size):
            w[i, j] = _cubic_aa((j + lo - center + 0.5) * invscale)
            total += w[i, j]
        if total != 0.0:
            for j in range(size):
                w[i, j] /= total
                wt_max = max(wt_max, w[i, j])
        xmin[i], xsize[i] = lo, size
    precision = 0
    while precision < 22:
        if int(0.5 + wt_max * (1 << (precision + 1))) >= (1 << 15):
            break
        precision += 1
    scaled = w * float(1 << precision)
    w16 = np.where(scaled < 0, scaled - 0.5, scaled + 0.5).astype(np.int64).astype(np.int16)   # C cast: toward zero
    return xmin, xsize, w16, precision


def _resample_axis_u8(img, out_size, axis):
    """One separable pass on a uint8 array: sum_j w16[j] * px[xmin + j], + 2^(p-1), >> p, clamp to [0, 255]."""
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    xmin, xsize, w16, prec = aa_weights_int16(in_size, out_size)
    src = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(src.shape[:-1] + (out_size,), np.uint8)
    for i in range(out_size):
        n = int(xsize[i])
        acc = (src[..., xmin[i]:xmin[i] + n] * w16[i, :n].astype(np.int64)).sum(-1) + (1 << (prec - 1))
        out[..., i] = np.clip(acc >> prec, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def resize_bicubic_aa_u8(img_chw, out_h, out_w):
    """uint8 [C,H,W] -> [C,out_h,out_w]; horizontal pass first, uint8 in between (upsample_avx_bilinear_bicubic_uint8)."""
    tmp = _resample_axis_u8(img_chw, out_w, axis=2)
    return _resample_axis_u8(tmp, out_h, axis=1)


# ---------------------------------------------------------------------------------------------------------------------
# the two pipelines (nodes.py:184-196)
# ---------------------------------------------------------------------------------------------------------------------
def resized_size_short_side(h, w, size):
    """torchvision _compute_resized_output_size for an int `size`: short side -> size, long side int(size * long / short)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def center_crop_offsets(h, w, ch, cw):
    """torchvision center_crop: top = int(round((h - ch) / 2.0)) (Python banker's rounding)."""
    return int(round((h - ch) / 2.0)), int(round((w - cw) / 2.0))


def normalize_u8(u8):
    """ToDtype(float32, scale=True) then Normalize(0.5, 0.5): x.float() * (1/255) in fp32, then (x - 0.5) / 0.5."""
    x = u8.astype(np.float32) * np.float32(1.0 / 255.0)
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)


def siglip2_preprocess(frame_u8_chw):
    """Resize((512, 512), bicubic, antialias) -> float -> normalize."""
    return normalize_u8(resize_bicubic_aa_u8(frame_u8_chw, 512, 512))


def syncformer_preprocess(frame_u8_chw):
    """Resize(224) on the short side (bicubic, antialias) -> CenterCrop(224) -> float -> normalize."""
    _, h, w = frame_u8_chw.shape
    nh, nw = resized_size_short_side(h, w, 224)
    r = resize_bicubic_aa_u8(frame_u8_chw, nh, nw)
    top, left = center_crop_offsets(nh, nw, 224, 224)
    return normalize_u8(r[:, top:top + 224, left:left + 224])


def preprocess_clip(image_nhwc_f32, duration, frame_rate):
    """The whole of nodes.py:293-317 + utils.py:270-273: IMAGE batch -> (siglip2 [T8,3,512,512], sync [T25,3,224,224])."""
    total = image_nhwc_f32.shape[0]
    n = int(duration * frame_rate)
    out = []
    for fps, fn in ((8, siglip2_preprocess), (25, syncformer_preprocess)):
        idx = source_frame(frame_indices(n, duration, fps), total)
        frames = [fn(np.transpose(to_uint8(image_nhwc_f32[i]), (2, 0, 1))) for i in idx]
        out.append(np.stack(frames) if frames else np.zeros((0, 3, 1, 1), np.float32))
    return out[0], out[1]
