"""Frame preprocessing for the condition encoders on the GPU: the host logic of reference
nodes.py:293-317 (frame count, hold-last-frame padding, 8 / 25 fps picks) over `foley_preprocess_frames`, which fuses
`(image*255).byte()` + torchvision's uint8 antialiased-bicubic Resize [+ CenterCrop] + ToDtype(scale) + Normalize(0.5, 0.5)
(nodes.py:184-196) into two HBM-bound kernels, bit-exact with the reference's per-frame CPU path.  No CPU fallback.
"""
from ctypes import c_int32, c_void_p

import torch

from .engine import FoleyError, _check, load_library

SIGLIP2_SIZE = 512      # v2.Resize((512, 512))                      nodes.py:185
SYNC_SIZE = 224         # v2.Resize(224) + v2.CenterCrop(224)        nodes.py:192-193


def resample_frame_indices(num_frames_to_process, duration, fps):
    """torch.linspace(0, n-1, int(duration*fps)).long() — the reference's frame pick (nodes.py:310,315)."""
    return torch.linspace(0, num_frames_to_process - 1, int(duration * fps)).long()


def resized_size_short_side(h, w, size):
    """torchvision `_compute_resized_output_size` for an int size: short side -> size, long side -> int(size*long/short)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def center_crop_offsets(h, w, ch, cw):
    """torchvision center_crop (no padding case): top = int(round((h - ch) / 2.0)), left likewise."""
    return int(round((h - ch) / 2.0)), int(round((w - cw) / 2.0))


def preprocess_frames(image, frame_idx, resize_hw, crop=None):
    """image: CUDA fp32 [N,H,W,3] in [0,1]; frame_idx: picks into N; resize_hw: (h, w) of the bicubic-antialias resize;
    crop: (top, left, h, w) window of the resized frame (None = all).  Returns CUDA fp32 [T,3,h,w], normalised."""
    lib = load_library()
    if image.device.type != "cuda":
        raise FoleyError("preprocess_frames needs the frames on the CUDA device (no CPU path exists)")
    if image.dim() != 4 or image.shape[-1] != 3:
        raise FoleyError(f"IMAGE must be [N,H,W,3], got {tuple(image.shape)}")
    image = image.to(torch.float32).contiguous()
    N, H, W, _ = image.shape
    idx = [int(i) for i in frame_idx]
    T = len(idx)
    rh, rw = int(resize_hw[0]), int(resize_hw[1])
    top, left, oh, ow = (0, 0, rh, rw) if crop is None else (int(c) for c in crop)
    out = torch.empty(T, 3, oh, ow, dtype=torch.float32, device=image.device)
    if T == 0:
        return out
    arr = (c_int32 * T)(*idx)
    with torch.cuda.device(image.device):
        st = c_void_p(torch.cuda.current_stream(image.device).cuda_stream)
        _check(lib.foley_preprocess_frames(c_void_p(image.data_ptr()), N, H, W, arr, T, rh, rw, top, left, oh, ow,
                                           c_void_p(out.data_ptr()), st))
    return out


def preprocess_video(image, duration, frame_rate, device):
    """reference nodes.py:293-317 + utils.py:270-273 on the GPU.  image: ComfyUI IMAGE [N,H,W,3] float in [0,1] (any
    device).  Returns (siglip2_in [T8,3,512,512], sync_in [T25,3,224,224], audio_len_in_s) on `device`."""
    total_input_frames = image.shape[0]
    num_frames_to_process = int(duration * frame_rate)
    # frames beyond the input hold the last frame (nodes.py:298-303): the picks are clamped instead of materialising copies
    idx8 = resample_frame_indices(num_frames_to_process, duration, 8).clamp_(max=total_input_frames - 1)
    idx25 = resample_frame_indices(num_frames_to_process, duration, 25).clamp_(max=total_input_frames - 1)
    # only the frames that are actually picked travel to the device
    used = torch.unique(torch.cat([idx8, idx25]))
    remap = {int(u): i for i, u in enumerate(used.tolist())}
    if used.numel() == total_input_frames:   # every frame is picked (fps <= 8..25): no host-side gather, one (async) upload
        frames = image.to(device, torch.float32, non_blocking=True)
    else:
        frames = image.index_select(0, used.to(image.device)).to(device, torch.float32)
    H, W = frames.shape[1], frames.shape[2]
    nh, nw = resized_size_short_side(H, W, SYNC_SIZE)
    top, left = center_crop_offsets(nh, nw, SYNC_SIZE, SYNC_SIZE)

    def run(picks, resize_hw, crop):
        # every DISTINCT picked frame is resized once (25 fps picks of an 8..16 fps video repeat each frame ~2-3 times);
        # the repeats are copies of the small output, not re-reads of the big input
        uniq, inverse = torch.unique(picks, return_inverse=True)
        out = preprocess_frames(frames, [remap[int(i)] for i in uniq], resize_hw, crop)
        return out if uniq.numel() == picks.numel() else out.index_select(0, inverse.to(out.device))

    pre8 = run(idx8, (SIGLIP2_SIZE, SIGLIP2_SIZE), None)
    pre25 = run(idx25, (nh, nw), (top, left, SYNC_SIZE, SYNC_SIZE))
    return pre8, pre25, idx25.numel() / 25.0
