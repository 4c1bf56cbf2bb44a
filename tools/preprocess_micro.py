"""Times foley_preprocess_frames on BASELINE-shaped input (5 s of 1080p video at the given frame rate -> 40 SigLIP2
frames 512x512 + 125 Synchformer frames 224x224) with CUDA events and reports the achieved HBM traffic rate, next to the
reference's per-frame torchvision CPU path on a bounded sample of the same frames.

    python tools/preprocess_micro.py [--height 1080 --width 1920 --fps 8 --seconds 5]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--fps", type=float, default=8.0)
ap.add_argument("--seconds", type=float, default=5.0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--cpu-frames", type=int, default=6)
a = ap.parse_args()
pp = ge.load_pkg("preprocess")
dev = torch.device("cuda", 0)
n = int(a.seconds * a.fps)
H, W = a.height, a.width
image = torch.rand(n, H, W, 3, device=dev)
idx8 = pp.resample_frame_indices(n, a.seconds, 8).tolist()
idx25 = pp.resample_frame_indices(n, a.seconds, 25).tolist()
nh, nw = pp.resized_size_short_side(H, W, 224)
top, left = pp.center_crop_offsets(nh, nw, 224, 224)


u25 = sorted(set(idx25))
inv25 = torch.tensor([u25.index(i) for i in idx25], device=dev)


def run():   # as preprocess_video does: distinct frames resized once, repeats are copies of the small output
    p8 = pp.preprocess_frames(image, idx8, (512, 512))
    p25 = pp.preprocess_frames(image, u25, (nh, nw), (top, left, 224, 224)).index_select(0, inv25)
    return p8, p25


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
# algorithmic bytes: every picked frame read once as fp32 RGB, uint8 intermediate written + read, fp32 output written
b8 = len(idx8) * (12 * H * W + 2 * 3 * H * 512 + 12 * 512 * 512)
b25 = len(u25) * (12 * H * W + 2 * 3 * H * 224 + 12 * 224 * 224) + len(idx25) * 2 * 12 * 224 * 224
gb = (b8 + b25) / 1e9
print(f"GPU: {len(idx8)} + {len(idx25)} frames of {H}x{W}: {ms:.3f} ms  -> {gb / ms * 1e3:.0f} GB/s algorithmic "
      f"({gb:.2f} GB), {(len(idx8) + len(idx25)) / ms * 1e3:.0f} frames/s")
# reference CPU path on a bounded sample
from torchvision.transforms import v2  # noqa: E402
sig = v2.Compose([v2.Resize((512, 512), interpolation=v2.InterpolationMode.BICUBIC, antialias=True),
                  v2.ToDtype(torch.float32, scale=True), v2.Normalize(mean=[0.5] * 3, std=[0.5] * 3)])
syn = v2.Compose([v2.Resize(224, interpolation=v2.InterpolationMode.BICUBIC, antialias=True), v2.CenterCrop(224),
                  v2.ToDtype(torch.float32, scale=True), v2.Normalize(mean=[0.5] * 3, std=[0.5] * 3)])
k = min(a.cpu_frames, n)
cpu = image[:k].cpu()
torch.set_num_threads(os.cpu_count())
t0 = time.perf_counter()
u8 = (cpu * 255.0).byte().permute(0, 3, 1, 2)
for f in u8:
    sig(f)
    syn(f)
dt = time.perf_counter() - t0
per_pair = dt / k
est = per_pair * (len(idx8) + len(idx25)) / 2
print(f"CPU reference path ({os.cpu_count()} cores): {per_pair * 1e3:.1f} ms per (siglip2 + sync) frame pair on {k} frames "
      f"-> ~{est * 1e3:.0f} ms for the clip (extrapolated)")
