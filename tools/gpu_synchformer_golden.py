"""Runs the reference's OWN MotionFormer (staged baseline/_ref, through tools/ref_shims.py: omegaconf / timm stand-ins) on the
GPU the way the reference's Sampler runs it — parameters moved to bf16 (nodes.py:283-284), torch.autocast(fp16) around the call
(feature_utils.py:100-102) — on seeded weights (tools/synthetic.py) and seeded frames, plus the fp32 module without autocast
(ground truth).  Writes tests/golden/synchformer_d{depth}.pt: {"out_autocast": [S, 8, 768], "out_fp32": ..., "frames_seed", ...}.
    python tools/gpu_synchformer_golden.py [--depths 12,2] [--frames 24]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import ref_shims as R  # noqa: E402
from tools import synthetic as SY  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--depths", default="12,2")
ap.add_argument("--frames", type=int, default=24)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
frames = SY.synth_sync_frames(a.frames, seed=0)
S = (a.frames - 16) // 8 + 1
segs = torch.stack([frames[i * 8: i * 8 + 16] for i in range(S)])[None].cuda()     # [1, S, 16, 3, 224, 224] (feature_utils.py:91-96)
x = segs.permute(0, 1, 3, 2, 4, 5)                                                   # Synchformer.forward (synchformer.py:46-47)
for depth in [int(d) for d in a.depths.split(",")]:
    mk = R.load_motionformer(depth)
    model = mk().eval()
    missing = model.load_state_dict(SY.synth_motionformer_state_dict(depth, seed=0), strict=True)
    model = model.cuda()
    with torch.inference_mode():
        out32 = model(x.float()).float()[0]                                          # [S, 8, 768]
        model = model.to(torch.bfloat16)
        with torch.autocast(device_type="cuda", enabled=True, dtype=torch.half):
            out16 = model(x)
        out_dtype = str(out16.dtype)
        out16 = out16.float()[0]
    rel = float((out16 - out32).norm() / out32.norm())
    print(f"depth {depth}: out {tuple(out16.shape)} dtype under autocast {out_dtype}, autocast vs fp32 rel-L2 {rel:.3e}, |out| {float(out32.abs().mean()):.3f}")
    torch.save({"out_autocast": out16.cpu(), "out_fp32": out32.cpu(), "n_frames": a.frames, "frames_seed": 0, "weights_seed": 0, "depth": depth,
                "out_dtype_under_autocast": out_dtype, "autocast_vs_fp32": rel, "torch": str(torch.__version__)},
               os.path.join(a.out, f"synchformer_d{depth}.pt"))
    del model
