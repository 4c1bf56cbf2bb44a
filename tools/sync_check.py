"""Bring-up / timing of the Synchformer visual extractor on the engine against the reference's own MotionFormer (staged
baseline/_ref through tools/ref_shims.py) on the same GPU: bf16 parameters under fp16 autocast, and the fp32 module.
    python tools/sync_check.py [--depths 1,2,12] [--frames 24] [--time]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg, rel_l2  # noqa: E402
from tools import ref_shims as R  # noqa: E402
from tools import synthetic as SY  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--depths", default="1,2,12")
ap.add_argument("--frames", type=int, default=24)
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
enc = load_pkg("encoders")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


frames = SY.synth_sync_frames(a.frames, seed=0).cuda()
S = (a.frames - 16) // 8 + 1
x = torch.stack([frames[i * 8: i * 8 + 16] for i in range(S)])[None].permute(0, 1, 3, 2, 4, 5)
for depth in [int(d) for d in a.depths.split(",")]:
    sd = SY.synth_motionformer_state_dict(depth, seed=0)
    model = R.load_motionformer(depth)().eval()
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    e = enc.SynchformerEncoder.from_state_dict(sd, num_layers=depth, **{k: v for k, v in enc.MOTIONFORMER_DIVIDED_224.items() if k != "num_layers"})
    got = e.encode(frames)[0].view(S, 8, 768)
    with torch.inference_mode():
        out32 = model(x.float()).float()[0]
        model = model.to(torch.bfloat16)
        with torch.autocast(device_type="cuda", enabled=True, dtype=torch.half):
            out16 = model(x).float()[0]
            t_ref = timed(lambda: model(x)) if a.time else None
    line = {"depth": depth, "segments": S, "engine_vs_fp32": rel_l2(got, out32), "engine_vs_autocast": rel_l2(got, out16),
            "autocast_vs_fp32": rel_l2(out16, out32), "finite": bool(torch.isfinite(got).all())}
    if a.time:
        line["ms_engine"], line["ms_reference_eager_autocast"] = timed(lambda: e.encode(frames)), t_ref
    print(json.dumps(line))
    del model, e
