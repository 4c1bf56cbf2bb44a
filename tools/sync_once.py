"""One Synchformer pass on the engine (5 s clip: 125 frames) and nothing else: the ncu target.
    python tools/sync_once.py [--frames 125]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
from tools import synthetic as SY  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=125)
a = ap.parse_args()
enc = load_pkg("encoders")
e = enc.SynchformerEncoder.from_state_dict(SY.synth_motionformer_state_dict(12, seed=0))
frames = SY.synth_sync_frames(24, seed=0).cuda().repeat(6, 1, 1, 1)[: a.frames].contiguous()
e.encode(frames)
torch.cuda.synchronize()
