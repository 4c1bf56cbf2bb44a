"""Where does the Euler step's time go INSIDE the replayed graph (PDL overlap, parallel branches, warm L2)?
ncu's per-kernel durations are cold and serialised, so this tool measures marginal costs instead: it replays the real
step with one class of kernels removed (engine option "debug_skip"; results are garbage, timing is not) and reports
the step time next to the full step.

    python tools/ablate_step.py --model xl [--batch 1]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from tools import synthetic as SY  # noqa: E402

MASKS = [("full step", 0), ("- attention", 1), ("- qk-norm/RoPE", 2), ("- combine/LN/mod", 4), ("- mod GEMM", 8),
         ("- visual-branch GEMMs", 16), ("- single qkv GEMM", 32), ("- single w1|w3 GEMM", 64), ("- single w2 GEMM", 128),
         ("- single linear1 GEMM", 256), ("- all triple blocks", 512), ("- all single blocks", 1024),
         ("- all row-wise + attention", 1 | 2 | 4), ("- all single GEMMs", 32 | 64 | 128 | 256)]

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="xl")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--duration", type=float, default=5.0)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--masks", default="")
a = ap.parse_args()
E, sampling, cfgmod = (ge.load_pkg(m) for m in ("engine", "sampling", "config"))
dev = torch.device("cuda", 0)
c = SY.model_config(a.model)
cfg = cfgmod.load_model_config(a.model) if a.model in ("xl", "xxl") else None
eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs) if cfg else c, device=dev)
sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
eng.load_state_dict(sd)
eng.finalize()
L, Lv, S = SY.clip_lengths(a.duration)
f = SY.synth_conditions(c, L, Lv, S, dtype=torch.bfloat16)
text = sampling._pad_or_trim_time(f["text_feat"], 77)
utext = sampling._pad_or_trim_time(f["uncond_text_feat"], 77)
clip = torch.cat([sd["empty_clip_feat"].cpu()[None].expand(1, Lv, -1), f["siglip2_feat"]])
sync = torch.cat([sd["empty_sync_feat"].cpu()[None].expand(1, S, -1), f["syncformer_feat"]])
del sd
eng.set_conditions(clip.to(dev), sync.to(dev), torch.cat([utext, text]).to(dev), L=L, batch=a.batch)
noise = torch.randn(a.batch, 128, L, device=dev)
sig = sampling.sigma_schedule(a.steps)
masks = MASKS if not a.masks else [(m, int(m)) for m in a.masks.split(",")]
base = None
for name, mask in masks:
    eng.set_option("debug_skip", mask)
    for _ in range(2):
        eng.denoise(noise, sig, 4.5)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.denoise(noise, sig, 4.5)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    base = best if base is None else base
    print(f"{name:32s} mask={mask:5d}  {best:7.3f} ms/step  (saves {base - best:6.3f} ms = {100 * (base - best) / base:5.1f} %)", flush=True)
eng.set_option("debug_skip", 0)
