"""Coordinate-descent search of the per-shape GEMM plans (tile width, K splits) with the MEASURED Euler step as the objective.
Every candidate is a fresh engine in this process (FOLEY_PLAN_OVERRIDE is read at foley_engine_create), same seeded weights
and conditions, 50-step denoise timed with CUDA events (best of `--reps`).  Prints the accepted moves and the final override
string; the result goes into Engine::create's built-in table.
    python tools/plan_search.py [--model xl] [--duration 5] [--batch 1] [--passes 2]
"""
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
ap.add_argument("--model", default="xl")
ap.add_argument("--duration", type=float, default=5.0)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--init", default="")
a = ap.parse_args()
E, sampling, cfgmod = load_pkg("engine"), load_pkg("sampling"), load_pkg("config")
dev = torch.device("cuda", 0)
c = SY.model_config(a.model)
cfg = cfgmod.load_model_config(a.model)
L, Lv, S = SY.clip_lengths(a.duration)
sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
feats = {k: v.to(dev) for k, v in SY.synth_conditions(c, L, Lv, S, dtype=torch.bfloat16).items()}
par = load_pkg("parallel")
noise = torch.randn((a.batch, 128, L), generator=torch.Generator().manual_seed(123), dtype=torch.bfloat16).to(dev).float()
sig = sampling.sigma_schedule(50, 1.0)


def cond_rows(f):
    clip = torch.cat([sd["empty_clip_feat"].to(dev).reshape(1, 1, -1).expand(1, Lv, -1), f["siglip2_feat"]])
    sync = torch.cat([sd["empty_sync_feat"].to(dev).reshape(1, 1, -1).expand(1, S, -1), f["syncformer_feat"]])
    T = 77
    pad = lambda x: torch.nn.functional.pad(x[:, :T], (0, 0, 0, T - min(T, x.shape[1])))
    text = torch.cat([pad(f["uncond_text_feat"]), pad(f["text_feat"])])
    return clip, sync, text


def measure(over):
    os.environ["FOLEY_PLAN_OVERRIDE"] = ";".join(f"{k[0]}:{k[1]}:{k[2]}:{k[3]}={v[0]}:{v[1]}" for k, v in over.items())
    eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev)
    eng.load_state_dict(sd)
    eng.finalize()
    eng.set_conditions(*cond_rows(feats), L=L, batch=a.batch)
    eng.denoise(noise, sig, 4.5)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.denoise(noise, sig, 4.5)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    plans = eng.debug_read("plans").cpu().view(-1, 6).long().tolist()
    eng.close()
    return best, {tuple(p[:4]): (p[4], p[5]) for p in plans}


over = {}
if a.init:
    for item in a.init.split(";"):
        k, v = item.split("=")
        over[tuple(int(x) for x in k.split(":"))] = tuple(int(x) for x in v.split(":"))
base, plans = measure(over)
print(f"baseline {base:.4f} ms/step; shapes the step plans: {len(plans)}", flush=True)
for k, v in sorted(plans.items()):
    print("   ", k, "->", v)
cur = dict(plans)
cur.update(over)
best = base
for p_ in range(a.passes):
    moved = False
    for key in sorted(plans, key=lambda k: -k[2] * k[3]):      # biggest GEMMs first
        rows, batch, n, kb = key
        bn0, s0 = cur[key]
        cands = []
        for bn in (128, 256):
            for s in sorted({1, 2, 3, 4, 6, 8, s0 - 1, s0 + 1}):
                if s < 1 or s > 8 or (s > 1 and kb // s < 4) or (bn, s) == (bn0, s0):
                    continue
                if abs(s - s0) > 2 and s not in (1,):
                    continue
                cands.append((bn, s))
        for cand in cands:
            trial = dict(cur)
            trial[key] = cand
            t, seen = measure(trial)
            if seen.get(key) != cand:      # the shape cannot split (or the cap clipped it): nothing to learn
                continue
            if t < best * 0.9985:
                print(f"pass {p_}: {key} {cur[key]} -> {cand}: {best:.4f} -> {t:.4f}", flush=True)
                best, cur, moved = t, trial, True
    if not moved:
        break
final = {k: v for k, v in cur.items() if plans.get(k) != v}
print(f"final {best:.4f} ms/step (baseline {base:.4f})")
print("FOLEY_PLAN_OVERRIDE=" + ";".join(f"{k[0]}:{k[1]}:{k[2]}:{k[3]}={v[0]}:{v[1]}" for k, v in sorted(final.items())))
