"""Would running the two CFG halves as two CONCURRENT chains beat the fused batch?  Probe without touching the engine: two engine
instances (separate weights: no L2 sharing, conservative), each a batch-1 no-CFG denoise (M = 250 rows) on its own stream,
enqueued back to back (foley_denoise without a progress callback is asynchronous) against ONE engine running the fused CFG pair
(M = 500 rows).  Same total work per Euler step.
    python tools/concurrency_probe.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
from tools import synthetic as SY  # noqa: E402

E, sampling, cfgmod = load_pkg("engine"), load_pkg("sampling"), load_pkg("config")
dev = torch.device("cuda", 0)
c = SY.model_config("xl")
cfg = cfgmod.load_model_config("xl")
L, Lv, S = SY.clip_lengths(5.0)
sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
f = {k: v.to(dev) for k, v in SY.synth_conditions(c, L, Lv, S, dtype=torch.bfloat16).items()}
pad = lambda x: torch.nn.functional.pad(x[:, :77], (0, 0, 0, 77 - min(77, x.shape[1])))
uclip = sd["empty_clip_feat"].to(dev).reshape(1, 1, -1).expand(1, Lv, -1)
usync = sd["empty_sync_feat"].to(dev).reshape(1, 1, -1).expand(1, S, -1)
noise = torch.randn((1, 128, L), generator=torch.Generator().manual_seed(123), dtype=torch.bfloat16).to(dev).float()
sig = sampling.sigma_schedule(50, 1.0)


def engine():
    e = E.FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev)
    e.load_state_dict(sd)
    e.finalize()
    return e


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    return best


fused = engine()
fused.set_conditions(torch.cat([uclip, f["siglip2_feat"]]), torch.cat([usync, f["syncformer_feat"]]),
                     torch.cat([pad(f["uncond_text_feat"]), pad(f["text_feat"])]), L=L, batch=1)
t_fused = timed(lambda: fused.denoise(noise, sig, 4.5))
a, b = engine(), engine()
a.set_conditions(uclip.contiguous(), usync.contiguous(), pad(f["uncond_text_feat"]), L=L, batch=1)
b.set_conditions(f["siglip2_feat"], f["syncformer_feat"], pad(f["text_feat"]), L=L, batch=1)
t_one = timed(lambda: a.denoise(noise, sig, 1.0))
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    sa.wait_stream(torch.cuda.current_stream())
    sb.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(sa):
        a.denoise(noise, sig, 1.0)
    with torch.cuda.stream(sb):
        b.denoise(noise, sig, 1.0)
    torch.cuda.current_stream().wait_stream(sa)
    torch.cuda.current_stream().wait_stream(sb)


t_both = timed(both)
print(f"fused CFG pair (M = 500): {t_fused:.3f} ms per Euler step; ONE half alone (M = 250): {t_one:.3f}; the two halves as concurrent chains: {t_both:.3f}")
