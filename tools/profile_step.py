"""Runs a few Euler steps of the benchmark model in eager (non-graph) mode so `ncu` sees every launch:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --model xl --steps 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from tools import synthetic as SY  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="xl")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--duration", type=float, default=5.0)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--dac", type=int, default=1)
a = ap.parse_args()
E, nodes, sampling, cfgmod = (ge.load_pkg(m) for m in ("engine", "nodes", "sampling", "config"))
dev = torch.device("cuda", 0)
c = SY.model_config(a.model)
cfg = cfgmod.load_model_config(a.model) if a.model in ("xl", "xxl") else None
eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs) if cfg else c, device=dev)
sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
eng.load_state_dict(sd)
eng.finalize()
eng.set_option("cuda_graph", a.graph)
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
lat = eng.denoise(noise, sig, 4.5)
torch.cuda.synchronize()
if a.dac:
    dsd = SY.synth_state_dict_cuda(SY.dac_param_specs(SY.DAC_CONFIG), 3, dev, torch.float32)
    dac = nodes.FoleyDAC.from_state_dict(dsd, device=dev)
    wav = dac.decode(lat)
    torch.cuda.synchronize()
print("done", float(lat.abs().mean()), eng.launch_count())
