"""Where does the end-to-end leg's overhead over the device-resident leg go?  Host timers with a synchronize around every phase
of sampling.denoise_process_with_generator's work (xl, 5 s, 50 steps, CFG 4.5, batch 1).   python tools/e2e_breakdown.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
from tools import synthetic as SY  # noqa: E402

E, nodes, sampling, cfgmod = load_pkg("engine"), load_pkg("nodes"), load_pkg("sampling"), load_pkg("config")
ops = load_pkg("torch_ops")
dev = torch.device("cuda", 0)
c = SY.model_config("xl")
cfg = cfgmod.load_model_config("xl")
L, Lv, S = SY.clip_lengths(5.0)
sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev)
eng.load_state_dict(sd)
eng.finalize()
model = nodes.FoleyModel(eng, sd["empty_clip_feat"].cpu(), sd["empty_sync_feat"].cpu(), cfg, dtype=torch.bfloat16)
dac = nodes.FoleyDAC.from_state_dict(SY.synth_state_dict_cuda(SY.dac_param_specs(SY.DAC_CONFIG), 3, dev, torch.float32), device=dev)
feats = {k: v.pin_memory() for k, v in SY.synth_conditions(c, L, Lv, S, dtype=torch.bfloat16).items()}
deps = cfgmod.AttributeDict({"dac_model": dac, "device": dev, "report_progress": True, "foley_model": model})
gen = torch.Generator(device="cpu").manual_seed(123)
sig = sampling.sigma_schedule(50, 1.0)


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def whole(progress):
    deps["report_progress"] = progress
    t0 = sync()
    audio, sr = sampling.denoise_process_with_generator({"siglip2_feat": feats["siglip2_feat"], "syncformer_feat": feats["syncformer_feat"]},
                                                        {"text_feat": feats["text_feat"], "uncond_text_feat": feats["uncond_text_feat"]},
                                                        5.0, deps, cfg, 4.5, 50, 1, "euler", generator=gen)
    out = audio.float().cpu()
    return (sync() - t0) * 1e3


for _ in range(2):
    whole(True)
print("whole job, progress on :", [round(whole(True), 2) for _ in range(3)])
print("whole job, progress off:", [round(whole(False), 2) for _ in range(3)])
# phases
clip = torch.cat([model.get_empty_clip_sequence(bs=1, len=Lv).to(dev, torch.bfloat16), feats["siglip2_feat"].to(dev)])
syn = torch.cat([model.get_empty_sync_sequence(bs=1, len=S).to(dev, torch.bfloat16), feats["syncformer_feat"].to(dev)])
text = torch.cat([sampling._pad_or_trim_time(feats["uncond_text_feat"].to(dev), 77), sampling._pad_or_trim_time(feats["text_feat"].to(dev), 77)])
noise = torch.randn((1, 128, L), generator=gen, dtype=torch.bfloat16)
for _ in range(2):
    t0 = sync(); ops.set_conditions(eng, clip, syn, text, L, 1); t1 = sync()
    lat = ops.denoise(eng, noise.to(dev, torch.float32), sig, 4.5, solver="euler", progress=None); t2 = sync()
    wav = dac.decode(lat); t3 = sync()
    out = wav.float().cpu(); t4 = sync()
    t5 = time.perf_counter(); n2 = sampling.prepare_latents_with_generator(None, 1, 128, L, torch.bfloat16, "cpu", gen); t6 = time.perf_counter()
print(f"set_conditions {1e3 * (t1 - t0):.2f} ms | denoise (50 steps, no progress) {1e3 * (t2 - t1):.2f} | dac {1e3 * (t3 - t2):.2f} | D2H {1e3 * (t4 - t3):.2f} | host noise draw {1e3 * (t6 - t5):.2f}")
