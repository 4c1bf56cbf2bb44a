"""2-GPU bring-up of parallel.denoise_sharded: where do the sharded results start to differ from single-GPU slices?"""
import os
import sys
import threading

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg, rel_l2  # noqa: E402
from test_gpu_multigpu import _objects  # noqa: E402

torch.cuda.set_device(0)
nodes, cfgmod, model, dac, deps, text, cfg = _objects()
sampling, par = load_pkg("sampling"), load_pkg("parallel")
duration, steps, B, seed = 2.0, 10, 3, 7
clip_len, sync_len = nodes.t2a_feature_lengths(duration)
visual = {"siglip2_feat": model.get_empty_clip_sequence(bs=1, len=clip_len).to("cpu", torch.bfloat16),
          "syncformer_feat": model.get_empty_sync_sequence(bs=1, len=sync_len).to("cpu", torch.bfloat16)}


def one(dev, lo, hi, decode):
    dev = torch.device("cuda", dev)
    with torch.cuda.device(dev):
        md = cfgmod.AttributeDict(dict(deps))
        md["foley_model"], md["dac_model"], md["device"] = model.on_device(dev), dac.on_device(dev), dev
        gen = torch.Generator(device="cpu").manual_seed(seed)
        w, _ = sampling.denoise_process_with_generator(visual, text, duration, md, cfg, 4.5, steps, B, "euler", generator=gen,
                                                       batch_slice=(lo, hi), decode=decode)
        torch.cuda.synchronize(dev)
        return w.float().cpu()


for decode in (False, True):
    tag = "waveform" if decode else "latents"
    a0 = one(0, 0, 2, decode)
    a0b = one(0, 0, 2, decode)
    a1 = one(0, 2, 3, decode)
    b0 = one(1, 0, 2, decode)        # the replica on GPU 1, alone
    b1 = one(1, 2, 3, decode)
    print(f"[{tag}] GPU0 repeat: equal={torch.equal(a0, a0b)}; GPU1 replica vs GPU0, rows 0-1: equal={torch.equal(a0, b0)} rel={rel_l2(b0, a0):.3e}; "
          f"row 2: equal={torch.equal(a1, b1)} rel={rel_l2(b1, a1):.3e}", flush=True)
    res = {}

    def work(dev, lo, hi):
        res[dev] = one(dev, lo, hi, decode)
    t = threading.Thread(target=work, args=(1, 2, 3))
    t.start()
    work(0, 0, 2)
    t.join()
    print(f"[{tag}] threaded: GPU0 rows 0-1 equal={torch.equal(res[0], a0)} rel={rel_l2(res[0], a0):.3e}; GPU1 row 2 equal={torch.equal(res[1], a1)} "
          f"rel={rel_l2(res[1], a1):.3e}", flush=True)

md = cfgmod.AttributeDict(dict(deps))
md["foley_model"], md["device"] = model, torch.device("cuda", 0)
gen = torch.Generator(device="cpu").manual_seed(seed)
full, _ = par.denoise_sharded(visual, text, duration, md, cfg, 4.5, steps, B, "euler", gen, [torch.device("cuda", 0), torch.device("cuda", 1)])
full = full.float().cpu()
a0, a1 = one(0, 0, 2, True), one(0, 2, 3, True)
print(f"denoise_sharded: rows 0-1 equal={torch.equal(full[:2], a0)} rel={rel_l2(full[:2], a0):.3e}; row 2 equal={torch.equal(full[2:], a1)} rel={rel_l2(full[2:], a1):.3e}")
