"""Condition encoders: this engine against the HF modules the reference runs (eager PyTorch, bf16 module, same GPU, same
seeded weights), per clip of `--seconds` seconds at 8 fps (SigLIP2-base-patch16-512) and per prompt pair (CLAP text).
One JSON line per encoder; CUDA events, warm-up first.
    python tools/encoder_bench.py [--seconds 5] [--iters 10] [--only siglip|clap] [--once]
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

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=5.0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--only", default="")
ap.add_argument("--once", action="store_true", help="one engine call per encoder and nothing else (ncu target)")
ap.add_argument("--tokens", type=int, default=77)
a = ap.parse_args()
enc = load_pkg("encoders")


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def siglip():
    from transformers import SiglipVisionConfig, SiglipVisionModel
    torch.manual_seed(0)
    model = SiglipVisionModel(SiglipVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                                 num_attention_heads=12, image_size=512, patch_size=16,
                                                 hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)).eval().cuda()
    T = int(a.seconds * 8)
    px = torch.rand(T, 3, 512, 512, device="cuda") * 2 - 1
    e = enc.SiglipVisionEncoder.from_hf(model)
    if a.once:
        e.encode(px)
        torch.cuda.synchronize()
        return
    n0 = e.launch_count()
    got = e.encode(px)
    launches = e.launch_count() - n0
    with torch.inference_mode():
        m16 = model.to(torch.bfloat16)
        ref = m16(pixel_values=px).pooler_output
        t_ref = timed(lambda: m16(pixel_values=px).pooler_output, max(3, a.iters // 2))
    t_eng = timed(lambda: e.encode(px), a.iters)
    e.set_option("att_tc", 0)
    t_eng_mma = timed(lambda: e.encode(px), a.iters)
    e.set_option("att_tc", 1)
    qkv = torch.randn(T, 1024, 2304, device="cuda").bfloat16()
    t_att = {impl: timed(lambda: enc.attention_d64(qkv[..., :768], qkv[..., 768:1536], qkv[..., 1536:], 12, impl=impl), a.iters) for impl in (0, 2, 4)}
    att_flops = 4 * 1024 * 1024 * 64 * 12 * T
    tokens = T * 1024
    flops = tokens * 12 * (2 * 768 * (4 * 768 + 2 * 3072) + 4 * 1024 * 768) + tokens * 2 * 768 * 768 * 3   # layers + patch + head kv
    print(json.dumps({"encoder": "siglip2-base-patch16-512 vision + pooling head", "frames": T, "ms_engine": t_eng, "ms_hf_eager_bf16": t_ref,
                      "speedup": t_ref / t_eng, "launches": launches, "tflops_engine": flops / t_eng * 1e-9,
                      "ms_engine_mma_sync_attention": t_eng_mma, "attention_ms_per_layer": {"mma.sync": t_att[0], "tcgen05 (2 threads per row)": t_att[2], "tcgen05 (1 thread per row)": t_att[4]},
                      "attention_tflops": {"mma.sync": att_flops / t_att[0] * 1e-9, "tcgen05 (2 threads per row)": att_flops / t_att[2] * 1e-9,
                                           "tcgen05 (1 thread per row)": att_flops / t_att[4] * 1e-9},
                      "rel_l2_vs_hf_bf16": rel_l2(got.float(), ref.float())}))


def clap():
    from transformers import ClapTextConfig, ClapTextModelWithProjection
    torch.manual_seed(0)
    model = ClapTextModelWithProjection(ClapTextConfig()).eval().cuda()
    T = a.tokens
    ids = torch.randint(3, 50265, (2, T))
    ids[:, 0] = 0
    ids[0, T // 2:] = 1
    mask = (ids != 1).long()
    e = enc.ClapTextEncoder.from_hf(model)
    if a.once:
        e.encode(ids, mask)
        torch.cuda.synchronize()
        return
    n0 = e.launch_count()
    got = e.encode(ids, mask)
    launches = e.launch_count() - n0
    with torch.inference_mode():
        m16 = model.to(torch.bfloat16)
        idc, mc = ids.cuda(), mask.cuda()
        ref = m16(input_ids=idc, attention_mask=mc, output_hidden_states=True).last_hidden_state
        t_ref = timed(lambda: m16(input_ids=idc, attention_mask=mc, output_hidden_states=True).last_hidden_state, a.iters)
    t_eng = timed(lambda: e.encode(ids, mask), a.iters)
    print(json.dumps({"encoder": "clap text (roberta-base) last_hidden_state", "tokens": [2, T], "ms_engine": t_eng, "ms_hf_eager_bf16": t_ref,
                      "speedup": t_ref / t_eng, "launches": launches, "rel_l2_vs_hf_bf16": rel_l2(got.float(), ref.float())}))


if a.only in ("", "siglip"):
    siglip()
if a.only in ("", "clap"):
    clap()
