"""The REFERENCE ITSELF on a B200 at the benchmarked configuration -> tests/golden/full_*.pt  (+ a parity report).

Run on the GPU box (the reference tree travels as baseline/_ref, see tools/stage_reference.py):

    python tools/gpu_reference_golden.py --out gpurun_out/golden_full

What it runs (reference utils.py:125-258 `denoise_process_with_generator`, hifi_foley.py:707-924, dac.py:280-303 —
the reference's own modules, imported through tools/ref_shims.py; nothing of the product on that side):

  xl  V2A 5 s, 50 Euler steps, CFG 4.5, seed 123, full depth (BASELINE.json configs[1])
      (a) bf16 weights under torch.autocast("cuda", bf16), batch_size 1      — the path users run
      (b) same, batch_size 2 (row 0 = same noise as (a))                      — the path's own run-to-run floor:
          a different batch composition makes cuBLAS / SDPA pick other kernels and summation orders
      (c) fp32 weights/activations, TF32 off (matmul + cuDNN)                 — ground truth
  xxl one CFG-pair forward at 5 s and at 30 s (L=1500, Lv=240, S=736), bf16-autocast and fp32

and, in the same process, the ENGINE on the same seeded weights / conditions / noise, so the report carries
engine-vs-reference distances measured on the same box.  Weights are the GPU-drawn synthetic weights the bench
uses (tools/synthetic.synth_state_dict_cuda: philox on the device, deterministic for a given GPU model + torch
build); a checksum of a few tensors is stored so a test can tell "different weights" from "different arithmetic".

Also timed (informational baselines for bench.py / DESIGN.md): the reference's eager CUDA-autocast step and DAC decode.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shims  # noqa: E402
import synthetic as SY  # noqa: E402


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def weight_checksum(sd):
    names = sorted(sd.keys())
    pick = [names[0], names[len(names) // 3], names[len(names) // 2], names[-1]]
    return {n: float(sd[n].double().sum().item()) for n in pick}


def ref_cfg(ns, name):
    cfg = ns.load_yaml(os.path.join(ns.config_dir, "hunyuanvideo-foley-xxl.yaml"))
    for k, v in SY.MODEL_CONFIGS[name].items():
        cfg.model_config.model_kwargs[k] = v
    return cfg


def build_ref_model(ns, name, sd, dtype, dev):
    """The reference loader's sequence (nodes.py:94-104): meta-init, to_empty, load_state_dict, .to(dtype)."""
    cfg = ref_cfg(ns, name)
    with torch.device("meta"):
        model = ns.HunyuanVideoFoley(cfg, dtype=torch.float32)
    model = model.to_empty(device=dev)
    res = model.load_state_dict({k: v.to(torch.float32) for k, v in sd.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.to(dtype).eval(), cfg


def build_ref_dac(ns, dsd, dev):
    dac = ns.DAC(**{**ns.utils._DAC_KWARGS, "encoder_dim": 16})   # encoder unused by decode (utils.py:32-44)
    res = dac.load_state_dict({k: v.float().cpu() for k, v in dsd.items()}, strict=False)
    assert not res.unexpected_keys
    assert all(not k.startswith(("decoder", "post_quant")) for k in res.missing_keys), res.missing_keys
    return dac.eval().float().to(dev)


def set_exact_fp32(exact):
    """torch defaults: fp32 matmul exact, cuDNN convolutions TF32 (what the reference's fp32 DAC decode runs with).
    exact=True switches TF32 off everywhere (ground truth)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = not exact


def run_ref_denoise(ns, model, dac, cfg, feats, duration, steps, guidance, batch, dev, dtype):
    captured = {}
    orig = dac.decode

    def spy(z):
        captured["latents"] = z.detach().float().clone()
        return orig(z)

    dac.decode = spy
    md = ns.AttributeDict({"foley_model": model, "dac_model": dac, "device": dev})
    if hasattr(model, "_text_len_fixed"):
        del model._text_len_fixed
    visual = {"siglip2_feat": feats["siglip2_feat"].to(dev, dtype), "syncformer_feat": feats["syncformer_feat"].to(dev, dtype)}
    text = {"text_feat": feats["text_feat"].to(dev, dtype), "uncond_text_feat": feats["uncond_text_feat"].to(dev, dtype)}
    gen = torch.Generator(device="cpu").manual_seed(123)
    orig_randn = ns.utils.randn_tensor

    def randn_bf16_valued(shape, generator=None, device=None, dtype=None, layout=None):
        # every run starts from the SAME noise: the bf16 draw of the users' path (utils.py:151-156), up-cast for fp32
        return orig_randn(shape, generator=generator, device=device, dtype=torch.bfloat16).to(dtype)

    ns.utils.randn_tensor = randn_bf16_valued
    try:
        audio, sr = ns.utils.denoise_process_with_generator(visual, text, duration, md, cfg, guidance_scale=guidance,
                                                            num_inference_steps=steps, batch_size=batch,
                                                            sampler="euler", generator=gen)
    finally:
        dac.decode = orig
        ns.utils.randn_tensor = orig_randn
    return captured["latents"], audio.float()


def engine_objects(name, sd, dsd, dev):
    import __graft_entry__ as ge
    E, nodes, cfgmod = ge.load_pkg("engine"), ge.load_pkg("nodes"), ge.load_pkg("config")
    cfg = cfgmod.load_model_config(name)
    eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev)
    eng.load_state_dict(sd)
    eng.finalize()
    model = nodes.FoleyModel(eng, sd["empty_clip_feat"].cpu(), sd["empty_sync_feat"].cpu(), cfg, dtype=torch.bfloat16)
    dac = nodes.FoleyDAC.from_state_dict(dsd, device=dev) if dsd is not None else None
    return eng, model, dac, cfg, cfgmod


def run_engine_denoise(model, dac, cfg, cfgmod, feats, duration, steps, guidance, batch, dev):
    import __graft_entry__ as ge
    sampling = ge.load_pkg("sampling")
    deps = cfgmod.AttributeDict({"dac_model": dac, "device": dev, "report_progress": False})
    deps["foley_model"] = model
    if hasattr(model, "_text_len_fixed"):
        del model._text_len_fixed
    f = {k: v.to(dev, torch.bfloat16) for k, v in feats.items()}
    visual = {"siglip2_feat": f["siglip2_feat"], "syncformer_feat": f["syncformer_feat"]}
    text = {"text_feat": f["text_feat"], "uncond_text_feat": f["uncond_text_feat"]}
    gen = torch.Generator(device="cpu").manual_seed(123)
    lat, _ = sampling.denoise_process_with_generator(visual, text, duration, deps, cfg, guidance_scale=guidance,
                                                     num_inference_steps=steps, batch_size=batch, sampler="euler",
                                                     generator=gen, decode=False)
    wav = dac.decode(lat)
    return lat.float(), wav.float()


def forward_inputs(c, L, Lv, S, dev, seed=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, c["audio_vae_latent_dim"], L, generator=g).bfloat16().float()
    t = torch.tensor([875.0, 875.0])
    cond = torch.randn(2, 77, c["condition_dim"], generator=g)
    cond[:, 9:] = 0
    clip = torch.randn(2, Lv, c["clip_dim"], generator=g)
    sync = torch.randn(2, S, c["sync_feat_dim"], generator=g)
    return [v.bfloat16().float() for v in (x, t, cond, clip, sync)]


def ref_forward(model, x, t, cond, clip, sync, dev, dtype):
    with torch.inference_mode():
        if dtype == torch.bfloat16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = model(x=x.to(dev, dtype), t=t.to(dev), cond=cond.to(dev, dtype), clip_feat=clip.to(dev, dtype),
                            sync_feat=sync.to(dev, dtype))["x"]
        else:
            out = model(x=x.to(dev), t=t.to(dev), cond=cond.to(dev), clip_feat=clip.to(dev), sync_feat=sync.to(dev))["x"]
    return out.float()


def time_ref_steps(ns, model, cfg, feats, duration, guidance, dev, n=6):
    """ms per eager CUDA-autocast Euler step of the reference (CFG pair), CUDA events, 2 warm-up + n timed steps."""
    class _NoDac:
        sample_rate = 48000

        def parameters(self):
            return iter([torch.zeros(1, device=dev)])

        def decode(self, z):
            return z[:, :1]

    md = ns.AttributeDict({"foley_model": model, "dac_model": _NoDac(), "device": dev})
    visual = {"siglip2_feat": feats["siglip2_feat"].to(dev, torch.bfloat16), "syncformer_feat": feats["syncformer_feat"].to(dev, torch.bfloat16)}
    text = {"text_feat": feats["text_feat"].to(dev, torch.bfloat16), "uncond_text_feat": feats["uncond_text_feat"].to(dev, torch.bfloat16)}

    def run(k):
        gen = torch.Generator(device="cpu").manual_seed(123)
        ns.utils.denoise_process_with_generator(visual, text, duration, md, cfg, guidance_scale=guidance,
                                                num_inference_steps=k, batch_size=1, sampler="euler", generator=gen)
    run(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden_full"))
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--skip-xxl", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ns = ref_shims.load_reference()
    report = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "reference_root": ref_shims.REF_ROOT}
    duration, guidance = 5.0, 4.5
    L, Lv, S = SY.clip_lengths(duration)

    # ------------------------------------------------------------------ xl, full loop
    c = SY.model_config("xl")
    sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
    dsd = SY.synth_state_dict_cuda(SY.dac_param_specs(SY.DAC_CONFIG), 3, dev, torch.float32)
    feats = {k: v.bfloat16().float() for k, v in SY.synth_conditions(c, L, Lv, S).items()}   # bf16-valued for every run
    ck = {"dit": weight_checksum(sd), "dac": weight_checksum(dsd)}
    ref_dac = build_ref_dac(ns, dsd, dev)

    set_exact_fp32(False)
    model16, cfg = build_ref_model(ns, "xl", sd, torch.bfloat16, dev)
    t0 = time.time()
    lat16_b1, wav16_b1 = run_ref_denoise(ns, model16, ref_dac, cfg, feats, duration, a.steps, guidance, 1, dev, torch.bfloat16)
    torch.cuda.synchronize()
    report["ref_bf16_loop_wall_s"] = time.time() - t0
    lat16_b2, wav16_b2 = run_ref_denoise(ns, model16, ref_dac, cfg, feats, duration, a.steps, guidance, 2, dev, torch.bfloat16)
    # one forward at the first timestep (single-step error, same inputs for everyone)
    fx = forward_inputs(c, L, Lv, S, dev)
    fwd16 = ref_forward(model16, *fx, dev, torch.bfloat16)
    report["ref_bf16_ms_per_step_eager"] = time_ref_steps(ns, model16, cfg, feats, duration, guidance, dev)
    # reference DAC decode timing (fp32, cuDNN TF32 default)
    z = lat16_b1.to(dev)
    with torch.inference_mode():
        ref_dac.decode(z); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ref_dac.decode(z); e1.record(); torch.cuda.synchronize()
    report["ref_dac_decode_ms"] = e0.elapsed_time(e1)
    del model16
    torch.cuda.empty_cache()

    set_exact_fp32(True)
    model32, _ = build_ref_model(ns, "xl", sd, torch.float32, dev)
    lat32, wav32 = run_ref_denoise(ns, model32, ref_dac, cfg, feats, duration, a.steps, guidance, 1, dev, torch.float32)
    fwd32 = ref_forward(model32, *fx, dev, torch.float32)
    with torch.inference_mode():
        wav16_truthdac = ref_dac.decode(lat16_b1.to(dev)).float()   # bf16-path latents through the exact-fp32 decoder
    del model32
    torch.cuda.empty_cache()
    set_exact_fp32(False)

    # ---- engine on the same everything
    eng, emodel, edac, ecfg, cfgmod = engine_objects("xl", sd, dsd, dev)
    elat_b1, ewav_b1 = run_engine_denoise(emodel, edac, ecfg, cfgmod, feats, duration, a.steps, guidance, 1, dev)
    elat_b2, _ = run_engine_denoise(emodel, edac, ecfg, cfgmod, feats, duration, a.steps, guidance, 2, dev)
    x, t, cond, clip, sync = fx
    eng.set_conditions(clip.to(dev), sync.to(dev), cond.to(dev), L=L, batch=1)
    efwd = eng.dit_forward(x.to(dev), t)
    ewav_on_ref_lat = edac.decode(lat16_b1.to(dev)).float()
    assert eng.debug_flags()[0] == 0

    floor = rel_l2(lat16_b2[:1], lat16_b1)
    report["xl_5s_50step"] = {
        "steps": a.steps, "guidance": guidance,
        "latents": {
            "ref_bf16_b2row0_vs_ref_bf16_b1 (floor)": floor,
            "ref_bf16_vs_ref_fp32": rel_l2(lat16_b1, lat32),
            "engine_vs_ref_bf16": rel_l2(elat_b1, lat16_b1),
            "engine_vs_ref_fp32": rel_l2(elat_b1, lat32),
            "engine_b2row0_vs_engine_b1": rel_l2(elat_b2[:1], elat_b1),
            "engine_b2row0_vs_ref_bf16_b2row0": rel_l2(elat_b2[:1], lat16_b2[:1]),
            "engine_b2row1_vs_ref_bf16_b2row1": rel_l2(elat_b2[1:], lat16_b2[1:]),
        },
        "forward_one_step": {
            "ref_bf16_vs_ref_fp32": rel_l2(fwd16, fwd32),
            "engine_vs_ref_bf16": rel_l2(efwd, fwd16),
            "engine_vs_ref_fp32": rel_l2(efwd, fwd32),
        },
        "waveform": {
            "ref_bf16_b2row0_vs_ref_bf16_b1 (floor)": rel_l2(wav16_b2[:1], wav16_b1),
            "ref_bf16_vs_ref_fp32": rel_l2(wav16_b1, wav32),
            "engine_vs_ref_bf16": rel_l2(ewav_b1, wav16_b1),
            "engine_vs_ref_fp32": rel_l2(ewav_b1, wav32),
            "dac_only: engine(tf32) vs ref fp32-exact, same latents": rel_l2(ewav_on_ref_lat, wav16_truthdac),
            "dac_only: ref cuDNN-TF32 vs ref fp32-exact, same latents": rel_l2(wav16_b1, wav16_truthdac),
        },
    }
    torch.save({"checksum": ck, "args": dict(model="xl", duration=duration, steps=a.steps, guidance=guidance, seed=123),
                "lat_ref_bf16_b1": lat16_b1.cpu(), "lat_ref_bf16_b2": lat16_b2.cpu(), "lat_ref_fp32": lat32.cpu(),
                "fwd_ref_bf16": fwd16.cpu().bfloat16(), "fwd_ref_fp32": fwd32.cpu(),
                "wav_ref_bf16_b1": wav16_b1.cpu().half(), "wav_ref_fp32": wav32.cpu().half(),
                "wav_ref_exactdac_on_bf16_lat": wav16_truthdac.cpu().half(),
                "floor_latents": floor},
               os.path.join(a.out, "full_xl_5s_50step.pt"))
    print(json.dumps(report, indent=1), flush=True)
    del eng, emodel, edac, sd, dsd, ref_dac
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ xxl, one forward at 5 s and 30 s
    if not a.skip_xxl:
        c = SY.model_config("xxl")
        sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
        ck2 = weight_checksum(sd)
        shapes = {"5s": SY.clip_lengths(5.0), "30s": SY.clip_lengths(30.0)}
        ins = {k: forward_inputs(c, *v, dev) for k, v in shapes.items()}
        out16, out32, oute = {}, {}, {}
        model16, _ = build_ref_model(ns, "xxl", sd, torch.bfloat16, dev)
        for k in shapes:
            out16[k] = ref_forward(model16, *ins[k], dev, torch.bfloat16)
        # floor of one forward: the same rows inside a batch of 4 (other kernels / split choices in cuBLAS)
        x, t, cond, clip, sync = ins["5s"]
        rep = lambda v: torch.cat([v, v])
        out16_b4 = ref_forward(model16, rep(x), rep(t), rep(cond), rep(clip), rep(sync), dev, torch.bfloat16)[:2]
        del model16
        torch.cuda.empty_cache()
        set_exact_fp32(True)
        model32, _ = build_ref_model(ns, "xxl", sd, torch.float32, dev)
        for k in shapes:
            out32[k] = ref_forward(model32, *ins[k], dev, torch.float32)
        del model32
        torch.cuda.empty_cache()
        set_exact_fp32(False)
        eng, _, _, _, _ = engine_objects("xxl", sd, None, dev)
        for k, (L_, Lv_, S_) in shapes.items():
            x, t, cond, clip, sync = ins[k]
            eng.set_conditions(clip.to(dev), sync.to(dev), cond.to(dev), L=L_, batch=1)
            oute[k] = eng.dit_forward(x.to(dev), t).float()
        assert eng.debug_flags()[0] == 0
        report["xxl_forward"] = {k: {"ref_bf16_vs_ref_fp32": rel_l2(out16[k], out32[k]),
                                     "engine_vs_ref_bf16": rel_l2(oute[k], out16[k]),
                                     "engine_vs_ref_fp32": rel_l2(oute[k], out32[k])} for k in shapes}
        report["xxl_forward"]["5s"]["ref_bf16_in_batch4_vs_alone (floor)"] = rel_l2(out16_b4, out16["5s"])
        torch.save({"checksum": ck2, "fwd_ref_bf16_5s": out16["5s"].cpu().bfloat16(), "fwd_ref_fp32_5s": out32["5s"].cpu().half(),
                    "fwd_ref_bf16_30s": out16["30s"].cpu().bfloat16(), "fwd_ref_fp32_30s": out32["30s"].cpu().half(),
                    "floor_5s": report["xxl_forward"]["5s"]["ref_bf16_in_batch4_vs_alone (floor)"]},
                   os.path.join(a.out, "full_xxl_forward.pt"))
    with open(os.path.join(a.out, "parity_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
