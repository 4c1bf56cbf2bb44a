"""Generates tests/golden/*.pt by running the REFERENCE's own modules (imported read-only from
/root/reference via tools/ref_shims.py) on the seeded synthetic weights of oracle/weights.py.

Run in the authoring container only (the reference does not exist on the GPU box):
    python tools/make_golden.py [--only NAME]

Fixtures (inputs are re-derived from seeds by the tests; only outputs are stored):
  dit_tiny_fp32.pt       HunyuanVideoFoley.forward, tiny config, fp32 CPU                  (pins oracle fp32)
  dit_tiny_cudabf16.pt   same modules with bf16 weights under an emulation of torch's CUDA autocast cast
                         policy on CPU (see CudaAutocastEmu)                               (pins oracle cuda_bf16)
  dit_small_fp32.pt      small config (3+4 blocks, hidden 512), Lv != multiple shapes
  denoise_tiny_*.pt      denoise_process_with_generator (utils.py:125-258): T2A CFG-off 10 steps (config #1
                         shape) and V2A CFG 4.5 4 steps, incl. DAC decode of a reduced-width DAC
  dac_full_L25.pt        DAC.decode with the full-size decoder (decoder_dim 2048) on 0.5 s of latents
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F
from torch.overrides import TorchFunctionMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shims  # noqa: E402
from oracle import weights as W  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class CudaAutocastEmu(TorchFunctionMode):
    """Applies torch's CUDA autocast cast policy to CPU tensors so the reference's dataflow under
    torch.autocast('cuda', bf16) can be reproduced without a GPU: lower_precision_fp ops (linear, conv1d,
    SDPA) take bf16 inputs; fp32-list ops (layer_norm, interpolate) take fp32 inputs and return fp32; ops
    without an autocast rule run in their input dtype with type promotion.  nn.RMSNorm(eps=None) is replaced
    by what the CUDA kernel computes for bf16 (profiles/r01_torch_probe.json: fp32 eps, single rounding)."""

    LOW = {F.linear, F.conv1d, F.scaled_dot_product_attention, torch.matmul, torch.bmm, F.conv_transpose1d}
    F32 = {F.layer_norm, F.interpolate}

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}

        def cast(x, dt):
            return x.to(dt) if isinstance(x, torch.Tensor) and x.is_floating_point() else x

        if func in self.LOW:
            args = tuple(cast(a, torch.bfloat16) for a in args)
            kwargs = {k: cast(v, torch.bfloat16) for k, v in kwargs.items()}
        elif func in self.F32:
            args = tuple(cast(a, torch.float32) for a in args)
            kwargs = {k: cast(v, torch.float32) for k, v in kwargs.items()}
        elif func is F.rms_norm:
            x, shape = args[0], args[1]
            w = kwargs.get("weight", args[2] if len(args) > 2 else None)
            xf = x.float()
            y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + torch.finfo(torch.float32).eps)
            if w is not None:
                y = y * w.float()
            return y.to(x.dtype)
        return func(*args, **kwargs)


def ref_cfg(ns, name):
    cfg = ns.load_yaml(os.path.join(ns.config_dir, "hunyuanvideo-foley-xxl.yaml"))
    mc = W.MODEL_CONFIGS[name]
    for k, v in mc.items():
        cfg.model_config.model_kwargs[k] = v
    return cfg


def build_ref_model(ns, name, dtype, seed=0):
    cfg = ref_cfg(ns, name)
    model = ns.HunyuanVideoFoley(cfg, dtype=torch.float32)
    sd = W.synth_dit_state_dict(W.model_config(name), seed=seed)
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    model = model.to(dtype).eval()
    return model, cfg, sd


def dit_inputs(c, B, L, Lv, S, seed=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, c["audio_vae_latent_dim"], L, generator=g)
    t = torch.tensor([875.0, 120.0, 500.0, 999.0][:B])
    cond = torch.randn(B, 77, c["condition_dim"], generator=g)
    cond[:, 9:] = 0  # zero-padded prompt tail, as the sampler feeds it (utils.py:186-188)
    clip = torch.randn(B, Lv, c["clip_dim"], generator=g)
    sync = torch.randn(B, S, c["sync_feat_dim"], generator=g)
    return x, t, cond, clip, sync


def gold_dit(ns, name, tag, B, L, Lv, S, policy):
    c = W.model_config(name)
    dtype = torch.float32 if policy == "fp32" else torch.bfloat16
    model, _, _ = build_ref_model(ns, name, dtype)
    x, t, cond, clip, sync = dit_inputs(c, B, L, Lv, S)
    with torch.inference_mode():
        if policy == "fp32":
            out = model(x=x, t=t, cond=cond, clip_feat=clip, sync_feat=sync)["x"]
        else:
            with CudaAutocastEmu():
                out = model(x=x.bfloat16(), t=t, cond=cond.bfloat16(), clip_feat=clip.bfloat16(),
                            sync_feat=sync.bfloat16())["x"]
    path = os.path.join(GOLD, f"dit_{tag}.pt")
    torch.save({"out": out.float().clone(), "out_dtype": str(out.dtype), "config": name,
                "shape": dict(B=B, L=L, Lv=Lv, S=S), "policy": policy}, path)
    print("wrote", path, tuple(out.shape), out.dtype, float(out.float().norm()))


def small_dac(ns, dcfg):
    dac = ns.DAC(encoder_dim=16, encoder_rates=[2, 3, 4, 5, 8], latent_dim=dcfg["latent_dim"],
                 decoder_dim=dcfg["decoder_dim"], decoder_rates=list(dcfg["decoder_rates"]), n_codebooks=9,
                 codebook_size=1024, codebook_dim=8, quantizer_dropout=False, sample_rate=48000, continuous=True)
    sd = W.synth_dac_state_dict(dcfg, seed=3)
    res = dac.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(not k.startswith(("decoder", "post_quant")) for k in res.missing_keys), res.missing_keys
    return dac.eval().float()


def gold_denoise(ns, tag, duration, steps, guidance, batch, v2a, sampler="euler"):
    name = "tiny"
    c = W.model_config(name)
    model, cfg, sd = build_ref_model(ns, name, torch.float32)
    dac = small_dac(ns, W.DAC_TINY)
    L, Lv, S = W.clip_lengths(duration)
    feats = W.synth_conditions(c, L, Lv, S)
    if v2a:
        visual = {"siglip2_feat": feats["siglip2_feat"], "syncformer_feat": feats["syncformer_feat"]}
    else:  # nodes.py:326-338: learned empty features
        visual = {"siglip2_feat": model.get_empty_clip_sequence(bs=1, len=Lv).detach(),
                  "syncformer_feat": model.get_empty_sync_sequence(bs=1, len=S).detach()}
    text = {"text_feat": feats["text_feat"], "uncond_text_feat": feats["uncond_text_feat"]}
    md = ns.AttributeDict({"foley_model": model, "dac_model": dac, "device": torch.device("cpu")})
    captured = {}
    orig_decode = dac.decode

    def decode_spy(z):
        captured["latents"] = z.detach().clone()
        return orig_decode(z)

    dac.decode = decode_spy
    gen = torch.Generator(device="cpu").manual_seed(123)
    audio, sr = ns.utils.denoise_process_with_generator(visual, text, duration, md, cfg, guidance_scale=guidance,
                                                        num_inference_steps=steps, batch_size=batch,
                                                        sampler=sampler, generator=gen)
    path = os.path.join(GOLD, f"denoise_{tag}.pt")
    torch.save({"latents": captured["latents"].float(), "audio": audio.float().half(), "sr": sr,
                "args": dict(duration=duration, steps=steps, guidance=guidance, batch=batch, v2a=v2a, sampler=sampler)}, path)
    print("wrote", path, tuple(captured["latents"].shape), tuple(audio.shape))


def gold_dac_full(ns):
    dac = ns.DAC(**{**ns.utils._DAC_KWARGS, "encoder_dim": 16})
    sd = W.synth_dac_state_dict(W.DAC_CONFIG, seed=3)
    res = dac.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    z = torch.randn(1, 128, 25, generator=torch.Generator().manual_seed(5))
    with torch.inference_mode():
        wav = dac.eval().float().decode(z)
    path = os.path.join(GOLD, "dac_full_L25.pt")
    torch.save({"wav": wav.float().clone()}, path)
    print("wrote", path, tuple(wav.shape), float(wav.abs().mean()))


def gold_fp8_wrap(ns):
    """The reference Model Loader's quantization path (nodes.py:106-121, utils.py:410-485) on the tiny config: which
    modules _wrap_fp8_inplace really replaces, and the forward of the wrapped model (fp32 compute on CPU so that the
    only difference to dit_tiny_fp32 is the FP8 storage of the weights)."""
    import json
    c = W.model_config("tiny")
    model, _, sd = build_ref_model(ns, "tiny", torch.bfloat16)      # loader: params in the compute dtype ...
    linear_like = [n for n, m in model.named_modules() if isinstance(m, (torch.nn.Linear, torch.nn.Conv1d, torch.nn.Conv2d))]
    counts, _ = ns.utils._wrap_fp8_inplace(model, quantization="fp8_e4m3fn", state_dict=sd)   # ... then the FP8 wrap
    wrapped = [n for n, m in model.named_modules() if type(m).__name__ == "FP8WeightWrapper"]
    with open(os.path.join(GOLD, "fp8_wrapped_tiny.json"), "w") as f:
        json.dump({"wrapped": wrapped, "linear_like": linear_like, "counts": counts}, f, indent=0)
    model = model.float()
    x, t, cond, clip, sync = dit_inputs(c, 2, 50, 8, 16)
    with torch.inference_mode():
        out = model(x=x, t=t, cond=cond, clip_feat=clip, sync_feat=sync)["x"]
    path = os.path.join(GOLD, "dit_tiny_fp8e4m3_fp32.pt")
    torch.save({"out": out.float().clone(), "config": "tiny", "shape": dict(B=2, L=50, Lv=8, S=16)}, path)
    print("wrote", path, len(wrapped), "of", len(linear_like), "modules wrapped", float(out.norm()))


def gold_preprocess(ns):
    """Frame preprocessing exactly as the reference's Sampler + Dependencies Loader do it (nodes.py:293-317, 184-196;
    utils.py:270-273), executed with torch / torchvision here: hold-last-frame padding, (x*255).byte(), linspace picks at
    8 / 25 fps, v2 Resize(bicubic, antialias) [+ CenterCrop] + ToDtype(scale) + Normalize on uint8 CHW views, per frame on
    the CPU.  Full tensors for the small resize cases, SHA-256 digests for the two full pipelines (512x512 outputs are
    too big for fixtures)."""
    import hashlib
    from torchvision.transforms import v2
    siglip2_preprocess = v2.Compose([v2.Resize((512, 512), interpolation=v2.InterpolationMode.BICUBIC, antialias=True),
                                     v2.ToDtype(torch.float32, scale=True), v2.Normalize(mean=[0.5] * 3, std=[0.5] * 3)])
    syncformer_preprocess = v2.Compose([v2.Resize(224, interpolation=v2.InterpolationMode.BICUBIC, antialias=True),
                                        v2.CenterCrop(224), v2.ToDtype(torch.float32, scale=True),
                                        v2.Normalize(mean=[0.5] * 3, std=[0.5] * 3)])
    out = {"cases": []}
    for (N, H, W, duration, frame_rate, seed) in [(6, 90, 160, 1.0, 6.0, 7), (3, 250, 140, 1.0, 8.0, 8), (20, 64, 64, 2.0, 10.0, 9)]:
        image = torch.rand(N, H, W, 3, generator=torch.Generator().manual_seed(seed))
        n = int(duration * frame_rate)
        image_slice = torch.cat((image, image[-1:].repeat(n - N, 1, 1, 1)), dim=0) if n > N else image[:n]
        image_slice = (image_slice * 255.0).byte().permute(0, 3, 1, 2)
        i8 = torch.linspace(0, n - 1, int(duration * 8)).long()
        i25 = torch.linspace(0, n - 1, int(duration * 25)).long()
        p8 = torch.stack([siglip2_preprocess(f) for f in image_slice.index_select(0, i8)])
        p25 = torch.stack([syncformer_preprocess(f) for f in image_slice.index_select(0, i25)])
        out["cases"].append({
            "args": dict(N=N, H=H, W=W, duration=duration, frame_rate=frame_rate, seed=seed),
            "idx8": i8, "idx25": i25,
            "siglip2_sha256": hashlib.sha256(p8.numpy().tobytes()).hexdigest(), "siglip2_shape": tuple(p8.shape),
            "sync_sha256": hashlib.sha256(p25.numpy().tobytes()).hexdigest(), "sync_shape": tuple(p25.shape),
            "sync_frame0": p25[0].clone(),            # one full 3x224x224 frame per case (600 KB fp32 -> stored as fp16-exact? no: keep fp32)
        })
    small = []
    for (H, W, oh, ow, seed) in [(90, 160, 45, 64, 1), (37, 53, 74, 106, 2), (64, 48, 64, 24, 3), (33, 200, 17, 31, 4)]:
        img = torch.randint(0, 256, (H, W, 3), generator=torch.Generator().manual_seed(seed), dtype=torch.uint8).permute(2, 0, 1)
        r = v2.functional.resize(img, [oh, ow], interpolation=v2.InterpolationMode.BICUBIC, antialias=True)
        small.append({"args": dict(H=H, W=W, oh=oh, ow=ow, seed=seed), "out": r.contiguous().clone()})
    out["resize_u8"] = small
    path = os.path.join(GOLD, "preprocess.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_shims.load_reference()
    torch.set_num_threads(os.cpu_count())
    jobs = {
        "dit_tiny_fp32": lambda: gold_dit(ns, "tiny", "tiny_fp32", 2, 50, 8, 16, "fp32"),
        "dit_tiny_cudabf16": lambda: gold_dit(ns, "tiny", "tiny_cudabf16", 2, 50, 8, 16, "cuda_bf16"),
        "dit_small_fp32": lambda: gold_dit(ns, "small", "small_fp32", 2, 125, 20, 48, "fp32"),
        "dit_small_cudabf16": lambda: gold_dit(ns, "small", "small_cudabf16", 2, 125, 20, 48, "cuda_bf16"),
        "denoise_t2a": lambda: gold_denoise(ns, "tiny_t2a_nocfg", 1.0, 10, 1.0, 1, False),
        "denoise_v2a": lambda: gold_denoise(ns, "tiny_v2a_cfg", 1.0, 4, 4.5, 2, True),
        "denoise_heun": lambda: gold_denoise(ns, "tiny_heun2", 1.0, 6, 4.5, 1, True, "heun-2"),
        "denoise_midpoint": lambda: gold_denoise(ns, "tiny_midpoint2", 1.0, 6, 1.0, 1, True, "midpoint-2"),
        "denoise_kutta": lambda: gold_denoise(ns, "tiny_kutta4", 1.0, 8, 4.5, 1, True, "kutta-4"),
        "dac_full": lambda: gold_dac_full(ns),
        "fp8_wrap": lambda: gold_fp8_wrap(ns),
        "preprocess": lambda: gold_preprocess(ns),
    }
    for k, fn in jobs.items():
        if a.only and a.only != k:
            continue
        fn()


if __name__ == "__main__":
    main()
