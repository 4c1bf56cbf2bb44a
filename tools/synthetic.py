"""TEST / BENCH INFRASTRUCTURE — not part of the product path (and not the oracle: no reference arithmetic here).

Synthetic, seed-deterministic state dicts with the reference's parameter names and shapes
(hifi_foley.py:455-527 for the DiT, dac.py:98-149 + utils.py:32-44 for the
DAC-VAE decoder).  No checkpoint exists offline, so every parity test and the benchmark run on these.
Many reference layers are zero-initialised (modulate_layers.py:12-13, mlp_layers.py:86-95), which would
make parity vacuous; here every tensor is random and the norm/gate paths are exercised.

Each tensor is drawn from its own generator seeded by (seed, crc32(name)) so single tensors can be
regenerated independently and the result does not depend on iteration order.
"""
import zlib
from collections import OrderedDict

import torch

MODEL_CONFIGS = {
    # configs/hunyuanvideo-foley-xxl.yaml / -xl.yaml (model_config.model_kwargs)
    "xxl": dict(hidden_size=1536, num_heads=12, depth_triple_blocks=18, depth_single_blocks=36, mlp_ratio=4),
    "xl": dict(hidden_size=1408, num_heads=11, depth_triple_blocks=12, depth_single_blocks=24, mlp_ratio=4),
    # reduced shapes for CPU-speed parity tests (same code paths, head_dim stays 128)
    "tiny": dict(hidden_size=256, num_heads=2, depth_triple_blocks=2, depth_single_blocks=2, mlp_ratio=4),
    "small": dict(hidden_size=512, num_heads=4, depth_triple_blocks=3, depth_single_blocks=4, mlp_ratio=4),
}
COMMON = dict(clip_dim=768, sync_feat_dim=768, condition_dim=768, audio_vae_latent_dim=128, freq_dim=256,
              rope_theta=10000.0, text_length=77, audio_frame_rate=50)


def model_config(name):
    c = dict(COMMON)
    c.update(MODEL_CONFIGS[name])
    c["name"] = name
    c["head_dim"] = c["hidden_size"] // c["num_heads"]
    c["mlp_hidden_triple"] = int(c["hidden_size"] * c["mlp_ratio"])
    c["mlp_hidden_single"] = convmlp_hidden(c["hidden_size"] * c["mlp_ratio"])
    c["sync_hidden"] = convmlp_hidden(c["hidden_size"] * 4)
    return c


def convmlp_hidden(hidden_dim, multiple_of=256):
    """ConvMLP's hidden width rule (mlp_layers.py:133-134)."""
    h = int(2 * hidden_dim / 3)
    return multiple_of * ((h + multiple_of - 1) // multiple_of)


def dit_param_specs(c):
    """[(name, shape, kind)] in the reference's state-dict naming; kind picks the distribution."""
    C, F = c["hidden_size"], c["mlp_hidden_triple"]
    Hs, Hy = c["mlp_hidden_single"], c["sync_hidden"]
    D = c["head_dim"]
    s = []

    def lin(name, out_f, in_f, bias=True, k=None):
        s.append((name + ".weight", (out_f, in_f) if k is None else (out_f, in_f, k), "w"))
        if bias:
            s.append((name + ".bias", (out_f,), "b"))

    lin("audio_embedder.proj", C, c["audio_vae_latent_dim"], k=1)
    lin("visual_proj.w1", C, c["clip_dim"], bias=False)
    lin("visual_proj.w2", C, C, bias=False)
    lin("visual_proj.w3", C, c["clip_dim"], bias=False)
    lin("cond_in.linear_1", C, c["condition_dim"])
    lin("cond_in.linear_2", C, C)
    lin("time_in.mlp.0", C, c["freq_dim"])
    lin("time_in.mlp.2", C, C)
    s.append(("sync_pos_emb", (1, 1, 8, c["sync_feat_dim"]), "e"))
    lin("sync_in.0", C, c["sync_feat_dim"])
    lin("sync_in.2.w1", Hy, C, bias=False, k=1)
    lin("sync_in.2.w2", C, Hy, bias=False, k=1)
    lin("sync_in.2.w3", Hy, C, bias=False, k=1)
    for i in range(c["depth_triple_blocks"]):
        p = f"triple_blocks.{i}."
        for st in ("audio", "v_cond"):
            lin(p + f"{st}_mod.linear", 9 * C, C)
        lin(p + "audio_self_attn_qkv", 3 * C, C)
        s.append((p + "audio_self_q_norm.weight", (D,), "n"))
        s.append((p + "audio_self_k_norm.weight", (D,), "n"))
        lin(p + "audio_self_proj", C, C)
        lin(p + "v_cond_attn_qkv", 3 * C, C)
        s.append((p + "v_cond_attn_q_norm.weight", (D,), "n"))
        s.append((p + "v_cond_attn_k_norm.weight", (D,), "n"))
        lin(p + "v_cond_self_proj", C, C)
        lin(p + "audio_cross_q", C, C)
        lin(p + "v_cond_cross_q", C, C)
        lin(p + "text_cross_kv", 2 * C, C)
        s.append((p + "audio_cross_q_norm.weight", (D,), "n"))
        s.append((p + "v_cond_cross_q_norm.weight", (D,), "n"))
        s.append((p + "text_cross_k_norm.weight", (D,), "n"))
        lin(p + "audio_cross_proj", C, C)
        lin(p + "v_cond_cross_proj", C, C)
        for st in ("audio", "v_cond"):
            lin(p + f"{st}_mlp.fc1", F, C)
            lin(p + f"{st}_mlp.fc2", C, F)
    for i in range(c["depth_single_blocks"]):
        p = f"single_blocks.{i}."
        lin(p + "modulation.linear", 6 * C, C)
        lin(p + "linear_qkv", 3 * C, C)
        lin(p + "linear1", C, C, k=3)
        lin(p + "linear2.w1", Hs, C, bias=False, k=3)
        lin(p + "linear2.w2", C, Hs, bias=False, k=3)
        lin(p + "linear2.w3", Hs, C, bias=False, k=3)
        s.append((p + "q_norm.weight", (D,), "n"))
        s.append((p + "k_norm.weight", (D,), "n"))
    lin("final_layer.linear", c["audio_vae_latent_dim"], C)
    lin("final_layer.adaLN_modulation.1", 2 * C, C)  # dead on this path (modulate_layers.py:20-22)
    s.append(("empty_clip_feat", (1, c["clip_dim"]), "e"))
    s.append(("empty_sync_feat", (1, c["sync_feat_dim"]), "e"))
    return s


# DAC-VAE decoder hyper-parameters (utils.py:32-44 _DAC_KWARGS)
DAC_CONFIG = dict(latent_dim=128, decoder_dim=2048, decoder_rates=(8, 5, 4, 3, 2), sample_rate=48000)
DAC_TINY = dict(latent_dim=128, decoder_dim=1024, decoder_rates=(8, 5, 4, 3, 2), sample_rate=48000)


def dac_param_specs(d):
    """Decoder + post_quant_conv parameters (dac.py:119-149, 231-233), weight-norm parametrised."""
    s = []

    def wn(name, shape):
        # weight_norm(dim=0): original0 = g [shape[0],1,1], original1 = v
        s.append((name + ".parametrizations.weight.original0", (shape[0], 1, 1), "g"))
        s.append((name + ".parametrizations.weight.original1", shape, "v"))

    L, C = d["latent_dim"], d["decoder_dim"]
    s.append(("post_quant_conv.weight", (L, L, 1), "v"))
    s.append(("post_quant_conv.bias", (L,), "b"))
    wn("decoder.model.0", (C, L, 7))
    s.append(("decoder.model.0.bias", (C,), "b"))
    out_dim = C
    for i, stride in enumerate(d["decoder_rates"]):
        in_dim, out_dim = C // 2 ** i, C // 2 ** (i + 1)
        p = f"decoder.model.{i + 1}.block."
        s.append((p + "0.alpha", (1, in_dim, 1), "a"))
        wn(p + "1", (in_dim, out_dim, 2 * stride))  # ConvTranspose1d weight is [C_in, C_out, k]
        s.append((p + "1.bias", (out_dim,), "b"))
        for j in range(3):
            q = p + f"{j + 2}.block."
            s.append((q + "0.alpha", (1, out_dim, 1), "a"))
            wn(q + "1", (out_dim, out_dim, 7))
            s.append((q + "1.bias", (out_dim,), "b"))
            s.append((q + "2.alpha", (1, out_dim, 1), "a"))
            wn(q + "3", (out_dim, out_dim, 1))
            s.append((q + "3.bias", (out_dim,), "b"))
    n = len(d["decoder_rates"]) + 1
    s.append((f"decoder.model.{n}.alpha", (1, out_dim, 1), "a"))
    wn(f"decoder.model.{n + 1}", (1, out_dim, 7))
    s.append((f"decoder.model.{n + 1}.bias", (1,), "b"))
    return s


def _draw(name, shape, kind, seed, dtype):
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFF)
    if kind == "w":      # linear / conv weights: fan-in scaled so activations stay O(1) through depth
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        t = torch.randn(shape, generator=g, dtype=torch.float32) * (0.7 / fan_in ** 0.5)
    elif kind == "b":
        t = torch.randn(shape, generator=g, dtype=torch.float32) * 0.05
    elif kind == "n":    # RMSNorm weights around 1
        t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
    elif kind == "e":    # embeddings / learned empty features
        t = torch.randn(shape, generator=g, dtype=torch.float32) * 0.5
    elif kind == "a":    # snake alpha, positive
        t = 0.5 + torch.rand(shape, generator=g, dtype=torch.float32)
    elif kind == "v":    # weight-norm direction
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        t = torch.randn(shape, generator=g, dtype=torch.float32) / fan_in ** 0.5
    elif kind == "g":    # weight-norm magnitude per dim-0 slice, ~ what keeps conv outputs O(1)
        t = 0.6 + 0.3 * torch.rand(shape, generator=g, dtype=torch.float32)
    else:
        raise ValueError(kind)
    return t.to(dtype)


def synth_dit_state_dict(cfg, seed=0, dtype=torch.float32):
    return OrderedDict((n, _draw(n, sh, k, seed, dtype)) for n, sh, k in dit_param_specs(cfg))


def synth_dac_state_dict(dcfg=DAC_CONFIG, seed=0, dtype=torch.float32):
    return OrderedDict((n, _draw(n, sh, k, seed, dtype)) for n, sh, k in dac_param_specs(dcfg))


def synth_conditions(cfg, L, Lv, S, T_prompt=9, T_neg=5, seed=1, dtype=torch.float32):
    """Synthetic condition features: SigLIP2 [1,Lv,768], Synchformer [1,S,768],
    CLAP hidden states for prompt / negative prompt [1,T,768]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32).to(dtype)
    return dict(siglip2_feat=r(1, Lv, cfg["clip_dim"]), syncformer_feat=r(1, S, cfg["sync_feat_dim"]),
                text_feat=r(1, T_prompt, cfg["condition_dim"]), uncond_text_feat=r(1, T_neg, cfg["condition_dim"]))


def clip_lengths(duration_s):
    """Token counts the Sampler derives from a duration (nodes.py:294-317, 326-331)."""
    L = int(duration_s * 50)
    Lv = int(duration_s * 8)
    n25 = int(duration_s * 25)
    S = ((n25 - 16) // 8 + 1) * 8
    return L, Lv, S


def synth_state_dict_cuda(specs, seed, device, dtype):
    """Same specs, drawn directly on the GPU (bench only: different values than the CPU draw, same statistics)."""
    import torch
    out = OrderedDict()
    g = torch.Generator(device=device)
    scale = {"b": 0.05, "e": 0.5}
    for name, shape, kind in specs:
        g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFF)
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        if kind == "w":
            t = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * (0.7 / fan_in ** 0.5)
        elif kind == "v":
            t = torch.randn(shape, generator=g, device=device, dtype=torch.float32) / fan_in ** 0.5
        elif kind == "n":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        elif kind == "a":
            t = 0.5 + torch.rand(shape, generator=g, device=device, dtype=torch.float32)
        elif kind == "g":
            t = 0.6 + 0.3 * torch.rand(shape, generator=g, device=device, dtype=torch.float32)
        else:
            t = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * scale[kind]
        out[name] = t.to(dtype)
    return out


def motionformer_param_specs(depth=12, C=768, F=3072):
    """State dict of the reference's MotionFormer as Synchformer.__init__ builds it (divided_224_16x4, spatial aggregation
    layer; models/synchformer/motionformer.py, video_model_builder.py): (name, shape, kind) in module order."""
    specs = [("cls_token", (1, 1, C), "e"), ("pos_embed", (1, 197, C), "e"), ("temp_embed", (1, 8, C), "e"),
             ("patch_embed.proj.weight", (C, 3, 16, 16), "w"), ("patch_embed.proj.bias", (C,), "b"),
             ("patch_embed_3d.proj.weight", (C, 3, 2, 16, 16), "w"), ("patch_embed_3d.proj.bias", (C,), "b")]
    for i in range(depth):
        p = f"blocks.{i}."
        specs += [(p + "norm1.weight", (C,), "n"), (p + "norm1.bias", (C,), "b"),
                  (p + "attn.qkv.weight", (3 * C, C), "w"), (p + "attn.qkv.bias", (3 * C,), "b"),
                  (p + "attn.proj.weight", (C, C), "w"), (p + "attn.proj.bias", (C,), "b"),
                  (p + "timeattn.qkv.weight", (3 * C, C), "w"), (p + "timeattn.qkv.bias", (3 * C,), "b"),
                  (p + "timeattn.proj.weight", (C, C), "w"), (p + "timeattn.proj.bias", (C,), "b"),
                  (p + "norm2.weight", (C,), "n"), (p + "norm2.bias", (C,), "b"),
                  (p + "mlp.fc1.weight", (F, C), "w"), (p + "mlp.fc1.bias", (F,), "b"),
                  (p + "mlp.fc2.weight", (C, F), "w"), (p + "mlp.fc2.bias", (C,), "b"),
                  (p + "norm3.weight", (C,), "n"), (p + "norm3.bias", (C,), "b")]
    a = "spatial_attn_agg."
    specs += [("norm.weight", (C,), "n"), ("norm.bias", (C,), "b"), (a + "cls_token", (1, 1, C), "e"),
              (a + "self_attn.in_proj_weight", (3 * C, C), "w"), (a + "self_attn.in_proj_bias", (3 * C,), "b"),
              (a + "self_attn.out_proj.weight", (C, C), "w"), (a + "self_attn.out_proj.bias", (C,), "b"),
              (a + "linear1.weight", (F, C), "w"), (a + "linear1.bias", (F,), "b"),
              (a + "linear2.weight", (C, F), "w"), (a + "linear2.bias", (C,), "b"),
              (a + "norm1.weight", (C,), "n"), (a + "norm1.bias", (C,), "b"),
              (a + "norm2.weight", (C,), "n"), (a + "norm2.bias", (C,), "b")]
    return specs


def synth_motionformer_state_dict(depth=12, seed=0, dtype=torch.float32):
    """Seeded MotionFormer weights (same values wherever they are drawn: CPU generator per tensor).  q/k/v projections get
    1.5 x the fan-in scale so that the attention maps are far from uniform."""
    sd = OrderedDict()
    for name, shape, kind in motionformer_param_specs(depth):
        t = _draw("mformer." + name, shape, kind, seed, torch.float32)
        if "qkv.weight" in name or "in_proj_weight" in name:
            t = t * 1.5
        if kind == "e":
            t = t * 0.2
        sd[name] = t.to(dtype)
    return sd


def synth_sync_frames(n_frames, seed=0):
    """Preprocessed 25 fps frames [n_frames, 3, 224, 224] in [-1, 1] with temporal structure (a drifting smooth pattern + noise)."""
    g = torch.Generator(device="cpu").manual_seed(1000 + seed)
    base = torch.rand(3, 28, 28, generator=g)
    fr = []
    for t in range(n_frames):
        img = torch.roll(base, shifts=(t // 2, t // 3), dims=(1, 2))
        img = torch.nn.functional.interpolate(img[None], size=(224, 224), mode="bilinear", align_corners=False)[0]
        fr.append((img + 0.1 * torch.rand(3, 224, 224, generator=g)).clamp(0, 1) * 2 - 1)
    return torch.stack(fr)
