"""Condition encoders through the C ABI against the HF modules the reference runs
(feature_utils.py:64-79 `siglip2_model.get_image_features(...).pooler_output`, :132-138 `clap_model(...).last_hidden_state`),
instantiated from their configs with seeded random weights (no checkpoint is available offline) and moved to the device in
bf16 exactly as the reference's Sampler does (nodes.py:283-284).

Tolerance (floating point, stated here): the reference path is a bf16 module, so its own distance to the fp32 module is the
floor.  Asserted: engine vs HF-fp32 <= 1.25 x (HF-bf16 vs HF-fp32) + 5e-4, and engine vs HF-bf16 <= 2 x that floor.  The
head_dim-64 attention kernels alone: <= 4e-3 against fp32 softmax(QK^T/8)V on the same bf16 operands (P and the output are
rounded to bf16, 2^-9 relative each)."""
import pytest
import torch

from conftest import load_pkg, rel_l2

pytestmark = pytest.mark.gpu


def _perturb(model, seed):
    """HF's init leaves LayerNorm at (1, 0) and biases at 0: move every parameter so that each one matters."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1 and ("norm" in name.lower()) and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                p.add_(0.01 * torch.randn(p.shape, generator=g))
    return model


def _ref_attn64(q, k, v, H, mask=None):
    B, Sq, Sk = q.shape[0], q.shape[1], k.shape[1]
    qf = q.float().view(B, Sq, H, 64).transpose(1, 2)
    kf = k.float().view(B, Sk, H, 64).transpose(1, 2)
    vf = v.float().view(B, Sk, H, 64).transpose(1, 2)
    s = qf @ kf.transpose(-1, -2) * 0.125
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    return (s.softmax(-1) @ vf).transpose(1, 2).reshape(B, Sq, H * 64)


@pytest.mark.parametrize("name,B,H,Sq,Sk,impl", [
    ("siglip_frame", 2, 12, 1024, 1024, 0), ("ragged", 2, 3, 130, 77, 0), ("one_tile", 1, 2, 64, 64, 0),
    ("tiny", 1, 1, 5, 3, 0), ("tc_siglip_frame", 2, 12, 1024, 1024, 2), ("tc_ragged", 2, 3, 130, 77, 2), ("tc_one_chunk", 1, 2, 128, 128, 2),
    ("tc_tiny", 1, 1, 5, 3, 0 + 2), ("tc_odd_chunks", 1, 2, 300, 333, 2), ("tc_many_ctas", 5, 12, 1024, 1024, 2),
    ("tc1_siglip_frame", 2, 12, 1024, 1024, 4), ("tc1_ragged", 2, 3, 130, 77, 4), ("tc1_odd_chunks", 1, 2, 300, 333, 4), ("clap", 2, 12, 77, 77, 1), ("clap_long", 1, 12, 300, 300, 1), ("pool", 3, 12, 1, 1024, 1),
    ("pool_blk", 3, 12, 1, 1024, 3), ("cls_blk", 2, 12, 1, 1569, 3), ("clap_blk_masked", 2, 3, 1, 77, 3)])
def test_attention_d64_vs_fp32(name, B, H, Sq, Sk, impl):
    enc = load_pkg("encoders")
    g = torch.Generator(device="cuda").manual_seed(len(name) * 7 + Sq)
    C = H * 64
    if Sq == Sk:     # operands read in place from a fused projection output [B, S, 3C]
        qkv = torch.randn(B, Sq, 3 * C, device="cuda", generator=g).bfloat16()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = torch.randn(B, Sq, C, device="cuda", generator=g).bfloat16()
        kv = torch.randn(B, Sk, 2 * C, device="cuda", generator=g).bfloat16()
        k, v = kv[..., :C], kv[..., C:]
    mask = None
    if impl in (1, 3) and name.startswith("clap"):
        mask = torch.ones(B, Sk, dtype=torch.int32, device="cuda")
        mask[0, Sk - Sk // 3:] = 0
    out = enc.attention_d64(q, k, v, H, key_mask=mask, impl=impl)
    want = _ref_attn64(q, k, v, H, mask)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), want) < 4e-3, name
    assert load_pkg("engine").load_library() and all(f == 0 for f in _flags()), "a pipeline wait of the kernel timed out"


def _flags():
    import ctypes
    lib = load_pkg("engine").load_library()
    buf = (ctypes.c_uint32 * 4)()
    lib.foley_debug_flags(buf)
    return list(buf)


def test_pool_attention_rounds_like_multihead_attention():
    """nn.MultiheadAttention with need_weights (the SigLIP pooling head's call) rounds the bmm scores and the softmax to bf16."""
    enc = load_pkg("encoders")
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, Sk = 4, 12, 1024
    C = H * 64
    mha = torch.nn.MultiheadAttention(C, H, batch_first=True).cuda().bfloat16().eval()
    probe = torch.randn(1, 1, C, device="cuda", generator=g).bfloat16()
    xs = torch.randn(B, Sk, C, device="cuda", generator=g).bfloat16()
    with torch.inference_mode():
        Wq, Wk, Wv = mha.in_proj_weight.chunk(3)
        bq, bk, bv = mha.in_proj_bias.chunk(3)
        q = torch.nn.functional.linear(probe, Wq, bq).expand(B, 1, C).contiguous()
        k = torch.nn.functional.linear(xs, Wk, bk)
        v = torch.nn.functional.linear(xs, Wv, bv)
        want = mha(probe.expand(B, 1, C), xs, xs)[0]
        for impl in (1, 3):
            got = torch.nn.functional.linear(enc.attention_d64(q, k, v, H, round_scores=True, impl=impl), mha.out_proj.weight, mha.out_proj.bias)
            assert rel_l2(got.float(), want.float()) < 4e-3, impl


def _siglip(layers, image, seed):
    from transformers import SiglipVisionConfig, SiglipVisionModel
    cfg = SiglipVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=layers, num_attention_heads=12,
                             image_size=image, patch_size=16, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)
    torch.manual_seed(seed)
    return _perturb(SiglipVisionModel(cfg).eval(), seed + 1)


def _gate(got, ref_bf16, ref_fp32, what):
    floor = rel_l2(ref_bf16.float(), ref_fp32)
    d32, d16 = rel_l2(got.float(), ref_fp32), rel_l2(got.float(), ref_bf16.float())
    print(f"{what}: engine vs fp32 {d32:.3e}, engine vs HF bf16 {d16:.3e}, HF bf16 vs fp32 (floor) {floor:.3e}")
    assert torch.isfinite(got.float()).all()
    assert d32 <= 1.25 * floor + 5e-4, (what, d32, floor)
    assert d16 <= 2.0 * floor + 5e-4, (what, d16, floor)


@pytest.mark.parametrize("layers,image,frames", [(2, 64, 5), (3, 512, 2), (12, 512, 3)])
def test_siglip_vision_vs_hf(layers, image, frames):
    enc = load_pkg("encoders")
    model = _siglip(layers, image, 11 * layers)
    g = torch.Generator().manual_seed(3)
    px = (torch.rand(frames, 3, image, image, generator=g) * 2 - 1).cuda()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.inference_mode():
        m32 = model.cuda().float()
        ref32 = m32(pixel_values=px).pooler_output.float()
        e = enc.SiglipVisionEncoder.from_hf(m32)
        m16 = m32.to(torch.bfloat16)
        ref16 = m16(pixel_values=px).pooler_output
    got = e.encode(px)
    torch.cuda.synchronize()
    assert got.shape == (frames, 768) and got.dtype == torch.bfloat16
    _gate(got, ref16, ref32, f"siglip L={layers} img={image}")
    # frames are independent: a one-frame call reproduces its row bit for bit (same kernels, same tiles per frame)
    one = e.encode(px[1:2])
    assert torch.equal(one[0], got[1])
    # the mma.sync attention kernel (option att_tc = 0) gives the same features within the attention kernels' own tolerance
    e.set_option("att_tc", 0)
    alt = e.encode(px)
    assert rel_l2(alt.float(), got.float()) < 1.5e-2
    _gate(alt, ref16, ref32, f"siglip L={layers} img={image} (mma.sync attention)")


def test_siglip_chunked_passes_match():
    """max_frames_per_pass only bounds activation memory: the result does not depend on it."""
    enc = load_pkg("encoders")
    model = _siglip(2, 64, 4)
    px = (torch.rand(7, 3, 64, 64, generator=torch.Generator().manual_seed(9)) * 2 - 1).cuda()
    cfg = dict(enc.SIGLIP2_BASE_512, num_layers=2, image_size=64)
    a = enc.SiglipVisionEncoder(cfg).load_state_dict(model.state_dict()).finalize()
    b = enc.SiglipVisionEncoder(cfg, max_frames_per_pass=3).load_state_dict(model.state_dict()).finalize()
    assert torch.equal(a.encode(px), b.encode(px))


def test_siglip_layer_taps_vs_hf():
    """Residual stream after every layer (option layers_run) against HF's hidden_states, bf16 module vs fp32 module."""
    enc = load_pkg("encoders")
    model = _siglip(4, 64, 21)
    px = (torch.rand(3, 3, 64, 64, generator=torch.Generator().manual_seed(2)) * 2 - 1).cuda()
    with torch.inference_mode():
        m32 = model.cuda().float()
        hs32 = m32(pixel_values=px, output_hidden_states=True).hidden_states
        e = enc.SiglipVisionEncoder.from_hf(m32)
        hs16 = m32.to(torch.bfloat16)(pixel_values=px, output_hidden_states=True).hidden_states
    n = 3 * 16 * 768
    for i in range(5):
        e.set_option("layers_run", i)
        e.encode(px)
        x = e.debug_read("x", n).view(3, 16, 768)
        _gate(x, hs16[i].cpu(), hs32[i].float().cpu(), f"siglip residual stream after {i} layers")


def _clap(layers, seed):
    from transformers import ClapTextConfig, ClapTextModelWithProjection
    cfg = ClapTextConfig(num_hidden_layers=layers)
    torch.manual_seed(seed)
    return _perturb(ClapTextModelWithProjection(cfg).eval(), seed + 1)


@pytest.mark.parametrize("layers,T", [(2, 9), (12, 24), (12, 77)])
def test_clap_text_vs_hf(layers, T):
    enc = load_pkg("encoders")
    model = _clap(layers, 5 + layers)
    g = torch.Generator().manual_seed(T)
    ids = torch.randint(3, 50265, (2, T), generator=g)
    ids[:, 0] = 0
    n0 = max(3, T // 2)                     # the negative prompt is shorter: right-padded with pad_token_id = 1
    ids[0, n0 - 1] = 2
    ids[0, n0:] = 1
    ids[1, T - 1] = 2
    mask = (ids != 1).long()
    with torch.inference_mode():
        m32 = model.cuda().float()
        ref32 = m32(input_ids=ids.cuda(), attention_mask=mask.cuda(), output_hidden_states=True).last_hidden_state.float()
        e = enc.ClapTextEncoder.from_hf(m32)
        ref16 = m32.to(torch.bfloat16)(input_ids=ids.cuda(), attention_mask=mask.cuda(), output_hidden_states=True).last_hidden_state
    got = e.encode(ids, mask)
    torch.cuda.synchronize()
    assert got.shape == (2, T, 768)
    _gate(got, ref16, ref32, f"clap L={layers} T={T}")
    # the rows the reference hands on (padded positions included) are all there and finite
    assert torch.isfinite(got[0, n0:].float()).all()


def test_encoder_errors():
    enc = load_pkg("encoders")
    E = load_pkg("engine")
    with pytest.raises(E.FoleyError):
        enc.SiglipVisionEncoder(dict(enc.SIGLIP2_BASE_512, num_heads=8))          # head_dim 96
    e = enc.SiglipVisionEncoder(dict(enc.SIGLIP2_BASE_512, num_layers=1, image_size=64))
    with pytest.raises(E.FoleyError, match="missing tensor"):
        e.finalize()
    c = enc.ClapTextEncoder(dict(enc.CLAP_TEXT_GENERAL, num_layers=1))
    with pytest.raises(E.FoleyError):
        c.encode(torch.zeros(1, 4, dtype=torch.long))                              # not finalized: weights missing


class _FakeTokenizer:
    """Stands in for HF's tokenizer (no vocabulary files offline): deterministic ids, right padding with pad id 1."""
    model_max_length = 512

    def __call__(self, texts, padding=True, return_tensors="pt"):
        rows = [[0] + [3 + (ord(c) * 31 + i) % 50000 for i, c in enumerate(t)][:60] + [2] for t in texts]
        T = max(len(r) for r in rows)
        ids = torch.tensor([r + [1] * (T - len(r)) for r in rows])
        return {"input_ids": ids, "attention_mask": (ids != 1).long()}


def test_sampler_runs_frames_to_waveform_on_the_engine_encoders():
    """Frames -> GPU preprocessing -> SigLIP2, Synchformer and CLAP text on the engine -> denoise -> DAC through
    HunyuanFoleySampler.generate_audio (no torch module, no reference package anywhere on the path); the features the Sampler consumed equal the HF modules'
    on the same preprocessed frames / token ids within the bf16 floor."""
    nodes, cfgmod, E, enc, fb = (load_pkg(m) for m in ("nodes", "config", "engine", "encoders", "feature_bridge"))
    from oracle import weights as W
    c = W.model_config("tiny")
    sd = W.synth_dit_state_dict(c, seed=0)
    cfg = cfgmod.load_model_config("xxl")
    for k in ("hidden_size", "num_heads", "depth_triple_blocks", "depth_single_blocks"):
        cfg.model_config.model_kwargs[k] = c[k]
    eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs))
    eng.load_state_dict(sd)
    eng.finalize()
    model = nodes.FoleyModel(eng, sd["empty_clip_feat"], sd["empty_sync_feat"], cfg, dtype=torch.bfloat16)
    dac = nodes.FoleyDAC.from_state_dict(W.synth_dac_state_dict(W.DAC_TINY, seed=3))
    hf_sig, hf_clap = _siglip(2, 512, 31).cuda(), _clap(2, 32).cuda()
    sig = enc.SiglipVisionEncoder.from_hf(hf_sig)
    clap = enc.ClapTextEncoder.from_hf(hf_clap)
    L, Lv, S = W.clip_lengths(1.0)
    seen = {}

    from tools import synthetic as SY
    sync = enc.SynchformerEncoder.from_state_dict(SY.synth_motionformer_state_dict(1, seed=5), **dict(enc.MOTIONFORMER_DIVIDED_224, num_layers=1))

    def sync_encode(frames):
        seen["sync_in"] = frames
        return sync.encode(frames[0])

    inner = fb.make_extract_features(sig, _FakeTokenizer(), clap, sync_encode, torch.device("cuda"))

    def extract(pre8, pre25, prompt, negative_prompt):
        out = inner(pre8, pre25, prompt, negative_prompt)
        seen["pre8"], seen["out"] = pre8, out
        return out

    deps = cfgmod.AttributeDict({"dac_model": dac, "extract_features": extract, "preprocessed_inputs": True,
                                 "clap_tokenizer": _FakeTokenizer(), "clap_model": clap})
    image = torch.rand(8, 48, 64, 3, generator=torch.Generator().manual_seed(4))
    first, batch = nodes.HunyuanFoleySampler().generate_audio(model, deps, 8.0, 1.0, "rain on a tin roof", "noisy", 4.5, 4,
                                                              "euler", 1, 0, True, image=image)
    assert batch["waveform"].shape == (1, 1, 48000) and torch.isfinite(batch["waveform"]).all()
    visual, text, audio_len = seen["out"]
    assert visual["siglip2_feat"].shape == (1, 8, 768) and seen["sync_in"].shape == (1, 25, 3, 224, 224) and audio_len == 1.0
    assert visual["syncformer_feat"].shape == (1, 16, 768) and torch.isfinite(visual["syncformer_feat"]).all()
    tok = _FakeTokenizer()(["noisy", "rain on a tin roof"])
    assert text["uncond_text_feat"].shape == (1, tok["input_ids"].shape[1], 768) and text["text_feat"].shape == text["uncond_text_feat"].shape
    with torch.inference_mode():
        want_sig = hf_sig.to(torch.bfloat16)(pixel_values=seen["pre8"]).pooler_output
        want_txt = hf_clap.to(torch.bfloat16)(input_ids=tok["input_ids"].cuda(), attention_mask=tok["attention_mask"].cuda()).last_hidden_state
    assert rel_l2(visual["siglip2_feat"][0].float(), want_sig.float()) < 2e-2
    assert rel_l2(torch.cat([text["uncond_text_feat"], text["text_feat"]]).float(), want_txt.float()) < 2e-2


# ---------------------------------------------------------------------------------------------- Synchformer (MotionFormer)
@pytest.mark.parametrize("depth", [2, 12])
def test_synchformer_vs_reference_golden(depth, golden_dir):
    """The Synchformer visual extractor on the engine against the reference's OWN MotionFormer run on a B200
    (tools/gpu_synchformer_golden.py: staged reference, bf16 parameters under fp16 autocast as the Sampler runs it —
    feature_utils.py:100-102, nodes.py:283-284 — and the fp32 module as ground truth), same seeded weights and frames.
    Tolerance: engine vs reference-autocast <= 2e-3 (measured 4.8e-4: same fp16 rounding points, other summation orders) and
    engine vs fp32 <= 1.1 x (reference-autocast vs fp32, 7.4e-3)."""
    import os
    enc = load_pkg("encoders")
    from tools import synthetic as SY
    g = torch.load(os.path.join(golden_dir, f"synchformer_d{depth}.pt"))
    frames = SY.synth_sync_frames(g["n_frames"], seed=g["frames_seed"]).cuda()
    cfg = dict(enc.MOTIONFORMER_DIVIDED_224, num_layers=depth)
    e = enc.SynchformerEncoder.from_state_dict(SY.synth_motionformer_state_dict(depth, seed=g["weights_seed"]), **cfg)
    got = e.encode(frames)
    torch.cuda.synchronize()
    S = (g["n_frames"] - 16) // 8 + 1
    assert got.shape == (1, S * 8, 768) and got.dtype == torch.float32 and torch.isfinite(got).all()
    got = got[0].view(S, 8, 768).cpu()
    d16, d32 = rel_l2(got, g["out_autocast"]), rel_l2(got, g["out_fp32"])
    print(f"synchformer depth {depth}: engine vs reference autocast {d16:.3e}, vs fp32 {d32:.3e}, reference autocast vs fp32 {g['autocast_vs_fp32']:.3e}")
    assert d16 <= 2e-3
    assert d32 <= 1.1 * g["autocast_vs_fp32"]
    assert all(f == 0 for f in _flags())
    # Synchformer state-dict names (vfeat_extractor.*) with the audio extractor's tensors in between are accepted / ignored
    sd = {"vfeat_extractor." + k: v for k, v in SY.synth_motionformer_state_dict(depth, seed=g["weights_seed"]).items()}
    sd["afeat_extractor.ast.embeddings.cls_token"] = torch.zeros(1, 1, 768)
    sd["transformer.pos_emb_cfg.pos_emb"] = torch.zeros(1, 8, 768)
    e2 = enc.SynchformerEncoder.from_state_dict(sd, **cfg)
    assert torch.equal(e2.encode(frames)[0].view(S, 8, 768).cpu(), got)
    # windows are independent: a clip with one more window reproduces the first ones bit for bit
    longer = SY.synth_sync_frames(g["n_frames"] + 8, seed=g["frames_seed"]).cuda()
    longer[: g["n_frames"]] = frames
    assert torch.equal(e.encode(longer)[0, : S * 8].cpu().view(S, 8, 768), got)


def test_synchformer_refuses_short_clips_and_missing_weights():
    enc, E = load_pkg("encoders"), load_pkg("engine")
    from tools import synthetic as SY
    cfg = dict(enc.MOTIONFORMER_DIVIDED_224, num_layers=1)
    e = enc.SynchformerEncoder.from_state_dict(SY.synth_motionformer_state_dict(1), **cfg)
    with pytest.raises(E.FoleyError, match="16 frames"):
        e.encode(torch.zeros(15, 3, 224, 224))
    sd = SY.synth_motionformer_state_dict(1)
    del sd["blocks.0.timeattn.qkv.weight"]
    with pytest.raises(E.FoleyError, match="missing tensor"):
        enc.SynchformerEncoder.from_state_dict(sd, **cfg)
