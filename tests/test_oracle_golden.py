"""Pins the CPU oracle (oracle/foley_oracle.py) against outputs of the reference's own modules
(tests/golden/*.pt, produced by tools/make_golden.py from /root/reference)."""
import os

import pytest
import torch

from conftest import rel_l2
from oracle import foley_oracle as O
from oracle import weights as W


def _dit_inputs(c, B, L, Lv, S, seed=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, c["audio_vae_latent_dim"], L, generator=g)
    t = torch.tensor([875.0, 120.0, 500.0, 999.0][:B])
    cond = torch.randn(B, 77, c["condition_dim"], generator=g)
    cond[:, 9:] = 0
    clip = torch.randn(B, Lv, c["clip_dim"], generator=g)
    sync = torch.randn(B, S, c["sync_feat_dim"], generator=g)
    return x, t, cond, clip, sync


@pytest.mark.parametrize("tag,policy,tol", [
    ("tiny_fp32", "fp32", 1e-5),
    ("small_fp32", "fp32", 1e-5),
    # reference under the emulated CUDA-autocast cast policy (bf16): the oracle places the same rounding
    # points; residual differences are fp32 summation order flipping a few bf16 roundings.
    ("tiny_cudabf16", "cuda_bf16", 4e-3),
    ("small_cudabf16", "cuda_bf16", 4e-3),
])
def test_dit_forward_matches_reference(golden_dir, tag, policy, tol):
    gold = torch.load(os.path.join(golden_dir, f"dit_{tag}.pt"))
    c = W.model_config(gold["config"])
    sd = W.synth_dit_state_dict(c, seed=0)
    sh = gold["shape"]
    x, t, cond, clip, sync = _dit_inputs(c, sh["B"], sh["L"], sh["Lv"], sh["S"])
    out = O.dit_forward(sd, c, x, t, cond, clip, sync, policy=policy)
    assert out.shape == gold["out"].shape
    assert rel_l2(out, gold["out"]) <= tol


def test_bf16_policy_is_closer_to_bf16_reference_than_fp32_is(golden_dir):
    """The cuda_bf16 policy must explain most of the bf16-vs-fp32 gap of the reference itself."""
    g32 = torch.load(os.path.join(golden_dir, "dit_small_fp32.pt"))["out"]
    g16 = torch.load(os.path.join(golden_dir, "dit_small_cudabf16.pt"))
    c = W.model_config("small")
    sd = W.synth_dit_state_dict(c, seed=0)
    sh = g16["shape"]
    out = O.dit_forward(sd, c, *_dit_inputs(c, sh["B"], sh["L"], sh["Lv"], sh["S"]), policy="cuda_bf16")
    gap_ref = rel_l2(g16["out"], g32)
    assert rel_l2(out, g16["out"]) < 0.5 * gap_ref


@pytest.mark.parametrize("tag", ["tiny_t2a_nocfg", "tiny_v2a_cfg", "tiny_heun2", "tiny_midpoint2", "tiny_kutta4"])
def test_denoise_loop_and_decode_match_reference(golden_dir, tag):
    gold = torch.load(os.path.join(golden_dir, f"denoise_{tag}.pt"))
    a = gold["args"]
    c = W.model_config("tiny")
    sd = W.synth_dit_state_dict(c, seed=0)
    L, Lv, S = W.clip_lengths(a["duration"])
    feats = W.synth_conditions(c, L, Lv, S)
    if not a["v2a"]:
        feats["siglip2_feat"] = sd["empty_clip_feat"][None].expand(1, Lv, -1)
        feats["syncformer_feat"] = sd["empty_sync_feat"][None].expand(1, S, -1)
    gen = torch.Generator(device="cpu").manual_seed(123)
    noise = torch.randn((a["batch"], 128, L), generator=gen, dtype=torch.float32)
    lat = O.denoise(sd, c, feats, noise, a["steps"], a["guidance"], policy="fp32", solver=a.get("sampler", "euler"))
    assert rel_l2(lat, gold["latents"]) <= 2e-5
    dsd = W.synth_dac_state_dict(W.DAC_TINY, seed=3)
    wav = O.dac_decode(dsd, lat)
    assert wav.shape == gold["audio"].shape
    assert rel_l2(wav, gold["audio"].float()) <= 2e-3   # fixture stored as fp16


def test_dac_decode_full_size_matches_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "dac_full_L25.pt"))
    dsd = W.synth_dac_state_dict(W.DAC_CONFIG, seed=3)
    z = torch.randn(1, 128, 25, generator=torch.Generator().manual_seed(5))
    wav = O.dac_decode(dsd, z)
    assert wav.shape == (1, 1, 25 * 960)
    assert rel_l2(wav, gold["wav"]) <= 1e-5


def test_interleaved_rope_positions_closed_form():
    """Closed form vs the reference's interpolate round trip (hifi_foley.py:35-60) on index tensors."""
    import torch.nn.functional as F
    for L, Lv in [(50, 8), (250, 40), (125, 20), (1500, 240), (3000, 480), (55, 8)]:
        up = F.interpolate(torch.arange(Lv, dtype=torch.float32)[None, None], size=L, mode="nearest-exact")[0, 0]
        slots = 2 * torch.arange(L, dtype=torch.float32) + 1           # odd slots carry the visual copies
        back_pos = F.interpolate(slots[None, None], size=Lv, mode="nearest-exact")[0, 0].long()
        back_tok = F.interpolate(up[None, None], size=Lv, mode="nearest-exact")[0, 0].long()
        a_pos, v_pos = O.interleaved_positions(L, Lv)
        assert torch.equal(v_pos, back_pos)
        assert torch.equal(back_tok, torch.arange(Lv))
        assert torch.equal(a_pos, 2 * torch.arange(L))


def _fp8_storage_state_dict(sd, mode="fp8_e4m3fn"):
    """What the reference loader leaves in the model with quantization != none (nodes.py:100-121): every parameter in
    the compute dtype (bf16), Linear / Conv weights additionally rounded through the FP8 storage format."""
    from conftest import load_pkg
    ck = load_pkg("checkpoint")
    out = {}
    for k, v in sd.items():
        v = v.to(torch.bfloat16)
        out[k] = (ck.round_through_fp8(v, mode) if ck.fp8_wraps(k, v.dim()) else v).float()
    return out


def test_fp8_weight_storage_matches_reference_wrapper(golden_dir):
    """Reference: bf16 model -> _wrap_fp8_inplace(e4m3fn) -> fp32 compute.  Oracle: the same forward on weights rounded
    through FP8 by the rule checkpoint.fp8_wraps — pins both the rule (which modules) and the rounding."""
    gold = torch.load(os.path.join(golden_dir, "dit_tiny_fp8e4m3_fp32.pt"))
    c = W.model_config("tiny")
    sd = _fp8_storage_state_dict(W.synth_dit_state_dict(c, seed=0))
    sh = gold["shape"]
    out = O.dit_forward(sd, c, *_dit_inputs(c, sh["B"], sh["L"], sh["Lv"], sh["S"]), policy="fp32")
    assert rel_l2(out, gold["out"]) <= 1e-5
    # and FP8 storage is not a no-op: the plain fp32 golden differs at the percent level
    plain = torch.load(os.path.join(golden_dir, "dit_tiny_fp32.pt"))["out"]
    assert rel_l2(gold["out"], plain) > 1e-3


def test_synchformer_golden_is_what_the_reference_module_computes(golden_dir):
    """tests/golden/synchformer_d2.pt was written on a B200 by tools/gpu_synchformer_golden.py.  Here the reference's own
    MotionFormer (models/synchformer/motionformer.py through the omegaconf / timm stand-ins of tools/ref_shims.py) is run
    again on the CPU in fp32 with the same seeded weights and frames: it must reproduce the golden's fp32 output (1e-4: CPU
    vs GPU fp32 summation order), which pins the weight / frame generators and the shims the GPU parity test relies on.
    Skipped when no reference tree is available (the GPU box: the goldens are committed)."""
    from tools import ref_shims as R
    from tools import synthetic as SY
    if not os.path.isdir(os.path.join(R.REF_ROOT, "hunyuanvideo_foley", "models", "synchformer")):
        pytest.skip("no reference tree")
    g = torch.load(os.path.join(golden_dir, "synchformer_d2.pt"))
    model = R.load_motionformer(g["depth"])().eval()
    model.load_state_dict(SY.synth_motionformer_state_dict(g["depth"], seed=g["weights_seed"]), strict=True)
    frames = SY.synth_sync_frames(g["n_frames"], seed=g["frames_seed"])
    S = (g["n_frames"] - 16) // 8 + 1
    x = torch.stack([frames[i * 8: i * 8 + 16] for i in range(S)])[None].permute(0, 1, 3, 2, 4, 5)
    with torch.inference_mode():
        out = model(x)[0]
    assert out.shape == g["out_fp32"].shape
    assert rel_l2(out, g["out_fp32"]) < 1e-4
    # and the autocast golden sits at the bf16-parameter / fp16-activation distance from it, not further
    assert 1e-3 < rel_l2(g["out_autocast"], g["out_fp32"]) < 2e-2


# ------------------------------------------------------------------------------------------------ condition encoders (§8f row 1)
def _perturbed(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1 and "norm" in name.lower() and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                p.add_(0.01 * torch.randn(p.shape, generator=g))
    return model.eval()


def test_siglip_oracle_matches_the_hf_module():
    """oracle/encoders_oracle.py siglip_pooler_output against HF SiglipVisionModel (the module the reference calls through
    get_image_features, feature_utils.py:71), seeded weights, CPU fp32."""
    from transformers import SiglipVisionConfig, SiglipVisionModel
    from oracle import encoders_oracle as EO
    torch.manual_seed(1)
    m = _perturbed(SiglipVisionModel(SiglipVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=2, num_attention_heads=12,
                                                        image_size=64, patch_size=16, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)), 2)
    px = torch.rand(3, 3, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1
    with torch.inference_mode():
        want = m(pixel_values=px).pooler_output
        got = EO.siglip_pooler_output({k: v for k, v in m.state_dict().items()}, px)
    assert got.shape == want.shape == (3, 768)
    assert rel_l2(got, want) < 1e-5


def test_clap_oracle_matches_the_hf_module():
    """clap_last_hidden_state against HF ClapTextModelWithProjection(...).last_hidden_state (feature_utils.py:134-137) with a
    right-padded prompt pair and its attention mask; position-id rule included."""
    from transformers import ClapTextConfig, ClapTextModelWithProjection
    from oracle import encoders_oracle as EO
    torch.manual_seed(4)
    m = _perturbed(ClapTextModelWithProjection(ClapTextConfig(num_hidden_layers=2)), 5)
    ids = torch.randint(3, 50265, (2, 11), generator=torch.Generator().manual_seed(6))
    ids[:, 0] = 0
    ids[0, 5] = 2
    ids[0, 6:] = 1
    ids[1, 10] = 2
    mask = (ids != 1).long()
    with torch.inference_mode():
        want = m(input_ids=ids, attention_mask=mask).last_hidden_state
        got = EO.clap_last_hidden_state({k: v for k, v in m.state_dict().items()}, ids, mask)
    assert EO.clap_position_ids(ids)[0].tolist() == [2, 3, 4, 5, 6, 7, 1, 1, 1, 1, 1]
    assert rel_l2(got, want) < 1e-5


def test_synchformer_oracle_matches_the_reference_golden(golden_dir):
    """synchformer_visual against the fp32 output of the reference's own MotionFormer (golden written on a B200 by
    tools/gpu_synchformer_golden.py; reproduced on the CPU by the test above), same seeded weights and frames."""
    from oracle import encoders_oracle as EO
    from tools import synthetic as SY
    g = torch.load(os.path.join(golden_dir, "synchformer_d2.pt"))
    sd = SY.synth_motionformer_state_dict(g["depth"], seed=g["weights_seed"])
    with torch.inference_mode():
        got = EO.synchformer_visual(sd, SY.synth_sync_frames(g["n_frames"], seed=g["frames_seed"]))
    want = g["out_fp32"].reshape(got.shape)
    assert rel_l2(got, want) < 1e-4
    # the Synchformer checkpoint's key names are accepted too
    got2 = EO.synchformer_visual({"vfeat_extractor." + k: v for k, v in sd.items()}, SY.synth_sync_frames(g["n_frames"], seed=g["frames_seed"]))
    assert torch.equal(got, got2)
