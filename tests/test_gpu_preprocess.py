"""GPU parity of foley_preprocess_frames (csrc/preprocess.cuh) with the CPU oracle and the torchvision-generated
fixtures: quantise + frame picks + uint8 antialiased bicubic resize + crop + normalise.  Integer / byte work up to the
final exact fp32 scaling, so every comparison is bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import preprocess_oracle as P

pytestmark = pytest.mark.gpu


def _oracle_frames(image_np, idx, resize_hw, crop=None):
    out = []
    for i in idx:
        u8 = np.transpose(P.to_uint8(image_np[i]), (2, 0, 1))
        r = P.resize_bicubic_aa_u8(u8, resize_hw[0], resize_hw[1])
        if crop is not None:
            t, l, h, w = crop
            r = r[:, t:t + h, l:l + w]
        out.append(P.normalize_u8(r))
    return np.stack(out)


def test_reference_fixtures_bit_exact(golden_dir):
    pp = load_pkg("preprocess")
    gold = torch.load(os.path.join(golden_dir, "preprocess.pt"))
    for case in gold["cases"]:
        a = case["args"]
        image = torch.rand(a["N"], a["H"], a["W"], 3, generator=torch.Generator().manual_seed(a["seed"]))
        pre8, pre25, audio_len = pp.preprocess_video(image, a["duration"], a["frame_rate"], torch.device("cuda", 0))
        assert tuple(pre8.shape) == tuple(case["siglip2_shape"]) and tuple(pre25.shape) == tuple(case["sync_shape"])
        assert hashlib.sha256(pre8.cpu().numpy().tobytes()).hexdigest() == case["siglip2_sha256"]
        assert hashlib.sha256(pre25.cpu().numpy().tobytes()).hexdigest() == case["sync_sha256"]
        assert torch.equal(pre25[0].cpu(), case["sync_frame0"])
        assert audio_len == int(a["duration"] * 25) / 25.0


@pytest.mark.parametrize("H,W,rh,rw,crop", [
    (90, 160, 45, 64, None),            # downscale both axes
    (37, 53, 74, 106, None),            # upscale both axes
    (64, 48, 64, 24, None),             # rows unchanged (ATen skips that pass)
    (33, 200, 17, 31, (3, 5, 9, 20)),   # ragged sizes + crop window
    (224, 224, 224, 224, None),         # identity resize
    (131, 517, 224, 884, (0, 330, 224, 224)),
])
def test_resize_crop_normalise_matches_oracle(H, W, rh, rw, crop):
    pp = load_pkg("preprocess")
    g = torch.Generator().manual_seed(H * 1000 + W)
    image = torch.rand(3, H, W, 3, generator=g)
    image[0, :2] = torch.tensor([-0.01, 1.004, 0.5])      # out-of-range pixels wrap like (x*255).byte()
    idx = [2, 0, 0, 1]
    got = pp.preprocess_frames(image.cuda(), idx, (rh, rw), crop).cpu().numpy()
    want = _oracle_frames(image.numpy(), idx, (rh, rw), crop)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_full_hd_frame_matches_oracle():
    """BASELINE shapes: 1080p input frames -> SigLIP2 512x512 and Synchformer 224 crop (21-tap filters)."""
    pp = load_pkg("preprocess")
    image = torch.rand(2, 1080, 1920, 3, generator=torch.Generator().manual_seed(5))
    dev = image.cuda()
    got8 = pp.preprocess_frames(dev, [1], (512, 512)).cpu().numpy()
    assert np.array_equal(got8, _oracle_frames(image.numpy(), [1], (512, 512)))
    nh, nw = pp.resized_size_short_side(1080, 1920, 224)
    top, left = pp.center_crop_offsets(nh, nw, 224, 224)
    got25 = pp.preprocess_frames(dev, [0], (nh, nw), (top, left, 224, 224)).cpu().numpy()
    assert np.array_equal(got25, _oracle_frames(image.numpy(), [0], (nh, nw), (top, left, 224, 224)))


def test_hold_last_frame_and_errors():
    pp, E = load_pkg("preprocess"), load_pkg("engine")
    image = torch.rand(3, 40, 60, 3, generator=torch.Generator().manual_seed(1))
    # 2 s at 8 fps requested from 3 frames: the last frame is held (nodes.py:298-303)
    pre8, pre25, _ = pp.preprocess_video(image, 2.0, 8.0, torch.device("cuda", 0))
    want_idx = P.source_frame(P.frame_indices(16, 2.0, 8), 3)
    want = _oracle_frames(image.numpy(), want_idx, (512, 512))
    assert np.array_equal(pre8.cpu().numpy(), want)
    assert pre25.shape == (50, 3, 224, 224)
    with pytest.raises(E.FoleyError):
        pp.preprocess_frames(image.cuda(), [3], (8, 8))                      # frame index out of range
    with pytest.raises(E.FoleyError):
        pp.preprocess_frames(image.cuda(), [0], (8, 8), (4, 4, 8, 8))        # crop outside the resized frame


def test_sampler_uses_gpu_preprocessing_when_the_extractors_ask_for_it():
    nodes, cfgmod, E = load_pkg("nodes"), load_pkg("config"), load_pkg("engine")
    from oracle import weights as W
    c = W.model_config("tiny")
    sd = W.synth_dit_state_dict(c, seed=0)
    cfg = cfgmod.load_model_config("xxl")
    for k in ("hidden_size", "num_heads", "depth_triple_blocks", "depth_single_blocks"):
        cfg.model_config.model_kwargs[k] = c[k]
    eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs))
    eng.load_state_dict(sd)
    eng.finalize()
    model = nodes.FoleyModel(eng, sd["empty_clip_feat"], sd["empty_sync_feat"], cfg, dtype=torch.float32)
    dac = nodes.FoleyDAC.from_state_dict(W.synth_dac_state_dict(W.DAC_TINY, seed=3))
    L, Lv, S = W.clip_lengths(1.0)
    feats = W.synth_conditions(c, L, Lv, S)
    seen = {}

    def extract(pre8, pre25, prompt, negative_prompt):
        seen["pre8"], seen["pre25"] = pre8, pre25
        return ({"siglip2_feat": feats["siglip2_feat"], "syncformer_feat": feats["syncformer_feat"]},
                {"text_feat": feats["text_feat"], "uncond_text_feat": feats["uncond_text_feat"]}, pre25.shape[0] / 25.0)

    deps = cfgmod.AttributeDict({"dac_model": dac, "extract_features": extract, "preprocessed_inputs": True})
    image = torch.rand(8, 48, 64, 3, generator=torch.Generator().manual_seed(4))
    first, batch = nodes.HunyuanFoleySampler().generate_audio(model, deps, 8.0, 1.0, "p", "n", 4.5, 10, "euler", 1, 0, True,
                                                              image=image)
    assert seen["pre8"].shape == (8, 3, 512, 512) and seen["pre25"].shape == (25, 3, 224, 224)
    assert seen["pre8"].is_cuda and float(seen["pre8"].abs().max()) <= 1.0
    assert batch["waveform"].shape == (1, 1, 48000)
