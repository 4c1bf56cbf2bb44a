"""GPU tests of the Model Loader pipeline: .safetensors -> engine layout directly
(foley_engine_load_safetensors), FP8 checkpoints de-quantised at load, and the reference's `quantization` setting
(FP8 weight-only storage, utils.py:316-485) reproduced as a rounding of the wrapped weights.

Tolerances: the two load paths must agree bit for bit; engine vs oracle(cuda_bf16) on the FP8-rounded weights uses the
same 4e-3 relative L2 as the plain forward test (tests/test_gpu_dit.py)."""
import pytest
import torch

from conftest import load_pkg, rel_l2
from oracle import foley_oracle as O
from oracle import weights as W
from test_gpu_dit import _inputs, make_engine
from test_oracle_golden import _fp8_storage_state_dict

pytestmark = pytest.mark.gpu


def _forward(eng, c, B=2, L=50, Lv=8, S=16):
    x, t, cond, clip, sync = _inputs(c, B, L, Lv, S)
    eng.set_conditions(clip.cuda(), sync.cuda(), cond.cuda(), L=L, batch=1)
    out = eng.dit_forward(x.cuda(), t).cpu()
    assert eng.debug_flags()[0] == 0
    return out, (x, t, cond, clip, sync)


@pytest.mark.parametrize("file_dtype", [torch.bfloat16, torch.float32, torch.float16])
def test_safetensors_direct_load_is_bit_identical_to_state_dict_load(tmp_path, file_dtype):
    E, ck = load_pkg("engine"), load_pkg("checkpoint")
    c = W.model_config("tiny")
    sd = {k: v.to(file_dtype) for k, v in W.synth_dit_state_dict(c, seed=0).items()}
    path = str(tmp_path / "tiny.safetensors")
    ck.write_safetensors(path, sd, metadata={"format": "pt"})
    ref_eng, _, _ = make_engine("tiny", sd=sd)
    want, _ = _forward(ref_eng, c)
    eng = E.FoleyEngine(c)
    n = eng.load_safetensors(path)
    assert n == sum(1 for k in sd if not k.startswith("final_layer.adaLN_modulation"))
    eng.finalize()
    got, _ = _forward(eng, c)
    assert torch.equal(got, want)


@pytest.mark.parametrize("mode", ["fp8_e4m3fn", "fp8_e5m2"])
def test_fp8_weight_storage_matches_oracle_and_prequantised_checkpoint(tmp_path, mode):
    E, ck = load_pkg("engine"), load_pkg("checkpoint")
    c = W.model_config("tiny")
    sd = W.synth_dit_state_dict(c, seed=0)
    # (a) bf16 checkpoint + quantization=<mode>: weights rounded on the device
    eng = E.FoleyEngine(c)
    eng.set_fp8_weight_storage(mode)
    eng.load_state_dict(sd)
    eng.finalize()
    got, (x, t, cond, clip, sync) = _forward(eng, c)
    sd_q = _fp8_storage_state_dict(sd, mode)
    want16 = O.dit_forward(sd_q, c, x, t, cond, clip, sync, policy="cuda_bf16")
    plain16 = O.dit_forward(sd, c, x, t, cond, clip, sync, policy="cuda_bf16")
    r, shift = rel_l2(got, want16), rel_l2(want16, plain16)
    print(f"\n[{mode}] engine vs oracle(cuda_bf16, fp8-rounded weights) {r:.3e}; effect of the fp8 storage itself {shift:.3e}")
    assert r <= 4e-3 and shift > 2 * r
    # (b) a checkpoint that already stores the wrapped weights in FP8 (what `_detect_ckpt_fp8` looks for), loaded
    # straight from the file: same bits as (a)
    qd = torch.float8_e5m2 if mode == "fp8_e5m2" else torch.float8_e4m3fn
    sd_file = {k: (v.bfloat16().to(qd) if ck.fp8_wraps(k, v.dim()) else v.bfloat16()) for k, v in sd.items()}
    path = str(tmp_path / "tiny_fp8.safetensors")
    ck.write_safetensors(path, sd_file)
    for storage in (mode, "none"):     # verbatim FP8 bytes need no second rounding: both settings give the same model
        eng2 = E.FoleyEngine(c)
        eng2.set_fp8_weight_storage(storage)
        eng2.load_safetensors(path)
        eng2.finalize()
        got2, _ = _forward(eng2, c)
        assert torch.equal(got2, got), storage


def test_model_loader_from_safetensors(tmp_path, monkeypatch):
    """FoleyModel.from_safetensors: header-only precision / fp8 detection, host copies of the empty-feature rows, and a
    working engine; `quantization=auto` on B200 means e4m3fn storage, as in the reference (nodes.py:108-119).
    precision=fp32 is REFUSED (the engine is the reference's bf16 path) unless the user opts into bf16 arithmetic."""
    nodes, cfgmod, ck = load_pkg("nodes"), load_pkg("config"), load_pkg("checkpoint")
    c = W.model_config("tiny")
    sd = {k: v.bfloat16() for k, v in W.synth_dit_state_dict(c, seed=0).items()}
    path = str(tmp_path / "m.safetensors")
    ck.write_safetensors(path, sd)
    cfg = cfgmod.load_model_config("xxl")
    for k in ("hidden_size", "num_heads", "depth_triple_blocks", "depth_single_blocks"):
        cfg.model_config.model_kwargs[k] = c[k]
    m_auto = nodes.FoleyModel.from_safetensors(path, precision="auto", quantization="auto", cfg=cfg)
    E = load_pkg("engine")
    with pytest.raises(E.FoleyError, match="not implemented by the B200 engine"):
        nodes.FoleyModel.from_safetensors(path, precision="fp32", quantization="none", cfg=cfg)
    monkeypatch.setenv("FOLEY_B200_PRECISION_FALLBACK", "bf16")
    m_none = nodes.FoleyModel.from_safetensors(path, precision="fp32", quantization="none", cfg=cfg)
    assert m_auto.dtype == torch.bfloat16 and m_none.dtype == torch.float32
    assert torch.equal(m_auto.empty_clip_feat, sd["empty_clip_feat"])
    assert m_auto.get_empty_sync_sequence(bs=1, len=16).shape == (1, 16, 768)
    out_auto, _ = _forward(m_auto.engine, c)
    out_none, _ = _forward(m_none.engine, c)
    ref, _, _ = make_engine("tiny")
    want_none, _ = _forward(ref, c)
    assert torch.equal(out_none, want_none)
    assert not torch.equal(out_auto, out_none)      # auto = fp8 storage emulation


def test_load_safetensors_errors(tmp_path):
    E, ck = load_pkg("engine"), load_pkg("checkpoint")
    c = W.model_config("tiny")
    eng = E.FoleyEngine(c)
    with pytest.raises(E.FoleyError, match="cannot open"):
        eng.load_safetensors(str(tmp_path / "nope.safetensors"))
    sd = W.synth_dit_state_dict(c, seed=0)
    del sd["single_blocks.0.linear1.weight"]
    path = str(tmp_path / "missing.safetensors")
    ck.write_safetensors(path, sd)
    eng.load_safetensors(path)
    with pytest.raises(E.FoleyError, match="missing tensor"):
        eng.finalize()
