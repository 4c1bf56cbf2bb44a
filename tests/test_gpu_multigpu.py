"""The Sampler node on TWO GPUs of one process (reference nodes.py:228 `batch_size` is the only parallel
axis): variations sharded over the visible GPUs by parallel.denoise_sharded — weights replicated, one broadcast of the
condition embeddings, one gather of the waveforms, one host thread per GPU.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import pytest
import torch

from conftest import load_pkg, rel_l2
from oracle import weights as W

pytestmark = pytest.mark.gpu


def _objects():
    E, nodes, cfgmod = load_pkg("engine"), load_pkg("nodes"), load_pkg("config")
    c = W.model_config("small")
    sd = W.synth_dit_state_dict(c, seed=0)
    cfg = cfgmod.load_model_config("xxl")
    for k in ("hidden_size", "num_heads", "depth_triple_blocks", "depth_single_blocks"):
        cfg.model_config.model_kwargs[k] = c[k]

    def make(dev):
        e = E.FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev)
        e.load_state_dict(sd)
        e.finalize()
        return e
    model = nodes.FoleyModel(make(torch.device("cuda", 0)), sd["empty_clip_feat"], sd["empty_sync_feat"], cfg, torch.bfloat16)
    model._make_engine = make
    dac = nodes.FoleyDAC.from_state_dict(W.synth_dac_state_dict(W.DAC_TINY, seed=3), device=torch.device("cuda", 0))
    g = torch.Generator().manual_seed(5)
    text = {"text_feat": torch.randn(1, 9, c["condition_dim"], generator=g).bfloat16(),
            "uncond_text_feat": torch.randn(1, 5, c["condition_dim"], generator=g).bfloat16()}
    deps = cfgmod.AttributeDict({"dac_model": dac, "report_progress": False,
                                 "extract_features": lambda f8, f25, prompt, neg: ({}, text, None)})
    return nodes, cfgmod, model, dac, deps, text, cfg


def test_sampler_shards_variations_over_two_gpus(monkeypatch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    torch.cuda.set_device(0)
    nodes, cfgmod, model, dac, deps, text, cfg = _objects()
    sampling = load_pkg("sampling")
    duration, steps, B, seed = 2.0, 10, 3, 7          # 3 variations over 2 GPUs: a ragged split (2 + 1)
    node = nodes.HunyuanFoleySampler()

    def run():
        first, batch = node.generate_audio(model, deps, 16, duration, "p", "n", 4.5, steps, "euler", B, seed, True)
        return batch["waveform"]

    monkeypatch.setenv("FOLEY_B200_GPUS", "auto")
    wav2 = run()
    assert wav2.shape[0] == B and torch.isfinite(wav2).all()
    assert 1 in model._replicas and 1 in dac._replicas          # engines exist on the second GPU
    assert model._replicas[1].engine.device == torch.device("cuda", 1)
    assert model._replicas[1].engine.launch_count() > 0          # and did the work

    # every shard alone on GPU 0, same host noise rows: bit for bit what the two GPUs produced
    clip_len, sync_len = nodes.t2a_feature_lengths(duration)
    visual = {"siglip2_feat": model.get_empty_clip_sequence(bs=1, len=clip_len).to("cpu", torch.bfloat16),
              "syncformer_feat": model.get_empty_sync_sequence(bs=1, len=sync_len).to("cpu", torch.bfloat16)}
    md = cfgmod.AttributeDict(dict(deps))
    md["foley_model"], md["device"] = model, torch.device("cuda", 0)
    par = load_pkg("parallel")
    for r in range(2):
        lo, hi = par.shard_range(B, 2, r)
        gen = torch.Generator(device="cpu").manual_seed(seed)
        w, _ = sampling.denoise_process_with_generator(visual, text, duration, md, cfg, 4.5, steps, B, "euler", generator=gen,
                                                       batch_slice=(lo, hi))
        assert torch.equal(w.float().cpu(), wav2[lo:hi]), f"shard {r} differs from its single-GPU run"

    # and the plain one-GPU batch: other tile plans (batch 3 instead of 2 + 1), so close, not identical
    monkeypatch.setenv("FOLEY_B200_GPUS", "1")
    wav1 = run()
    err = rel_l2(wav2, wav1)
    print(f"2-GPU sharded vs 1-GPU batch of {B}: waveform rel-L2 {err:.3e}")
    assert err <= 3e-2


def test_default_stream_denoise_then_decode_is_ordered():
    """Regression (round 2): called on the legacy default stream without a progress callback, the DiT engine enqueues the
    whole loop on ITS blocking stream and returns; the DAC engine then decodes on ANOTHER blocking stream.  The C ABI now
    makes the legacy stream wait for the engine's work, so the decode is ordered behind the denoise: repeated calls give
    the same waveform (before the fix only the first call did, thanks to the cudaMalloc of the decoder's buffers)."""
    torch.cuda.set_device(0)
    nodes, cfgmod, model, dac, deps, text, cfg = _objects()
    sampling = load_pkg("sampling")
    duration = 2.0
    clip_len, sync_len = nodes.t2a_feature_lengths(duration)
    visual = {"siglip2_feat": model.get_empty_clip_sequence(bs=1, len=clip_len).to("cpu", torch.bfloat16),
              "syncformer_feat": model.get_empty_sync_sequence(bs=1, len=sync_len).to("cpu", torch.bfloat16)}
    md = cfgmod.AttributeDict(dict(deps))
    md["foley_model"], md["device"] = model, torch.device("cuda", 0)
    outs = []
    for _ in range(3):
        gen = torch.Generator(device="cpu").manual_seed(7)
        w, _ = sampling.denoise_process_with_generator(visual, text, duration, md, cfg, 4.5, 10, 2, "euler", generator=gen)
        outs.append(w.float().cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
