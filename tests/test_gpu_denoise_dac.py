"""GPU parity of the Euler loop (foley_denoise), the DAC-VAE decoder (foley_dac_decode) and the node-level
call against the CPU oracle and the reference-generated golden fixtures.

Tolerances (relative L2):
  * final latents after N Euler steps: the engine must sit as close to the reference's fp32 run as the
    bf16-rounding-point oracle does (gap = oracle(cuda_bf16) vs reference fp32, ~2e-3 without CFG; CFG 4.5
    amplifies every bf16 rounding of (cond - uncond) 4.5x, ~5e-3 at 4 steps): both distances <= 1.5*gap + 1e-3.
  * DAC waveform vs the fp32 oracle: 3e-3.  The engine multiplies on TF32 tensor cores, as the reference does
    on GPU through cuDNN's default allow_tf32 (profiles/r01_torch_probe.json: 2.9e-4 per conv vs fp64).
"""
import os

import pytest
import torch

from conftest import load_pkg, rel_l2
from oracle import foley_oracle as O
from oracle import weights as W
from test_gpu_dit import make_engine

pytestmark = pytest.mark.gpu


def _feats(c, sd, duration, v2a):
    L, Lv, S = W.clip_lengths(duration)
    feats = W.synth_conditions(c, L, Lv, S)
    if not v2a:
        feats["siglip2_feat"] = sd["empty_clip_feat"][None].expand(1, Lv, -1).contiguous()
        feats["syncformer_feat"] = sd["empty_sync_feat"][None].expand(1, S, -1).contiguous()
    return feats, L, Lv, S


@pytest.mark.parametrize("tag", ["tiny_t2a_nocfg", "tiny_v2a_cfg"])
@pytest.mark.parametrize("graph", [0, 1])
def test_denoise_matches_oracle_and_reference(golden_dir, tag, graph):
    gold = torch.load(os.path.join(golden_dir, f"denoise_{tag}.pt"))
    a = gold["args"]
    eng, c, sd = make_engine("tiny")
    eng.set_option("cuda_graph", graph)
    feats, L, Lv, S = _feats(c, sd, a["duration"], a["v2a"])
    gen = torch.Generator(device="cpu").manual_seed(123)
    noise = torch.randn((a["batch"], 128, L), generator=gen, dtype=torch.float32)
    clip, sync, text = O.build_cfg_batch(sd, feats, 1, a["guidance"])     # one row per condition
    eng.set_conditions(clip.cuda(), sync.cuda(), text.cuda(), L=L, batch=a["batch"])
    sig = O.sigma_schedule(a["steps"])
    steps_seen = []
    lat = eng.denoise(noise.cuda(), sig, a["guidance"], progress=steps_seen.append).cpu()
    assert steps_seen == list(range(1, a["steps"] + 1))
    assert eng.debug_flags()[0] == 0
    want16 = O.denoise(sd, c, feats, noise, a["steps"], a["guidance"], policy="cuda_bf16")
    r16, r32 = rel_l2(lat, want16), rel_l2(lat, gold["latents"])
    print(f"\n[{tag} graph={graph}] latents: engine vs oracle(cuda_bf16) {r16:.3e} | vs reference fp32 {r32:.3e} | "
          f"oracle bf16 vs reference fp32 {rel_l2(want16, gold['latents']):.3e}")
    gap = rel_l2(want16, gold["latents"])
    assert r16 <= 1.5 * gap + 1e-3
    assert r32 <= 1.5 * gap + 1e-3
    # a second run on the same engine (graph replay, fresh step counter) is bit-identical
    lat2 = eng.denoise(noise.cuda(), sig, a["guidance"]).cpu()
    assert torch.equal(lat, lat2)


def test_dac_decode_full_size_matches_reference(golden_dir):
    E = load_pkg("engine")
    gold = torch.load(os.path.join(golden_dir, "dac_full_L25.pt"))
    dsd = W.synth_dac_state_dict(W.DAC_CONFIG, seed=3)
    eng = E.FoleyEngine(W.model_config("tiny"))
    eng.load_state_dict(dsd, prefix="dac.")
    z = torch.randn(1, 128, 25, generator=torch.Generator().manual_seed(5))
    wav = eng.dac_decode(z.cuda()).cpu()
    assert wav.shape == (1, 1, 25 * 960)
    r = rel_l2(wav, gold["wav"])
    print(f"\nDAC full-size L=25: engine (tf32) vs reference fp32 {r:.3e}")
    assert r <= 3e-3
    assert eng.debug_flags()[0] == 0


@pytest.mark.parametrize("B,L", [(1, 50), (3, 37), (2, 250)])
def test_dac_decode_batches_and_ragged_lengths(B, L):
    """Batch > 1 (per-sample zero halo via TMA OOB), lengths that are not tile multiples."""
    E = load_pkg("engine")
    dsd = W.synth_dac_state_dict(W.DAC_TINY, seed=3)
    eng = E.FoleyEngine(W.model_config("tiny"))
    eng.load_state_dict(dsd, prefix="dac.")
    z = torch.randn(B, 128, L, generator=torch.Generator().manual_seed(7))
    wav = eng.dac_decode(z.cuda()).cpu()
    want = O.dac_decode(dsd, z)
    assert wav.shape == want.shape == (B, 1, L * 960)
    r = rel_l2(wav, want)
    print(f"\nDAC tiny B={B} L={L}: {r:.3e}")
    assert r <= 3e-3
    # decoding each sample alone gives the same waveform: no bleed between batch items
    solo = eng.dac_decode(z[:1].cuda()).cpu()
    assert rel_l2(solo[0], wav[0]) <= 1e-6


def test_sampler_node_end_to_end(golden_dir):
    """The Sampler node (generate_audio) with stubbed condition encoders reproduces the reference's
    denoise_process_with_generator output (golden: T2A, CFG off, 10 steps, reduced-width DAC)."""
    nodes = load_pkg("nodes")
    cfgmod = load_pkg("config")
    E = load_pkg("engine")
    gold = torch.load(os.path.join(golden_dir, "denoise_tiny_t2a_nocfg.pt"))
    a = gold["args"]
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
    feats, L, Lv, S = _feats(c, sd, a["duration"], False)

    def extract(frames8, frames25, prompt, negative_prompt):
        return {}, {"text_feat": feats["text_feat"], "uncond_text_feat": feats["uncond_text_feat"]}, None

    deps = cfgmod.AttributeDict({"dac_model": dac, "extract_features": extract})
    first, batch = nodes.HunyuanFoleySampler().generate_audio(
        model, deps, frame_rate=8.0, duration=a["duration"], prompt="p", negative_prompt="n", cfg_scale=a["guidance"],
        steps=a["steps"], sampler="euler", batch_size=a["batch"], seed=123, force_offload=True)
    assert first["sample_rate"] == 48000 and batch["waveform"].shape == (a["batch"], 1, 48000)
    assert batch["waveform"].dtype == torch.float32 and batch["waveform"].device.type == "cpu"
    r = rel_l2(batch["waveform"], gold["audio"].float())
    print(f"\nSampler node waveform vs reference: {r:.3e}")
    assert r <= 1e-2
    with pytest.raises(ValueError):
        nodes.HunyuanFoleySampler().generate_audio(model, deps, 8.0, 1.0, "p", "n", 1.0, 10, "dpm++", 1, 0, True)


@pytest.mark.parametrize("tag", ["tiny_heun2", "tiny_midpoint2", "tiny_kutta4"])
def test_multistage_solvers_match_reference(golden_dir, tag):
    """heun-2 / midpoint-2 / kutta-4 (reference scheduler :299-373 incl. its inner-stage quirk) through the public
    host function, against the reference's own run."""
    nodes, cfgmod, sampling = load_pkg("nodes"), load_pkg("config"), load_pkg("sampling")
    gold = torch.load(os.path.join(golden_dir, f"denoise_{tag}.pt"))
    a = gold["args"]
    eng, c, sd = make_engine("tiny")
    cfg = cfgmod.load_model_config("xxl")
    model = nodes.FoleyModel(eng, sd["empty_clip_feat"], sd["empty_sync_feat"], cfg, dtype=torch.float32)
    feats, L, Lv, S = _feats(c, sd, a["duration"], a["v2a"])
    visual = {"siglip2_feat": feats["siglip2_feat"], "syncformer_feat": feats["syncformer_feat"]}
    text = {"text_feat": feats["text_feat"], "uncond_text_feat": feats["uncond_text_feat"]}
    md = cfgmod.AttributeDict({"foley_model": model, "device": torch.device("cuda", 0)})
    gen = torch.Generator(device="cpu").manual_seed(123)
    lat, _ = sampling.denoise_process_with_generator(visual, text, a["duration"], md, cfg, a["guidance"], a["steps"],
                                                     a["batch"], a["sampler"], generator=gen, decode=False)
    noise = torch.randn((a["batch"], 128, L), generator=torch.Generator(device="cpu").manual_seed(123))
    want16 = O.denoise(sd, c, feats, noise, a["steps"], a["guidance"], policy="cuda_bf16", solver=a["sampler"])
    gap = rel_l2(want16, gold["latents"])
    r16, r32 = rel_l2(lat.cpu(), want16), rel_l2(lat.cpu(), gold["latents"])
    print(f"\n[{tag}] engine vs oracle(cuda_bf16) {r16:.3e} | vs reference fp32 {r32:.3e} | oracle gap {gap:.3e}")
    assert r16 <= 1.5 * gap + 1e-3 and r32 <= 1.5 * gap + 1e-3
    # The loop above ran inside the engine (foley_denoise_solver: one graph per model call, stage table on the device).
    # The reference scheduler's stage machine on the host around foley_dit_forward must give the same bits.
    md_host = cfgmod.AttributeDict({"foley_model": model, "device": torch.device("cuda", 0), "host_solver": True})
    gen = torch.Generator(device="cpu").manual_seed(123)
    lat_host, _ = sampling.denoise_process_with_generator(visual, text, a["duration"], md_host, cfg, a["guidance"],
                                                          a["steps"], a["batch"], a["sampler"], generator=gen,
                                                          decode=False)
    assert torch.equal(lat.cpu(), lat_host.cpu())
    assert eng.debug_flags()[0] == 0


@pytest.mark.parametrize("solver", ["heun-2", "midpoint-2", "kutta-4"])
@pytest.mark.parametrize("n_calls,guidance,graph", [(7, 4.5, 1), (5, 1.0, 0)])
def test_engine_solver_ragged_call_counts(solver, n_calls, guidance, graph):
    """Call counts that stop in the middle of a solver's stage group (the reference loop allows any `steps`), with and
    without CFG / CUDA graph: foley_denoise_solver == host stage machine, bit for bit; and the torch op surface."""
    sampling, ops = load_pkg("sampling"), load_pkg("torch_ops")
    eng, c, sd = make_engine("tiny")
    eng.set_option("cuda_graph", graph)
    feats, L, Lv, S = _feats(c, sd, 1.0, True)
    clip, sync, text = O.build_cfg_batch(sd, feats, 1, guidance)
    ops.set_conditions(eng, clip.cuda(), sync.cuda(), text.cuda(), L, 2)
    noise = torch.randn((2, 128, L), generator=torch.Generator().manual_seed(11))
    sig = sampling.sigma_schedule(n_calls)
    seen = []
    got = ops.denoise(eng, noise.cuda(), sig, guidance, solver=solver, progress=seen.append)
    assert seen == list(range(1, n_calls + 1))
    want = sampling._denoise_multistage(eng, noise.cuda(), sig, guidance, solver, clip.shape[0], None)
    assert torch.equal(got.cpu(), want.cpu())
    # euler through the generic entry point == foley_denoise
    e1 = eng.denoise_solver(noise.cuda(), sig, guidance, "euler")
    e2 = eng.denoise(noise.cuda(), sig, guidance)
    assert torch.equal(e1, e2)
    with pytest.raises(ValueError):
        eng.denoise_solver(noise.cuda(), sig, guidance, "dpm++")
    assert eng.debug_flags()[0] == 0


def test_torch_ops_surface():
    """torch.ops.foley_b200.* call the same C-ABI entry points as the ctypes binding; CPU tensors are refused."""
    ops = load_pkg("torch_ops")
    E = load_pkg("engine")
    eng, c, sd = make_engine("tiny")
    h = ops.register_engine(eng)
    assert ops.register_engine(eng) == h
    feats, L, Lv, S = _feats(c, sd, 1.0, True)
    clip, sync, text = O.build_cfg_batch(sd, feats, 1, 4.5)
    torch.ops.foley_b200.set_conditions(h, clip.cuda(), sync.cuda(), text.cuda(), L, 1)
    x = torch.randn(2, 128, L, generator=torch.Generator().manual_seed(3)).cuda()
    t = torch.tensor([500.0])
    y_op = torch.ops.foley_b200.dit_forward(h, x, t)
    assert torch.equal(y_op, eng.dit_forward(x, 500.0))
    sig = O.sigma_schedule(3)
    lat_op = torch.ops.foley_b200.denoise(h, x[:1], sig, 4.5)
    assert torch.equal(lat_op, eng.denoise(x[:1], sig, 4.5))
    with pytest.raises(E.FoleyError):
        torch.ops.foley_b200.dit_forward(h, x.cpu(), t)
    with pytest.raises(E.FoleyError):
        torch.ops.foley_b200.dit_forward(h + 1000, x, t)
