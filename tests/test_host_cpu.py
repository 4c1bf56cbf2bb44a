"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host-side mirror of the
reference interface behaves like the reference (node surface, text bucket, noise draw, frame resampling), the
product path fails loudly without a GPU, and the multi-GPU sharding logic works at world_size 2 over gloo."""
import ctypes
import json
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import PKG_DIR, ROOT, load_pkg

LIB = os.path.join(PKG_DIR, "libfoley_b200.so")


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "foley_b200.h")).read()
    return sorted(set(re.findall(r"^\s*(?:foley_status|void|int64_t|const char\*)\s+(foley_[a-z_0-9]+)\(", hdr, re.M)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(LIB)
    syms = _declared_symbols()
    assert len(syms) >= 14 and "foley_denoise" in syms and "foley_dac_decode" in syms
    for s in syms:
        assert hasattr(lib, s), s
    lib.foley_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.foley_version()


def test_library_is_tcgen05_tma_code():
    """SASS evidence that the hot GEMM is Blackwell-native (B200_PROFILING.md): UTC*MMA, LDTM, UTMALDG present."""
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out and "LDTM" in out and "UTMALDG" in out
    assert "HGMMA" not in out


def test_product_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    E = load_pkg("engine")
    with pytest.raises(E.FoleyError, match="no CPU path"):
        E.FoleyEngine({"hidden_size": 256, "num_heads": 2, "depth_triple_blocks": 1, "depth_single_blocks": 1})


def test_no_product_module_imports_the_oracle():
    for root, _, files in os.walk(PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "foley_oracle" not in src or f in ("rowwise.cuh", "engine.cu"), f  # comments citing it only


def test_node_surface_matches_reference():
    """Ids, display names, sockets and widget lists of reference nodes.py:57-683."""
    pkg = load_pkg()
    ids = ["HunyuanModelLoader", "HunyuanDependenciesLoader", "HunyuanFoleySampler", "HunyuanFoleyTorchCompile",
           "HunyuanBlockSwap", "SelectAudioFromBatch"]
    assert list(pkg.NODE_CLASS_MAPPINGS) == ids
    assert pkg.NODE_DISPLAY_NAME_MAPPINGS["HunyuanFoleySampler"] == "Hunyuan-Foley Sampler"
    assert pkg.NODE_DISPLAY_NAME_MAPPINGS["HunyuanBlockSwap"] == "Hunyuan-Foley BlockSwap Settings"
    S = pkg.NODE_CLASS_MAPPINGS["HunyuanFoleySampler"]
    it = S.INPUT_TYPES()
    assert list(it["required"]) == ["hunyuan_model", "hunyuan_deps", "frame_rate", "duration", "prompt", "negative_prompt",
                                    "cfg_scale", "steps", "sampler", "batch_size", "seed", "force_offload"]
    assert list(it["optional"]) == ["image", "torch_compile_cfg", "block_swap_args"]
    assert it["required"]["hunyuan_model"] == ("HUNYUAN_MODEL",) and it["required"]["hunyuan_deps"] == ("HUNYUAN_DEPS",)
    assert it["optional"]["torch_compile_cfg"][0] == "TORCH_COMPILE_CFG" and it["optional"]["block_swap_args"][0] == "BLOCKSWAPARGS"
    assert it["required"]["cfg_scale"][1]["default"] == 4.5 and it["required"]["steps"][1]["default"] == 50
    assert it["required"]["sampler"][0] == ["euler", "heun-2", "midpoint-2", "kutta-4"]
    assert S.RETURN_TYPES == ("AUDIO", "AUDIO") and S.RETURN_NAMES == ("audio_first", "audio_batch")
    assert S.FUNCTION == "generate_audio" and S.CATEGORY == "audio/HunyuanFoley"
    L = pkg.NODE_CLASS_MAPPINGS["HunyuanModelLoader"]
    assert L.RETURN_TYPES == ("HUNYUAN_MODEL",) and L.FUNCTION == "build_model"
    assert list(L.INPUT_TYPES()["required"]) == ["model_name", "precision", "quantization"]
    D = pkg.NODE_CLASS_MAPPINGS["HunyuanDependenciesLoader"]
    assert D.RETURN_TYPES == ("HUNYUAN_DEPS",) and D.FUNCTION == "load_dependencies"
    T = pkg.NODE_CLASS_MAPPINGS["HunyuanFoleyTorchCompile"]
    cfg, = T().make_config("inductor", "default", "None", False, 64)
    assert cfg == {"backend": "inductor", "mode": "default", "dynamic": None, "fullgraph": False, "dynamo_cache_limit": 64}
    B = pkg.NODE_CLASS_MAPPINGS["HunyuanBlockSwap"]
    assert B().set_args(blocks_to_swap=30, prefetch_blocks=1) == ({"blocks_to_swap": 30, "prefetch_blocks": 1},)


def test_saved_workflow_widgets_still_bind():
    """The reference's example workflow stores the Sampler widgets positionally; the order must be unchanged."""
    path = "/root/reference/example_workflows/HunyuanVideoFoleyExample.json"
    if not os.path.exists(path):
        pytest.skip("reference workflow not available here")
    wf = json.load(open(path))
    pkg = load_pkg()
    types = {n["type"] for n in wf["nodes"]}
    ours = set(pkg.NODE_CLASS_MAPPINGS)
    assert {"HunyuanModelLoader", "HunyuanDependenciesLoader", "HunyuanFoleySampler"} <= types
    assert {t for t in types if t.startswith("Hunyuan")} <= ours
    sampler = next(n for n in wf["nodes"] if n["type"] == "HunyuanFoleySampler")
    widgets = [k for k, v in pkg.NODE_CLASS_MAPPINGS["HunyuanFoleySampler"].INPUT_TYPES()["required"].items()
               if not (isinstance(v[0], str) and v[0].isupper() and v[0] not in ("FLOAT", "INT", "STRING", "BOOLEAN"))]
    # saved list = widgets + ComfyUI's control_after_generate entry after `seed`
    assert len(sampler["widgets_values"]) == len(widgets) + 1
    inputs = {i["name"]: i.get("type") for i in sampler["inputs"]}
    assert inputs.get("torch_compile_cfg") == "TORCH_COMPILE_CFG" and inputs.get("block_swap_args") == "BLOCKSWAPARGS"


def test_select_audio_from_batch_clamps_like_reference():
    pkg = load_pkg()
    node = pkg.NODE_CLASS_MAPPINGS["SelectAudioFromBatch"]()
    batch = {"waveform": torch.arange(12.0).view(3, 1, 4), "sample_rate": 48000}
    out, = node.select_audio(batch, 1)
    assert out["waveform"].shape == (1, 1, 4) and torch.equal(out["waveform"][0], batch["waveform"][1])
    out, = node.select_audio(batch, 7)   # out of range -> last item (reference nodes.py:653-656)
    assert torch.equal(out["waveform"][0], batch["waveform"][2])


def test_sampler_host_helpers_match_reference_rules():
    S = load_pkg("sampling")
    nodes = load_pkg("nodes")
    x = torch.randn(2, 9, 4)
    assert S._pad_or_trim_time(x, 9) is x
    p = S._pad_or_trim_time(x, 12)
    assert p.shape == (2, 12, 4) and torch.equal(p[:, :9], x) and float(p[:, 9:].abs().sum()) == 0.0
    assert torch.equal(S._pad_or_trim_time(x, 5), x[:, :5])
    sig = S.sigma_schedule(50)
    assert sig.shape == (51,) and sig[0] == 1 and sig[-1] == 0 and torch.equal(sig, torch.linspace(1, 0, 51))
    assert torch.allclose(S.sigma_schedule(10, 3.0), (3.0 * torch.linspace(1, 0, 11)) / (1 + 2.0 * torch.linspace(1, 0, 11)))
    # noise: CPU generator, drawn in the target dtype (reference utils.py:114-121 / diffusers.randn_tensor)
    g1 = torch.Generator(device="cpu").manual_seed(123)
    g2 = torch.Generator(device="cpu").manual_seed(123)
    a = S.prepare_latents_with_generator(None, 3, 128, 250, torch.bfloat16, "cpu", g1)
    b = torch.randn((3, 128, 250), generator=g2, dtype=torch.bfloat16)
    assert a.dtype == torch.bfloat16 and torch.equal(a, b)
    # T2A token counts (nodes.py:326-331) and frame picks (nodes.py:310,315)
    assert nodes.t2a_feature_lengths(5.0) == (40, 112) and nodes.t2a_feature_lengths(1.0) == (8, 16)
    assert nodes.t2a_feature_lengths(30.0) == (240, 736)
    idx = nodes.resample_frame_indices(80, 5.0, 8)
    assert idx.shape == (40,) and idx[0] == 0 and idx[-1] == 79
    with pytest.raises(ValueError):
        S.denoise_process_with_generator({}, {}, 1.0, None, None, 1.0, 10, 1, "dpm++")


def test_config_yaml_and_attribute_dict():
    cfgmod = load_pkg("config")
    for size, (C, H, nt, ns) in {"xxl": (1536, 12, 18, 36), "xl": (1408, 11, 12, 24)}.items():
        cfg = cfgmod.load_model_config(size)
        kw = cfg.model_config.model_kwargs
        assert (kw.hidden_size, kw.num_heads, kw.depth_triple_blocks, kw.depth_single_blocks) == (C, H, nt, ns)
        assert kw.audio_frame_rate == 50 and kw.audio_vae_latent_dim == 128 and kw.get("text_length") == 77
        assert cfg.diffusion_config.sample_flow_shift == 1.0
    with pytest.raises(FileNotFoundError):
        cfgmod.load_yaml("/nonexistent.yaml")
    E = load_pkg("engine")
    ec = E.engine_config(dict(cfgmod.load_model_config("xxl").model_config.model_kwargs))
    assert (ec.mlp_hidden_triple, ec.mlp_hidden_single, ec.sync_hidden) == (6144, 4096, 4096)
    ec = E.engine_config(dict(cfgmod.load_model_config("xl").model_config.model_kwargs))
    assert (ec.mlp_hidden_triple, ec.mlp_hidden_single) == (5632, 3840)


def test_shard_range_covers_batch_exactly():
    par = load_pkg("parallel")
    for gb in (0, 1, 5, 8, 32):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(gb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from conftest import load_pkg
par = load_pkg("parallel"); S = load_pkg("sampling")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
shapes = [(1, 8, 768), (1, 16, 768), (1, 9, 768), (1, 5, 768)]
feats = None
if rank == 0:
    g = torch.Generator().manual_seed(1)
    feats = {k: torch.randn(s, generator=g) for k, s in zip(par.COND_KEYS, shapes)}
got = par.broadcast_conditions(feats, shapes, "cpu", src=0, dtype=torch.float32)
g = torch.Generator().manual_seed(1)
want = {k: torch.randn(s, generator=g) for k, s in zip(par.COND_KEYS, shapes)}
assert all(torch.equal(got[k], want[k]) for k in par.COND_KEYS), "broadcast mismatch"
# one host noise draw for the global batch of 3; each rank takes its rows (ragged: 2 + 1)
gen = torch.Generator(device="cpu").manual_seed(123)
noise = S.prepare_latents_with_generator(None, 3, 128, 50, torch.float32, "cpu", gen)
lo, hi = par.shard_range(3, 2, rank)
local = noise[lo:hi]
wav = local.mean(dim=1, keepdim=True).repeat(1, 1, 4)          # stand-in for denoise+decode: [b, 1, T]
full = par.gather_waveforms(wav, 3, dst=0)
if rank == 0:
    assert full.shape == (3, 1, 200)
    assert torch.equal(full, noise.mean(dim=1, keepdim=True).repeat(1, 1, 4)), "gather order mismatch"
else:
    assert full is None
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
'''


def test_world_size_2_gloo_broadcast_shard_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "OK 0" in outs[0] and "OK 1" in outs[1]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints one JSON line with the contract keys (tiny model keeps it fast): the staged
    reference's own modules when baseline/_ref (or /root/reference) is present — complete Euler steps + a full decode,
    extrapolation declared — else the oracle port."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny",
                          "--steps", "1", "--warmup", "0", "--duration", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "audio_seconds_per_sec" and line["higher_is_better"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1
    if cb["kind"] == "reference":
        assert cb["extrapolated"] is True and cb["euler_steps_timed"] >= 1 and cb["t_dac_decode_s"] > 0
        assert abs(cb["extrapolation_factor"] - 50 / cb["euler_steps_timed"]) < 1e-9 and line["extrapolated"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["value"] > 0 and "workload" in line["config"]


def test_torch_ops_are_registered_and_refuse_cpu_tensors():
    """torch.ops.foley_b200.* (the C ABI surfaced as torch ops): schemas exist, the CPU key raises instead of falling
    back, the Meta key gives shapes."""
    ops = load_pkg("torch_ops")
    E = load_pkg("engine")
    for name in ("set_conditions", "dit_forward", "denoise", "denoise_solver", "dac_decode"):
        assert hasattr(torch.ops.foley_b200, name)
    x = torch.zeros(2, 128, 10)
    with pytest.raises(E.FoleyError):
        torch.ops.foley_b200.dit_forward(1, x, torch.tensor([1.0]))
    with pytest.raises(E.FoleyError):
        torch.ops.foley_b200.denoise_solver(1, x, torch.linspace(1, 0, 5), 4.5, 3)
    assert torch.ops.foley_b200.dac_decode(1, x.to("meta")).shape == (2, 1, 9600)
    assert torch.ops.foley_b200.denoise(1, x.to("meta"), torch.linspace(1, 0, 5).to("meta"), 4.5).shape == x.shape
    with pytest.raises(ValueError):
        ops.denoise(object.__new__(E.FoleyEngine), x, torch.linspace(1, 0, 5), 4.5, solver="dpm++")
    assert ops.SOLVER_IDS == E.FoleyEngine.SOLVERS


def test_solver_stage_machine_matches_reference_scheduler_rules():
    """The host stage machine used as the cross-check of foley_denoise_solver: with a constant derivative field every
    solver must land on sample + v * (sigma_k - sigma_0) after a whole number of stage groups, and the stage count per
    sigma interval is 1 / 2 / 2 / 4 (scheduling_flow_match_discrete.py:299-373)."""
    sampling = load_pkg("sampling")
    sig = sampling.sigma_schedule(8)
    v = torch.full((1, 4, 3), 2.0)
    for solver, stages in sampling.FlowMatchSolverState.STAGES.items():
        st = sampling.FlowMatchSolverState(solver, sig)
        x = torch.zeros(1, 4, 3)
        for _ in range(8):
            x = st.step(v, x)
        assert st.step_index == 8 // stages
        want = 2.0 * float(sig[8 // stages] - sig[0])
        assert torch.allclose(x, torch.full_like(x, want), atol=1e-6), solver


def test_fp8_wrap_rule_matches_what_the_reference_wraps(golden_dir=os.path.join(ROOT, "tests", "golden")):
    """tests/golden/fp8_wrapped_tiny.json lists the modules the reference's _wrap_fp8_inplace replaced on its own tiny
    model (all 56 Linear / Conv1d — its deny list never matches).  foley_fp8_wraps (C) and checkpoint.fp8_wraps (host
    mirror) must select exactly those modules' weights from the state dict."""
    from oracle import weights as W
    E, ck = load_pkg("engine"), load_pkg("checkpoint")
    lib = E.load_library()
    gold = json.load(open(os.path.join(golden_dir, "fp8_wrapped_tiny.json")))
    assert sorted(gold["wrapped"]) == sorted(gold["linear_like"])
    sd = W.synth_dit_state_dict(W.model_config("tiny"), seed=0)
    picked_c = sorted(k[:-len(".weight")] for k, v in sd.items() if lib.foley_fp8_wraps(k.encode(), v.dim()))
    picked_py = sorted(k[:-len(".weight")] for k, v in sd.items() if ck.fp8_wraps(k, v.dim()))
    assert picked_c == picked_py == sorted(gold["wrapped"])
    assert not lib.foley_fp8_wraps(b"dac.decoder.model.0.weight", 3)
    assert not lib.foley_fp8_wraps(b"triple_blocks.0.audio_self_q_norm.weight", 1)
    assert not lib.foley_fp8_wraps(b"sync_pos_emb", 3)


def test_loader_precision_and_quantization_resolution_match_reference():
    """utils.py:492-515 and nodes.py:106-121."""
    ck = load_pkg("checkpoint")
    bf, f16, f32, e4, e5 = torch.bfloat16, torch.float16, torch.float32, torch.float8_e4m3fn, torch.float8_e5m2
    assert ck.detect_fp8([bf, f32]) is None
    assert ck.detect_fp8([bf, e5, e4]) == "fp8_e5m2" and ck.detect_fp8([e4, e5]) == "fp8_e4m3fn"
    assert ck.detect_major_precision([(bf, 10), (f32, 11), (e4, 1000)]) == f32
    assert ck.detect_major_precision([(e4, 5)]) == bf and ck.detect_major_precision([]) == bf
    assert ck.detect_major_precision([(f16, 3), (bf, 3)]) == bf      # ties: first key of the reference's dict
    assert ck.resolve_quantization("none", "fp8_e5m2") is None
    assert ck.resolve_quantization("auto", None) == "fp8_e4m3fn"      # B200: capability 10 >= 9
    assert ck.resolve_quantization("auto", "fp8_e5m2") == "fp8_e5m2"
    assert ck.resolve_quantization("auto", None, capability_major=8) == "fp8_e5m2"
    assert ck.resolve_quantization("fp8_e5m2", "fp8_e4m3fn") == "fp8_e5m2"
    w = torch.tensor([0.3, -0.02, 1.7, 448.0, 0.0009765625 * 3])
    assert torch.equal(ck.round_through_fp8(w, "fp8_e4m3fn").float(), w.bfloat16().to(e4).float())
    assert torch.equal(ck.round_through_fp8(w.to(e5), "fp8_e5m2").float(), w.to(e5).float())


def test_safetensors_probe_and_header_parser(tmp_path):
    """foley_safetensors_probe parses the container without a GPU; the host writer/reader round-trips; malformed
    files are refused with an error instead of a crash."""
    E, ck = load_pkg("engine"), load_pkg("checkpoint")
    lib = E.load_library()
    g = torch.Generator().manual_seed(0)
    tensors = {
        "a.weight": torch.randn(8, 4, generator=g).bfloat16(),
        'odd "name"\\x': torch.randn(3, generator=g),
        "b.weight": torch.randn(2, 4, 3, generator=g).half(),
        "q.weight": torch.randn(16, 8, generator=g).to(torch.float8_e4m3fn),
        "r.weight": torch.randn(16, generator=g).to(torch.float8_e5m2),
        "steps": torch.arange(3),
        "empty": torch.zeros(0, 5),
    }
    path = str(tmp_path / "t.safetensors")
    ck.write_safetensors(path, tensors, metadata={"format": "pt", "note": "x,{}[]"})
    from safetensors.torch import load_file   # independent reader agrees with our writer
    back = load_file(path)
    assert set(back) == set(tensors) and all(torch.equal(back[k].view(torch.uint8), tensors[k].view(torch.uint8)) for k in tensors)
    n, nbytes, by = ctypes.c_int64(), ctypes.c_int64(), (ctypes.c_int64 * 5)()
    assert lib.foley_safetensors_probe(path.encode(), ctypes.byref(n), ctypes.byref(nbytes), by) == 0
    assert n.value == len(tensors)
    assert nbytes.value == sum(t.numel() * t.element_size() for t in tensors.values())
    assert list(by) == [32, 3, 24, 128, 16]
    hdr, start = ck.read_header(path)
    assert torch.equal(ck.read_tensor(path, "b.weight", hdr, start), tensors["b.weight"])
    # malformed: truncated payload, garbage header, header length beyond the file
    raw = open(path, "rb").read()
    bad = {"trunc": raw[:-5], "garbage": raw[:8] + b"not json" + raw[16:], "hlen": b"\xff" * 8 + raw[8:], "short": b"abc"}
    for name, blob in bad.items():
        p = str(tmp_path / f"{name}.safetensors")
        open(p, "wb").write(blob)
        assert lib.foley_safetensors_probe(p.encode(), None, None, None) == 1, name
        assert b"safetensors" in lib.foley_last_error()
    assert lib.foley_safetensors_probe(str(tmp_path / "missing.safetensors").encode(), None, None, None) == 1


def test_engine_solver_stage_table_replays_the_reference_scheduler():
    """foley_solver_table (the host half of foley_denoise_solver) applied with numpy in separately rounded fp32 must give
    the same numbers, bit for bit, as the reference scheduler's stage machine (sampling.FlowMatchSolverState) on a
    random derivative sequence — for call counts that end mid-stage too."""
    import numpy as np
    sampling, E = load_pkg("sampling"), load_pkg("engine")
    lib = E.load_library()
    lib.foley_solver_table.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_float), ctypes.c_int32, ctypes.POINTER(ctypes.c_float)]
    g = torch.Generator().manual_seed(0)
    for solver, sid in (("heun-2", 1), ("midpoint-2", 2), ("kutta-4", 3)):
        for n in (1, 4, 7, 10):
            sig = sampling.sigma_schedule(n, 1.0 if n % 2 else 3.0).float()
            arr = (ctypes.c_float * (n + 1))(*sig.tolist())
            out = (ctypes.c_float * (9 * n))()
            assert lib.foley_solver_table(sid, arr, n, out) == 0
            tab = np.array(out, dtype=np.float32).reshape(n, 9)
            st = sampling.FlowMatchSolverState(solver, sig)
            x_ref = torch.randn(2, 4, 5, generator=g)
            x = x_ref.numpy().copy()
            d = [None, None, None]
            base = None
            for i in range(n):
                mo = torch.randn(2, 4, 5, generator=g)
                x_ref = st.step(mo, x_ref)
                dt, c0, c1, c2, cm, kind, slot, save, use_base = tab[i]
                m = mo.numpy()
                if save:
                    base = x.copy()
                if slot >= 0:
                    d[int(slot)] = m.copy()
                if kind == 1:
                    der = np.float32(0.5) * (d[0] + m)
                elif kind == 2:
                    der = ((c0 * d[0] + c1 * d[1]) + c2 * d[2]) + cm * m
                else:
                    der = m
                x = ((base if use_base else x) + der * dt).astype(np.float32)
                assert np.array_equal(x, x_ref.numpy()), (solver, n, i)
    assert lib.foley_solver_table(0, arr, n, out) == 1     # Euler has no stage table


def test_safetensors_parser_survives_mutated_files(tmp_path):
    """Native header parser hardening: a few thousand mutated / truncated / re-headered files must be either accepted or
    refused with an error — never crash, never report sizes beyond the file."""
    import random
    import struct
    E, ck = load_pkg("engine"), load_pkg("checkpoint")
    lib = E.load_library()
    g = torch.Generator().manual_seed(0)
    base = str(tmp_path / "b.safetensors")
    ck.write_safetensors(base, {"a.weight": torch.randn(8, 4, generator=g).bfloat16(), "b": torch.randn(3, generator=g),
                                "c.weight": torch.randn(2, 4, 3, generator=g).half()}, metadata={"format": "pt"})
    raw = bytearray(open(base, "rb").read())
    (hlen,) = struct.unpack("<Q", raw[:8])
    rng = random.Random(1)
    p = str(tmp_path / "m.safetensors")
    seen = {0: 0, 1: 0}
    for _ in range(1500):
        b = bytearray(raw)
        mode = rng.randrange(6)
        if mode == 0:
            for _k in range(rng.randrange(1, 6)):
                b[8 + rng.randrange(hlen)] = rng.randrange(256)
        elif mode == 1:
            b = b[:rng.randrange(0, len(b))]
        elif mode == 2:
            b[:8] = struct.pack("<Q", rng.choice([0, 1, 7, hlen - 1, hlen + 1, len(b), 2 ** 40, 2 ** 63, rng.randrange(0, 2 * len(b))]))
        elif mode == 3:
            i = 8 + rng.randrange(hlen)
            del b[i:min(len(b), i + rng.randrange(1, 20))]
        elif mode == 4:
            i = 8 + rng.randrange(hlen)
            b[i:i] = bytes(rng.randrange(256) for _k in range(rng.randrange(1, 10)))
        else:
            hdr = json.loads(bytes(raw[8:8 + hlen]))
            k = rng.choice([k for k in hdr if k != "__metadata__"])
            ch = rng.randrange(5)
            if ch == 0:
                hdr[k]["data_offsets"] = [rng.randrange(0, 2 ** 62), rng.randrange(0, 2 ** 62)]
            elif ch == 1:
                hdr[k]["shape"] = [rng.randrange(0, 2 ** 40) for _k in range(rng.randrange(0, 6))]
            elif ch == 2:
                hdr[k]["dtype"] = rng.choice(["F8_E4M3", "I64", "X", "", "F32"])
            elif ch == 3:
                hdr["__metadata__"] = {"a": {"b": [1, 2, {"c": None}]}, "x": "y\\\"z"}
            else:
                hdr[k] = rng.choice([[], 3, "s", {"dtype": "F32"}])
            hj = json.dumps(hdr).encode()
            b = bytearray(struct.pack("<Q", len(hj)) + hj + bytes(raw[8 + hlen:]))
        open(p, "wb").write(b)
        n, nb = ctypes.c_int64(), ctypes.c_int64()
        st = lib.foley_safetensors_probe(p.encode(), ctypes.byref(n), ctypes.byref(nb), None)
        assert st in (0, 1)
        if st == 0:
            assert 0 <= nb.value <= len(b)
        seen[st] += 1
    assert seen[0] > 0 and seen[1] > 0


def test_safetensors_parser_refuses_deep_nesting_without_overflowing_the_stack(tmp_path):
    """a header whose __metadata__ (or an unknown key of a tensor entry) nests millions of brackets used to
    recurse once per level and overflow the stack; the parser now refuses anything deeper than 64 levels."""
    import struct
    E = load_pkg("engine")
    lib = E.load_library()
    entry = '"a":{"dtype":"F32","shape":[1],"data_offsets":[0,4]'
    for depth, inside_entry in ((2_000_000, False), (500_000, True), (40, False)):
        nest = "[" * depth + "]" * depth
        hdr = ('{"__metadata__":' + nest + "," + entry + "}}") if not inside_entry else ("{" + entry + ',"x":' + nest + "}}")
        hj = hdr.encode()
        p = str(tmp_path / f"deep{depth}.safetensors")
        open(p, "wb").write(struct.pack("<Q", len(hj)) + hj + b"\0\0\0\0")
        n, nb = ctypes.c_int64(), ctypes.c_int64()
        st = lib.foley_safetensors_probe(p.encode(), ctypes.byref(n), ctypes.byref(nb), None)
        assert st == (0 if depth <= 64 else 1)
        if st == 0:
            assert n.value == 1 and nb.value == 4


def test_precision_other_than_bf16_is_refused_not_substituted(monkeypatch):
    """Reference utils.py:222-239 computes in the loader's dtype; the engine implements the bf16 path only and must
    say so ('never silently substitute'), with an explicit opt-in for the bf16 fallback."""
    nodes, E = load_pkg("nodes"), load_pkg("engine")
    monkeypatch.delenv("FOLEY_B200_PRECISION_FALLBACK", raising=False)
    nodes._check_precision("bf16")
    nodes._check_precision("auto", torch.bfloat16)
    for p, dt in (("fp32", None), ("fp16", None), ("auto", torch.float32), ("auto", torch.float16)):
        with pytest.raises(E.FoleyError, match="precision"):
            nodes._check_precision(p, dt)
    monkeypatch.setenv("FOLEY_B200_PRECISION_FALLBACK", "bf16")
    nodes._check_precision("fp32")
    nodes._check_precision("auto", torch.float16)


def test_feature_bridge_wires_encoders_like_the_reference():
    """Host logic of feature_bridge.make_extract_features with fake encoders: prompts are tokenised as [negative, prompt]
    with padding (feature_utils.py:134, utils.py:284), row 0 becomes uncond_text_feat and row 1 text_feat
    (utils.py:286), SigLIP features get the leading batch axis (feature_utils.py:77-78), the audio length follows the
    25 fps stream (utils.py:281), and a missing Synchformer only blocks video-to-audio."""
    fb, E = load_pkg("feature_bridge"), load_pkg("engine")
    calls = {}

    class Tok:
        def __call__(self, texts, padding=True, return_tensors="pt"):
            calls["texts"], calls["padding"] = list(texts), padding
            ids = torch.tensor([[0, 5, 2, 1, 1], [0, 7, 8, 9, 2]])
            return {"input_ids": ids, "attention_mask": (ids != 1).long()}

    class Sig:
        def encode(self, px):
            return torch.full((px.shape[0], 768), 2.0)

    class Clap:
        def encode(self, ids, mask):
            calls["mask"] = mask
            return torch.arange(2.0).view(2, 1, 1).expand(2, ids.shape[1], 768)

    dev = torch.device("cpu")
    ex = fb.make_extract_features(Sig(), Tok(), Clap(), lambda fr: torch.zeros(1, 8 * ((fr.shape[1] - 16) // 8 + 1), 768), dev)
    visual, text, alen = ex(torch.zeros(8, 3, 512, 512), torch.zeros(25, 3, 224, 224), "a dog barks", "noisy")
    assert calls["texts"] == ["noisy", "a dog barks"] and calls["padding"] is True
    assert visual["siglip2_feat"].shape == (1, 8, 768) and visual["syncformer_feat"].shape == (1, 16, 768) and alen == 1.0
    assert float(text["uncond_text_feat"].max()) == 0.0 and float(text["text_feat"].min()) == 1.0
    assert text["text_feat"].shape == (1, 5, 768) and calls["mask"].tolist() == [[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]]
    # text-to-audio needs no visual encoder at all
    ex_t2a = fb.make_extract_features(Sig(), Tok(), Clap(), "reference package not installed", dev)
    visual, text, alen = ex_t2a(None, None, "p", "n")
    assert visual == {} and alen is None and text["text_feat"].shape == (1, 5, 768)
    with pytest.raises(E.FoleyError, match="Synchformer"):
        ex_t2a(torch.zeros(8, 3, 512, 512), torch.zeros(25, 3, 224, 224), "p", "n")
    with pytest.raises(E.FoleyError, match="tokens"):
        fb.make_extract_features(Sig(), Tok(), Clap(), "x", dev, max_text_tokens=4)(None, None, "p", "n")
