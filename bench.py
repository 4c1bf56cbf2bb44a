#!/usr/bin/env python
"""Benchmark of the denoise hot path (BASELINE.json): audio-seconds/sec and denoise-steps/sec at
5 s / 50 Euler steps / CFG 4.5, synthetic conditions + seeded random weights of the named architecture.

    python bench.py --gpus N --steps K --warmup W [--impl b200|reference] [--model xl|xxl] [--batch B]

One bench "step" = one whole job of the hot path for the local batch: 50 x (CFG-pair DiT forward + CFG combine +
Euler update) + one DAC-VAE decode.  `value` is measured with conditions / noise already resident in HBM;
`e2e` goes through the public host API (sampling.denoise_process_with_generator, what the Sampler node calls)
with pinned HOST buffers, H2D of conditions + noise and D2H of the waveforms inside the timed region.
Multi-GPU: one process per GPU (torchrun), variations sharded across ranks (weak scaling: --batch per GPU), one
NCCL broadcast of the condition embeddings and one gather of the decoded waveforms in the e2e leg.

--impl reference times the reference's own CPU path on the host cores: the UNMODIFIED reference modules staged under
baseline/_ref (tools/stage_reference.py; imported through tools/ref_shims.py, which only fakes the ComfyUI / diffusers
packages the reference imports), fp32, all host threads, through the reference's own `denoise_process_with_generator`
— complete Euler steps (every block, embedders, final layer, CFG combine, scheduler step) and one full-length
`DAC.decode` — on a bounded sample of the workload: `--ref-steps` Euler steps per bench step,
extrapolated linearly to the 50 of the job (every step has identical shapes; `extrapolated` and the factor are in the
line).  Only when no staged reference is present does the arm fall back to the oracle port (kind "port").
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from tools import synthetic as SY  # noqa: E402

# Algorithmic (reference-executed) TFLOP per Euler step for one CFG pair at 5 s (FlopCounterMode)
STEP_TFLOP_5S = {"xxl": 3.8437, "xl": 2.1865}
DAC_GFLOP_5S = 577.3
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def load_pkg(sub):
    import __graft_entry__ as ge
    return ge.load_pkg(sub)


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_sample(model, batch, duration, n_steps, guidance, threads=None):
    """Times the oracle port (fp32 CPU) on one triple block + one single block + embed/final + a short DAC decode at
    the benchmark shapes and extrapolates to the whole job.  Returns (audio_s_per_s, steps_per_s, detail)."""
    from oracle import foley_oracle as O   # checker / baseline only; never on the product path
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    c = SY.model_config(model)
    L, Lv, S = SY.clip_lengths(duration)
    B2 = batch * (2 if guidance > 1.0 else 1)
    C, D, T = c["hidden_size"], c["head_dim"], 77
    # one block's worth of weights (shapes only matter for timing); reuse for the extrapolated blocks
    specs = [s for s in SY.dit_param_specs(c) if s[0].startswith(("triple_blocks.0.", "single_blocks.0."))]
    sd = {n: torch.randn(sh) * 0.02 for n, sh, _ in specs}
    p = O.Policy("fp32")
    g = torch.Generator().manual_seed(0)
    audio = torch.randn(B2, L, C, generator=g)
    v_cond = torch.randn(B2, Lv, C, generator=g)
    cond = torch.randn(B2, T, C, generator=g)
    vec = torch.randn(B2, C, generator=g)
    vec_tok = torch.randn(B2, L, C, generator=g)
    a_pos, v_pos = O.interleaved_positions(L, Lv)
    rope_av = (O.rope_tables(a_pos, D), O.rope_tables(v_pos, D))
    rope_a, rope_v, rope_t = O.rope_tables(torch.arange(L), D), O.rope_tables(torch.arange(Lv), D), O.rope_tables(torch.arange(T), D)

    def timed(fn, reps=2):
        fn()  # warm-up
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    with torch.inference_mode():
        t_triple = timed(lambda: O.triple_block(p, sd, "triple_blocks.0.", c, audio, cond, v_cond, vec, rope_av, rope_a, rope_v, rope_t))
        t_single = timed(lambda: O.single_block(p, sd, "single_blocks.0.", c, audio, vec_tok, rope_a))
        dsd = SY.synth_dac_state_dict(SY.DAC_CONFIG, seed=3)
        Ld = 10
        z = torch.randn(1, 128, Ld, generator=g)
        t_dac = timed(lambda: O.dac_decode(dsd, z), reps=1) * (L / Ld) * batch
    t_step = c["depth_triple_blocks"] * t_triple + c["depth_single_blocks"] * t_single
    t_job = n_steps * t_step + t_dac
    detail = {"t_triple_block_s": t_triple, "t_single_block_s": t_single, "t_dac_decode_s": t_dac,
              "t_euler_step_s": t_step, "t_job_s": t_job}
    return batch * duration / t_job, 1.0 / t_step, detail


_REF = {}


def staged_reference():
    """tools/ref_shims + tools/gpu_reference_golden helpers when the reference tree is staged (baseline/_ref), else None."""
    if "G" not in _REF:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ref_shims
            if not ref_shims.available():
                raise ImportError("no staged reference")
            import gpu_reference_golden as G
            _REF["G"], _REF["ns"] = G, ref_shims.load_reference()
        except Exception as e:   # noqa: BLE001
            _REF["G"], _REF["ns"], _REF["why"] = None, None, repr(e)
    return _REF["G"], _REF["ns"]


def cpu_reference_measured(model, batch, duration, n_steps, guidance, k_steps=1, threads=None):
    """The reference's OWN modules on the host cores (fp32, eager): k complete Euler steps through its
    `denoise_process_with_generator` (utils.py:125-258) + one full-length `DAC.decode`, extrapolated to the n_steps of the
    job.  Returns (audio_s_per_s, steps_per_s, detail) or None when no reference tree is staged."""
    G, ns = staged_reference()
    if G is None:
        return None
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cpu = torch.device("cpu")
    c = SY.model_config(model)
    L, Lv, S = SY.clip_lengths(duration)
    key = ("cpu", model)
    if key not in _REF:
        t0 = time.time()
        if torch.cuda.is_available():    # same seeded weights as the GPU arm (philox on the device), moved to the host
            sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, torch.device("cuda", torch.cuda.current_device()), torch.bfloat16)
        else:
            sd = SY.synth_dit_state_dict(c, seed=0)
        ref_model, ref_cfg = G.build_ref_model(ns, model, sd, torch.float32, cpu)
        del sd
        ref_dac = G.build_ref_dac(ns, SY.synth_dac_state_dict(SY.DAC_CONFIG, seed=3), cpu)
        _REF[key] = (ref_model, ref_cfg, ref_dac, time.time() - t0)
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    ref_model, ref_cfg, ref_dac, t_build = _REF[key]
    feats = {k: v.bfloat16().float() for k, v in SY.synth_conditions(c, L, Lv, S).items()}
    t0 = time.perf_counter()
    lat, _audio = G.run_ref_denoise(ns, ref_model, ref_dac, ref_cfg, feats, duration, k_steps, guidance, batch, cpu, torch.float32)
    t_k = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.inference_mode():
        ref_dac.decode(lat)
    t_dec = time.perf_counter() - t0
    t_step = max(t_k - t_dec, 1e-9) / k_steps
    t_job = n_steps * t_step + t_dec
    detail = {"t_euler_step_s": t_step, "t_dac_decode_s": t_dec, "t_job_s": t_job, "t_sample_s": t_k + t_dec,
              "euler_steps_timed": k_steps, "extrapolated": True, "extrapolation_factor": n_steps / k_steps,
              "t_model_build_s": t_build}
    return batch * duration / t_job, 1.0 / t_step, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count()
    vals, steps_s = [], []
    kind = "reference" if staged_reference()[0] is not None else "port"
    t_wall0 = time.time()
    for i in range(args.warmup + args.steps):
        if kind == "reference":
            v, s, detail = cpu_reference_measured(args.model, args.batch * args.gpus, args.duration, args.denoise_steps, args.cfg,
                                                  k_steps=args.ref_steps)
        else:
            v, s, detail = cpu_reference_sample(args.model, args.batch * args.gpus, args.duration, args.denoise_steps, args.cfg)
        if i >= args.warmup:
            vals.append(v)
            steps_s.append(s)
    value = sum(vals) / len(vals)
    if kind == "reference":
        sample = (f"the reference's own modules (baseline/_ref, fp32, eager, {cores} host threads), per bench step: "
                  f"{args.ref_steps} complete Euler step(s) through its denoise_process_with_generator (all blocks, embedders, "
                  f"final layer, CFG combine, scheduler step) + one full-length DAC.decode; extrapolated linearly x"
                  f"{args.denoise_steps / args.ref_steps:g} to {args.denoise_steps} Euler steps")
    else:
        sample = ("oracle port (fp32 torch-CPU restatement of the reference; no staged reference tree: " + _REF.get("why", "") + "), "
                  "per bench step: 1 triple block + 1 single block + DAC decode of 10 latent frames at the benchmark shapes, "
                  f"extrapolated linearly to {args.denoise_steps} Euler steps x all blocks + full decode")
    line = {
        "impl": "reference", "metric": "audio_seconds_per_sec", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * args.batch * args.gpus * args.duration / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "denoise_steps_per_sec": sum(steps_s) / len(steps_s),
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample, **detail},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "extrapolated": True,
        "timed_region_s": time.time() - t_wall0,
        "note": "value = job throughput implied by the timed sample (ms_per_step is the implied time of ONE whole job, not "
                "the wall time of a bench step; the wall time of the whole arm is timed_region_s)",
    }
    print(json.dumps(line))
    return 0


def workload_config(args):
    L, Lv, S = SY.clip_lengths(args.duration)
    return {"workload": f"V2A {args.model} {args.duration:g}s@8fps synthetic, {args.denoise_steps} Euler steps, CFG {args.cfg:g}, "
                        f"bf16, batch_size={args.batch}/GPU",
            "model_size": args.model, "duration_s": args.duration, "denoise_steps": args.denoise_steps,
            "cfg_scale": args.cfg, "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
            "tokens": {"audio_L": L, "clip_Lv": Lv, "sync_S": S, "text_T": 77},
            "parallelism": f"variations sharded x{args.gpus} (weights replicated)",
            "l2_policy": "weights streamed per step are 5.8 GB (xl) / 10.3 GB (xxl) >> 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, nodes, sampling, cfgmod = load_pkg("engine"), load_pkg("nodes"), load_pkg("sampling"), load_pkg("config")
    par = load_pkg("parallel")
    peaks = load_peaks()

    c = SY.model_config(args.model)
    cfg = cfgmod.load_model_config(args.model)
    L, Lv, S = SY.clip_lengths(args.duration)
    B, U = args.batch, (2 if args.cfg > 1.0 else 1)
    t0 = time.time()
    sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)   # same seed on every rank
    eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev)
    eng.load_state_dict(sd)
    eng.finalize()
    model = nodes.FoleyModel(eng, sd["empty_clip_feat"].cpu(), sd["empty_sync_feat"].cpu(), cfg, dtype=torch.bfloat16)
    del sd
    dsd = SY.synth_state_dict_cuda(SY.dac_param_specs(SY.DAC_CONFIG), 3, dev, torch.float32)
    dac = nodes.FoleyDAC.from_state_dict(dsd, device=dev)
    del dsd
    torch.cuda.empty_cache()
    setup_s = time.time() - t0

    # synthetic conditions on rank 0's pinned host memory, one host noise draw for the global batch
    feats = SY.synth_conditions(c, L, Lv, S, dtype=torch.bfloat16)
    feats = {k: v.pin_memory() for k, v in feats.items()}
    gen = torch.Generator(device="cpu").manual_seed(123)
    noise_all = torch.randn((B * world, 128, L), generator=gen, dtype=torch.bfloat16).pin_memory()
    sig = sampling.sigma_schedule(args.denoise_steps, 1.0)
    stream = torch.cuda.Stream(device=dev)

    cond_shapes = [tuple(feats[k].shape) for k in par.COND_KEYS]

    def share_conditions(f_host):
        """rank 0 -> all: ONE broadcast of the packed condition embeddings (NCCL over NVLink when world > 1)."""
        return par.broadcast_conditions(f_host if rank == 0 else None, cond_shapes, dev, src=0)

    def cond_rows(f):
        text = sampling._pad_or_trim_time(f["text_feat"], 77)
        utext = sampling._pad_or_trim_time(f["uncond_text_feat"], 77)
        if U == 2:
            uclip = model.get_empty_clip_sequence(bs=1, len=Lv).to(dev)
            usync = model.get_empty_sync_sequence(bs=1, len=S).to(dev)
            return torch.cat([uclip, f["siglip2_feat"]]), torch.cat([usync, f["syncformer_feat"]]), torch.cat([utext, text])
        return f["siglip2_feat"], f["syncformer_feat"], text

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.cuda.stream(stream):
        # ================= device-resident leg (`value`) =================
        f_dev = share_conditions(feats)
        eng.set_conditions(*cond_rows(f_dev), L=L, batch=B)
        noise_dev = noise_all[rank * B:(rank + 1) * B].to(dev).float()
        for _ in range(args.warmup):
            lat = eng.denoise(noise_dev, sig, args.cfg)
            wav = dac.decode(lat)
        torch.cuda.synchronize()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
        launches0 = eng.launch_count() + dac.engine.launch_count()
        torch.cuda.synchronize()
        ev[0].record()
        for k in range(args.steps):
            lat = eng.denoise(noise_dev, sig, args.cfg)
            ev[2 * k + 1].record()
            wav = dac.decode(lat)
            ev[2 * k + 2].record()
        torch.cuda.synchronize()
        barrier()
        launches = eng.launch_count() + dac.engine.launch_count() - launches0
        total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
        den_ms = max_over_ranks(sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)))
        dac_ms = max_over_ranks(sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(args.steps)))
        if rank == 0:
            sampler.stop_flag.set()
            sampler.join(timeout=3)

        # ================= end-to-end leg (`e2e`): public host API, host buffers =================
        # progress reporting ON, as in the node's default: the engine publishes the step counter in mapped host memory and
        # the calling thread polls it (no per-step stream synchronize); only rank 0's bar is driven in the product
        deps = cfgmod.AttributeDict({"dac_model": dac, "device": dev, "report_progress": True})
        deps["foley_model"] = model
        h2d = sum(v.numel() * v.element_size() for v in feats.values()) + B * 128 * L * 2
        d2h = B * world * L * 960 * 4

        def e2e_once():
            f = share_conditions(feats)                       # H2D on rank 0 (+ NCCL broadcast)
            visual = {"siglip2_feat": f["siglip2_feat"], "syncformer_feat": f["syncformer_feat"]}
            text = {"text_feat": f["text_feat"], "uncond_text_feat": f["uncond_text_feat"]}
            g = torch.Generator(device="cpu").manual_seed(123)
            audio, _sr = sampling.denoise_process_with_generator(
                visual, text, args.duration, deps, cfg, guidance_scale=args.cfg, num_inference_steps=args.denoise_steps,
                batch_size=B * world, sampler="euler", generator=g, batch_slice=(rank * B, (rank + 1) * B))
            full = par.gather_waveforms(audio.float(), B * world, dst=0)   # ONE gather of decoded waveforms
            return sampling.to_host(full) if full is not None else None

        for _ in range(max(1, args.warmup - 1)):
            e2e_once()
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = e2e_once()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1))

        # ================= dominant kernel, timed alone (roofline) =================
        roof = None
        if rank == 0:
            roof = time_dominant_gemm(eng, c, B * U, L, peaks)

        # ================= parity gate: this build against the reference's own B200 run =================
        parity = parity_vs_reference_golden(args, model, dac, cfg, cfgmod, c, dev) if rank == 0 else None

        # ================= the other named configurations, per-GPU shapes (BASELINE.json configs 4 and 5) =================
        extra = {}
        if args.extra_configs:
            for tag, (m_, b_, d_) in {"config4_xl_5s_batch4_per_gpu": ("xl", 4, 5.0), "config5_xxl_30s_batch1_per_gpu": ("xxl", 1, 30.0)}.items():
                r = measure_config(m_, b_, d_, args, dev, world, rank, (eng if m_ == args.model else None), E, nodes, cfgmod, sampling,
                                   barrier, max_over_ranks, peaks)
                if rank == 0:
                    extra[tag] = r

        # ================= informational: the condition encoders of the same clip =================
        enc_info = condition_encoders(args, dev) if rank == 0 and world == 1 and not args.no_encoders else None

        # ================= informational: the reference itself on this GPU, eager under CUDA autocast =================
        ref_gpu = None
        if rank == 0 and world == 1 and not args.no_ref_gpu:
            ref_gpu = reference_gpu_eager(args, c, dev)

    if rank == 0:
        gb = B * world
        value = gb * args.duration * args.steps / (total_ms / 1e3)
        steps_per_sec = args.denoise_steps * args.steps / (den_ms / 1e3)
        e2e_val = gb * args.duration * args.steps / (e2e_ms / 1e3)
        step_tflop = STEP_TFLOP_5S.get(args.model, 0) * B if abs(args.duration - 5.0) < 1e-6 and U == 2 else None
        line = {
            "metric": "audio_seconds_per_sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args),
            "denoise_steps_per_sec": steps_per_sec, "sample_steps_per_sec": steps_per_sec * gb,
            "ms_per_denoise_step": den_ms / args.steps / args.denoise_steps, "ms_dac_decode": dac_ms / args.steps,
            "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roof,
            "setup_s": setup_s,
        }
        if step_tflop:
            ach = step_tflop / (den_ms / args.steps / args.denoise_steps / 1e3)
            line["step_roofline"] = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"],
                                     "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"],
                                     "peak_source": peaks["_source"] + " sustained",
                                     "flops": "algorithmic TFLOP per Euler step the reference executes (BASELINE.md §3)"}
        if parity is not None:
            line["parity"] = parity
        if ref_gpu is not None:
            line["reference_gpu_eager"] = ref_gpu
        if enc_info is not None:
            line["encoders"] = enc_info
        if extra:
            line["extra_configs"] = extra
        if world == 1 and not args.no_cpu_baseline:
            got = cpu_reference_measured(args.model, B, args.duration, args.denoise_steps, args.cfg, k_steps=1)
            if got is not None:
                v, s, detail = got
                line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "reference",
                                        "denoise_steps_per_sec": s,
                                        "sample": "the reference's own modules (baseline/_ref), fp32 eager on the host cores: 1 complete "
                                                  "Euler step through its denoise_process_with_generator + one full-length DAC.decode, "
                                                  f"extrapolated linearly to {args.denoise_steps} steps", **detail}
            else:
                v, s, detail = cpu_reference_sample(args.model, B, args.duration, args.denoise_steps, args.cfg)
                line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                                        "denoise_steps_per_sec": s,
                                        "sample": "oracle port fp32 (no staged reference): 1 triple block + 1 single block + DAC decode of 10 "
                                                  "frames at the benchmark shapes, extrapolated linearly to the whole job", **detail}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def parity_vs_reference_golden(args, model, dac, cfg, cfgmod, c, dev):
    """rel-L2 of this build's final latents / waveform against tests/golden/full_xl_5s_50step.pt — outputs of the REFERENCE'S
    OWN modules run on a B200 (tools/gpu_reference_golden.py) on the same seeded weights, conditions, noise and sigma
    schedule — next to the reference's own floors stored with the golden."""
    path = os.path.join(ROOT, "tests", "golden", "full_xl_5s_50step.pt")
    if not os.path.exists(path):
        return {"unavailable": "no golden"}
    gold = torch.load(path)
    a = gold["args"]
    if not (args.model == a["model"] and abs(args.duration - a["duration"]) < 1e-6 and args.denoise_steps == a["steps"] and
            abs(args.cfg - a["guidance"]) < 1e-6):
        return {"unavailable": "golden exists for xl / 5 s / 50 steps / CFG 4.5 only"}
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity_reference_fullsize import _checksum, _same_weights
    from conftest import rel_l2
    sd_probe = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
    same = _same_weights(_checksum(sd_probe), gold["checksum"]["dit"])
    del sd_probe
    torch.cuda.empty_cache()
    if not same:
        return {"unavailable": "device-drawn synthetic weights differ from the golden's (other GPU model / torch build)"}
    L, Lv, S = SY.clip_lengths(args.duration)
    feats = {k: v.to(dev, torch.bfloat16) for k, v in SY.synth_conditions(c, L, Lv, S).items()}
    sampling = load_pkg("sampling")
    deps = cfgmod.AttributeDict({"dac_model": dac, "device": dev, "report_progress": False})
    deps["foley_model"] = model
    gen = torch.Generator(device="cpu").manual_seed(a["seed"])
    lat, _ = sampling.denoise_process_with_generator(
        {"siglip2_feat": feats["siglip2_feat"], "syncformer_feat": feats["syncformer_feat"]},
        {"text_feat": feats["text_feat"], "uncond_text_feat": feats["uncond_text_feat"]}, a["duration"], deps, cfg,
        guidance_scale=a["guidance"], num_inference_steps=a["steps"], batch_size=1, sampler="euler", generator=gen, decode=False)
    wav = dac.decode(lat).float().cpu()
    lat = lat.float().cpu()
    ref16, ref32 = gold["lat_ref_bf16_b1"], gold["lat_ref_fp32"]
    w16, w32 = gold["wav_ref_bf16_b1"].float(), gold["wav_ref_fp32"].float()
    return {
        "latents_rel_l2": rel_l2(lat, ref16), "latents_rel_l2_vs_ref_fp32": rel_l2(lat, ref32),
        "waveform_rel_l2": rel_l2(wav, w16), "waveform_rel_l2_vs_ref_fp32": rel_l2(wav, w32),
        "ref_bf16_vs_fp32": rel_l2(ref16, ref32), "ref_bf16_vs_fp32_waveform": rel_l2(w16, w32),
        "ref_bf16_run_to_run_floor": rel_l2(gold["lat_ref_bf16_b2"][:1], ref16),
        "golden": "tests/golden/full_xl_5s_50step.pt: the reference's own denoise_process_with_generator + DAC.decode on a B200 "
                  "(CUDA autocast bf16; fp32 with TF32 off), same seeded weights / conditions / noise / sigma schedule",
        "gate": "latents <= 1.5 x run-to-run floor of the reference's bf16 path + 2e-4, and <= 1.1 x (ref bf16 vs fp32) against fp32 "
                "(tests/test_gpu_parity_reference_fullsize.py); BASELINE.json's 1e-3 is below the reference's own floor on a B200",
    }


def measure_config(model_name, B, duration, args, dev, world, rank, eng, E, nodes, cfgmod, sampling, barrier, max_over_ranks, peaks):
    """One more named configuration at its per-GPU shape (weak scaling: every rank runs it): device-timed denoise loop."""
    c = SY.model_config(model_name)
    cfg = cfgmod.load_model_config(model_name)
    L, Lv, S = SY.clip_lengths(duration)
    U = 2 if args.cfg > 1.0 else 1
    own = eng is None
    if own:
        sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
        eng = E.FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev)
        eng.load_state_dict(sd)
        eng.finalize()
        empty_clip, empty_sync = sd["empty_clip_feat"], sd["empty_sync_feat"]
        del sd
        torch.cuda.empty_cache()
    else:
        empty_clip = empty_sync = None
    f = {k: v.to(dev) for k, v in SY.synth_conditions(c, L, Lv, S, dtype=torch.bfloat16).items()}
    text = sampling._pad_or_trim_time(f["text_feat"], 77)
    utext = sampling._pad_or_trim_time(f["uncond_text_feat"], 77)
    if U == 2:
        if empty_clip is None:
            sdp = SY.synth_state_dict_cuda([s_ for s_ in SY.dit_param_specs(c) if s_[0] in ("empty_clip_feat", "empty_sync_feat")], 0, dev, torch.bfloat16)
            empty_clip, empty_sync = sdp["empty_clip_feat"], sdp["empty_sync_feat"]
        uclip = empty_clip.to(dev).view(1, 1, -1).expand(1, Lv, -1)
        usync = empty_sync.to(dev).view(1, 1, -1).expand(1, S, -1)
        rows = (torch.cat([uclip, f["siglip2_feat"]]), torch.cat([usync, f["syncformer_feat"]]), torch.cat([utext, text]))
    else:
        rows = (f["siglip2_feat"], f["syncformer_feat"], text)
    eng.set_conditions(*[r.contiguous() for r in rows], L=L, batch=B)
    g = torch.Generator(device="cpu").manual_seed(123)
    noise = torch.randn((B, 128, L), generator=g, dtype=torch.bfloat16).to(dev).float()
    sig = sampling.sigma_schedule(args.denoise_steps, 1.0)
    eng.denoise(noise, sig, args.cfg)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 2
    e0.record()
    for _ in range(n):
        eng.denoise(noise, sig, args.cfg)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / n
    ms_step = ms / args.denoise_steps
    tflop = {("xl", 4, 5.0): 8.746, ("xxl", 1, 30.0): 24.303, ("xl", 1, 30.0): 13.885}.get((model_name, B, duration))
    out = {"workload": f"V2A {model_name} {duration:g}s, {args.denoise_steps} Euler steps, CFG {args.cfg:g}, bf16, batch_size={B}/GPU x{world} GPUs",
           "ms_per_denoise_step": ms_step, "sample_steps_per_sec": world * B * 1e3 / ms_step,
           "audio_seconds_per_sec_denoise_only": world * B * duration / (ms / 1e3), "tokens": {"audio_L": L, "clip_Lv": Lv, "sync_S": S}}
    if tflop and U == 2:
        ach = tflop / (ms_step / 1e3)
        out["step_roofline"] = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                "frac": ach / peaks["bf16_tflops_sustained"], "flops": "BASELINE.md §3 per-GPU TFLOP per Euler step"}
    if own:
        del eng
        torch.cuda.empty_cache()
    return out


def condition_encoders(args, dev):
    """Informational: SigLIP2 vision tower (+ pooling head) on the clip's 8 fps frames and the CLAP text tower on the prompt
    pair, on the engine (encoders.py) and as the HF modules the reference runs (eager, bf16 module: nodes.py:283-284) on the
    same GPU with the same seeded weights.  CUDA events after warm-up."""
    try:
        from transformers import ClapTextConfig, ClapTextModelWithProjection, SiglipVisionConfig, SiglipVisionModel
        enc = load_pkg("encoders")

        def timed(fn, iters=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / iters

        torch.manual_seed(0)
        T = int(args.duration * 8)
        px = torch.rand(T, 3, 512, 512, device=dev) * 2 - 1
        ids = torch.randint(3, 50265, (2, 77))
        ids[:, 0] = 0
        ids[0, 40:] = 1
        mask = (ids != 1).long()
        out = {}
        with torch.inference_mode():
            sig = SiglipVisionModel(SiglipVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                                                       image_size=512, patch_size=16, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)).eval().to(dev)
            e_sig = enc.SiglipVisionEncoder.from_hf(sig, device=dev)
            sig = sig.to(torch.bfloat16)
            got, want = e_sig.encode(px).float(), sig(pixel_values=px).pooler_output.float()
            out["siglip2_vision"] = {"frames": T, "ms": timed(lambda: e_sig.encode(px)), "ms_hf_eager_bf16": timed(lambda: sig(pixel_values=px).pooler_output, 3),
                                     "rel_l2_vs_hf_bf16": float((got - want).norm() / want.norm())}
            del sig, e_sig
            clap = ClapTextModelWithProjection(ClapTextConfig()).eval().to(dev)
            e_clap = enc.ClapTextEncoder.from_hf(clap, device=dev)
            clap = clap.to(torch.bfloat16)
            idc, mc = ids.to(dev), mask.to(dev)
            got, want = e_clap.encode(ids, mask).float(), clap(input_ids=idc, attention_mask=mc).last_hidden_state.float()
            out["clap_text"] = {"tokens": [2, 77], "ms": timed(lambda: e_clap.encode(ids, mask)),
                                "ms_hf_eager_bf16": timed(lambda: clap(input_ids=idc, attention_mask=mc).last_hidden_state),
                                "rel_l2_vs_hf_bf16": float((got - want).norm() / want.norm())}
            del clap, e_clap
            # Synchformer visual extractor: 25 fps frames of the clip; the reference's own module (staged tree) when available
            T25 = int(args.duration * 25)
            frames = SY.synth_sync_frames(min(T25, 64), seed=0).to(dev)
            frames = frames.repeat((T25 + frames.shape[0] - 1) // frames.shape[0], 1, 1, 1)[:T25].contiguous()
            sd = SY.synth_motionformer_state_dict(12, seed=0)
            e_sync = enc.SynchformerEncoder.from_state_dict(sd, device=dev)
            got = e_sync.encode(frames)
            out["synchformer"] = {"frames": T25, "segments": got.shape[1] // 8, "ms": timed(lambda: e_sync.encode(frames))}
            try:
                from tools import ref_shims as R
                ref = R.load_motionformer(12)().eval()
                ref.load_state_dict(sd, strict=True)
                ref = ref.to(dev).to(torch.bfloat16)
                S = got.shape[1] // 8
                x = torch.stack([frames[i * 8: i * 8 + 16] for i in range(S)])[None].permute(0, 1, 3, 2, 4, 5)
                with torch.autocast(device_type="cuda", enabled=True, dtype=torch.half):
                    want = ref(x).float().reshape(1, S * 8, -1)
                    out["synchformer"]["ms_reference_eager_fp16_autocast"] = timed(lambda: ref(x), 3)
                out["synchformer"]["rel_l2_vs_reference_autocast"] = float((got - want).norm() / want.norm())
                del ref
            except Exception as e:   # noqa: BLE001 — no staged reference tree on this box
                out["synchformer"]["reference"] = "unavailable: " + repr(e)[:120]
            del e_sync
        torch.cuda.empty_cache()
        out["what"] = ("SigLIP2-base-patch16-512 vision tower + pooling head and CLAP text tower on the engine vs the HF modules in bf16, "
                       "Synchformer visual extractor on the engine vs the reference's own module under fp16 autocast; eager, same GPU, "
                       "seeded random weights")
        return out
    except Exception as e:   # noqa: BLE001
        return {"unavailable": repr(e)}


def reference_gpu_eager(args, c, dev):
    """Informational: the UNMODIFIED reference on this GPU — bf16 weights, eager, torch.autocast("cuda", bf16) —
    ms per Euler step of its own denoise loop, CUDA events.  None when no reference tree is staged."""
    G, ns = staged_reference()
    if G is None:
        return None
    try:
        L, Lv, S = SY.clip_lengths(args.duration)
        sd = SY.synth_state_dict_cuda(SY.dit_param_specs(c), 0, dev, torch.bfloat16)
        ref_model, ref_cfg = G.build_ref_model(ns, args.model, sd, torch.bfloat16, dev)
        del sd
        feats = {k: v.bfloat16().float() for k, v in SY.synth_conditions(c, L, Lv, S).items()}
        ms = G.time_ref_steps(ns, ref_model, ref_cfg, feats, args.duration, args.cfg, dev, n=6)
        del ref_model
        torch.cuda.empty_cache()
        return {"ms_per_denoise_step": ms, "denoise_steps_per_sec": 1e3 / ms,
                "what": "the reference's own denoise_process_with_generator (baseline/_ref), eager PyTorch under CUDA autocast bf16, "
                        "batch_size 1, same GPU, same weights: 2 warm-up + 6 timed Euler steps"}
    except Exception as e:   # noqa: BLE001
        return {"unavailable": repr(e)}


def time_dominant_gemm(eng, c, B2, L, peaks):
    """The single-stream ConvMLP w1/w3 k=3 conv GEMM (+SwiGLU epilogue): 37 % of the step's FLOPs (the three ConvMLP convs
    together are 53 %).  Timed alone, back to back, with CUDA events on the launching stream."""
    lib = eng.lib
    i64, i32, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p
    lib.foley_gemm.argtypes = [vp, i32, i64, i64, i64, i64, i64, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, i64,
                               i64, i64, vp]
    C, Hs = c["hidden_size"], c["mlp_hidden_single"]
    dev = eng.device
    a = torch.randn(B2, L, C, device=dev).bfloat16()
    w = (torch.randn(2 * Hs, 3 * C, device=dev) * 0.02).bfloat16()
    out = torch.empty(B2, L, Hs, device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream(dev)

    # same tile-width rule as Engine::pick_bn
    m_tiles = ((L + 127) // 128) * B2
    bn = 256 if (2 * Hs) % 256 == 0 and m_tiles * (2 * Hs // 256) >= (148 * 3) // 4 else 128

    def launch():
        s = lib.foley_gemm(a.data_ptr(), 0, B2, L, C, C, L * C, w.data_ptr(), 2 * Hs, 3, -1, 1, 1, bn, 1, 0, None,
                           out.data_ptr(), Hs, L * Hs, 0, vp(st.cuda_stream))
        assert s == 0, lib.foley_last_error()

    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    iters = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        launch()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    flops = 2.0 * B2 * L * 3 * C * 2 * Hs
    ach = flops / (us * 1e-6) / 1e12
    # DRAM traffic per launch from the committed `ncu --set full` capture of this kernel; the capture names the source hash of
    # the build it profiled, so a stale number is visible (traffic_build != this build)
    traffic, traffic_build = None, None
    prof = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            traffic, traffic_build = pj.get("dram_bytes_per_launch"), pj.get("src_hash")
        except Exception:
            traffic = None
    this_build = lib.foley_version().decode().rsplit("src:", 1)[-1]
    return {"kernel": f"gemm_tcgen05_kernel<{bn},bf16> single-block ConvMLP w1|w3 conv(k=3)+SwiGLU", "bound": "tensor",
            "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
            "peak_source": peaks["_source"] + " burst", "us_per_launch": us, "flops_per_launch": flops,
            "traffic": traffic, "traffic_build": traffic_build, "this_build": this_build,
            "algorithmic_bytes": 2.0 * (2 * Hs * 3 * C + B2 * L * C + B2 * L * Hs)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="xl", choices=["xl", "xxl", "small", "tiny"])
    ap.add_argument("--batch", type=int, default=1, help="variations per GPU")
    ap.add_argument("--duration", type=float, default=5.0)
    ap.add_argument("--denoise-steps", type=int, default=50)
    ap.add_argument("--cfg", type=float, default=4.5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encoders", action="store_true", help="skip the informational condition-encoder timing")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the informational eager-reference-on-GPU timing")
    ap.add_argument("--extra-configs", type=int, default=1, help="also time BASELINE.json configs 4 and 5 at their per-GPU shapes")
    ap.add_argument("--ref-steps", type=int, default=1, help="--impl reference: complete Euler steps timed per bench step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
