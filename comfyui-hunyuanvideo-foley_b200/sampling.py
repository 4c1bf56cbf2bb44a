"""Host side of the denoise hot path: mirrors the reference's pipeline helpers (reference utils.py:104-258)
name for name, with the loop body and the decoder delegated to the CUDA engine.

  prepare_latents_with_generator  utils.py:114-121   noise draw on the CPU generator in the model dtype
  denoise_process_with_generator  utils.py:125-258   CFG batch build -> engine.denoise -> engine.dac_decode
"""
import torch
import torch.nn.functional as F

try:  # ComfyUI progress bar when running inside ComfyUI (utils.py:201,247)
    from comfy.utils import ProgressBar
except Exception:  # pragma: no cover - outside ComfyUI
    class ProgressBar:
        def __init__(self, total):
            self.total, self.current = total, 0

        def update(self, n):
            self.current += n


SOLVERS = ("euler", "heun-2", "midpoint-2", "kutta-4")


def sigma_schedule(num_inference_steps, shift=1.0):
    """FlowMatchDiscreteScheduler.set_timesteps (scheduling_flow_match_discrete.py:131-155), reverse=True."""
    sigmas = torch.linspace(1, 0, num_inference_steps + 1)
    if shift != 1.0:
        sigmas = (shift * sigmas) / (1 + (shift - 1) * sigmas)
    return sigmas


class FlowMatchSolverState:
    """The reference scheduler's `step` for its four solvers (scheduling_flow_match_discrete.py:210-373), quirk
    included: heun-2 / midpoint-2 / kutta-4 treat consecutive calls as inner stages, `step_index` (so sigma and
    sigma_next) only advances after the last stage, while the caller keeps feeding the next `timesteps` entry to the
    model.  Euler runs inside the engine (foley_denoise); the multi-stage solvers use this host state machine around
    foley_dit_forward."""

    STAGES = {"euler": 1, "heun-2": 2, "midpoint-2": 2, "kutta-4": 4}

    def __init__(self, solver, sigmas):
        self.solver, self.sigmas = solver, sigmas.float()
        self.step_index = 0
        self.derivatives, self.dt, self.sample = [], None, None

    def step(self, model_output, sample):
        mo, sample = model_output.float(), sample.float()
        sigma, sigma_next = float(self.sigmas[self.step_index]), float(self.sigmas[self.step_index + 1])
        dt_full = torch.tensor(sigma_next, dtype=torch.float32) - torch.tensor(sigma, dtype=torch.float32)
        stage, n_stages = len(self.derivatives), self.STAGES[self.solver]
        if n_stages == 1:
            derivative, dt, last = mo, dt_full, True
        elif stage == 0:
            self.derivatives, self.dt, self.sample = [mo], dt_full, sample
            derivative, last = mo, False
            dt = self.dt if self.solver == "heun-2" else self.dt / 2
        elif stage < n_stages - 1:
            self.derivatives.append(mo)
            derivative, last = mo, False
            dt = self.dt / 2 if stage == 1 else self.dt
        else:
            d = self.derivatives
            if self.solver == "heun-2":
                derivative = 0.5 * (d[0] + mo)
            elif self.solver == "midpoint-2":
                derivative = mo
            else:
                derivative = 1 / 6 * d[0] + 1 / 3 * d[1] + 1 / 3 * d[2] + 1 / 6 * mo
            dt, sample, last = self.dt, self.sample, True
            self.derivatives, self.dt, self.sample = [], None, None
        if last:
            self.step_index += 1
        return sample + derivative * dt.to(sample.device)


def _denoise_multistage(engine, latents, sigmas, guidance_scale, solver, n_cond, progress):
    """utils.py:203-247 for the multi-stage solvers: one foley_dit_forward per `timesteps` entry."""
    timesteps = (sigmas[:-1] * 1000).to(torch.float32)
    state = FlowMatchSolverState(solver, sigmas)
    lat = latents.float()
    for i in range(timesteps.numel()):
        x = torch.cat([lat] * 2) if n_cond == 2 else lat
        out = engine.dit_forward(x, float(timesteps[i]))          # fp32 tensor holding the bf16 model output
        if n_cond == 2:
            u, c = out.bfloat16().chunk(2)
            out = (u + guidance_scale * (c - u)).float()          # bf16 arithmetic, utils.py:241-243
        lat = state.step(out, lat)
        if progress is not None:
            progress(i + 1)
    return lat


def _pad_or_trim_time(x, T_fixed):
    """utils.py:104-111."""
    T_cur = x.shape[1]
    if T_cur == T_fixed:
        return x
    if T_cur > T_fixed:
        return x[:, :T_fixed, :]
    return F.pad(x, (0, 0, 0, T_fixed - T_cur))


def _caps(model_dict, cfg):
    """utils.py:97-102."""
    tokmax = int(getattr(getattr(model_dict, "clap_tokenizer", None), "model_max_length", 10 ** 9) or 10 ** 9)
    posmax = int(getattr(getattr(getattr(model_dict, "clap_model", None), "config", None),
                         "max_position_embeddings", 10 ** 9) or 10 ** 9)
    cfgmax = int(cfg.model_config.model_kwargs.get("text_length", 10 ** 9))
    return min(tokmax, posmax, cfgmax)


def to_host(t):
    """Device -> host through a pinned staging buffer (a pageable `.cpu()` of the 0.96 MB waveform of a 5 s clip takes
    0.5 ms, this 0.05 ms).  Returns a pinned CPU tensor."""
    if not t.is_cuda:
        return t
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out


def prepare_latents_with_generator(scheduler, batch_size, num_channels_latents, length, dtype, device, generator=None):
    """utils.py:114-121 + diffusers.randn_tensor: with a CPU generator the draw happens on the CPU in the
    target dtype and is then moved, which is what makes seeds reproducible across devices.  Multi-GPU shards
    slice this one host draw."""
    shape = (batch_size, num_channels_latents, int(length))
    gen_device = generator.device if generator is not None else torch.device("cpu")
    latents = torch.randn(shape, generator=generator, device=gen_device, dtype=dtype).to(device)
    if scheduler is not None and hasattr(scheduler, "init_noise_sigma"):
        latents = latents * scheduler.init_noise_sigma
    return latents


def denoise_process_with_generator(visual_feats, text_feats, audio_len_in_s, model_dict, cfg, guidance_scale,
                                   num_inference_steps, batch_size, sampler, generator=None, batch_slice=None,
                                   decode=True):
    """Drop-in for reference utils.py:125-258.  `model_dict.foley_model` is a FoleyModel (nodes.py),
    `model_dict.dac_model` a FoleyDAC.  `batch_slice=(lo, hi)` keeps only variations [lo, hi) of the one host
    noise draw — the multi-GPU sharding hook; the reference has no equivalent."""
    if sampler not in SOLVERS:
        raise ValueError(f"Solver {sampler} not supported. Supported solvers: {list(SOLVERS)}")
    foley_model = model_dict.foley_model
    engine = foley_model.engine
    device = model_dict.device
    target_dtype = foley_model.dtype
    kw = cfg.model_config.model_kwargs
    sigmas = sigma_schedule(num_inference_steps, cfg.diffusion_config.sample_flow_shift)

    L = int(audio_len_in_s * kw.audio_frame_rate)
    local_batch = batch_size if batch_slice is None else len(range(batch_size)[batch_slice[0]:batch_slice[1]])

    # text bucket (utils.py:164-188): 77 normally, 128 for long prompts, capped by tokenizer / model / YAML
    T_cur_len = int(text_feats["text_feat"].shape[1])
    cap = _caps(model_dict, cfg)
    T_fixed = min(77, cap) if T_cur_len <= 77 else min(128, cap)
    if not hasattr(foley_model, "_text_len_fixed"):
        foley_model._text_len_fixed = T_fixed
    else:
        foley_model._text_len_fixed = max(foley_model._text_len_fixed, T_fixed)
    T_fixed = foley_model._text_len_fixed

    dt = target_dtype
    clip = visual_feats["siglip2_feat"].to(device, dt)[:1]
    sync = visual_feats["syncformer_feat"].to(device, dt)[:1]
    text = _pad_or_trim_time(text_feats["text_feat"].to(device, dt)[:1], T_fixed)
    utext = _pad_or_trim_time(text_feats["uncond_text_feat"].to(device, dt)[:1], T_fixed)
    if guidance_scale > 1.0:   # unconditional rows FIRST (utils.py:192-199)
        uclip = foley_model.get_empty_clip_sequence(bs=1, len=clip.shape[1]).to(device, dt)
        usync = foley_model.get_empty_sync_sequence(bs=1, len=sync.shape[1]).to(device, dt)
        clip, sync, text = torch.cat([uclip, clip]), torch.cat([usync, sync]), torch.cat([utext, text])

    # the engine shares the (identical, `.repeat`-ed in the reference) condition rows between variations
    from . import torch_ops as ops   # registers torch.ops.foley_b200.* (the C ABI surfaced as torch ops)
    ops.set_conditions(engine, clip, sync, text, L, local_batch)
    # the ONE host noise draw for the global batch (reference utils.py:155-157), drawn while the device works on the
    # conditions enqueued above (the draw does not depend on them: same values as drawing first)
    latents = prepare_latents_with_generator(None, batch_size, kw.audio_vae_latent_dim, L, target_dtype, "cpu",
                                             generator)
    if batch_slice is not None:
        latents = latents[batch_slice[0]:batch_slice[1]]
    pbar = ProgressBar(num_inference_steps)
    progress = (lambda step: pbar.update(1)) if model_dict.get("report_progress", True) else None
    with torch.inference_mode():
        if model_dict.get("host_solver", False) and sampler != "euler":
            # the reference scheduler's stage machine on the host around foley_dit_forward (cross-check path)
            latents = _denoise_multistage(engine, latents.to(device), sigmas, guidance_scale, sampler,
                                          2 if guidance_scale > 1.0 else 1, progress)
        else:   # whole loop inside the engine, one CUDA graph replayed per model call, for all four solvers
            latents = ops.denoise(engine, latents.to(device, torch.float32), sigmas, guidance_scale, solver=sampler,
                                  progress=progress)
        if not decode:
            return latents, model_dict.dac_model.sample_rate if "dac_model" in model_dict else 48000
        audio = model_dict.dac_model.decode(latents)
    # the reference slices dim 1 (the channel dim, size 1) here, a no-op kept for shape parity (utils.py:257)
    audio = audio[:, :int(audio_len_in_s * model_dict.dac_model.sample_rate)]
    return audio, model_dict.dac_model.sample_rate
