"""TEST INFRASTRUCTURE — CPU oracle for the HunyuanVideo-Foley denoise hot path.  NOT a product path:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

A functional (state-dict in, tensors out) restatement in plain torch-CPU ops of what the reference
computes on the path SURVEY.md §8a lists.  Each function cites the reference lines it follows
(paths relative to the reference repository root).

Pinning: tools/make_golden.py runs the *reference's own modules* (imported from /root/reference through
tools/ref_shims.py) on the seeded weights of oracle/weights.py and stores outputs under tests/golden/;
tests/test_oracle_golden.py checks this file against those fixtures (fp32 policy: rel-L2 <= 1e-5).
The reference ships no tests or golden vectors of its own, so that is the only pin.

Two arithmetic policies:
  "fp32"      — everything in float32, the reference's CPU / fp32 path (config #1).
  "cuda_bf16" — the rounding points of the reference's benchmarked path: bf16 weights under
                torch.autocast("cuda", bf16) (utils.py:229-234).  Which tensors are bf16 and which fp32
                follows torch's CUDA autocast policy as recorded on the B200 box in
                profiles/r01_torch_probe.json (layer_norm and nearest-exact interpolate return fp32, so the
                audio residual stream is fp32 from hifi_foley.py:839 on; nn.RMSNorm(eps=None) on bf16
                uses fp32 machine epsilon and rounds once).  All sums are accumulated in fp32.
"""
import math

import torch
import torch.nn.functional as F

F32_EPS = float(torch.finfo(torch.float32).eps)


class Policy:
    def __init__(self, name="fp32"):
        assert name in ("fp32", "cuda_bf16")
        self.name = name
        self.bf16 = name == "cuda_bf16"

    def r(self, x):
        """Round to the model dtype (bf16) and return as float32; identity in the fp32 policy."""
        return x.bfloat16().float() if self.bf16 else x

    def w(self, t):
        """Weights as the model stores them."""
        return t.float().bfloat16().float() if self.bf16 else t.float()


# ------------------------------------------------------------------------------------------------ pieces
def linear(p, x, w, b=None):
    """F.linear under autocast: bf16 inputs, fp32 accumulate, one rounding of (Wx+b)."""
    return p.r(F.linear(p.r(x), p.w(w), p.w(b) if b is not None else None))


def conv1d_cl(p, x, w, b=None, padding=0):
    """ChannelLastConv1d (mlp_layers.py:104-110): x [B,L,C] -> [B,L,C_out], zero padding per sample."""
    y = F.conv1d(p.r(x).transpose(1, 2), p.w(w), p.w(b) if b is not None else None, padding=padding)
    return p.r(y.transpose(1, 2))


def silu(p, x):
    return p.r(F.silu(x))


def layer_norm(x, eps):
    """nn.LayerNorm(elementwise_affine=False); autocast runs it in fp32 and returns fp32."""
    return F.layer_norm(x.float(), (x.shape[-1],), eps=eps)


def modulate_ln(p, x, shift, scale, eps):
    """LN(x) * (1 + scale) + shift (modulate_layers.py:19-30; hifi_foley.py:368,387).
    (1 + scale) is a bf16 op on the bf16 modulation tensor; the rest is fp32; the consumer (a Linear under
    autocast) rounds the result to bf16."""
    return p.r(layer_norm(x, eps) * p.r(1.0 + scale) + shift)


def rms_norm_custom(p, x, w, eps=1e-6):
    """models/nn/norm_layers.py:37-52 (triple blocks): fp32 norm, cast to x dtype, then * weight (bf16 op)."""
    xf = x.float()
    n = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return p.r(p.r(n) * p.w(w))


def rms_norm_torch(p, x, w, eps=None):
    """torch.nn.RMSNorm(dim, eps=None) (hifi_foley.py:360-361): eps = finfo(fp32).eps on this box for bf16
    and fp32 inputs alike (profiles/r01_torch_probe.json), weight applied inside, one rounding."""
    xf = x.float()
    e = F32_EPS if eps is None else eps
    return p.r(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + e) * p.w(w))


def rope_tables(positions, dim=128, theta=10000.0):
    """get_1d_rotary_pos_embed(use_real=True) (posemb_layers.py:123-172): cos/sin [P, dim], each frequency
    repeated for the (2k, 2k+1) pair."""
    pos = positions.float()
    idx = torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2]
    freqs = torch.pow(torch.tensor(theta, dtype=torch.float32).expand_as(idx), -(idx / dim))
    ang = torch.outer(pos, freqs)
    return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)


def apply_rope(p, x, cos, sin):
    """apply_rotary_emb / rotate_half (attn_layers.py:112-148), x [B,S,H,D], tables [S,D]; fp32 then cast."""
    xf = x.float()
    xr = xf.reshape(*xf.shape[:-1], -1, 2)
    rot = torch.stack([-xr[..., 1], xr[..., 0]], dim=-1).flatten(-2)
    return p.r(xf * cos[None, :, None, :] + rot * sin[None, :, None, :])


def sdpa(p, q, k, v):
    """attention(mode='torch') (attn_layers.py:418-422,452-456): q,k,v [B,S,H,D] -> [B,S,H*D]; no mask."""
    q, k, v = (t.transpose(1, 2) for t in (p.r(q), p.r(k), p.r(v)))
    o = F.scaled_dot_product_attention(q, k, v)
    o = p.r(o).transpose(1, 2)
    return o.reshape(o.shape[0], o.shape[1], -1)


def interleaved_positions(L, Lv):
    """RoPE positions of the interleaved audio/visual sequence (hifi_foley.py:35-60, 236-251): audio token i
    sits at 2i; visual token j is up-sampled to L (nearest-exact), interleaved at odd slots, rotated, then
    down-sampled back to Lv (nearest-exact), i.e. it picks slot 2*src(dst(j))+1."""
    a_pos = 2 * torch.arange(L)
    if Lv == L:
        return a_pos, a_pos + 1
    pick = nearest_exact_index(L, Lv)          # down-sample L -> Lv picks up-sampled slot pick[j]
    # the slot must hold visual token j itself (true whenever L >= Lv, the only shapes the nodes produce)
    assert bool((nearest_exact_index(Lv, L)[pick] == torch.arange(Lv)).all()), "unsupported L/Lv ratio"
    return a_pos, 2 * pick + 1


def nearest_exact_index(n_in, n_out):
    """Source index of F.interpolate(mode='nearest-exact') (hifi_foley.py:43,58,760-762).  ATen computes
    floorf((dst + 0.5f) * scale) with scale = (float)n_in / n_out, all in fp32."""
    i = torch.arange(n_out, dtype=torch.float32)
    scale = torch.tensor(float(n_in), dtype=torch.float32) / torch.tensor(float(n_out), dtype=torch.float32)
    return torch.floor((i + 0.5) * scale).long().clamp_(max=n_in - 1)


def timestep_embedding(t, dim=256, max_period=10000):
    """embed_layers.py:76-103."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# ------------------------------------------------------------------------------------------------ blocks
def triple_block(p, sd, pre, cfg, audio, cond, v_cond, vec, rope_av, rope_a_plain, rope_v_plain, rope_text):
    """TwoStreamCABlock.forward (hifi_foley.py:179-333).  audio is the fp32 (cuda_bf16 policy) residual
    stream, v_cond the bf16 one."""
    H = cfg["num_heads"]
    g = lambda n: sd[pre + n]
    B, L, C = audio.shape
    Lv = v_cond.shape[1]
    D = C // H

    def mod(name):
        m = linear(p, silu(p, vec), g(name + ".linear.weight"), g(name + ".linear.bias"))
        return [c[:, None, :] for c in m.chunk(9, dim=-1)]

    a_m = mod("audio_mod")
    v_m = mod("v_cond_mod")

    def qkv(x, m, wname, qn, kn):
        h = modulate_ln(p, x, m[0], m[1], 1e-6)
        o = linear(p, h, g(wname + ".weight"), g(wname + ".bias"))
        q, k, v = o.reshape(B, -1, 3, H, D).unbind(2)
        return rms_norm_custom(p, q, g(qn + ".weight")), rms_norm_custom(p, k, g(kn + ".weight")), v

    aq, ak, av = qkv(audio, a_m, "audio_self_attn_qkv", "audio_self_q_norm", "audio_self_k_norm")
    vq, vk, vv = qkv(v_cond, v_m, "v_cond_attn_qkv", "v_cond_attn_q_norm", "v_cond_attn_k_norm")
    (a_cos, a_sin), (v_cos, v_sin) = rope_av
    aq, ak = apply_rope(p, aq, a_cos, a_sin), apply_rope(p, ak, a_cos, a_sin)
    vq, vk = apply_rope(p, vq, v_cos, v_sin), apply_rope(p, vk, v_cos, v_sin)
    attn = sdpa(p, torch.cat((vq, aq), 1), torch.cat((vk, ak), 1), torch.cat((vv, av), 1))
    v_attn, a_attn = attn[:, :Lv], attn[:, Lv:]
    audio = audio + p.r(linear(p, a_attn, g("audio_self_proj.weight"), g("audio_self_proj.bias")) * a_m[2])
    v_cond = p.r(v_cond + p.r(linear(p, v_attn, g("v_cond_self_proj.weight"), g("v_cond_self_proj.bias")) * v_m[2]))

    # cross attention to text (hifi_foley.py:271-319)
    def cross_q(x, m, wname, qn, rope):
        h = modulate_ln(p, x, m[3], m[4], 1e-6)
        q = linear(p, h, g(wname + ".weight"), g(wname + ".bias")).reshape(B, -1, H, D)
        return apply_rope(p, rms_norm_custom(p, q, g(qn + ".weight")), *rope)

    a_q = cross_q(audio, a_m, "audio_cross_q", "audio_cross_q_norm", rope_a_plain)
    v_q = cross_q(v_cond, v_m, "v_cond_cross_q", "v_cond_cross_q_norm", rope_v_plain)
    kv = linear(p, cond, g("text_cross_kv.weight"), g("text_cross_kv.bias"))
    tk, tv = kv.reshape(B, -1, 2, H, D).unbind(2)
    tk = apply_rope(p, rms_norm_custom(p, tk, g("text_cross_k_norm.weight")), *rope_text)
    cross = sdpa(p, torch.cat((v_q, a_q), 1), tk, tv)
    v_cross, a_cross = cross[:, :Lv], cross[:, Lv:]
    audio = audio + p.r(linear(p, a_cross, g("audio_cross_proj.weight"), g("audio_cross_proj.bias")) * a_m[5])
    v_cond = p.r(v_cond + p.r(linear(p, v_cross, g("v_cond_cross_proj.weight"), g("v_cond_cross_proj.bias")) * v_m[5]))

    # MLPs (hifi_foley.py:321-331; mlp_layers.py:44-51, GELU-tanh)
    def mlp(x, m, name):
        h = modulate_ln(p, x, m[6], m[7], 1e-6)
        h = linear(p, h, g(name + ".fc1.weight"), g(name + ".fc1.bias"))
        h = p.r(F.gelu(h, approximate="tanh"))
        h = linear(p, h, g(name + ".fc2.weight"), g(name + ".fc2.bias"))
        return p.r(h * m[8])

    audio = audio + mlp(audio, a_m, "audio_mlp")
    v_cond = p.r(v_cond + mlp(v_cond, v_m, "v_cond_mlp"))
    return audio, v_cond


def single_block(p, sd, pre, cfg, x, vec_tok, rope_plain):
    """SingleStreamBlock.forward (hifi_foley.py:364-390); vec_tok [B,L,C] per-token condition."""
    H = cfg["num_heads"]
    g = lambda n: sd[pre + n]
    B, L, C = x.shape
    D = C // H
    m = linear(p, p.r(F.silu(vec_tok)), g("modulation.linear.weight"), g("modulation.linear.bias"))
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = m.chunk(6, dim=-1)
    h = modulate_ln(p, x, sh_a, sc_a, 1e-5)
    qkv = linear(p, h, g("linear_qkv.weight"), g("linear_qkv.bias"))
    q, k, v = qkv.reshape(B, L, H, D, 3).unbind(-1)          # "(H D K)" channel order (hifi_foley.py:362)
    q = rms_norm_torch(p, q, g("q_norm.weight"))
    k = rms_norm_torch(p, k, g("k_norm.weight"))
    q, k = apply_rope(p, q, *rope_plain), apply_rope(p, k, *rope_plain)
    o = sdpa(p, q, k, v)
    x = x + p.r(conv1d_cl(p, o, g("linear1.weight"), g("linear1.bias"), padding=1) * g_a)
    h = modulate_ln(p, x, sh_m, sc_m, 1e-5)
    u = p.r(silu(p, conv1d_cl(p, h, g("linear2.w1.weight"), padding=1)) *
            conv1d_cl(p, h, g("linear2.w3.weight"), padding=1))
    x = x + p.r(conv1d_cl(p, u, g("linear2.w2.weight"), padding=1) * g_m)
    return x


# ------------------------------------------------------------------------------------------------ model
def dit_forward(sd, cfg, x, t, cond, clip_feat, sync_feat, policy="fp32", taps=None):
    """HunyuanVideoFoley.forward (hifi_foley.py:707-924) for the shipped configs (interleaved RoPE,
    add_sync_feat_to_audio, no attention mask).  x [B,128,L], t [B], cond [B,T,768], clip_feat [B,Lv,768],
    sync_feat [B,S,768] -> [B,128,L].  `taps` (dict) receives intermediate tensors for per-block tests."""
    p = policy if isinstance(policy, Policy) else Policy(policy)
    g = lambda n: sd[n]
    C, H = cfg["hidden_size"], cfg["num_heads"]
    D = C // H
    B, _, L = x.shape
    Lv, T = clip_feat.shape[1], cond.shape[1]
    x, cond, clip_feat, sync_feat = p.r(x.float()), p.r(cond.float()), p.r(clip_feat.float()), p.r(sync_feat.float())

    # time embedding (embed_layers.py:105-136)
    e = p.r(timestep_embedding(t.float(), cfg["freq_dim"]))
    vec = linear(p, silu(p, linear(p, e, g("time_in.mlp.0.weight"), g("time_in.mlp.0.bias"))),
                 g("time_in.mlp.2.weight"), g("time_in.mlp.2.bias"))
    # sync features (hifi_foley.py:755-762)
    S = sync_feat.shape[1]
    assert S % 8 == 0
    s = p.r(sync_feat.view(B, S // 8, 8, -1) + p.w(g("sync_pos_emb"))).view(B, S, -1)
    s = silu(p, linear(p, s, g("sync_in.0.weight"), g("sync_in.0.bias")))
    s = conv1d_cl(p, p.r(silu(p, conv1d_cl(p, s, g("sync_in.2.w1.weight"))) * conv1d_cl(p, s, g("sync_in.2.w3.weight"))),
                  g("sync_in.2.w2.weight"))
    a_sync = s[:, nearest_exact_index(S, L)]                       # fp32 in the cuda_bf16 policy
    # text / audio / clip embedders (hifi_foley.py:765-770)
    cond = linear(p, silu(p, linear(p, cond, g("cond_in.linear_1.weight"), g("cond_in.linear_1.bias"))),
                  g("cond_in.linear_2.weight"), g("cond_in.linear_2.bias"))
    audio = conv1d_cl(p, x.transpose(1, 2), g("audio_embedder.proj.weight"), g("audio_embedder.proj.bias"))
    v_cond = linear(p, p.r(silu(p, linear(p, clip_feat, g("visual_proj.w1.weight"))) *
                           linear(p, clip_feat, g("visual_proj.w3.weight"))), g("visual_proj.w2.weight"))
    # RoPE tables (hifi_foley.py:797-803, 151-166, 865)
    a_pos, v_pos = interleaved_positions(L, Lv)
    rope_av = (rope_tables(a_pos, D, cfg["rope_theta"]), rope_tables(v_pos, D, cfg["rope_theta"]))
    rope_a = rope_tables(torch.arange(L), D, cfg["rope_theta"])
    rope_v = rope_tables(torch.arange(Lv), D, cfg["rope_theta"])
    rope_t = rope_tables(torch.arange(T), D, cfg["rope_theta"])
    if taps is not None:
        taps.update(vec=vec, a_sync=a_sync, cond=cond, audio0=audio, v_cond0=v_cond)

    audio = audio + a_sync                                         # layer 0 (hifi_foley.py:838-839)
    for i in range(cfg["depth_triple_blocks"]):
        audio, v_cond = triple_block(p, sd, f"triple_blocks.{i}.", cfg, audio, cond, v_cond, vec,
                                     rope_av, rope_a, rope_v, rope_t)
        if taps is not None:
            taps[f"triple{i}.audio"], taps[f"triple{i}.v_cond"] = audio, v_cond
    vec_tok = a_sync + vec[:, None, :]                             # hifi_foley.py:866-867
    xs = audio
    for i in range(cfg["depth_single_blocks"]):
        xs = single_block(p, sd, f"single_blocks.{i}.", cfg, xs, vec_tok, rope_a)
        if taps is not None:
            taps[f"single{i}.x"] = xs
    # final layer: adaLN is a no-op because modulate() drops 3-D shift/scale (modulate_layers.py:20-22)
    y = linear(p, p.r(layer_norm(xs, 1e-6)), g("final_layer.linear.weight"), g("final_layer.linear.bias"))
    return y.transpose(1, 2).contiguous()                          # unpatchify1d (hifi_foley.py:926-936)


def sigma_schedule(n_steps, shift=1.0):
    """FlowMatchDiscreteScheduler.set_timesteps (scheduling_flow_match_discrete.py:131-155)."""
    sig = torch.linspace(1, 0, n_steps + 1)
    if shift != 1.0:
        sig = (shift * sig) / (1 + (shift - 1) * sig)
    return sig


def pad_or_trim(x, T):
    """_pad_or_trim_time (utils.py:104-111)."""
    if x.shape[1] == T:
        return x
    if x.shape[1] > T:
        return x[:, :T]
    return F.pad(x, (0, 0, 0, T - x.shape[1]))


def build_cfg_batch(sd, feats, batch, guidance, T=77):
    """utils.py:158-199: repeat per variation, pad/trim text to the fixed bucket, learned empty features for
    the unconditional half, unconditional rows FIRST.  The bucket is min(77 or 128, caps) and caps includes
    the YAML's text_length = 77 (utils.py:97-102,168-183), so it is always 77."""
    clip = feats["siglip2_feat"].float().repeat(batch, 1, 1)
    sync = feats["syncformer_feat"].float().repeat(batch, 1, 1)
    text = pad_or_trim(feats["text_feat"].float().repeat(batch, 1, 1), T)
    utext = pad_or_trim(feats["uncond_text_feat"].float().repeat(batch, 1, 1), T)
    if guidance > 1.0:
        uclip = sd["empty_clip_feat"].float()[None].expand(batch, clip.shape[1], -1)
        usync = sd["empty_sync_feat"].float()[None].expand(batch, sync.shape[1], -1)
        return torch.cat([uclip, clip]), torch.cat([usync, sync]), torch.cat([utext, text])
    return clip, sync, text


class SolverState:
    """FlowMatchDiscreteScheduler.step for all four solvers (scheduling_flow_match_discrete.py:210-373), including
    its quirk: the multi-stage solvers treat consecutive `step()` calls as *inner* stages — the caller keeps feeding
    the next entry of `timesteps` to the model while `step_index` (hence sigma, sigma_next) only advances after the
    last stage, so N calls perform N/2 (heun-2, midpoint-2) or N/4 (kutta-4) real steps."""

    STAGES = {"euler": 1, "heun-2": 2, "midpoint-2": 2, "kutta-4": 4}

    def __init__(self, solver, sigmas):
        if solver not in self.STAGES:
            raise ValueError(f"Solver {solver} not supported. Supported solvers: {list(self.STAGES)}")
        self.solver, self.sigmas = solver, sigmas.float()
        self.step_index = 0
        self.d, self.dt, self.sample = [], None, None

    def step(self, model_output, sample):
        mo, sample = model_output.float(), sample.float()
        sigma, sigma_next = self.sigmas[self.step_index], self.sigmas[self.step_index + 1]
        stage, n_stages = len(self.d), self.STAGES[self.solver]
        if n_stages == 1:
            derivative, dt, last = mo, sigma_next - sigma, True
        elif stage == 0:
            self.d, self.dt, self.sample = [mo], sigma_next - sigma, sample
            derivative, last = mo, False
            dt = self.dt if self.solver == "heun-2" else self.dt / 2
        elif stage < n_stages - 1:                 # kutta-4 stages 2 and 3
            self.d.append(mo)
            derivative, last = mo, False
            dt = self.dt / 2 if stage == 1 else self.dt
        else:
            if self.solver == "heun-2":
                derivative = 0.5 * (self.d[0] + mo)
            elif self.solver == "midpoint-2":
                derivative = mo
            else:
                derivative = 1 / 6 * self.d[0] + 1 / 3 * self.d[1] + 1 / 3 * self.d[2] + 1 / 6 * mo
            dt, sample, last = self.dt, self.sample, True
            self.d, self.dt, self.sample = [], None, None
        if last:
            self.step_index += 1
        return sample + derivative * dt


def denoise(sd, cfg, feats, latents, n_steps, guidance, policy="fp32", step_callback=None, solver="euler"):
    """denoise_process_with_generator's loop (utils.py:203-247) with the scheduler's step
    (scheduling_flow_match_discrete.py:210-373).  latents [B,128,L] initial noise -> final latents (fp32)."""
    p = policy if isinstance(policy, Policy) else Policy(policy)
    B = latents.shape[0]
    clip, sync, text = build_cfg_batch(sd, feats, B, guidance)
    sig = sigma_schedule(n_steps)
    ts = (sig[:-1] * 1000).float()
    state = SolverState(solver, sig)
    lat = p.r(latents.float())      # noise is drawn in the model dtype (utils.py:151-156)
    for i in range(n_steps):
        x = torch.cat([lat] * 2) if guidance > 1.0 else lat
        t = ts[i].expand(x.shape[0])
        out = dit_forward(sd, cfg, p.r(x), t, text, clip, sync, p)
        if guidance > 1.0:
            u, c = out.chunk(2)
            out = p.r(u + p.r(guidance * p.r(c - u)))              # utils.py:241-243 (model dtype)
        lat = state.step(out, lat)                                 # fp32 update
        if step_callback is not None:
            step_callback(i, lat)
    return lat


# ------------------------------------------------------------------------------------------------ DAC-VAE
def fold_weight_norm(g, v):
    """torch.nn.utils.parametrizations.weight_norm(dim=0): w = g * v / ||v|| over all dims but 0
    (dac_vae/nn/layers.py:9-14).  For ConvTranspose1d dim 0 is C_in."""
    n = v.float().flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return g.float() * v.float() / n


def snake(x, alpha):
    """dac_vae/nn/layers.py:18-24."""
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def dac_decode(sd, z, rates=(8, 5, 4, 3, 2), prefix=""):
    """DAC.decode (dac.py:280-303) for continuous=True: post_quant_conv then Decoder (dac.py:119-149)."""
    g = lambda n: sd[prefix + n].float()
    wn = lambda n: fold_weight_norm(sd[prefix + n + ".parametrizations.weight.original0"],
                                    sd[prefix + n + ".parametrizations.weight.original1"])
    x = F.conv1d(z.float(), g("post_quant_conv.weight"), g("post_quant_conv.bias"))
    x = F.conv1d(x, wn("decoder.model.0"), g("decoder.model.0.bias"), padding=3)
    for i, s in enumerate(rates):
        pre = f"decoder.model.{i + 1}.block."
        x = snake(x, g(pre + "0.alpha"))
        x = F.conv_transpose1d(x, wn(pre + "1"), g(pre + "1.bias"), stride=s, padding=math.ceil(s / 2),
                               output_padding=s % 2)
        for j, dil in enumerate((1, 3, 9)):
            q = pre + f"{j + 2}.block."
            y = snake(x, g(q + "0.alpha"))
            y = F.conv1d(y, wn(q + "1"), g(q + "1.bias"), dilation=dil, padding=3 * dil)
            y = snake(y, g(q + "2.alpha"))
            y = F.conv1d(y, wn(q + "3"), g(q + "3.bias"))
            x = x + y
    n = len(rates) + 1
    x = snake(x, g(f"decoder.model.{n}.alpha"))
    x = F.conv1d(x, wn(f"decoder.model.{n + 1}"), g(f"decoder.model.{n + 1}.bias"), padding=3)
    return torch.tanh(x)
