"""Per-kernel parity of the row-wise kernels of the DiT step (csrc/rowwise.cuh), each called alone through the C ABI and
compared with an fp32 torch restatement of the reference arithmetic on identical bf16 inputs:
  foley_qk_norm_rope    attn_layers.py:112-148 (apply_rotary_emb), norm_layers.py:49-51 (RMSNorm), hifi_foley.py:376-381
  foley_combine_ln_mod  hifi_foley.py:216-331 / 364-390 (gate * proj + residual), modulate_layers.py (LayerNorm, modulate)
  foley_cfg_euler       utils.py:241-243 (CFG combine), scheduling_flow_match_discrete.py:262-297 (Euler step)
Tolerances: every value passes through the same bf16 rounding points; the only differences are rsqrtf vs torch.rsqrt and
fp32 summation order, which flip a bf16 rounding (one ulp = 2^-8 relative) on < 2 % of the elements => rel-L2 <= 1e-3.
The CFG / Euler arithmetic has no such freedom: bit-exact."""
import ctypes

import pytest
import torch

from conftest import load_pkg, rel_l2

pytestmark = pytest.mark.gpu
vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


def bf(t):
    return t.bfloat16().float()


def _lib():
    E = load_pkg("engine")
    lib = E.load_library()
    lib.foley_qk_norm_rope.argtypes = [vp, vp, i32, vp, i64, i32, i32, i32, i32, i32, f32, ctypes.POINTER(vp), vp, vp,
                                       ctypes.POINTER(vp), i32, i32, vp]
    lib.foley_combine_ln_mod.argtypes = [vp, i32, vp, vp, i64, i64, i32, i32, i32, vp, vp, i32, vp, f32, i32, i32, i32, vp]
    lib.foley_cfg_euler.argtypes = [vp, vp, vp, i32, i32, i32, i32, f32, vp, vp, vp]
    return lib, E


def _rope_tables(L, theta=10000.0):
    k = torch.arange(64, dtype=torch.float32)
    ang = torch.arange(L, dtype=torch.float32)[:, None] * torch.pow(torch.tensor(theta), -(2 * k) / 128.0)[None, :]
    return ang.cos().repeat_interleave(2, 1).contiguous(), ang.sin().repeat_interleave(2, 1).contiguous()   # [L,128] pairs duplicated


def _norm_rope_ref(x, w, cos, sin, kind, eps):
    """x [B, L, H, 128] fp32 (bf16-valued)."""
    rstd = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    y = bf(bf(x * rstd) * w) if kind == 0 else bf((x * rstd) * w)
    y0, y1 = y[..., 0::2], y[..., 1::2]
    c, s = cos[None, :, None, 0::2], sin[None, :, None, 0::2]
    return bf(torch.stack([y0 * c - y1 * s, y1 * c + y0 * s], -1).flatten(-2))


@pytest.mark.parametrize("kind,n_parts,splits", [(0, 3, 0), (1, 3, 0), (0, 1, 0), (0, 3, 2), (1, 3, 3)])
def test_qk_norm_rope(kind, n_parts, splits):
    lib, E = _lib()
    B, L, H, Lv = 2, 61, 3, 7
    S_total, ld = L + Lv, n_parts * H * 128
    eps = 1e-6 if kind == 0 else float(torch.finfo(torch.float32).eps)
    g = torch.Generator().manual_seed(3 + kind + splits)
    ws = [(1 + 0.3 * torch.randn(128, generator=g)).bfloat16() for _ in range(2)]
    cos, sin = _rope_tables(L)
    if splits:
        parts = torch.randn(splits, B * L, ld, generator=g)
        bias = (0.1 * torch.randn(ld, generator=g)).bfloat16()
        acc = parts[0].clone()
        for s in range(1, splits):
            acc = acc + parts[s]
        src = bf(acc + bias.float())
        src_dev, part_dev, bias_dev = None, parts.cuda(), bias.cuda()
    else:
        src = bf(torch.randn(B * L, ld, generator=g) * 1.7)
        src_dev, part_dev, bias_dev = src.bfloat16().cuda(), None, None
    dsts = [torch.zeros(B, H, S_total, 128, dtype=torch.bfloat16, device="cuda") for _ in range(n_parts)]
    w_dev = [w.cuda() for w in ws]
    normed_parts = 2 if n_parts == 3 else 1                                # q, k normalised + rotated; v copied
    norm_ptrs = (vp * 3)(*[w_dev[p].data_ptr() if p < normed_parts else None for p in range(3)])
    dst_ptrs = (vp * 3)(*[d.data_ptr() for d in dsts] + [None] * (3 - n_parts))
    cos_d, sin_d = cos.cuda(), sin.cuda()
    st = lib.foley_qk_norm_rope(src_dev.data_ptr() if src_dev is not None else None, part_dev.data_ptr() if part_dev is not None else None,
                                splits, bias_dev.data_ptr() if bias_dev is not None else None, ld, n_parts, B, L, H, kind, eps,
                                norm_ptrs, cos_d.data_ptr(), sin_d.data_ptr(), dst_ptrs, S_total, Lv, None)
    assert st == 0, lib.foley_last_error()
    torch.cuda.synchronize()
    x = src.view(B, L, n_parts, H, 128)
    for p in range(n_parts):
        normed = norm_ptrs[p] is not None
        want = _norm_rope_ref(x[:, :, p], ws[p].float(), cos, sin, kind, eps) if normed else x[:, :, p]
        got = dsts[p].float().cpu()[:, :, Lv:].permute(0, 2, 1, 3)           # [B, L, H, 128]
        err = rel_l2(got, want)
        flips = (got != want).float().mean().item()
        print(f"[kind {kind} parts {n_parts} splits {splits}] part {p} normed={normed}: rel-L2 {err:.2e}, {100 * flips:.2f} % of elements differ")
        assert err <= (1e-3 if normed else 0.0) and flips <= 0.02
        assert (dsts[p][:, :, :Lv] == 0).all()                               # rows before seq_offset untouched


@pytest.mark.parametrize("case", ["gate_mod_splits", "plain_ln_round_x", "x_init_no_partials", "per_token_mod"])
def test_combine_ln_mod(case):
    lib, E = _lib()
    B, L, C = 2, 37, 384
    g = torch.Generator().manual_seed(hash(case) % 100)
    splits = {"gate_mod_splits": 4, "plain_ln_round_x": 1, "x_init_no_partials": 0, "per_token_mod": 2}[case]
    per_tok = case == "per_token_mod"
    n_chunks = 6
    mod = (0.5 * torch.randn(B, L if per_tok else 1, n_chunks * C, generator=g)).bfloat16()
    gate_c, shift_c, scale_c = (2, 3, 4) if case in ("gate_mod_splits", "per_token_mod") else (-1, -1, -1)
    if case == "x_init_no_partials":
        shift_c, scale_c = 0, 1
    x0 = torch.randn(B * L, C, generator=g) * 2
    round_x = int(case == "plain_ln_round_x")
    if round_x:
        x0 = bf(x0)
    x_init = (torch.randn(B * L, C, generator=g)).bfloat16() if case == "x_init_no_partials" else None
    parts = torch.randn(splits, B * L, C, generator=g) if splits else None
    bias = (0.2 * torch.randn(C, generator=g)).bfloat16() if splits else None
    eps = 1e-6
    x_dev = x0.clone().cuda()
    h_dev = torch.zeros(B * L, C, dtype=torch.bfloat16, device="cuda")
    mod_dev = mod.cuda()
    hold = [t.cuda() if t is not None else None for t in (parts, bias, x_init)]
    st = lib.foley_combine_ln_mod(hold[0].data_ptr() if parts is not None else None, splits, hold[1].data_ptr() if bias is not None else None,
                                  mod_dev.data_ptr(), mod.stride(0), mod.stride(1) if per_tok else 0, gate_c, shift_c, scale_c,
                                  x_dev.data_ptr(), hold[2].data_ptr() if x_init is not None else None, round_x, h_dev.data_ptr(), eps,
                                  B, L, C, None)
    assert st == 0, lib.foley_last_error()
    torch.cuda.synchronize()
    # fp32 restatement
    m = mod.float().expand(B, L, n_chunks * C).reshape(B * L, n_chunks, C)
    x = x_init.float() if x_init is not None else x0.clone()
    if parts is not None:
        acc = parts[0].clone()
        for s in range(1, splits):
            acc = acc + parts[s]
        y = bf(acc + bias.float())
        if gate_c >= 0:
            y = bf(y * m[:, gate_c])
        x = x + y
        if round_x:
            x = bf(x)
    mean = x.mean(-1, keepdim=True)
    xc = x - mean
    o = xc * torch.rsqrt(xc.pow(2).mean(-1, keepdim=True) + eps)
    if shift_c >= 0:
        o = o * bf(1.0 + m[:, scale_c]) + m[:, shift_c]
    want_h = bf(o)
    got_x, got_h = x_dev.cpu(), h_dev.float().cpu()
    if parts is not None or x_init is not None:
        assert torch.allclose(got_x, x, rtol=1e-6, atol=1e-6), "residual stream"
    else:
        assert torch.equal(got_x, x0)
    err = rel_l2(got_h, want_h)
    flips = (got_h != want_h).float().mean().item()
    print(f"[{case}] h rel-L2 {err:.2e}, {100 * flips:.2f} % of elements differ")
    assert err <= 1e-3 and flips <= 0.02


@pytest.mark.parametrize("n_cond,guidance", [(2, 4.5), (1, 1.0)])
def test_cfg_euler_bit_exact(n_cond, guidance):
    lib, E = _lib()
    B, ch, L = 3, 128, 77
    g = torch.Generator().manual_seed(9)
    y = torch.randn(n_cond * B, L, ch, generator=g).bfloat16()
    lat = torch.randn(B, ch, L, generator=g)
    sig = torch.tensor([1.0, 0.93, 0.81, 0.0])
    step = 1
    lat_dev, y_dev = lat.clone().cuda(), y.cuda()
    xn = torch.zeros(n_cond * B, L, ch, dtype=torch.bfloat16, device="cuda")
    sig_dev, step_dev = sig.cuda(), torch.tensor([step], dtype=torch.int32, device="cuda")
    st = lib.foley_cfg_euler(y_dev.data_ptr(), lat_dev.data_ptr(), xn.data_ptr(), B, n_cond, ch, L, guidance, sig_dev.data_ptr(),
                             step_dev.data_ptr(), None)
    assert st == 0, lib.foley_last_error()
    torch.cuda.synchronize()
    yf = y.float()
    if n_cond == 2:
        u, t = yf[:B], yf[B:]
        v = bf(u + bf(torch.tensor(guidance) * bf(t - u)))      # bf16 arithmetic of the reference's CFG line
    else:
        v = yf
    dt = sig[step + 1] - sig[step]
    want = lat + (v.permute(0, 2, 1) * dt)                        # fp32 multiply, then add (scheduler.step)
    assert torch.equal(lat_dev.cpu(), want)
    assert torch.equal(xn.float().cpu(), bf(want).permute(0, 2, 1).repeat(n_cond, 1, 1))
