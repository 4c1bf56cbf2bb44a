"""Stand-alone parity of the tcgen05 / TMEM attention kernel (csrc/attention_tc.cuh) through the C ABI (foley_attention):
against fp32 softmax(QK^T/sqrt(d))V on the same bf16 operands for the shapes of the path (joint 290 keys, single 250,
cross 77 with a condition map, 30 s joint 1740 = the chunked online-softmax mode), and with the q/k RMSNorm + RoPE folded
into the operand load against an fp32 restatement of the reference's modules (attn_layers.py:112-148 apply_rotary_emb,
norm_layers.py:49-51 RMSNorm, hifi_foley.py:376-381 nn.RMSNorm) on identical inputs.  Tolerance: P and the output are
rounded to bf16 (2^-9 relative each), measured 2.2e-3; asserted 4e-3."""
import pytest
import torch

from conftest import load_pkg, rel_l2

pytestmark = pytest.mark.gpu
TOL = 4e-3


def _prepared(t):
    return dict(t=t, batch_stride=t.stride(0), head_stride=t.stride(1), row_stride=128, rows=t.shape[2], batch=t.shape[0])


def _ref_attn(q, k, v, kv_map=None):
    q, k, v = q.float(), k.float(), v.float()
    if kv_map is not None:
        k, v = k[kv_map.long()], v[kv_map.long()]
    s = torch.einsum("bhqd,bhkd->bhqk", q, k) * 128 ** -0.5
    return torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), v).permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[2], -1)


@pytest.mark.parametrize("name,B,H,Sq,Sk,kvB", [
    ("joint_5s", 2, 11, 290, 290, None), ("single_5s", 2, 11, 250, 250, None), ("cross_5s", 4, 11, 290, 77, 2),
    ("one_chunk_max", 1, 3, 320, 320, None), ("first_long", 1, 3, 321, 321, None), ("tiny", 1, 2, 40, 16, None),
    ("ragged", 2, 2, 129, 33, None), ("joint_30s", 1, 4, 1740, 1740, None), ("single_30s_b2", 2, 2, 1500, 1500, None)])
def test_attention_vs_fp32(name, B, H, Sq, Sk, kvB):
    E = load_pkg("engine")
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 1000)
    q = torch.randn(B, H, Sq, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(kvB or B, H, Sk, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(kvB or B, H, Sk, 128, device="cuda", generator=g).bfloat16()
    kv_map = (torch.arange(B, device="cuda", dtype=torch.int32) % kvB) if kvB else None
    want = _ref_attn(q, k, v, kv_map)
    out = torch.full((B, Sq, H * 128), float("nan"), device="cuda", dtype=torch.bfloat16)
    E.attention(_prepared(q), _prepared(k), _prepared(v), out, H, kv_batch_map=kv_map)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    err = rel_l2(out.float(), want)
    print(f"[{name}] tcgen05 attention vs fp32: {err:.3e}")
    assert err <= TOL
    if Sk <= 320:   # the round-1 mma.sync kernel on the same operands: same rounding points, same tolerance
        out1 = torch.zeros_like(out)
        E.attention(_prepared(q), _prepared(k), _prepared(v), out1, H, kv_batch_map=kv_map, impl=1)
        torch.cuda.synchronize()
        assert rel_l2(out1.float(), want) <= TOL


def _rope_table(pos, theta=10000.0):
    """[n, 64, 2] fp32 (cos, sin), posemb_layers.py:160-168."""
    k = torch.arange(64, dtype=torch.float32)
    freq = torch.pow(torch.tensor(theta, dtype=torch.float32), -(2 * k) / 128.0)
    ang = pos.float()[:, None] * freq[None, :]
    return torch.stack([ang.cos(), ang.sin()], -1).contiguous()


def _norm_rope_ref(x, w, rope, kind, eps):
    """x [B, S, H, 128] bf16-valued fp32 -> normalised + rotated, rounded to bf16 where the reference rounds."""
    bf = lambda t: t.bfloat16().float()
    rstd = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    y = bf(bf(x * rstd) * w.float()) if kind == 0 else bf(x * rstd * w.float())
    y0, y1 = y[..., 0::2], y[..., 1::2]
    c, s = rope[None, :, None, :, 0], rope[None, :, None, :, 1]
    o = torch.stack([y0 * c - y1 * s, y1 * c + y0 * s], -1).flatten(-2)
    return bf(o)


@pytest.mark.parametrize("kind,Lv,L", [(0, 40, 250), (1, 0, 250), (0, 8, 50)])
def test_attention_fused_norm_rope(kind, Lv, L):
    """Operands read straight from a projection output [B, S, 3*C] (row stride 3C, head stride 128); q / k rows are
    RMS-normalised and rotated inside the kernel: two norm groups (visual rows first, then audio rows) with different
    weights and tables for the joint attention of a triple block (kind 0), one group for a single block (kind 1)."""
    E = load_pkg("engine")
    B, H = 2, 11
    C, S = H * 128, Lv + L
    eps = 1e-6 if kind == 0 else torch.finfo(torch.float32).eps
    g = torch.Generator(device="cuda").manual_seed(11 + kind)
    qkv = (torch.randn(B, S, 3 * C, device="cuda", generator=g) * 1.5).bfloat16()
    wq = [(1 + 0.2 * torch.randn(128, device="cuda", generator=g)).bfloat16() for _ in range(2)]
    wk = [(1 + 0.2 * torch.randn(128, device="cuda", generator=g)).bfloat16() for _ in range(2)]
    if Lv:   # interleaved positions of the joint attention (hifi_foley.py:151-166): audio 2i, visual 2*pick+1
        pos = [2 * torch.linspace(0, L - 1, Lv).round() + 1, 2 * torch.arange(L)]
        rows0 = Lv
    else:
        pos = [torch.arange(L), torch.arange(L)]
        rows0 = L
    rope = [_rope_table(p).cuda() for p in pos]

    def operand(part, norm_w):
        return dict(t=qkv, off=part * C, batch_stride=S * 3 * C, head_stride=128, row_stride=3 * C, rows=S, batch=B, rows0=rows0,
                    norm=[(norm_w[0], rope[0]), (norm_w[1], rope[1])] if norm_w else [])
    out = torch.zeros(B, S, C, device="cuda", dtype=torch.bfloat16)
    E.attention(operand(0, wq), operand(1, wk), operand(2, None), out, H, norm_kind=kind, eps=eps)
    torch.cuda.synchronize()

    x = qkv.float().view(B, S, 3, H, 128)

    def prep(part, ws):
        xs = x[:, :, part]
        if ws is None:
            return xs
        lo = _norm_rope_ref(xs[:, :rows0], ws[0], rope[0], kind, eps)
        if rows0 == S:
            return lo
        return torch.cat([lo, _norm_rope_ref(xs[:, rows0:], ws[1], rope[1], kind, eps)], 1)
    qn, kn, vn = prep(0, wq), prep(1, wk), prep(2, None)
    want = _ref_attn(qn.permute(0, 2, 1, 3), kn.permute(0, 2, 1, 3), vn.permute(0, 2, 1, 3))
    err = rel_l2(out.float(), want)
    print(f"[fused kind={kind} Lv={Lv} L={L}] vs fp32 restatement: {err:.3e}")
    assert err <= TOL


@pytest.mark.parametrize("env", [{"FOLEY_ATT_TC": "1"}, {"FOLEY_ATT_TC": "1", "FOLEY_ATT_FUSED": "1"}])
def test_engine_forward_on_the_tcgen05_attention(env, monkeypatch):
    """The whole DiT forward with every attention call on the tcgen05 kernel (prepared operands, and with the
    q/k-norm + RoPE folded into its operand load: no qk_norm_rope_kernel launch) against the oracle, same tolerance as
    the default path (tests/test_gpu_dit.py), and against the default path itself."""
    from oracle import foley_oracle as O
    from test_gpu_dit import _inputs, make_engine
    eng0, c, sd = make_engine("small")
    x, t, cond, clip, sync = _inputs(c, 2, 125, 20, 48)
    eng0.set_conditions(clip.cuda(), sync.cuda(), cond.cuda(), L=125, batch=1)
    base = eng0.dit_forward(x.cuda(), t).cpu()
    n0 = eng0.launch_count()
    for k, v in env.items():
        monkeypatch.setenv(k, v)          # read at engine creation
    eng, _, _ = make_engine("small", sd=sd)
    eng.set_conditions(clip.cuda(), sync.cuda(), cond.cuda(), L=125, batch=1)
    out = eng.dit_forward(x.cuda(), t).cpu()
    assert eng.debug_flags()[0] == 0
    want16 = O.dit_forward(sd, c, x, t, cond, clip, sync, policy="cuda_bf16")
    r16, rb = rel_l2(out, want16), rel_l2(out, base)
    print(f"\n[{env}] engine vs oracle(cuda_bf16) {r16:.3e} | vs the default attention path {rb:.3e} | launches {eng.launch_count()} vs {n0}")
    assert r16 <= 4e-3 and rb <= 4e-3
    if "FOLEY_ATT_FUSED" in env:
        assert eng.launch_count() < n0       # the q/k-norm launches are gone
