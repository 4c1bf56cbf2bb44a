"""Bring-up / timing of the tcgen05 attention kernel through foley_attention against fp32 torch attention.
    python tools/attn_check.py [--iters 200]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg, rel_l2  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--only", default="", help="run one case only: joint5 | single5 | cross5 | joint30 | single30 (ncu target)")
ap.add_argument("--probe", action="store_true", help="clock64 timeline of CTA 0 (tcgen05 kernel)")
ap.add_argument("--fused", action="store_true", help="also time the operand load with the q/k-norm + RoPE folded in")
ap.add_argument("--variants", action="store_true", help="try the alternative MN-major descriptor stride assignments")
a = ap.parse_args()
E = load_pkg("engine")


def prepared(t, S):
    return dict(t=t, batch_stride=t.stride(0), head_stride=t.stride(1), row_stride=128, rows=S, batch=t.shape[0])


def ref_attn(q, k, v, kv_map=None):
    q, k, v = q.float(), k.float(), v.float()
    if kv_map is not None:
        k, v = k[kv_map.long()], v[kv_map.long()]
    s = torch.einsum("bhqd,bhkd->bhqk", q, k) * 128 ** -0.5
    return torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), v).permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[2], -1)


def run(name, B, H, Sq, Sk, kvB=None, dbg=(0, 0, 0, 0), time_it=True, pattern=None):
    g = torch.Generator(device="cuda").manual_seed(7)
    q = torch.randn(B, H, Sq, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(kvB or B, H, Sk, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(kvB or B, H, Sk, 128, device="cuda", generator=g).bfloat16()
    if pattern == "v_ones":
        v = torch.ones_like(v)
    if pattern == "v_keyidx":    # V[key, d] = key index / 64 + d / 1024: exposes a wrong V layout
        v = (torch.arange(Sk, device="cuda").view(1, 1, Sk, 1) / 64.0 + torch.arange(128, device="cuda").view(1, 1, 1, 128) / 1024.0).expand_as(v).bfloat16().contiguous()
    kv_map = torch.arange(B, device="cuda", dtype=torch.int32) % kvB if kvB else None
    want = ref_attn(q, k, v, kv_map)
    res = {}
    for impl in (0, 1):
        out = torch.zeros(B, Sq, H * 128, device="cuda", dtype=torch.bfloat16)
        E.attention(prepared(q, Sq), prepared(k, Sk), prepared(v, Sk), out, H, kv_batch_map=kv_map, impl=impl, dbg=dbg)
        torch.cuda.synchronize()
        res[impl] = rel_l2(out.float(), want)
        us = float("nan")
        if time_it:   # 20 launches per CUDA graph: device time, not the ctypes call
            st = torch.cuda.Stream()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.stream(st):
                with torch.cuda.graph(gr, stream=st):
                    for _ in range(20):
                        E.attention(prepared(q, Sq), prepared(k, Sk), prepared(v, Sk), out, H, kv_batch_map=kv_map, impl=impl, dbg=dbg,
                                    stream=st.cuda_stream)
                gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(max(a.iters // 20, 1)):
                    gr.replay()
                e1.record(st)
                torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (max(a.iters // 20, 1) * 20)
        print(f"{name:22s} B={B} H={H} Sq={Sq} Sk={Sk} impl={'tcgen05' if impl == 0 else 'mma.sync'} dbg={dbg[:2]} {pattern or ''}: "
              f"rel-L2 vs fp32 {res[impl]:.3e}   {us:7.2f} us", flush=True)
    if a.probe:
        import ctypes
        lib = E.load_library()
        lib.foley_debug_times.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
        for chunk in sorted({0, min(1, (Sk - 1) // 128 if Sk > 320 else 0)}):
            out = torch.zeros(B, Sq, H * 128, device="cuda", dtype=torch.bfloat16)
            for _ in range(3):
                E.attention(prepared(q, Sq), prepared(k, Sk), prepared(v, Sk), out, H, kv_batch_map=kv_map, impl=0, dbg=(chunk + 1, 0, 0, 0))
            torch.cuda.synchronize()
            t = (ctypes.c_uint64 * 16)()
            lib.foley_debug_times(t)
            t = list(t)
            ghz = (t[13] - t[0]) / max(t[14] - t[15], 1)
            names = ["pdl wait done", "loads issued / chunk start", "Q,K landed", "norm + barrier", "S complete", "max + exchange",
                     "P written", "V landed + barrier", "PV complete", "O read", "loop done", "stored", "TMEM released"]
            print(f"      CTA0 timeline (chunk {chunk}; {ghz:.2f} GHz, total {t[14] - t[15]} ns); cumulative ns: " +
                  ", ".join(f"{n} {(t[i + 1] - t[0]) / ghz:.0f}" for i, n in enumerate(names)))
    return res


def rope_table(n):
    k = torch.arange(64, dtype=torch.float32)
    ang = torch.arange(n, dtype=torch.float32)[:, None] * torch.pow(torch.tensor(10000.0), -(2 * k) / 128.0)[None, :]
    return torch.stack([ang.cos(), ang.sin()], -1).contiguous().cuda()


def run_fused(name, kind, Lv, L, B=2, H=11, normed=True):
    """operands read from a projection output [B, S, 3C]; q / k normalised + rotated in the kernel (timing + timeline)"""
    import ctypes
    C, S = H * 128, Lv + L
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(B, S, 3 * C, device="cuda", generator=g).bfloat16()
    w = torch.ones(128, device="cuda").bfloat16()
    rope = rope_table(S)
    rows0 = Lv if Lv else S

    def operand(part, nm):
        return dict(t=qkv, off=part * C, batch_stride=S * 3 * C, head_stride=128, row_stride=3 * C, rows=S, batch=B, rows0=rows0,
                    norm=[(w, rope), (w, rope)] if (nm and normed) else [])
    out = torch.zeros(B, S, C, device="cuda", dtype=torch.bfloat16)
    call = lambda st=None, dbg=(0, 0, 0, 0): E.attention(operand(0, True), operand(1, True), operand(2, False), out, H, norm_kind=kind,
                                                         eps=1e-6, stream=st, dbg=dbg)
    call()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            for _ in range(20):
                call(st.cuda_stream)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            gr.replay()
        e1.record(st)
        torch.cuda.synchronize()
    print(f"{name:22s} fused from [B,S,3C] kind={kind} Lv={Lv} L={L} normed={normed}: {e0.elapsed_time(e1) * 1e3 / 200:7.2f} us", flush=True)
    if a.probe:
        lib = E.load_library()
        lib.foley_debug_times.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
        for _ in range(3):
            call(None, (1, 0, 0, 0))
        torch.cuda.synchronize()
        t = (ctypes.c_uint64 * 16)()
        lib.foley_debug_times(t)
        t = list(t)
        ghz = (t[13] - t[0]) / max(t[14] - t[15], 1)
        names = ["pdl wait done", "chunk start", "Q,K landed", "norm + barrier + S issue", "S complete", "max + exchange",
                 "P written", "V landed + barrier", "PV complete", "O read", "loop done", "stored", "TMEM released"]
        print(f"      CTA0 timeline ({ghz:.2f} GHz, total {t[14] - t[15]} ns); cumulative ns: " +
              ", ".join(f"{n} {(t[i + 1] - t[0]) / ghz:.0f}" for i, n in enumerate(names)))


CASES = {"joint5": ("joint 5s", 2, 11, 290, 290, None), "single5": ("single 5s", 2, 11, 250, 250, None),
         "cross5": ("cross 5s (77 text)", 2, 11, 290, 77, 2), "short": ("short", 1, 2, 40, 16, None),
         "joint30": ("joint 30s xxl", 2, 12, 1740, 1740, None), "single30": ("single 30s xxl", 2, 12, 1500, 1500, None)}
if a.only:
    n_, B_, H_, Sq_, Sk_, kvB_ = CASES[a.only]
    run(n_, B_, H_, Sq_, Sk_, kvB=kvB_, time_it=False)
    sys.exit(0)
if a.fused:
    run_fused("single 5s", 1, 0, 250)
    run_fused("single 5s", 1, 0, 250, normed=False)
    run_fused("joint 5s", 0, 40, 250)
    run_fused("joint 5s", 0, 40, 250, normed=False)
if a.variants:
    for pat in ("v_ones", "v_keyidx"):
        run("joint 5s", 2, 11, 290, 290, time_it=False, pattern=pat)
for key in ("joint5", "single5", "cross5", "short", "joint30", "single30"):
    n_, B_, H_, Sq_, Sk_, kvB_ = CASES[key]
    run(n_, B_, H_, Sq_, Sk_, kvB=kvB_)
