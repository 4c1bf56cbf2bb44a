"""Micro-benchmark / ncu target for single GEMM shapes of the DiT step through foley_gemm.
    python tools/gemm_micro.py --shape w13 --model xl [--bn 128] [--iters 50]
"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synthetic as SY  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="xl")
ap.add_argument("--shape", default="w13")
ap.add_argument("--bn", type=int, default=128)
ap.add_argument("--splits", type=int, default=0)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--all", action="store_true")
ap.add_argument("--pair", type=int, default=-1, help="-1 default policy, 0 single-CTA tiles, 1 CTA-pair (cta_group::2) tiles")
ap.add_argument("--mcast", type=int, default=0, help="weight-tile multicast over this many m-tiles (0 default policy, 1 off, 2, 4)")
ap.add_argument("--sweep", action="store_true", help="every shape x {single, pair} x {full, TMA only, TMA+MMA} at --bn")
ap.add_argument("--dbg", type=int, default=0, help="bring-up modes: 1 = setup only, 2 = TMA only, 3 = TMA+MMA (no epilogue), 4 = MMA only on a re-used "
                                                       "ring, 8 = full kernel + clock64 timeline of CTA 0")
a = ap.parse_args()
lib = ctypes.CDLL(os.environ.get("FOLEY_B200_LIB", os.path.join(ROOT, "comfyui-hunyuanvideo-foley_b200", "libfoley_b200.so")))
lib.foley_last_error.restype = ctypes.c_char_p
i64, i32, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p
if hasattr(lib, "foley_debug_times"):   # (absent from older builds loaded through FOLEY_B200_LIB for A/B runs)
    lib.foley_debug_times.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
lib.foley_gemm.argtypes = [vp, i32, i64, i64, i64, i64, i64, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, i64, i64,
                           i64, vp]
c = SY.model_config(a.model)
C, F, Hs = c["hidden_size"], c["mlp_hidden_triple"], c["mlp_hidden_single"]
B2, L = 2 * a.batch, 250
# name: (K, taps, N, mode, default splits)
SHAPES = {"w13": (C, 3, 2 * Hs, 1, 1), "w2": (Hs, 3, C, 2, 3), "qkv": (C, 1, 3 * C, 0, 1), "fc1": (C, 1, F, 0, 1),
          "fc2": (F, 1, C, 2, 3), "lin1": (C, 3, C, 2, 3), "proj": (C, 1, C, 2, 3),
          "mod": (C, 1, 6 * C * c["depth_single_blocks"], 0, 1)}
dev = "cuda"


def run(name, bn, pair=None, dbg=None):
    pair = a.pair if pair is None else pair
    dbg = a.dbg if dbg is None else dbg
    K, taps, N, mode, sp = SHAPES[name]
    sp = a.splits or sp
    x = torch.randn(B2, L, K, device=dev).bfloat16()
    w = (torch.randn(N, taps * K, device=dev) * 0.02).bfloat16()
    if mode == 2:
        out = torch.empty(sp, B2, L, N, device=dev, dtype=torch.float32)
    else:
        out = torch.empty(B2, L, N // (2 if mode == 1 else 1), device=dev, dtype=torch.bfloat16)
    ldo = out.shape[-1]

    def launch():
        s = lib.foley_gemm(x.data_ptr(), 0, B2, L, K, K, L * K, w.data_ptr(), N, taps, -(taps // 2), 1, sp, bn | (dbg << 16) | ((pair + 1) << 20) | (a.mcast << 22), mode, 0,
                           None, out.data_ptr(), ldo, L * ldo, B2 * L * ldo, None)
        assert s == 0, lib.foley_last_error()

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        launch()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / a.iters
    fl = 2.0 * B2 * L * K * taps * N
    kb = K * taps // 64 // sp
    print(f"{name:5s} K={K}x{taps} N={N} bn={bn} splits={sp} pair={pair} dbg={dbg}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  "
          f"({us / kb:.3f} us per k-block of {kb})", flush=True)
    if dbg == 8:
        t = (ctypes.c_uint64 * 16)()
        lib.foley_debug_times(t)
        t = list(t)
        wall_ns = t[14] - t[15]
        cyc = t[8] - t[0]
        ghz = cyc / max(wall_ns, 1)
        names = ["prologue", "A released (PDL wait)", "first stage landed", "last MMA issued", "accumulator complete",
                 "epilogue done", "CTA joined", "TMEM released"]
        marks = [t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]]
        print(f"      CTA0 timeline: {wall_ns} ns wall, {cyc} cycles ({ghz:.2f} GHz); cumulative ns:")
        for nm, m in zip(names, marks):
            print(f"        {nm:28s} {(m - t[0]) / ghz:8.0f}")
        if t[9]:
            print("      epilogue, first chunk (ns after 'accumulator complete'): " +
                  ", ".join(f"{nm} {(t[i] - t[5]) / ghz:.0f}" for nm, i in (("tcgen05.ld done", 9), ("math + st.shared", 10), ("all chunks done", 13))))


if a.sweep:
    for n in ("w13", "w2", "qkv", "fc1", "fc2", "lin1"):
        for pair in (0, 1):
            for dbg in (0, 2, 3):
                run(n, a.bn, pair, dbg)
elif a.all:
    for n in SHAPES:
        for bn in (128, 256):
            run(n, bn)
else:
    run(a.shape, a.bn)
