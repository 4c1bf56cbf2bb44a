"""CTA-pair (cta_group::2) GEMM tiles and weight-tile multicast clusters (2 / 4 m-tiles) against plain single-CTA tiles
through foley_gemm: same operands, every epilogue mode — the results must be bit-identical (same MMA shapes per row,
same accumulation order), then timings of all of them.
    python tools/pair_check.py [--iters 200]
"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
lib = ctypes.CDLL(os.path.join(ROOT, "comfyui-hunyuanvideo-foley_b200", "libfoley_b200.so"))
lib.foley_last_error.restype = ctypes.c_char_p
i64, i32, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p
lib.foley_gemm.argtypes = [vp, i32, i64, i64, i64, i64, i64, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, i64, i64,
                           i64, vp]
C, F, Hs = 1408, 5632, 3840
B2, L = 2 * a.batch, 250
SHAPES = {"w13": (C, 3, 2 * Hs, 1, 1), "w2": (Hs, 3, C, 2, 3), "qkv": (C, 1, 3 * C, 0, 1), "fc1": (C, 1, F, 0, 1),
          "fc2": (F, 1, C, 2, 3), "lin1": (C, 3, C, 2, 6), "proj": (C, 1, C, 2, 4)}
ok_all = True
for name, (K, taps, N, mode, sp) in SHAPES.items():
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B2, L, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, taps * K, device="cuda", generator=g) * 0.02).bfloat16()
    res = {}
    for bn in (256, 128):
        for pair, mc in ((0, 1), (0, 2), (0, 4), (1, 1)):
            if mode == 2:
                out = torch.zeros(sp, B2, L, N, device="cuda", dtype=torch.float32)
            else:
                out = torch.zeros(B2, L, N // (2 if mode == 1 else 1), device="cuda", dtype=torch.bfloat16)
            ldo = out.shape[-1]

            def launch():
                s = lib.foley_gemm(x.data_ptr(), 0, B2, L, K, K, L * K, w.data_ptr(), N, taps, -(taps // 2), 1, sp,
                                   bn | ((pair + 1) << 20) | (mc << 22), mode, 0, None, out.data_ptr(), ldo, L * ldo, B2 * L * ldo, None)
                assert s == 0, lib.foley_last_error()
            launch()
            torch.cuda.synchronize()
            res[(bn, pair, mc)] = out.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                launch()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.iters
            fl = 2.0 * B2 * L * K * taps * N
            same = torch.equal(res[(bn, pair, mc)], res[(bn, 0, 1)])
            ok_all &= same
            print(f"{name:5s} K={K}x{taps} N={N} splits={sp} bn={bn} pair={pair} mcast={mc}: {us:7.2f} us {fl / us / 1e6:7.1f} TFLOP/s"
                  f"  {'bit-identical to plain tiles' if same else 'MISMATCH max|d|=%g' % (res[(bn, pair, mc)].float() - res[(bn, 0, 1)].float()).abs().max().item()}",
                  flush=True)
print("PAIR_CHECK", "OK" if ok_all else "FAILED")
