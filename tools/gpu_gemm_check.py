"""GPU check of the tcgen05 GEMM entry point (foley_gemm) against torch matmul on the same inputs.

Run on the GPU box: python tools/gpu_gemm_check.py
"""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "comfyui-hunyuanvideo-foley_b200", "libfoley_b200.so"))
lib.foley_last_error.restype = ctypes.c_char_p
i64, i32, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p
lib.foley_gemm.argtypes = [vp, i32, i64, i64, i64, i64, i64, vp, i64, i32, i32, i32, i32, i32, i32, i32,
                           vp, vp, i64, i64, i64, vp]
lib.foley_gemm.restype = i32


def gemm(a, w, taps=1, off0=0, stride=1, splits=1, bn=128, mode=0, act=0, bias=None):
    """a: [B, R, K] (bf16|f32), w: [N, taps*K] -> out per mode."""
    B, R, K = a.shape
    N = w.shape[0]
    dt = 0 if a.dtype == torch.bfloat16 else 1
    if mode == 0:
        out = torch.zeros(B, R, N, dtype=torch.bfloat16, device=a.device)
    elif mode == 1:
        out = torch.zeros(B, R, N // 2, dtype=torch.bfloat16, device=a.device)
    else:
        out = torch.zeros(splits, B, R, N, dtype=torch.float32, device=a.device)
    ldo = out.shape[-1]
    st = lib.foley_gemm(a.data_ptr(), dt, B, R, K, a.stride(1), a.stride(0), w.data_ptr(), N, taps, off0, stride,
                        splits, bn, mode, act, bias.data_ptr() if bias is not None else None, out.data_ptr(),
                        ldo, R * ldo, B * R * ldo, None)
    if st != 0:
        raise RuntimeError(lib.foley_last_error().decode())
    return out


def ref_conv(a, w, taps, off0, stride):
    """fp32 reference of the multi-tap form with per-sample zero halo."""
    B, R, K = a.shape
    af = a.float()
    wf = w.float()
    out = torch.zeros(B, R, w.shape[0], dtype=torch.float32, device=a.device)
    for t in range(taps):
        sh = off0 + t * stride
        shifted = torch.zeros_like(af)
        lo, hi = max(0, -sh), min(R, R - sh)
        if hi > lo:
            shifted[:, lo:hi] = af[:, lo + sh:hi + sh]
        out += shifted @ wf[:, t * K:(t + 1) * K].T
    return out


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def flags():
    buf = (ctypes.c_uint32 * 4)()
    lib.foley_debug_flags(buf)
    return [hex(x) for x in buf]


def stage(n):
    """Bring-up: run one tiny GEMM with a partial pipeline (1 setup, 2 +TMA, 3 +MMA, 0 full)."""
    a = torch.randn(1, 128, 64, device="cuda").bfloat16()
    w = torch.randn(128, 64, device="cuda").bfloat16()
    out = gemm(a, w, bn=128 | (n << 16), mode=2)
    torch.cuda.synchronize()
    print(f"stage {n}: ok flags={flags()} rel={rel(out[0], a.float() @ w.float().T):.3e}", flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "stage":
        stage(int(sys.argv[2]))
        return 0
    if len(sys.argv) > 1 and sys.argv[1] == "bringup":
        import subprocess
        for n in (1, 2, 3, 0):
            r = subprocess.run([sys.executable, __file__, "stage", str(n)], capture_output=True, text=True)
            print(r.stdout.strip()[-400:], r.stderr.strip()[-600:] if r.returncode else "", flush=True)
        return 0
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    cases = [
        # (B, R, K, N, taps, off0, stride, splits, bn, dtype)
        (1, 128, 64, 128, 1, 0, 1, 1, 128, torch.bfloat16),
        (1, 128, 256, 128, 1, 0, 1, 1, 128, torch.bfloat16),
        (2, 250, 1536, 4608, 1, 0, 1, 1, 128, torch.bfloat16),
        (2, 250, 1536, 1536, 3, -1, 1, 3, 128, torch.bfloat16),
        (2, 250, 1536, 1536, 1, 0, 1, 1, 64, torch.bfloat16),
        (2, 40, 1536, 4608, 1, 0, 1, 1, 256, torch.bfloat16),
        (2, 250, 4096, 1536, 3, -1, 1, 4, 128, torch.bfloat16),
        (1, 300, 128, 256, 7, -27, 9, 1, 128, torch.float32),
        (2, 250, 64, 64, 7, -3, 1, 1, 64, torch.float32),
        (1, 251, 256, 1024, 2, 0, -1, 1, 256, torch.float32),
    ]
    for (B, R, K, N, taps, off0, stride, splits, bn, dt) in cases:
        a = torch.randn(B, R, K, device=dev).to(dt)
        w = (torch.randn(N, taps * K, device=dev) * 0.05).to(dt)
        out = gemm(a, w, taps, off0, stride, splits, bn, mode=2)
        torch.cuda.synchronize()
        got = out.sum(0)
        want = ref_conv(a, w, taps, off0, stride)
        r = rel(got, want)
        fl = flags()
        if fl[0] != "0x0":
            print("  mbarrier timeout code", fl, flush=True)
        tol = 2e-3 if dt == torch.float32 else 1e-5
        flag = "OK " if r < tol else "BAD"
        ok &= r < tol
        print(f"{flag} f32-partials B={B} R={R} K={K} N={N} taps={taps} off0={off0} st={stride} splits={splits} "
              f"bn={bn} {str(dt)[6:]} rel={r:.3e}", flush=True)

    # bf16 epilogues
    a = torch.randn(2, 250, 1536, device=dev).bfloat16()
    w = (torch.randn(4608, 1536, device=dev) * 0.02).bfloat16()
    bias = (torch.randn(4608, device=dev) * 0.1).bfloat16()
    want = torch.nn.functional.linear(a.float(), w.float(), bias.float())
    for act, fn in [(0, lambda x: x), (1, torch.nn.functional.silu),
                    (2, lambda x: torch.nn.functional.gelu(x, approximate="tanh"))]:
        out = gemm(a, w, mode=0, act=act, bias=bias)
        ref = want.bfloat16() if act == 0 else fn(want.bfloat16().float()).bfloat16()
        r = rel(out, ref)
        mism = (out != ref).float().mean().item()
        flag = "OK " if r < 3e-3 else "BAD"
        ok &= r < 3e-3
        print(f"{flag} bf16 epilogue act={act} rel={r:.3e} mismatch_frac={mism:.4f}", flush=True)
    # swiglu: interleave w1/w3 rows
    w1 = (torch.randn(4096, 1536, device=dev) * 0.02).bfloat16()
    w3 = (torch.randn(4096, 1536, device=dev) * 0.02).bfloat16()
    wi = torch.stack([w1, w3], dim=1).reshape(8192, 1536).contiguous()
    out = gemm(a, wi, mode=1)
    g = (a.float() @ w1.float().T).bfloat16()
    u = (a.float() @ w3.float().T).bfloat16()
    ref = (torch.nn.functional.silu(g.float()).bfloat16().float() * u.float()).bfloat16()
    r = rel(out, ref)
    flag = "OK " if r < 3e-3 else "BAD"
    ok &= r < 3e-3
    print(f"{flag} swiglu rel={r:.3e} mismatch_frac={(out != ref).float().mean().item():.4f}", flush=True)

    # timing of the big shapes (cold-ish; informational)
    for (B, R, K, N, taps, splits, bn) in [(2, 250, 1536, 8192, 3, 1, 128), (2, 250, 4096, 1536, 3, 3, 128),
                                           (2, 250, 1536, 9216, 1, 1, 128), (16, 250, 1536, 8192, 3, 1, 128),
                                           (16, 250, 1536, 8192, 3, 1, 256)]:
        a = torch.randn(B, R, K, device=dev).bfloat16()
        w = (torch.randn(N, taps * K, device=dev) * 0.02).bfloat16()
        for _ in range(3):
            gemm(a, w, taps, -1 if taps == 3 else 0, 1, splits, bn, mode=2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        outbuf = None
        e0.record()
        iters = 20
        for _ in range(iters):
            outbuf = gemm(a, w, taps, -1 if taps == 3 else 0, 1, splits, bn, mode=2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * B * R * K * taps * N
        print(f"time B={B} R={R} K={K}x{taps} N={N} splits={splits} bn={bn}: {ms*1e3:.1f} us  "
              f"{fl/ms/1e9:.1f} TFLOP/s (incl. torch.zeros alloc)", flush=True)
    print("ALL_OK" if ok else "SOME_BAD")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
