"""The persistent multi-tile GEMM (csrc/gemm_persistent.cuh: grids of more than ~1.5 waves of 128x256 tiles) against an
fp32 torch reference and, bit for bit, against the one-tile-per-CTA kernel (same problem issued sample by sample, which
stays below the persistent threshold).  Tolerances: fp32 partial sums of bf16 products 1e-5 relative L2; bf16 epilogues
3e-3 (one bf16 rounding of the output)."""
import ctypes

import pytest
import torch

from conftest import load_pkg, rel_l2

pytestmark = pytest.mark.gpu
i64, i32, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p


def _lib():
    # (conftest.py sets FOLEY_GEMM_PERSIST=2 before the library is loaded: fp32 partials take the persistent kernel too)
    lib = load_pkg("engine").load_library()
    lib.foley_gemm.argtypes = [vp, i32, i64, i64, i64, i64, i64, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, i64,
                               i64, i64, vp]
    return lib


def _gemm(lib, a, w, taps, splits, mode, act=0, bias=None):
    B, R, K = a.shape
    N = w.shape[0]
    if mode == 0:
        out = torch.zeros(B, R, N, dtype=torch.bfloat16, device=a.device)
    elif mode == 1:
        out = torch.zeros(B, R, N // 2, dtype=torch.bfloat16, device=a.device)
    else:
        out = torch.zeros(splits, B, R, N, dtype=torch.float32, device=a.device)
    ldo = out.shape[-1]
    st = lib.foley_gemm(a.data_ptr(), 0, B, R, K, K, R * K, w.data_ptr(), N, taps, -(taps // 2), 1, splits, 256, mode, act,
                        bias.data_ptr() if bias is not None else None, out.data_ptr(), ldo, R * ldo, B * R * ldo, None)
    assert st == 0, lib.foley_last_error()
    torch.cuda.synchronize()
    return out


def _ref(a, w, taps):
    B, R, K = a.shape
    af, wf = a.float(), w.float()
    out = torch.zeros(B, R, w.shape[0], device=a.device)
    for t in range(taps):
        sh = t - taps // 2
        shifted = torch.zeros_like(af)
        lo, hi = max(0, -sh), min(R, R - sh)
        shifted[:, lo:hi] = af[:, lo + sh:hi + sh]
        out += shifted @ wf[:, t * K:(t + 1) * K].T
    return out


@pytest.mark.parametrize("B,R,K,N,taps,splits", [(8, 250, 1408, 4224, 1, 1), (8, 250, 1408, 1408, 3, 3), (16, 250, 512, 2304, 1, 2),
                                                 (2, 250, 256, 40960, 1, 1)])
def test_fp32_partials(B, R, K, N, taps, splits):
    lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(B * 7 + taps)
    a = torch.randn(B, R, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, taps * K, device="cuda", generator=g) * 0.05).bfloat16()
    out = _gemm(lib, a, w, taps, splits, 2)
    assert rel_l2(out.sum(0).cpu(), _ref(a, w, taps).cpu()) <= 1e-5
    # one-tile-per-CTA kernel on each sample alone: identical bits (same per-tile arithmetic, K ranges and split order)
    solo = torch.cat([_gemm(lib, a[b:b + 1], w, taps, splits, 2) for b in range(B)], dim=1)
    assert torch.equal(out, solo)


@pytest.mark.parametrize("act", [0, 1, 2])
def test_bf16_epilogue_with_bias_and_activation(act):
    lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(act)
    a = torch.randn(8, 250, 1408, device="cuda", generator=g).bfloat16()
    w = (torch.randn(5632, 1408, device="cuda", generator=g) * 0.02).bfloat16()
    bias = (torch.randn(5632, device="cuda", generator=g) * 0.1).bfloat16()
    out = _gemm(lib, a, w, 1, 1, 0, act, bias)
    want = torch.nn.functional.linear(a.float(), w.float(), bias.float()).bfloat16().float()
    want = [want, torch.nn.functional.silu(want), torch.nn.functional.gelu(want, approximate="tanh")][act].bfloat16()
    assert rel_l2(out.float().cpu(), want.float().cpu()) <= 3e-3
    solo = torch.cat([_gemm(lib, a[b:b + 1], w, 1, 1, 0, act, bias) for b in range(8)], dim=0)
    assert torch.equal(out, solo)


def test_swiglu_pairs():
    lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(8, 250, 1408, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(3840, 3 * 1408, device="cuda", generator=g) * 0.02).bfloat16()
    w3 = (torch.randn(3840, 3 * 1408, device="cuda", generator=g) * 0.02).bfloat16()
    wi = torch.stack([w1, w3], dim=1).reshape(7680, 3 * 1408).contiguous()
    out = _gemm(lib, a, wi, 3, 1, 1)
    gate, up = _ref(a, w1, 3).bfloat16().float(), _ref(a, w3, 3).bfloat16().float()
    want = (torch.nn.functional.silu(gate).bfloat16().float() * up).bfloat16()
    assert rel_l2(out.float().cpu(), want.float().cpu()) <= 3e-3
    solo = torch.cat([_gemm(lib, a[b:b + 1], wi, 3, 1, 1) for b in range(8)], dim=0)
    assert torch.equal(out, solo)
    flags = (ctypes.c_uint32 * 4)()
    lib.foley_debug_flags(flags)
    assert flags[0] == 0


@pytest.mark.parametrize("R,K,N,act,bn", [(21966, 768, 2304, 0, 256), (1569, 768, 3072, 4, 256), (112, 3072, 768, 0, 64), (300, 1536, 768, 0, 128)])
def test_fp16_operands(R, K, N, act, bn):
    """fp16 I/O mode (FOLEY_DT_F16: the Synchformer's Linear layers under fp16 autocast) on the persistent and the one-tile
    kernel against fp32 matmul of the same fp16 operands rounded once to fp16 (tolerance: one fp16 rounding, 2^-11: 5e-4);
    act 4 = nn.GELU()."""
    lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(R + N)
    a = torch.randn(1, R, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    b = (torch.randn(N, device="cuda", generator=g) * 0.1).half()
    out = torch.zeros(1, R, N, dtype=torch.float16, device="cuda")
    st = lib.foley_gemm(a.data_ptr(), 2, 1, R, K, K, R * K, w.data_ptr(), N, 1, 0, 1, 1, bn, 0, act, b.data_ptr(), out.data_ptr(), N, R * N, R * N, None)
    assert st == 0, lib.foley_last_error()
    torch.cuda.synchronize()
    want = (a[0].float() @ w.float().T + b.float()).half().float()
    if act == 4:
        want = torch.nn.functional.gelu(want).half().float()
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out[0].float(), want) < 5e-4
