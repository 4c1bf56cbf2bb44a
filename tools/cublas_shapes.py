"""What does cuBLAS (torch.matmul, bf16) achieve on the GEMM shapes of the XL step at M = 500?  Informational ceiling for
csrc/gemm.cuh (library kernels are not on the product path).  CUDA-graph timing, 20 launches per graph."""
import os

import torch

M = 500
SHAPES = {"w13 (conv k=3 as K=4224)": (4224, 7680), "w2 (K=11520)": (11520, 1408), "qkv": (1408, 4224), "fc1": (1408, 5632),
          "fc2": (5632, 1408), "lin1 (K=4224)": (4224, 1408), "proj": (1408, 1408)}
for name, (K, N) in SHAPES.items():
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    torch.matmul(a, w.t(), out=out)
    torch.cuda.synchronize()
    if os.environ.get("ONE_LAUNCH"):      # ncu target: which kernel does the library pick for this shape?
        print(name, flush=True)
        continue
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(20):
                torch.matmul(a, w.t(), out=out)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            g.replay()
        e1.record(st)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 200
    print(f"cuBLAS bf16 {name:26s} M={M} K={K} N={N}: {us:7.2f} us  {2.0 * M * K * N / us / 1e6:7.1f} TFLOP/s", flush=True)
