"""Times single ViT GEMM shapes through the C ABI (CUDA events, 20 repetitions after 5 warm-ups):
   python scripts/gemm_probe.py            -> fc1 with and without GELU, qkv, fc2/proj residual epilogues."""
import sys

import torch

sys.path.insert(0, ".")
from stamp_b200 import ops  # noqa: E402

dev = torch.device("cuda")
M = 192 * 197


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def run(name, N, K, act, store=ops.ST_16, resid=False):
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) * 0.03).half()
    bias = torch.randn(N, device=dev) * 0.1
    if resid:
        out = torch.zeros(M, N, device=dev)
        gamma = torch.ones(N, device=dev)
        us = timed(lambda: ops.gemm_tn(a, w, out=out, bias=bias, gamma=gamma, store=ops.ST_RESID32))
    else:
        out = torch.empty(M, N if store != ops.ST_SWIGLU16 else N // 2, device=dev, dtype=torch.float16)
        us = timed(lambda: ops.gemm_tn(a, w, out=out, bias=bias, act=act, store=store))
    tf = 2.0 * M * N * K / us / 1e6
    print(f"{name:28s} {us:8.1f} us  {tf:7.1f} TFLOP/s")


run("fc1 + GELU", 4096, 1024, ops.ACT_GELU)
run("fc1, no activation", 4096, 1024, ops.ACT_NONE)
run("fc1 + ReLU", 4096, 1024, ops.ACT_RELU)
run("qkv", 3072, 1024, ops.ACT_NONE)
run("fc2 (residual)", 1024, 4096, ops.ACT_NONE, resid=True)
run("proj (residual)", 1024, 1024, ops.ACT_NONE, resid=True)
run("virchow fc1 SwiGLU", 6832, 1280, ops.ACT_NONE, store=ops.ST_SWIGLU16)


def run_cublas(name, N, K):
    """The library GEMM (cuBLAS through torch.matmul, fp16 in / fp16 out, no bias, no activation) at the same shape:
    what "peak at this shape" means for a K = 1024 product, next to the 8192^3 figure of MEASURED_PEAKS.json."""
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) * 0.03).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    us = timed(lambda: torch.matmul(a, w.t(), out=out))
    print(f"cuBLAS {name:21s} {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")


run_cublas("fc1 shape", 4096, 1024)
run_cublas("qkv shape", 3072, 1024)
run_cublas("fc2 shape", 1024, 4096)
run_cublas("proj shape", 1024, 1024)
a = torch.randn(8192, 8192, device=dev).half()
b = torch.randn(8192, 8192, device=dev).half()
o = torch.empty(8192, 8192, device=dev, dtype=torch.float16)
us = timed(lambda: torch.matmul(a, b, out=o))
print(f"cuBLAS 8192^3 (burst)        {us:8.1f} us  {2.0 * 8192**3 / us / 1e6:7.1f} TFLOP/s")
