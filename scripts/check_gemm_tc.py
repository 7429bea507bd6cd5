"""Standalone check + timing of the tcgen05 GEMM against torch (run under `timeout` on the GPU box)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gram_b200 import _cabi  # noqa: E402

lib = _cabi.load_library()
torch.backends.cuda.matmul.allow_tf32 = False


def run(M, N, K, epi, impl=1, check=True, iters=0):
    g = torch.Generator(device="cpu").manual_seed(M + N + K + epi)
    A = (torch.randn(M, K, generator=g)).cuda().to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).cuda().to(torch.bfloat16)
    if epi in (0, 1):
        Cd = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    else:
        Cd = torch.ones(M, N, device="cuda", dtype=torch.float32)
    args = (0, 1, impl, epi, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(Cd.data_ptr()), M, N, K, None)
    rc = lib.gram_op_gemm(*args)
    if rc != 0:
        print(f"  rc={rc} {lib.gram_last_error(None)}")
        return False
    torch.cuda.synchronize()
    ok = True
    if check:
        ref = A.float() @ W.float().t()
        if epi == 1:
            ref = ref.clamp_min(0)
        if epi == 2:
            ref = ref + 1.0
        err = ((Cd.float() - ref).abs().max() / ref.abs().max()).item()
        tol = 8e-3 if epi in (0, 1) else 2e-5
        ok = err < tol
        print(f"  M={M} N={N} K={K} epi={epi} impl={impl} rel_err={err:.3e} {'ok' if ok else 'FAIL'}", flush=True)
    if iters:
        for _ in range(3):
            lib.gram_op_gemm(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            lib.gram_op_gemm(*args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"  M={M} N={N} K={K} epi={epi} impl={impl}: {ms * 1000:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


if __name__ == "__main__":
    allok = True
    print("correctness (tcgen05 vs torch fp32 matmul of the same bf16 operands)")
    for (M, N, K) in [(128, 128, 64), (128, 128, 512), (1, 128, 64), (200, 512, 512), (77, 192, 64), (333, 1536, 512),
                      (1000, 2048, 512), (513, 512, 2048), (129, 32128, 512), (700, 6144, 512), (4097, 768, 3072),
                      (38000, 512, 512), (38001, 784, 128), (20000, 32128, 64)]:   # the last three take the 128x256 tile
        for epi in (0, 1, 2, 3):
            allok &= run(M, N, K, epi)
    print("CTA pairs (impl 1) against single-CTA tiles (impl 2) on shapes large enough for 256x256 pair tiles")
    for (M, N, K) in [(38000, 512, 512), (38001, 784, 128), (20000, 32128, 64), (37889, 1536, 512), (75776, 2048, 512),
                      (40000, 512, 2048), (18880, 512, 512), (18880, 2048, 512)]:
        for epi in (0, 1, 2, 3):
            allok &= run(M, N, K, epi, impl=1)
            allok &= run(M, N, K, epi, impl=2)
    print("ALL OK" if allok else "SOME FAILED")
    if "--time" in sys.argv:
        print("timing")
        for (M, N, K, epi) in [(262144, 1536, 512, 0), (262144, 512, 512, 2), (262144, 2048, 512, 1), (262144, 512, 2048, 2),
                               (262144, 6144, 512, 0), (1048576, 1536, 512, 0), (1048576, 2048, 512, 1), (18880, 32128, 512, 3), (18880, 512, 512, 2), (18880, 2048, 512, 1), (5120, 1536, 512, 0), (5120, 512, 512, 2), (5120, 2048, 512, 1),
                               (5120, 512, 2048, 2), (5120, 32128, 512, 3)]:
            run(M, N, K, epi, impl=1, check=False, iters=10)
            run(M, N, K, epi, impl=2, check=False, iters=10)
            if "--simt" in sys.argv:
                run(M, N, K, epi, impl=0, check=False, iters=2)
    sys.exit(0 if allok else 1)
