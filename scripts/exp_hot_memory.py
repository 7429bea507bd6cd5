"""Does the power state the tensor-bound encoder phase leaves behind slow down HBM-bound work as such?  A plain device copy
(torch copy_, 2 x 1 GiB) and the cross-attention kernel (1,888 Beauty users, 20 beam rows), each timed cold and right after
~100 ms of back-to-back bf16 matmuls at the power cap -- the situation of the decode phase inside a step.

    python scripts/exp_hot_memory.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gram_b200 import _cabi  # noqa: E402
from gram_b200.data import GramTestData  # noqa: E402

lib = _cabi.load_library()
users, H, dk, K = 1888, 8, 64, 20
HD = H * dk
data = GramTestData("Beauty")
lens = np.sort(np.array([data.valid_tokens([u]) for u in range(users)]))[::-1].copy()
ustart = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
T = int(ustart[-1])
kv = torch.randn(T + 256, 2 * HD, device="cuda", dtype=torch.bfloat16)
us = torch.from_numpy(ustart).cuda()
valid = torch.ones(T + 256, dtype=torch.uint8, device="cuda")
q = (torch.randn(users * K, HD, device="cuda") * 0.3).to(torch.bfloat16)
out = torch.zeros(users * K, HD, device="cuda", dtype=torch.bfloat16)
args = (0, 1, 1, C.c_void_p(q.data_ptr()), C.c_void_p(kv.data_ptr()), T, C.c_void_p(us.data_ptr()),
        C.c_void_p(valid.data_ptr()), C.c_void_p(out.data_ptr()), users, K, H, dk, None)
src = torch.empty(1 << 29, device="cuda", dtype=torch.bfloat16)     # 1 GiB
dst = torch.empty_like(src)
ga = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
gb = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)


def xattn():
    assert lib.gram_op_cross_attention(*args) == 0


def copy():
    dst.copy_(src)


def timed(fn, heat_ms, reps, n_inner):
    ts = []
    for _ in range(reps):
        if heat_ms:
            for _ in range(int(heat_ms / 0.8)):          # ~0.8 ms per 8192^3 matmul at the sustained rate
                torch.matmul(ga, gb)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(n_inner + 1)]
        e[0].record()
        for i in range(n_inner):
            fn()
            e[i + 1].record()
        torch.cuda.synchronize()
        ts.append([e[i].elapsed_time(e[i + 1]) for i in range(n_inner)])
    return np.median(np.array(ts), axis=0)


for name, fn, nbytes in (("copy 1 GiB (read + write)", copy, 2 * src.numel() * 2), ("cross-attention K=20", xattn, T * 2 * HD * 2)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    cold = timed(fn, 0, 8, 4)
    hot = timed(fn, 100, 8, 4)
    print(f"{name}: cold {[f'{nbytes / t / 1e6:.0f}' for t in cold]} GB/s; launches 1..4 right after 100 ms of matmuls: "
          f"{[f'{nbytes / t / 1e6:.0f}' for t in hot]} GB/s", flush=True)
