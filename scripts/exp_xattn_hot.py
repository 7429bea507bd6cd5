"""Cross-attention kernel ALONE (gram_op_cross_attention, persistent kernel) at the bench's shape -- 1,888 Beauty users, longest
first, 8 heads -- cold and between bursts of tensor-core work (the power state it meets inside a decode step), for 1 and 20
beam rows per user.  Testbed for kernel variants (GRAM_XATTN_WARPS=8): the in-step times are 765 us at step 0 and 905 us at
steps >= 1 against 720 us alone under ncu (profiles/r2b_instep_launch_times.txt).

    [GRAM_XATTN_WARPS=8] python scripts/exp_xattn_hot.py
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
users, H, dk = 1888, 8, 64
HD = H * dk
data = GramTestData("Beauty")
lens = np.sort(np.array([data.valid_tokens([u]) for u in range(users)]))[::-1].copy()
ustart = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
T = int(ustart[-1])
kv = torch.randn(T + 256, 2 * HD, device="cuda", dtype=torch.bfloat16)
us = torch.from_numpy(ustart).cuda()
valid = torch.ones(T + 256, dtype=torch.uint8, device="cuda")
ga = torch.randn(4096, 4096, device="cuda").to(torch.bfloat16)
gb = torch.randn(4096, 4096, device="cuda").to(torch.bfloat16)
print(f"variant warps={os.environ.get('GRAM_XATTN_WARPS', '4')}  {users} users, {T} tokens, {T * 2 * HD * 2 / 1e9:.2f} GB of K/V per launch", flush=True)
for K in (1, 20):
    q = (torch.randn(users * K, HD, device="cuda") * 0.3).to(torch.bfloat16)
    out = torch.zeros(users * K, HD, device="cuda", dtype=torch.bfloat16)
    args = (0, 1, 1, C.c_void_p(q.data_ptr()), C.c_void_p(kv.data_ptr()), T, C.c_void_p(us.data_ptr()),
            C.c_void_p(valid.data_ptr()), C.c_void_p(out.data_ptr()), users, K, H, dk, None)
    for _ in range(3):
        assert lib.gram_op_cross_attention(*args) == 0, lib.gram_last_error(None)
    for mode, burst in (("cold", 0), ("hot-0.3ms", 4), ("hot-0.6ms", 8), ("hot-1.2ms", 16)):
        evs = []
        for it in range(40):
            for _ in range(burst):
                torch.matmul(ga, gb)                  # ~75 us of tensor-core work each at 4096^3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.gram_op_cross_attention(*args)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = np.median([a.elapsed_time(b) for a, b in evs[8:]])
        print(f"  K={K:2d} {mode:10s}: {ms * 1000:7.1f} us  {T * 2 * HD * 2 / ms / 1e6:6.0f} GB/s", flush=True)
