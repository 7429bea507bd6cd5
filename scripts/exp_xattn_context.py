"""Cross-attention kernel timed ALONE through the C ABI (gram_op_cross_attention), cold and in the power state the
encoder phase leaves behind (XA_HOT=1 runs 8 ms of bf16 matmuls before every launch).  Explains why the kernel's
in-step bandwidth (bench.py: roofline_cross_attention) is below what it reaches in isolation.

    XA_HOT=0|1 XA_VALID=0|1 python scripts/exp_xattn_context.py

Measured on B200 (944 users, beam 20, 8 heads): 5.9-6.0 TB/s cold, 4.9-5.3 TB/s hot; back-to-back launches without
a sync in between reach 6.3-6.4 TB/s.  A 12 KiB row pitch (six layers interleaved, as the K/V projection writes
them) costs 2 % against a 2 KiB pitch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gram_b200 import _cabi  # noqa: E402

lib = _cabi.load_library()
mult = 1
users, K, H, dk = 944, 20, 8, 64
HD = H * dk
from gram_b200.data import GramTestData  # noqa: E402
data = GramTestData("Beauty")
real = np.array([data.valid_tokens([u]) for u in range(3 * users, 4 * users)])
hot = os.environ.get("XA_HOT", "0") == "1"
ga = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
gb = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
use_valid = os.environ.get("XA_VALID", "0") == "1"
for name, lens in (("equal", np.full(users, 1088)), ("ragged", np.random.default_rng(0).integers(256, 2689, size=users)),
                   ("beauty", real), ("beauty-sorted", np.sort(real)[::-1].copy())):
    ustart = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    T = int(ustart[-1])
    q = (torch.randn(users * K, HD, device="cuda") * 0.3).to(torch.bfloat16)
    kv = torch.randn(T + 256, mult * 2 * HD, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(users * K, HD, device="cuda", dtype=torch.bfloat16)
    us = torch.from_numpy(ustart).cuda()
    valid = torch.ones(T + 256, dtype=torch.uint8, device="cuda")
    args = (0, 1, 1, C.c_void_p(q.data_ptr()), C.c_void_p(kv.data_ptr()), T, C.c_void_p(us.data_ptr()),
            C.c_void_p(valid.data_ptr()) if use_valid else None,
            C.c_void_p(out.data_ptr()), users, K, H, dk, None)
    for _ in range(3):
        assert lib.gram_op_cross_attention(*args) == 0, lib.gram_last_error(None)
    tot = 0.0
    for _ in range(20):
        if hot:
            for _ in range(8):
                torch.matmul(ga, gb)          # ~1 ms of tensor-core work each: the power state of the encoder phase
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.gram_op_cross_attention(*args)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 20
    print(f"pitch {mult * 2 * HD * 2} B valid={use_valid} hot={hot} {name} (mean {lens.mean():.0f}, max {lens.max()}): {ms * 1000:.1f} us  {T * 2 * HD * 2 / ms / 1e6:.0f} GB/s", flush=True)
    del kv
