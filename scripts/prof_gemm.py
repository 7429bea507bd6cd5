"""A handful of tcgen05 GEMM launches at the encoder shapes, for `ncu` (profiling helper, not a benchmark).

    python scripts/prof_gemm.py [impl]        impl 1 = CTA pairs where eligible (default), 2 = single-CTA tiles
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gram_b200 import _cabi  # noqa: E402

lib = _cabi.load_library()
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M = 1048576
for (N, K, epi) in [(1536, 512, 0), (2048, 512, 1), (512, 2048, 2), (512, 512, 2)]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    Cd = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if epi < 2 else torch.float32)
    for _ in range(2):
        rc = lib.gram_op_gemm(0, 1, impl, epi, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(Cd.data_ptr()),
                              M, N, K, None)
        assert rc == 0, lib.gram_last_error(None)
    torch.cuda.synchronize()
    del A, W, Cd
print("ok")
