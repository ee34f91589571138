"""ncu driver: the fold consumer GEMM and the default GEMM on the same shape (python tests/gpu_prof_fold.py M N K)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import _lib as L  # noqa: E402

lib = L.load()
L.check(lib.sprc_set_act_dtype(1))
M, N, K = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else (112184, 2304, 768)
dev = torch.device("cuda:0")
s = torch.randn(M, K, device=dev) * 1.7
s16 = s.half()
xs = s.view(M, K // 64, 64)
m = xs.mean(-1)
st = torch.stack([m, ((xs - m[..., None]) ** 2).sum(-1)], -1).permute(1, 0, 2).contiguous()
W = (torch.randn(N, K, device=dev) * K ** -0.5).half()
b = torch.randn(N, device=dev)
c = torch.randn(N, device=dev)
out = torch.zeros(M, N, device=dev, dtype=torch.float16)
f = L.SprcGemmFold()
f.split, f.eps = 0, 1e-12
f.st_in, f.c = st.data_ptr(), c.data_ptr()
for _ in range(3):
    L.check(lib.sprc_op_gemm(L.ptr(s16), L.ptr(W), M, N, K, K, K, 0, 0, L.ptr(b), None, None, L.ptr(out), N, 0, 0,
                             L.cur_stream()))
    L.check(lib.sprc_op_gemm_fold(L.ptr(s16), L.ptr(W), None, M, 0, N, K, L.ptr(b), None, 0, None, L.ptr(out), f,
                                  L.cur_stream()))
torch.cuda.synchronize()
print("done")
