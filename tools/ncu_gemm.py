"""Driver for ncu captures of the tcgen05 GEMM kernels on the unit-embedding shape (2M x 128 x 128):
   ncu --set full --clock-control none --import-source on -k regex:'atmem|wgrad_kernel' -c 4 -o gpurun_out/gemm python tools/ncu_gemm.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dotaclient_b200 import ops  # noqa: E402

d = torch.device("cuda", 0)
M = 131072 * 16
a = torch.randn(M, 128, device=d)
w = torch.randn(128, 128, device=d) * 0.1
b = torch.randn(128, device=d)
out = torch.empty(M, 128, device=d)
for _ in range(2):
    ops.gemm_tf32x3(a, w, b, out=out)
    ops.gemm_wgrad_tf32x3(out, a)
torch.cuda.synchronize()
