"""One GEMM shape, a few launches: target for ncu captures (python tools/prof_gemm.py M N K [tmpl])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from a2f_b200 import ops, lib as L

M, N, K = [int(v) for v in sys.argv[1:4]]
mode = sys.argv[4] if len(sys.argv) > 4 else "bf16"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
a = torch.randn(M, K, generator=g).to(dev).bfloat16()
w = (torch.randn(N, K, generator=g) * K ** -0.5).to(dev).bfloat16()
b = torch.randn(N, generator=g).to(dev)
if mode == "head":
    out = torch.empty((M, N), device=dev)
    tm = torch.randn((M + 299) // 300, N, generator=g).to(dev)
    fn = lambda: ops.gemm(a, w, out, bias=b, tmpl=tm, rows_per_tmpl=300, backend=L.TCGEN05)
else:
    out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, out, bias=b, backend=L.TCGEN05)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("done")
