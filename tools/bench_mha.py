"""Encoder attention micro-benchmark: the mma.sync kernel (attention.cu) against the tcgen05 kernel (attention_tc.cu).

    python tools/bench_mha.py [B,T ...]        e.g.  python tools/bench_mha.py 32,150 32,300 32,600 8,1800 8,3600

CUDA events over 20 launches after 5 warm-ups (inputs of one launch: B*T*2304 bf16; L2 is not flushed -- inside a
forward the QKV GEMM has just written them).  TFLOP/s = 4*T^2*768*B / time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from a2f_b200 import lib as L, ops


def main():
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(32, 150), (32, 300), (32, 600), (16, 1200),
                                                                              (8, 1800), (8, 3600), (1, 3600)]
    lib = L.load()
    dev = torch.device("cuda:0")
    print(f"{'B':>4} {'T':>6} {'mma.sync us':>12} {'TFLOP/s':>9} {'tcgen05 us':>12} {'TFLOP/s':>9} {'speedup':>8} {'auto us':>9}  (auto: short-clip kernel for T <= 160)")
    for B, T in shapes:
        qkv = torch.randn(B, T, 2304, device=dev).bfloat16()
        out = torch.empty(B, T, 768, device=dev, dtype=torch.bfloat16)
        res = []
        for impl in (1, 2, 0):
            L.check(lib.a2f_debug_set_umma_field(6, impl))
            for _ in range(5):
                ops.mha(qkv, out, B, T)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(20):
                ops.mha(qkv, out, B, T)
            e.record()
            torch.cuda.synchronize()
            res.append(s.elapsed_time(e) / 20 * 1e3)
        lib.a2f_debug_set_umma_field(6, 0)
        fl = 4.0 * T * T * 768 * B
        print(f"{B:4d} {T:6d} {res[0]:12.1f} {fl / res[0] / 1e6:9.1f} {res[1]:12.1f} {fl / res[1] / 1e6:9.1f} {res[0] / res[1]:8.2f} {res[2]:9.1f}")


if __name__ == "__main__":
    main()
