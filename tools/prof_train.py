"""One FaceFormer bf16 training step at the config-4 shape (B=8 x 5 s, 60 fps), a few times: ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from a2f_b200 import modules, trainer as tr
from oracle import inputs as oin, weights as ow

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
m = modules.Faceformer(15069, 12)
m.load_state_dict(ow.make_state_dict("faceformer", 13), strict=True)
m = m.to(dev).eval().set_precision("bf16")
t = tr.FaceformerTrainer(m, fps=60)
tp = oin.batch_templates(B, 1, scale=100.0)
args = [oin.audio(B, 80000, 1).to(dev), oin.one_hot(B, 12, 1).to(dev), tp.to(dev),
        oin.gt_like((B, 300, 5023, 3), tp[:, None], 2, scale=100.0).to(dev)]
import time
for i in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t.step(*args)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"step {i}: host enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms")
print("done")
