"""One FaceFormer bf16 forward at the bench shape, a few times: target for ncu launch lists / captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from a2f_b200 import modules
from oracle import inputs as oin, weights as ow

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
fps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
m = modules.Faceformer(15069, 12)
m.load_state_dict(ow.make_state_dict("faceformer", 13), strict=True)
m = m.to(dev).eval().set_precision("bf16")
audio, oh, tp = oin.audio(B, 80000, 1).to(dev), oin.one_hot(B, 12, 1).to(dev), oin.batch_templates(B, 1, scale=100.0).to(dev)
with torch.no_grad():
    for _ in range(reps - 1):
        m(audio, oh, tp, fps=fps)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()      # ncu --profile-from-start off: only the last forward is captured
    m(audio, oh, tp, fps=fps)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
