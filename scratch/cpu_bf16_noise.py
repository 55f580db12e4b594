import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from oracle import inputs as oin, ref_train as ort, weights as ow
from test_faceformer_train_gpu import _train_inputs
sd = ow.make_state_dict("faceformer", seed=13)
audio, oh, tp, gt = _train_inputs(8000, 41)
tot, want = ort.faceformer_loss_and_grads(sd, audio, oh, tp, gt)
with torch.autocast("cpu", dtype=torch.bfloat16):
    tot2, got = ort.faceformer_loss_and_grads(sd, audio, oh, tp, gt)
print(tot, tot2)
gmax = max(float(g.norm()) for g in want.values())
rows = []
for k, g in want.items():
    n = float(g.norm())
    rel = float((got[k].double() - g.double()).norm()) / max(n, 1e-30)
    rows.append((rel, k, n / gmax))
rows.sort(reverse=True)
for r in rows[:60]:
    print("  %.3e  %-70s |want|/gmax %.2e" % r)
