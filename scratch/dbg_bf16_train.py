import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import inputs as oin, ref_train as ort, weights as ow
sys.path.insert(0, "tests")
from test_faceformer_train_gpu import _train_inputs, _run_gpu
dev = torch.device("cuda:0")
sd = ow.make_state_dict("faceformer", seed=13)
audio, oh, tp, gt = _train_inputs(8000, 41)
tot, want = ort.faceformer_loss_and_grads(sd, audio, oh, tp, gt)
for prec in ("fp32", "bf16"):
    loss, grads = _run_gpu(dev, sd, prec, audio, oh, tp, gt)
    print(prec, loss, tot)
    gmax = max(float(g.norm()) for g in want.values())
    rows = []
    for k, g in want.items():
        n = float(g.norm())
        rel = float((grads[k].double() - g.double()).norm()) / max(n, 1e-30)
        rows.append((rel, k, n / gmax, float(grads[k].norm()) / gmax))
    rows.sort(reverse=True)
    for r in rows[:40]:
        print("  %.3e  %-70s |want|/gmax %.2e |got|/gmax %.2e" % r)
