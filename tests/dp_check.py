"""Data-parallel training check, run under torchrun on N >= 2 GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_check.py

Every rank takes its shard of a global batch, runs FaceformerTrainer.forward_backward (staged NCCL all-reduce started
from inside the backward); the averaged gradient must equal the gradient of the whole batch computed by one process
(mean of per-utterance losses), with and without overlap, and the replicas must stay bit-identical after the step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from a2f_b200 import modules, trainer as tr
from oracle import inputs as oin, weights as ow


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    per = 2
    Bg, n = per * world, 6400                                  # T = 24 frames
    audio, oh = oin.audio(Bg, n, 51), oin.one_hot(Bg, 12, 51)
    tp = oin.batch_templates(Bg, 51, scale=100.0)
    gt = oin.gt_like((Bg, 24, 5023, 3), tp[:, None], 52, scale=100.0)
    sd = ow.make_state_dict("faceformer", seed=13)
    sl = slice(rank * per, (rank + 1) * per)

    def make(overlap):
        m = modules.Faceformer(15069, 12)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).eval().set_precision(precision)
        return tr.FaceformerTrainer(m, lr=1e-4, fps=60, overlap=overlap)

    grads = {}
    for overlap in (True, False):
        t = make(overlap)
        out3 = t.forward_backward(audio[sl].to(dev), oh[sl].to(dev), tp[sl].to(dev), gt[sl].to(dev))
        t.flat.finish_all_reduce()
        torch.cuda.synchronize()
        grads[overlap] = (t.flat.reduced_grads().float() / world).clone()   # bf16 precision: the bf16 wire mirror
        if overlap:
            t.optimizer_step()
            torch.cuda.synchronize()
            mine = t.flat.params.clone()
            other = mine.clone()
            dist.broadcast(other, src=0)
            if not torch.equal(mine, other):
                g0 = t.flat.reduced_grads().float().clone()
                dist.broadcast(g0, src=0)
                for name, p, off, nn_, st in t.flat.entries:
                    dg = float((g0[off:off + nn_] - t.flat.reduced_grads()[off:off + nn_].float()).abs().max())
                    dp = float((other[off:off + nn_] - mine[off:off + nn_]).abs().max())
                    if (dg > 0 or dp > 0) and rank == 1:
                        print(f"  DIVERGED stage {st:2d} {name}: max|dgrad| {dg:.3e} max|dparam| {dp:.3e} "
                              f"|g|max {float(g0[off:off + nn_].abs().max()):.3e}", flush=True)
                bad = ((other != mine) | torch.isnan(mine)).nonzero().reshape(-1)
                if rank == 1:
                    print("  DIVERGED elements", bad.numel(), "first", bad[:8].tolist(), flush=True)
                    for name, p, off, nn_, st in t.flat.entries:
                        end = (off + nn_ + 63) // 64 * 64
                        k = int(((bad >= off + nn_) & (bad < end)).sum())
                        if k:
                            print(f"  DIVERGED padding after {name} (stage {st}, numel {nn_}, shape {tuple(p.shape)}): {k} elements; "
                                  f"grads there {t.flat.grads[off + nn_:end][:8].tolist()}", flush=True)
                raise AssertionError("replicas diverged after the optimizer step")
            loss_local = out3.clone()
    d = float((grads[True] - grads[False]).abs().max())
    assert d == 0.0 or precision != "fp32" or d < 1e-6 * float(grads[False].abs().max()), d

    # whole batch on one process (every rank does it; cheap at this size)
    t = make(False)
    full = t.forward_backward(audio.to(dev), oh.to(dev), tp.to(dev), gt.to(dev))
    torch.cuda.synchronize()
    want = t.flat.grads
    rel = float((grads[True] - want).norm() / want.norm())
    worst = 0.0
    for name, p, off, nn_, st in t.flat.entries:
        w = want[off:off + nn_]
        if float(w.norm()) > 1e-6 * float(want.norm()):
            worst = max(worst, float((grads[True][off:off + nn_] - w).norm() / w.norm()))
    losses = loss_local.clone()
    dist.all_reduce(losses)
    losses /= world
    lrel = abs(float(losses[0]) - float(full[0])) / abs(float(full[0]))
    tol = 1e-4 if precision == "fp32" else 5e-2
    if rank == 0:
        print(f"dp_check[{precision}] world={world}: grad rel {rel:.2e} (worst tensor {worst:.2e}), loss rel {lrel:.2e}")
    assert rel < tol and lrel < 1e-5, (rel, lrel)
    dist.barrier()
    if rank == 0:
        print("DP_CHECK_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
