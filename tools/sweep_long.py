"""BASELINE.json configs[4]: long-sequence sweep of FaceFormer autoregressive decode.

    python tools/sweep_long.py [--seconds 10,20,30,45,60] [--batches 1,2,4,...,128] [--fps 60] [--steps 5]

For every (utterance length L, batch B) one bf16 forward of the drop-in module is timed on the device (CUDA events,
3 warm-ups, L2 flushed between steps) and reported as frames/s, ms/step and fraction of the dense-bf16 roofline using
the FLOP model of SURVEY.md App. C (KV-cached decode).  Batches whose activations would exceed `--max-chunk-seconds`
of audio per launch are run by the module in utterance chunks (Faceformer.forward(..., max_chunk_seconds=...)).
Multi-GPU: run under torchrun; every rank takes B utterances (weak scaling, no collective), rank 0 prints the
aggregate.  Output: one JSON line per point + a table; summaries are committed under profiles/."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from a2f_b200 import modules
from bench import ff_flops_per_utt, peaks
from oracle import inputs as oin, weights as ow          # input / weight generators only


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", default="10,20,30,45,60")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128")
    ap.add_argument("--fps", type=int, default=60)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--max-chunk-seconds", type=float, default=480.0)
    ap.add_argument("--max-total-seconds", type=float, default=7680.0, help="skip points with B*L above this")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = modules.Faceformer(15069, 12)
    m.load_state_dict(ow.make_state_dict("faceformer", 13), strict=True)
    m = m.to(dev).eval().set_precision("bf16")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pk = peaks()
    rows = []
    for L in [float(s) for s in args.seconds.split(",")]:
        n = int(16000 * L)
        T = n * args.fps // 16000
        for B in [int(b) for b in args.batches.split(",")]:
            if B * L > args.max_total_seconds:
                continue
            audio = oin.audio(1, n, 7 + rank).to(dev).expand(B, n).contiguous()
            audio *= torch.linspace(0.7, 1.3, B, device=dev)[:, None]
            oh = oin.one_hot(B, 12, 7).to(dev)
            tp = oin.batch_templates(1, 7, scale=100.0).to(dev).expand(B, 5023, 3).contiguous()
            with torch.no_grad():
                for _ in range(3):
                    out = m(audio, oh, tp, fps=args.fps, max_chunk_seconds=args.max_chunk_seconds)
                assert bool(torch.isfinite(out[:, -1]).all())
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                for s, e in ev:
                    flush.zero_()
                    s.record()
                    out = m(audio, oh, tp, fps=args.fps, max_chunk_seconds=args.max_chunk_seconds)
                    e.record()
                torch.cuda.synchronize()
            sec = sum(s.elapsed_time(e) for s, e in ev) * 1e-3
            if world > 1:
                t = torch.tensor([sec], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t[0])
            del out
            fps_total = world * B * T * args.steps / sec
            tfl = world * B * ff_flops_per_utt(n, T) * args.steps / sec / 1e12
            row = {"seconds": L, "batch_per_gpu": B, "n_gpus": world, "frames_per_utt": T, "ms_per_step": 1e3 * sec / args.steps,
                   "frames_per_s": fps_total, "tflops": tfl, "frac_of_bf16_sustained": tfl / (pk["bf16_sustained"] * world)}
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    if rank == 0:
        print(f"# FaceFormer long-sequence sweep, {args.fps} fps, bf16, {world} GPU(s); roofline = {pk['bf16_sustained']} TFLOP/s "
              f"sustained per GPU ({pk['source']})")
        print(f"{'L(s)':>5} {'B/gpu':>6} {'T':>5} {'ms/step':>10} {'frames/s':>12} {'TFLOP/s':>9} {'frac':>6}")
        for r in rows:
            print(f"{r['seconds']:5g} {r['batch_per_gpu']:6d} {r['frames_per_utt']:5d} {r['ms_per_step']:10.2f} "
                  f"{r['frames_per_s']:12.0f} {r['tflops']:9.1f} {r['frac_of_bf16_sustained']:6.3f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
