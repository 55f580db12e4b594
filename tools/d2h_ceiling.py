"""Raw pinned-memory copy ceiling of the box for 1..N concurrent GPUs (is the end-to-end leg bound by
the box or by the staging?).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_ceiling.py [--mb 289]

Every rank owns one GPU, one pinned host buffer of --mb megabytes (the fp32 vertices of a 32 x 5 s @ 30 fps FaceFormer
step are 289 MB) and times plain `cudaMemcpyAsync` device->host and host->device copies with CUDA events -- first alone
(ranks take turns), then all ranks at once (barrier, same instant).  Before allocating, the rank binds itself to the CPUs
NVML reports as local to its GPU (first-touch places the pinned pages on that NUMA node).  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def bind_local_cpus(index: int) -> str:
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = sorted(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        return f"{len(before)} -> {len(after)} cpus ({after[0]}..{after[-1]})"
    except Exception as exc:  # noqa: BLE001
        return f"unchanged ({type(exc).__name__})"


def timed_copies(dst, src, reps):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in ev:
        s.record()
        dst.copy_(src, non_blocking=True)
        e.record()
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in ev)
    return src.numel() * src.element_size() / (ms[len(ms) // 2] * 1e-3) / 1e9      # GB/s, median


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=289.3)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    aff = bind_local_cpus(local)
    n = int(args.mb * 1e6) // 4
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    host.fill_(1.0)
    gpu = torch.ones(n, dtype=torch.float32, device=dev)
    timed_copies(host, gpu, 2)
    timed_copies(gpu, host, 2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    alone = torch.zeros(world, 2, dtype=torch.float64, device=dev)
    for r in range(world):                       # one rank at a time
        barrier()
        if r == rank:
            alone[r, 0] = timed_copies(host, gpu, args.reps)
            alone[r, 1] = timed_copies(gpu, host, args.reps)
    barrier()
    together = torch.zeros(world, 2, dtype=torch.float64, device=dev)
    together[rank, 0] = timed_copies(host, gpu, args.reps)
    barrier()
    together[rank, 1] = timed_copies(gpu, host, args.reps)
    barrier()
    if world > 1:
        dist.all_reduce(alone)
        dist.all_reduce(together)
        affs = [None] * world
        dist.all_gather_object(affs, aff)
    else:
        affs = [aff]
    if rank == 0:
        print(json.dumps({
            "what": "cudaMemcpyAsync pinned-host <-> device ceiling", "n_gpus": world, "bytes": n * 4, "reps": args.reps,
            "cpu_binding": affs,
            "d2h_alone_gbs": [round(float(x), 2) for x in alone[:, 0]], "h2d_alone_gbs": [round(float(x), 2) for x in alone[:, 1]],
            "d2h_concurrent_gbs": [round(float(x), 2) for x in together[:, 0]],
            "h2d_concurrent_gbs": [round(float(x), 2) for x in together[:, 1]],
            "d2h_concurrent_total_gbs": round(float(together[:, 0].sum()), 1),
            "h2d_concurrent_total_gbs": round(float(together[:, 1].sum()), 1),
            "frames_per_s_ceiling_fp32_vertices": round(float(together[:, 0].sum()) * 1e9 / 60276.0),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
