#!/usr/bin/env python
"""Host-link ceiling of the box with N GPUs copying AT THE SAME TIME, for reading bench.py's end-to-end scaling:
every rank moves exactly what one bench step moves (393 MB pinned host -> device, 224 MB device -> pinned host), in the
chunk shapes sdrb_bank_process_host_async uses (4 stream groups x 4 callbacks = 16 two-dimensional copies per direction)
and as two single copies, with nothing else running. Launch like the bench:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe_n.py
Rank 0 prints one JSON object: per-rank and aggregate GB/s, max over ranks of the device-timed duration (barrier before)."""
import json
import os
import sys

import torch
import torch.distributed as dist

S, NB, BLOCK, PCM = 128, 4, 384000, 219000          # streams, callbacks, complex samples per callback, int16 per callback record
H2D, D2H = S * NB * BLOCK * 2, S * NB * PCM * 2


def run(dev, mode, reps=8):
    hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
    hout = torch.empty(D2H, dtype=torch.uint8).pin_memory()
    din = torch.empty(H2D, dtype=torch.uint8, device=dev)
    dout = torch.empty(D2H, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    times = []
    for _ in range(reps):
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        if mode == "single":
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        else:
            # the library's chunks: for each callback, for each of 4 stream groups, a strided block of 32 rows
            hi, di = hin.view(S, NB, BLOCK * 2), din.view(S, NB, BLOCK * 2)
            ho, do = hout.view(S, NB, PCM * 2), dout.view(S, NB, PCM * 2)
            for cb in range(NB):
                for g in range(4):
                    with torch.cuda.stream(s1):
                        di[32 * g:32 * g + 32, cb].copy_(hi[32 * g:32 * g + 32, cb], non_blocking=True)
                    with torch.cuda.stream(s2):
                        ho[32 * g:32 * g + 32, cb].copy_(do[32 * g:32 * g + 32, cb], non_blocking=True)
        e1.record(s1)
        e2.record(s2)
        torch.cuda.synchronize()
        times.append(max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
    times.sort()
    return times[len(times) // 2]


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"n_gpus": world, "h2d_bytes_per_rank": H2D, "d2h_bytes_per_rank": D2H}
    for mode in ("single", "chunks"):
        ms = run(dev, mode)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            ts = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(ts, t)
            per = [float(x.item()) for x in ts]
        else:
            per = [ms]
        worst = max(per)
        out[mode] = {"ms_per_rank": per, "ms_max": worst, "aggregate_h2d_gbs": world * H2D / worst / 1e6,
                     "aggregate_d2h_gbs": world * D2H / worst / 1e6,
                     "bench_step_equivalent_gsps": world * S * NB * BLOCK / worst / 1e6}
    if rank == 0:
        json.dump(out, sys.stdout, indent=1)
        print()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
