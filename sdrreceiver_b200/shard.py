"""Multi-GPU host logic. The channelizer shards by stream (independent dongles): stream g of the
job lives on rank g % world, nothing is exchanged on the data path, every rank publishes its own
ZMQ endpoint. The only collective is an all_gather of a few numbers per rank at the end of a run:
device time, samples processed and an output digest (NCCL on GPUs; gloo in the CPU tests)."""
import numpy as np


def stream_ids(rank, world, per_rank):
    """Global stream indices owned by `rank` (round robin)."""
    return [rank + world * j for j in range(per_rank)]


def owner(stream, world):
    return stream % world, stream // world          # (rank, local slot)


def rank_endpoint(address, rank):
    """Per-rank ZMQ endpoint: the reference binds one PUB socket per process (vfo.cpp:160-165);
    here every rank binds base port + rank."""
    head, sep, port = address.rpartition(":")
    if not sep or not port.isdigit():
        return address
    return "%s:%d" % (head, int(port) + rank)


def pcm_digest(pcm):
    """Order-independent digest of an int16 output block: (sum, sum of squares mod 2^62, xor)."""
    a = np.asarray(pcm).astype(np.int64).reshape(-1)
    x = np.bitwise_xor.reduce(a.astype(np.int16).view(np.uint16).astype(np.int64)) if a.size else 0
    return [int(a.sum()), int((a * a).sum() % (1 << 62)), int(x)]


def gather(stats, digest, device=None):
    """all_gather of per-rank float stats and int digest. Returns (stats [world, n], digests)."""
    import torch
    import torch.distributed as dist
    s = torch.tensor(stats, dtype=torch.float64, device=device)
    d = torch.tensor(digest, dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return s.cpu().numpy()[None, :], [[int(v) for v in d.tolist()]]
    world = dist.get_world_size()
    ss = [torch.zeros_like(s) for _ in range(world)]
    dd = [torch.zeros_like(d) for _ in range(world)]
    dist.all_gather(ss, s)
    dist.all_gather(dd, d)
    return torch.stack(ss).cpu().numpy(), [[int(v) for v in t.tolist()] for t in dd]


def parity_verdict(digests):
    """Parity verdict of a run from the gathered per-rank rows:
        [rank digest (3 ints), canary digest (3 ints), canary max |LSB difference| vs the reference's golden output or -1]
    The canary is local stream 0 of every rank: identical bytes and history everywhere, so its digest must agree
    across ranks; every rank's own comparison with the golden file must be within +-1 LSB."""
    rows = [list(r) for r in digests]
    canary = [tuple(r[3:6]) for r in rows]
    lsb = [int(r[6]) for r in rows]
    equal = all(c == canary[0] for c in canary)
    checked = all(v >= 0 for v in lsb)
    within = checked and all(v <= 1 for v in lsb)
    return {"ok": bool(equal and within), "canary_digests_equal": bool(equal), "ranks": len(rows),
            "canary_max_lsb_vs_reference": (max(lsb) if checked else None), "checked_against_golden": bool(checked)}


def aggregate(stats_all, steps):
    """Whole-job numbers from the gathered rows [device_ms, e2e_ms, samples_per_step]:
    time = max over ranks, samples = sum over ranks."""
    dev_ms = float(stats_all[:, 0].max())
    e2e_ms = float(stats_all[:, 1].max())
    total = float(stats_all[:, 2].sum()) * steps
    return {"dev_ms": dev_ms, "e2e_ms": e2e_ms, "total_samples": total,
            "value_msps": total / (dev_ms * 1e-3) / 1e6 if dev_ms > 0 else None,
            "e2e_msps": total / (e2e_ms * 1e-3) / 1e6 if e2e_ms > 0 else None}
