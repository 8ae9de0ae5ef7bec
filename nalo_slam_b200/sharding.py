"""Multi-GPU sharding of the naturally parallel parts of the path.

Only two things shard: the motion candidates of FullSystem::trackNewCoarse (FullSystem.cpp:502-699) and batches of
independent frame-pair alignments. Both are embarrassingly parallel: every rank tracks its share on its own GPU with no
data-path collective, and ONE small all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests) collects the
per-candidate / per-pair records (32 doubles each). The sequential winner rule is then replayed on the gathered
records (nalo_winner_rule), which reproduces the reference's sequential loop exactly.
"""
from __future__ import annotations

import numpy as np

REC = 32  # doubles per record


def shard_range(n: int, rank: int, world: int):
    """Contiguous block partition; the first n % world ranks get one extra unit."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def pack_records(res: dict) -> np.ndarray:
    """track_multi / Batch.track result dict -> [n, 32] float64 records."""
    n = len(res["ok"])
    out = np.zeros((n, REC))
    out[:, 0] = res["ok"]
    out[:, 1:8] = res["poses"]
    out[:, 8:10] = res["affs"]
    out[:, 10:15] = res["lastRes"]
    if "flow" in res:
        out[:, 15:18] = res["flow"]
    if "pass_lvl" in res:
        out[:, 18:24] = res["pass_lvl"]
        out[:, 24:30] = res["pass_res"]
    return out


def unpack_records(rec: np.ndarray) -> dict:
    rec = np.asarray(rec, dtype=np.float64)
    return dict(
        ok=np.ascontiguousarray(rec[:, 0].astype(np.int32)),
        poses=np.ascontiguousarray(rec[:, 1:8]),
        affs=np.ascontiguousarray(rec[:, 8:10]),
        lastRes=np.ascontiguousarray(rec[:, 10:15]),
        flow=np.ascontiguousarray(rec[:, 15:18]),
        pass_lvl=np.ascontiguousarray(rec[:, 18:24].astype(np.int32)),
        pass_res=np.ascontiguousarray(rec[:, 24:30]),
    )


def all_gather_records(local: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """One all_gather of the per-unit records; every rank returns the full [n_total, 32] array in unit order.
    Ranks are padded to the largest shard so a single fixed-size collective suffices."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    sizes = shard_sizes(n_total, world)
    m = max(sizes)
    buf = torch.zeros((m, REC), dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = [out[r][: sizes[r]].cpu().numpy() for r in range(world)]
    return np.concatenate(parts, axis=0)


class _DevArray:
    """Minimal __cuda_array_interface__ wrapper so torch can view a raw device pointer owned by the library."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


def all_gather_device_records(dev_ptr: int, n_local: int, n_total: int, device, width: int = 16) -> np.ndarray:
    """all_gather of per-unit records that already sit in device memory (nalo_batch_results_dev: float64 [n_local][16]),
    without the host round trip of all_gather_records: one device copy into the padded send buffer, one NCCL
    all_gather_into_tensor over NVLink, one D2H of the gathered block."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    sizes = shard_sizes(n_total, world)
    m = max(sizes)
    # send / receive / pinned host buffers are kept between calls (a gather per step must not pay three allocations)
    key = (m, width, world, str(device))
    bufs = _gather_cache.get(key)
    if bufs is None:
        bufs = (torch.zeros((m, width), dtype=torch.float64, device=device), torch.empty((world * m, width), dtype=torch.float64, device=device),
                torch.empty((world * m, width), dtype=torch.float64).pin_memory())
        _gather_cache.clear()
        _gather_cache[key] = bufs
    buf, out, host_t = bufs
    if n_local:
        buf[:n_local].copy_(torch.as_tensor(_DevArray(dev_ptr, (n_local, width)), device=device))
    dist.all_gather_into_tensor(out, buf)
    host_t.copy_(out, non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    host = host_t.numpy().reshape(world, m, width)
    return np.concatenate([host[r, : sizes[r]] for r in range(world)], axis=0)


_gather_cache = {}
