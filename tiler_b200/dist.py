"""Multi-GPU plumbing for the stages that shard: one process per GPU, torch.distributed for the only
exchange step the path has.

  * match / reconstruct : keyframe sequences are independent units (tilingencoder.pas:1496, 1668) -> whole sequences
    per rank, dictionary and palettes replicated, NO collective; tilemaps are gathered on the host.
  * k-means             : points sharded, centroids replicated, one all-reduce(sum) of the per-cluster sums [k, dim]
    and counts [k] per Lloyd iteration (NCCL over NVLink on GPUs; gloo in the CPU tests).
  * palette quantisation / dithering : palettes and (tile, palette) pairs are independent -> shard, no collective.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_sequences(frame_counts, world):
    """Greedy longest-first assignment of keyframe sequences to ranks. -> list (per rank) of sequence indices."""
    order = sorted(range(len(frame_counts)), key=lambda i: (-frame_counts[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += frame_counts[i]
    return [sorted(s) for s in out]


def shard_rows(n, rank, world):
    """Contiguous row range [lo, hi) of rank `rank` when n rows are split as evenly as possible."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_partials(sums, counts, changed, inertia):
    """Sum the per-rank k-means partials over the default process group (no-op without one). Tensors in, tensors out."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return sums, counts, changed, inertia
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    extra = torch.tensor([float(changed), float(inertia)], dtype=torch.float64, device=sums.device)
    dist.all_reduce(extra, op=dist.ReduceOp.SUM)
    return sums, counts, int(extra[0].item()), float(extra[1].item())


def kmeans_fit_sharded(x_shard, init, max_iter=300, nan_empty=False):
    """Lloyd with the points sharded over the ranks of the default process group (x_shard: this rank's CUDA rows,
    init: the same [k, dim] on every rank).  Same stopping rule as the single-GPU tm_kmeans_fit: stop when no label
    changed anywhere or after max_iter updates.  -> labels (this shard), centroids, inertia, iterations."""
    from . import api
    cent = init.clone()
    labels = torch.full((x_shard.shape[0],), -1, dtype=torch.int32, device=x_shard.device)
    it = 0
    while True:
        labels, sums, counts, changed, inertia = api.kmeans_partial_step(x_shard, cent, labels)
        sums, counts, changed, inertia = allreduce_partials(sums, counts, changed, inertia)
        if changed == 0 or it >= max_iter:
            break
        it += 1
        cent = api.kmeans_finish_step(sums, counts, cent, nan_empty=nan_empty)
    return labels, cent, inertia, it


def kmeans_fit_i16_sharded(x_shard, init, max_iter=300, nan_empty=False, check_every=4, timings=None):
    """Same as kmeans_fit_sharded for int16 tile vectors (config C).  The shard's points are split into limb rows once
    (api.KmeansI16Shard); every iteration is: assignment on the tensor cores + per-cluster partial sums (one library call, no
    host copy) -> all-reduce of the [k,192] f64 sums and of [k] counts + 2 counters (NCCL over NVLink) -> centroid update.
    The stopping test (no label changed on any rank) reads the all-reduced counter on the host only every `check_every`
    iterations: converged iterations in between change nothing, so the result equals the every-iteration test's.
    timings (optional dict): per-iteration device times in ms of the three phases, from CUDA events."""
    from . import api
    dev = x_shard.device
    k = int(init.shape[0])
    cent = init.clone()
    labels = torch.full((x_shard.shape[0],), -1, dtype=torch.int32, device=dev)
    shard = api.KmeansI16Shard(x_shard, k)
    sums = torch.empty((k, 192), dtype=torch.float64, device=dev)
    meta = torch.zeros(k + 2, dtype=torch.int64, device=dev)        # [counts (k) | labels changed | exact-scan points]
    multi = dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    ev = []
    it = 0
    inertia = torch.zeros(1, dtype=torch.float64, device=dev)
    while True:
        meta[k:].zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timings is not None else None
        if e: e[0].record()
        shard.step(cent, labels, sums=sums, counts=meta[:k], stats=meta[k:], inertia=inertia)
        if e: e[1].record()
        if multi:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            dist.all_reduce(meta, op=dist.ReduceOp.SUM)
        if e: e[2].record()
        last = it >= max_iter
        if last or it % check_every == 0:
            if int(meta[k].item()) == 0 or last:       # the only host read of the loop
                if e: e[3].record(); ev.append(e)
                break
        it += 1
        cent = api.kmeans_finish_step(sums, meta[:k], cent, nan_empty=nan_empty)
        if e: e[3].record(); ev.append(e)
    if multi:
        dist.all_reduce(inertia, op=dist.ReduceOp.SUM)
    shard.close()
    if timings is not None:
        torch.cuda.synchronize()
        timings["assign_ms"] = [a[0].elapsed_time(a[1]) for a in ev]
        timings["allreduce_ms"] = [a[1].elapsed_time(a[2]) for a in ev]
        timings["update_ms"] = [a[2].elapsed_time(a[3]) for a in ev]
    return labels, cent, float(inertia.item()), it


def gather_tilemaps(local, world_shards):
    """All-gather per-rank tilemap arrays (numpy, keyed by sequence index) onto every rank via the object collective."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    merged = {}
    for d in out:
        merged.update(d)
    assert sorted(merged) == sorted(i for s in world_shards for i in s)
    return merged


def gather_rows(local_rows, n_total):
    """All-gather row shards produced by shard_rows (numpy [hi - lo, ...]) into the full [n_total, ...] array on every rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert local_rows.shape[0] == n_total
        return local_rows
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local_rows)
    full = np.concatenate(out, axis=0)
    assert full.shape[0] == n_total
    return full


def device_collectives():
    """True inside an initialised NCCL group: row shards and tilemaps are then exchanged as device tensors over NVLink
    (disjoint rows into a zero-filled full-size tensor + one all-reduce(sum) = an all-gather that needs no equal shard sizes)
    instead of pickled host objects."""
    return (dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            and dist.get_backend() == "nccl")


def allgather_rows_device(local_rows, lo, n_total, device):
    """local_rows: this rank's rows [hi - lo, ...] (numpy or CUDA tensor) of a [n_total, ...] array -> the full array as a CUDA tensor
    on every rank."""
    loc = local_rows if torch.is_tensor(local_rows) else torch.from_numpy(np.ascontiguousarray(local_rows))
    full = torch.zeros((n_total,) + tuple(loc.shape[1:]), dtype=loc.dtype, device=device)
    full[lo:lo + loc.shape[0]] = loc.to(device, non_blocking=True)
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    return full


def world_info():
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(), dist.get_world_size()
