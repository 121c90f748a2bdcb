"""Host-side plumbing for one-process-per-GPU runs (torch.distributed; NCCL on the GPU box, gloo in
the CPU tests).  The hot path shards three ways:

  * chains  -- independent units, no communication while sampling;
  * samples -- the predictor splits stored samples across ranks and merges per-row
               (count, mean, M2) once at the end (Chan et al. pairwise update);
  * rows    -- one chain, training rows split across ranks, one all-reduce of the likelihood
               partial gradient per leapfrog step inside libtbnn.so (tbnn_comm_init); host code
               only distributes the ncclUniqueId and the identical seeds.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous balanced block [lo, hi) of n units for `rank`."""
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi


def merge_moments(count, mean, m2, group=None):
    """Chan/Welford merge of per-rank (count, mean, M2) tensors of identical shape across all ranks.
    Every rank returns the merged triple (identical bits: the merge runs in rank order)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return count, mean, m2
    world = dist.get_world_size(group)
    pack = torch.stack([count.to(mean.dtype), mean, m2])
    gathered = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(gathered, pack, group=group)
    n, mu, s = gathered[0][0].clone(), gathered[0][1].clone(), gathered[0][2].clone()
    for r in range(1, world):
        nb, mb, sb = gathered[r][0], gathered[r][1], gathered[r][2]
        tot = n + nb
        delta = mb - mu
        safe = torch.where(tot > 0, tot, torch.ones_like(tot))
        mu = mu + delta * nb / safe
        s = s + sb + delta * delta * n * nb / safe
        n = tot
    return n, mu, s


def merge_moments_reduce(count, mean, m2, group=None):
    """The same merge as ONE all-reduce: (n, n*mean, M2 + n*mean^2) are plain sums over ranks (carried in float64, so
    the subtraction M2 = S2 - N*mean^2 loses nothing at float32 output precision).  No gather, no loop over ranks:
    on NVLink / NVSwitch the 3*M*8 bytes reduce in one collective whatever the number of ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return count, mean, m2
    n = count.to(torch.float64)
    mu = mean.to(torch.float64)
    pack = torch.stack([n, n * mu, m2.to(torch.float64) + n * mu * mu])
    dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    tot = pack[0]
    safe = torch.where(tot > 0, tot, torch.ones_like(tot))
    mean_t = pack[1] / safe
    m2_t = torch.clamp(pack[2] - tot * mean_t * mean_t, min=0.0)
    return tot.to(mean.dtype), mean_t.to(mean.dtype), m2_t.to(mean.dtype)


def broadcast_bytes(payload, src=0, group=None, device=None):
    """Broadcasts a bytes object of known length (e.g. the 128-byte ncclUniqueId) from `src`."""
    n = len(payload)
    t = torch.tensor(list(payload), dtype=torch.uint8, device=device) if dist.get_rank(group) == src \
        else torch.zeros(n, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def attach_row_sharding(engine, group=None, device=None):
    """Creates the library's own NCCL communicator for row-sharded sampling: rank 0 makes the
    ncclUniqueId, torch.distributed carries it to the other ranks, every rank calls tbnn_comm_init."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = engine.comm_unique_id() if rank == 0 else bytes(128)
    uid = broadcast_bytes(uid, 0, group, device)
    engine.comm_init(uid, rank, world)
    return rank, world
