"""Exchange steps of the multi-GPU path, written against torch.distributed only so the same code
runs over NCCL/NVLink on the GPUs and over gloo in the CPU tests.

The interval table is replicated; candidate generation and the pair tests are sharded by query read.  With host inputs the
upload is sharded too (gather_column_inplace).  Three exchanges exist on the data path:
  1. sum-all-reduce of the per-read partner counters (decides which reads are saturating), 4 bytes per query read;
  2. all-gather of the recorded pairs of the saturating reads (8 bytes per pair; the replay runs replicated);
  3. all-gather of every rank's spanning forest (<= n_query_reads - 1 edges each), then one final union-find.
"""
import torch
import torch.distributed as dist


def exchange_counts(counts, group=None):
    """In-place sum over ranks of the int32 per-read counts."""
    if dist.is_initialized() and dist.get_world_size(group) > 1 and counts.numel() > 0:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def exchange_forests(local_edges, group=None):
    """local_edges: int32 [n, 2] (or flat [2n]) tensor of this rank's forest edges.
    Returns the concatenation over ranks as a flat int32 [2 * total] tensor plus the per-rank edge counts."""
    flat = local_edges.reshape(-1).contiguous()
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return flat, [int(flat.numel() // 2)]
    world = dist.get_world_size(group)
    dev = flat.device
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([flat.numel()], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, mine, group=group)
    sizes_h = [int(x) for x in sizes.cpu().tolist()]
    mx = max(max(sizes_h), 2)
    send = torch.zeros(mx, dtype=torch.int32, device=dev)
    send[:flat.numel()] = flat
    recv = torch.empty(world * mx, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = [recv[r * mx: r * mx + sizes_h[r]] for r in range(world)]
    return torch.cat(parts).contiguous(), [s // 2 for s in sizes_h]


def row_slice(n_rows, rank, world):
    """Rows [lo, hi) of the table that rank `rank` uploads; chunks are equal (the last one is padded in the exchange)."""
    chunk = (n_rows + world - 1) // world
    lo = min(rank * chunk, n_rows)
    return lo, min(lo + chunk, n_rows), chunk


def gather_column_inplace(src, dst, n, rank, world, group=None):
    """src: host tensor [>= n] (pinned on the GPU path); dst: tensor [chunk * world] on this rank's device (a CPU tensor under
    gloo).  This rank moves elements row_slice(rank) of the column into their place in dst; an all-gather whose send buffer
    IS that slice of dst (NCCL's in-place form) completes the column on every rank: no staging buffer, no copy afterwards."""
    lo, hi, chunk = row_slice(n, rank, world)
    if world == 1 or not dist.is_initialized():
        dst[:n].copy_(src[:n], non_blocking=True)
        return
    mine = dst[rank * chunk:(rank + 1) * chunk]
    if hi > lo:
        mine[:hi - lo].copy_(src[lo:hi], non_blocking=True)
    if not dst.is_cuda:
        mine = mine.clone()                                   # (gloo: keep send and receive buffers apart)
    dist.all_gather_into_tensor(dst.view(torch.uint8), mine.view(torch.uint8), group=group)   # bytes: any column width


def shard_of_read(q, world):
    """Rank owning query read q (query rank): groups of 256 consecutive query ranks, round robin (k_pair)."""
    return (q >> 8) % world
