"""Exchange steps of the multi-GPU path, written against torch.distributed only so the same code
runs over NCCL/NVLink on the GPUs and over gloo in the CPU tests.

The interval table is replicated; the pair space is sharded.  Two exchanges exist:
  1. sum-all-reduce of the per-read passing-candidate counts (decides which reads are saturating);
  2. all-gather of every rank's spanning forest (<= n_query_reads - 1 edges each), then one final union-find.
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


def shard_of_position(i, world):
    """Rank owning sorted interval position i: tiles of 256 consecutive positions, round robin (k_pair)."""
    return (i >> 8) % world
