"""world_size-2 gloo test (CPU) of the multi-GPU exchange logic: each rank evaluates the pair relation for its shard of
query reads (tests/proto_model.py stands in for the device kernels), counts are sum-all-reduced,
local spanning forests are all-gathered, and the merged components must equal the oracle's clusters."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fslr_b200 import synth
    from fslr_b200.sharded import exchange_counts, exchange_forests, gather_column_inplace, row_slice, shard_of_read
    from fslr_b200.table import ClusterParams, ColumnarTable
    from tests.proto_model import Model
    t = ColumnarTable.from_synth(synth.make_config("C1", 0.12))
    p = ClusterParams.from_options(t, edge_threshold=3)
    m = Model(t, p)
    # ---- sharded upload: each rank holds 1/world of the rows, an all-gather completes every column (DeviceWireTable.upload)
    host = {"rstart": torch.from_numpy(t.rstart.copy()), "chrom_u8": torch.from_numpy(t.chrom.astype(np.uint8)),
            "rspan_i16": torch.from_numpy((t.rend - t.rstart).astype(np.int16))}
    for k, src in host.items():
        dst = torch.zeros(row_slice(t.n_rows, 0, world)[2] * world, dtype=src.dtype)
        gather_column_inplace(src, dst, t.n_rows, rank, world)
        assert torch.equal(dst[:t.n_rows], src), k
    # ---- sharded phase A: this rank evaluates only the (a, b) pairs of the query reads `a` it owns
    m.degub = np.zeros(m.Q, np.int64)
    m.later, m.cond = [], []
    for a in range(m.Q):
        done = set()
        for f in m.lists[a]:
            for pos in m.candidates_at(f):
                b = m.q[m.sidx[pos]]
                if b == a or b in done:
                    continue
                done.add(b)
                if shard_of_read(a, world) != rank:
                    continue
                if m.passes(a, b)[1]:
                    m.degub[a] += 1
                    (m.later if b > a else m.cond).append((a, b))
    counts = torch.from_numpy(m.degub.astype(np.int32))
    exchange_counts(counts)
    m.degub = counts.numpy().astype(np.int64)
    m.isP = m.degub >= m.Tedge
    # ---- replay replicated (sequential), edges of the replay contributed by rank 0 only
    m.stop = np.array([m.chrom_lo[c] for c in m.chrom], dtype=np.int64)
    m.final = ~m.isP
    P = [a for a in range(m.Q) if m.isP[a]]
    pedges = []
    for a in P:                                   # ascending rank: every dependency is final
        my, dep = m.replay(a, emit=pedges)
        assert not dep
        for f, s in my.items():
            m.stop[f] = s
        m.final[a] = True
    edges = [(a, b) for (a, b) in m.later if not m.isP[a]]
    edges += [(a, b) for (a, b) in m.cond if not m.isP[a] and m.isP[b] and not m.visited(b, a)]
    if rank == 0:
        edges += pedges
    # ---- local union-find -> spanning forest -> all-gather -> final union-find
    parent = list(range(m.Q))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]; x = parent[x]
        return x
    for a, b in edges:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    forest = [(x, find(x)) for x in range(m.Q) if find(x) != x]
    loc = torch.tensor(forest, dtype=torch.int32).reshape(-1, 2)
    allf, per_rank = exchange_forests(loc)
    assert sum(per_rank) * 2 == allf.numel() and per_rank[rank] == len(forest)
    parent = list(range(m.Q))
    ing = np.zeros(m.Q, bool)
    for a, b in allf.reshape(-1, 2).tolist():
        ing[a] = ing[b] = True
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    root = np.array([find(x) for x in range(m.Q)])
    if rank == 0:
        np.save(out, np.stack([root, ing.astype(np.int64), m.rid_of_q]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_oracle(tmp_path):
    from fslr_b200 import synth
    from fslr_b200.table import ClusterParams, ColumnarTable
    from oracle import oracle as orc
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    root, ing, rid_of_q = np.load(out)
    t = ColumnarTable.from_synth(synth.make_config("C1", 0.12))
    p = ClusterParams.from_options(t, edge_threshold=3)
    ocl, onr, ost = orc.oracle_cluster(t, p)
    # same partition of the clustered reads, and components ordered by their smallest query rank
    ncl = ost["components"]
    roots = sorted(set(root[ing == 1].tolist()))
    assert len(roots) == ncl
    cid = {r: i for i, r in enumerate(roots)}
    for q in range(root.shape[0]):
        if ing[q]:
            assert ocl[rid_of_q[q]] == cid[root[q]]
        else:
            assert ocl[rid_of_q[q]] >= ncl
