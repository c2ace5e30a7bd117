"""Worker of tests/test_gpu_multirank.py: one process per GPU under torchrun, real NCCL.  Every rank uploads its slice of the
HOST columns (DeviceWireTable.upload), runs the sharded step (Engine.run_sharded) and compares its own copy of the result with
the CPU oracle; then the checksum of the results is compared across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fslr_b200 import synth
    from fslr_b200.engine import DeviceWireTable, Engine, PinnedTable
    from fslr_b200.table import ClusterParams, ColumnarTable
    from oracle import oracle as orc
    eng = Engine(local)
    cases = [("C2", 1.0, dict()), ("C3", 0.2, dict(jaccard_cutoffs="0.34")), ("C5", 0.05, dict()), ("C2", 0.3, dict(edge_threshold=2)),
             ("C1", 1.0, dict(overlap=0.0))]
    for name, scale, kw in cases:
        t = ColumnarTable.from_synth(synth.make_config(name, scale))
        p = ClusterParams.from_options(t, cluster_mask=synth.CONFIG_MASK[name], **kw)
        ocl, onr, ost = orc.oracle_cluster(t, p)
        for compact in (False, True):
            ptab = PinnedTable(t, compact=compact)
            dtab = DeviceWireTable(ptab, eng.device, world)
            dtab.upload(ptab, rank, world)
            st = eng.run_sharded(dtab, t, p, rank, world)
            cl, nr = dtab.out_cluster[:t.n_reads].cpu().numpy(), dtab.out_n_reads[:t.n_reads].cpu().numpy()
            assert np.array_equal(cl, ocl), "rank %d %s compact=%s: cluster ids differ from the oracle" % (rank, name, compact)
            assert np.array_equal(nr, onr), "rank %d %s: n_reads differ" % (rank, name)
            assert st["components"] == ost["components"]
            chk = torch.tensor([int((cl.astype(np.int64) * np.arange(1, cl.size + 1)).sum())], device=eng.device)
            allc = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(allc, chk)
            assert all(int(c) == int(allc[0]) for c in allc)
        if rank == 0:
            print("mg ok", name, scale, kw, "world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MG_WORKER_OK", flush=True)


if __name__ == "__main__":
    main()
