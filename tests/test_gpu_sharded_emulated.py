"""The multi-GPU C-ABI stages (fslrc_mg_*) with world_size 2 and 3 EMULATED on one GPU: one context per rank, the
ranks' kernels run one after the other (nothing waits across contexts) and the three exchange steps — sum of the per-read
partner counters, concatenation of the recorded pairs of saturating reads, concatenation of the spanning forests — are done
here with torch ops standing in for the NCCL all-reduce / all-gathers of fslr_b200/sharded.py.  Every rank must end with the single-GPU (= oracle) result."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_emulated(table, params, world):
    from fslr_b200 import _native
    from fslr_b200.engine import DeviceTable, Engine, _DevView
    engines = [Engine(0) for _ in range(world)]
    dev = engines[0].device
    dtabs = [DeviceTable(table, dev) for _ in range(world)]
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ps = [e._params(table, params) for e in engines]
    ts = [Engine._table(d) for d in dtabs]
    for e, t, p in zip(engines, ts, ps):
        e._check(e.lib.fslrc_mg_prepare(e.ctx, C.byref(t), C.byref(p), stream))
    counts = []
    for r, e in enumerate(engines):
        ptr, n = C.c_void_p(), C.c_int64()
        e._check(e.lib.fslrc_mg_pair(e.ctx, r, world, C.byref(ptr), C.byref(n)))
        counts.append(torch.as_tensor(_DevView(ptr.value, (n.value,)), device=dev) if n.value > 0 else None)
    if counts[0] is not None:
        total = torch.stack([c.clone() for c in counts]).sum(0).to(torch.int32)      # all-reduce SUM
        for c in counts:
            c.copy_(total)
    torch.cuda.synchronize()
    pairs = []
    for r, e in enumerate(engines):
        ptr, n = C.c_void_p(), C.c_int64()
        e._check(e.lib.fslrc_mg_partners(e.ctx, r, world, C.byref(ptr), C.byref(n)))
        pairs.append(torch.as_tensor(_DevView(ptr.value, (2 * n.value,)), device=dev).clone() if n.value > 0
                     else torch.zeros(0, dtype=torch.int32, device=dev))
    allp = torch.cat(pairs).contiguous()                                                 # all-gather
    torch.cuda.synchronize()
    forests = []
    for r, e in enumerate(engines):
        ptr, n = C.c_void_p(), C.c_int64()
        e._check(e.lib.fslrc_mg_replay(e.ctx, r, world, allp.data_ptr() if allp.numel() else None, allp.numel() // 2,
                                       C.byref(ptr), C.byref(n)))
        forests.append(torch.as_tensor(_DevView(ptr.value, (2 * n.value,)), device=dev).clone() if n.value > 0
                       else torch.zeros(0, dtype=torch.int32, device=dev))
    allf = torch.cat(forests).contiguous()                                               # all-gather
    out = []
    for e, d in zip(engines, dtabs):
        st = _native.Stats()
        e._check(e.lib.fslrc_mg_finish(e.ctx, allf.data_ptr() if allf.numel() else None, allf.numel() // 2,
                                       d.out_cluster.data_ptr(), d.out_n_reads.data_ptr(), C.byref(st)))
        n = table.n_reads
        out.append((d.out_cluster[:n].cpu().numpy(), d.out_n_reads[:n].cpu().numpy(), st.as_dict(e.lib)))
    for e in engines:
        e.close()
    return out


@pytest.mark.parametrize("name,scale,T,world", [("C2", 0.3, 10, 2), ("C2", 0.1, 2, 3), ("C5", 0.02, 10, 2), ("C3", 0.05, 10, 2)])
def test_emulated_ranks_match_oracle(name, scale, T, world):
    from fslr_b200 import synth
    from fslr_b200.table import ClusterParams, ColumnarTable
    from oracle import oracle as orc
    t = ColumnarTable.from_synth(synth.make_config(name, scale))
    p = ClusterParams.from_options(t, cluster_mask=synth.CONFIG_MASK[name], edge_threshold=T)
    ocl, onr, ost = orc.oracle_cluster(t, p)
    for cl, nr, st in _run_emulated(t, p, world):
        assert np.array_equal(cl, ocl)
        assert np.array_equal(nr, onr)
        assert st["components"] == ost["components"]
