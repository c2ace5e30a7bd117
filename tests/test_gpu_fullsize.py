"""Full-size configs (BASELINE.json configs 3-5) on the GPU against the CPU oracle — bit-exact ids — plus
size-independent properties of the result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _props(res, n_reads):
    cl, nr = res.cluster, res.n_reads
    assert cl.min() == 0 and cl.max() == n_reads - 1 - (int((nr > 1).sum()) - int(res.stats["components"]))
    sizes = np.bincount(cl)
    assert np.array_equal(sizes[cl], nr)                       # n_reads is the size of the read's cluster
    ncl = int(res.stats["components"])
    assert (sizes[:ncl] >= 2).all() and (sizes[ncl:] == 1).all()   # clusters first, singletons after
    single = np.nonzero(nr == 1)[0]
    assert np.array_equal(cl[single], ncl + np.arange(single.size))   # singletons numbered in bed order


@pytest.mark.parametrize("name", ["C3", "C5", "C4"])
def test_fullsize_vs_oracle(name):
    from fslr_b200 import synth
    from fslr_b200.engine import get_engine
    from fslr_b200.table import ClusterParams, ColumnarTable
    from oracle import oracle as orc
    t = ColumnarTable.from_synth(synth.make_config(name))
    p = ClusterParams.from_options(t, cluster_mask=synth.CONFIG_MASK[name])
    res = get_engine(0).cluster(t, p)
    print(name, {k: v for k, v in res.stats.items() if k != "stage_ms"}, {k: round(v, 2) for k, v in res.stats["stage_ms"].items()})
    _props(res, t.n_reads)
    ocl, onr, ost = orc.oracle_cluster(t, p)
    assert np.array_equal(res.cluster, ocl)
    assert np.array_equal(res.n_reads, onr)
    assert res.stats["components"] == ost["components"]
    # idempotence: a second run gives the same answer (atomics only change internal edge order)
    res2 = get_engine(0).cluster(t, p)
    assert np.array_equal(res2.cluster, res.cluster)


@pytest.fixture(scope="module")
def c3_table():
    from fslr_b200 import synth
    from fslr_b200.table import ColumnarTable
    return ColumnarTable.from_synth(synth.make_config("C3"))


def _sweep():
    from fslr_b200 import synth
    return list(synth.C3_CUTOFF_SWEEP)


@pytest.mark.parametrize("cut", _sweep())
def test_c3_fullsize_cutoff_sweep(c3_table, cut):
    """BASELINE config 3 at full size (1M reads, --cluster-mask subtelomere,L1_TALEN) under every --jaccard-cutoffs list of the
    sweep (main.py:219), bit-exact against the oracle."""
    from fslr_b200 import synth
    from fslr_b200.engine import get_engine
    from fslr_b200.table import ClusterParams
    from oracle import oracle as orc
    p = ClusterParams.from_options(c3_table, cluster_mask=synth.CONFIG_MASK["C3"], jaccard_cutoffs=cut)
    res = get_engine(0).cluster(c3_table, p)
    _props(res, c3_table.n_reads)
    ocl, onr, ost = orc.oracle_cluster(c3_table, p)
    assert np.array_equal(res.cluster, ocl)
    assert np.array_equal(res.n_reads, onr)
    assert res.stats["components"] == ost["components"]


def test_c5_job_board_is_repeatable():
    """C5's giant clique drives the WALK replay's job board (kernels_replay.cuh: wide_job_* / wide_help): warps without tickets
    work off chunks of the tail walks, with hand-overs through global memory whose timing differs from run to run.  The result
    must not: eight more runs of the same table give the ids of the first one (which test_fullsize_vs_oracle pins to the oracle)."""
    from fslr_b200 import synth
    from fslr_b200.engine import get_engine
    from fslr_b200.table import ClusterParams, ColumnarTable
    t = ColumnarTable.from_synth(synth.make_config("C5"))
    p = ClusterParams.from_options(t, cluster_mask=synth.CONFIG_MASK["C5"])
    eng = get_engine(0)
    first = eng.cluster(t, p)
    _props(first, t.n_reads)
    assert first.stats["saturating_reads"] > 500000                # the hotspot is there
    for _ in range(8):
        res = eng.cluster(t, p)
        assert np.array_equal(res.cluster, first.cluster)
        assert np.array_equal(res.n_reads, first.n_reads)
        assert res.stats["edges"] == first.stats["edges"]
