"""Parity of the CUDA path (through the C ABI) with the reference: golden fixtures produced by the unmodified
cluster.py, and the oracle on seeded inputs.  Bit-exact: integer cluster ids and sizes."""
import numpy as np
import pytest

from tests import golden_io

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from fslr_b200.engine import get_engine
    return get_engine(0)


def _check_case(engine, case):
    res = engine.cluster(case.table, case.params, order=case.order)
    assert res.no_clusters == case.no_clusters, case.name
    if not case.no_clusters:
        ecl, enr = case.expected
        bad = np.nonzero(res.cluster != ecl)[0]
        assert bad.size == 0, "%s: %d reads differ, first %s" % (case.name, bad.size, bad[:5])
        assert np.array_equal(res.n_reads, enr), case.name


def test_golden_small_cases(engine):
    cases = golden_io.load("small_cases.npz")
    failed = []
    for c in cases:
        try:
            _check_case(engine, c)
        except AssertionError as e:
            failed.append(str(e))
    assert not failed, "%d/%d cases differ: %s" % (len(failed), len(cases), failed[:5])


def test_golden_wide_cases(engine):
    """> 4 fillings per read (general device paths), --overlap <= 0, many saturating reads."""
    cases = golden_io.load("wide_cases.npz")
    failed = []
    for c in cases:
        try:
            _check_case(engine, c)
        except AssertionError as e:
            failed.append(str(e))
    assert not failed, "%d/%d cases differ: %s" % (len(failed), len(cases), failed[:5])


@pytest.mark.parametrize("idx", range(8))
def test_golden_config_cases(engine, idx):
    cases = golden_io.load("config_cases.npz")
    _check_case(engine, cases[idx])


@pytest.mark.parametrize("name,scale,T", [("C1", 1.0, 10), ("C2", 1.0, 10), ("C2", 0.2, 1), ("C2", 0.2, 3), ("C3", 0.1, 10),
                                          ("C5", 0.02, 10), ("C5", 0.02, 2), ("C4", 0.02, 10)])
def test_oracle_stable_order(engine, name, scale, T):
    """Production mode (GPU stable sort) against the oracle with the same tie rule."""
    from fslr_b200 import synth
    from fslr_b200.table import ClusterParams, ColumnarTable
    from oracle import oracle as orc
    t = ColumnarTable.from_synth(synth.make_config(name, scale))
    p = ClusterParams.from_options(t, cluster_mask=synth.CONFIG_MASK[name], edge_threshold=T)
    res = engine.cluster(t, p)
    ocl, onr, ost = orc.oracle_cluster(t, p)
    assert np.array_equal(res.cluster, ocl)
    assert np.array_equal(res.n_reads, onr)
    assert res.stats["components"] == ost["components"]
    assert res.stats["n_intervals"] == ost["n_data"] and res.stats["n_fillings"] == ost["n_fillings"]


def test_resident_path_matches_host_path(engine):
    from fslr_b200 import synth
    from fslr_b200.engine import DeviceTable
    from fslr_b200.table import ClusterParams, ColumnarTable
    t = ColumnarTable.from_synth(synth.make_config("C1"))
    p = ClusterParams.from_options(t)
    a = engine.cluster(t, p)
    d = DeviceTable(t, engine.device)
    engine.run_resident(d, t, p)
    assert np.array_equal(d.out_cluster[:t.n_reads].cpu().numpy(), a.cluster)
    assert np.array_equal(d.out_n_reads[:t.n_reads].cpu().numpy(), a.n_reads)


def test_errors(engine):
    from fslr_b200 import synth
    from fslr_b200._native import FslrError
    from fslr_b200.table import ClusterParams, ColumnarTable
    t = ColumnarTable.from_synth(synth.make_config("C1", 0.1))
    p = ClusterParams.from_options(t)
    bad = ColumnarTable(**{**t.__dict__})
    bad.aln_size = t.aln_size.copy(); bad.aln_size[:] = 0
    with pytest.raises(FslrError) as e:
        engine.cluster(bad, p)
    assert e.value.code == -3
    mixed = ColumnarTable(**{**t.__dict__})
    mixed.n_alignments = t.n_alignments.copy()
    rows = np.nonzero(t.n_alignments >= 4)[0]
    mixed.n_alignments[rows[1]] += 1                               # a filling row whose n_alignments differs from its read's
    with pytest.raises(FslrError) as e:
        engine.cluster(mixed, p)
    assert e.value.code == -5
    empty = ColumnarTable(**{**t.__dict__})
    for k in ("read_id", "chrom", "rstart", "rend", "aln_size", "qstart", "qend", "n_alignments"):
        setattr(empty, k, np.zeros(0, np.int32))
    empty.n_reads = 0
    r = engine.cluster(empty, p)
    assert r.no_clusters and r.cluster.shape[0] == 0


def test_host_pipeline_concurrent_contexts():
    """Two library contexts in flight on one device (HostPipeline) with DIFFERENT tables and options: no state is shared
    between calls, every result equals the oracle's."""
    from fslr_b200 import synth
    from fslr_b200.engine import HostPipeline, PinnedTable
    from fslr_b200.table import ClusterParams, ColumnarTable
    from oracle import oracle as orc
    jobs = []
    for name, scale, kw in (("C2", 0.3, dict()), ("C3", 0.05, dict(jaccard_cutoffs="0.34", edge_threshold=3)),
                            ("C5", 0.01, dict(overlap=0.5)), ("C1", 1.0, dict(jaccard_cutoffs="1,1,1"))):
        t = ColumnarTable.from_synth(synth.make_config(name, scale))
        p = ClusterParams.from_options(t, cluster_mask=synth.CONFIG_MASK[name], **kw)
        jobs.append((t, p, PinnedTable(t)))
    pipe = HostPipeline(0, depth=2)
    for _ in range(3):
        futs = [pipe.submit(pt, t, p) for t, p, pt in jobs]
        for f in futs:
            f.result()
        for t, p, pt in jobs:
            ocl, onr, _ = orc.oracle_cluster(t, p)
            assert np.array_equal(pt.out_cluster[:t.n_reads].numpy(), ocl)
            assert np.array_equal(pt.out_n_reads[:t.n_reads].numpy(), onr)
    pipe.close()


def test_narrow_wire_columns_match():
    """The wire format (PinnedTable(compact=True): narrow columns, derived aln_size, run lengths) gives the int32 result, through the
    host-buffer call and through the device-resident call (narrow columns already in HBM)."""
    from fslr_b200 import synth
    from fslr_b200.engine import PinnedTable, get_engine
    from fslr_b200.table import ClusterParams, ColumnarTable
    t = ColumnarTable.from_synth(synth.make_config("C3", 0.05))
    p = ClusterParams.from_options(t, cluster_mask=synth.CONFIG_MASK["C3"])
    eng = get_engine(0)
    wide, narrow = PinnedTable(t), PinnedTable(t, compact=True)
    assert set(narrow.narrow) == {"chrom_u8", "n_alignments_u16", "rspan_i16", "qstart_u16", "qend_u16"} and set(narrow.cols) == {"rstart"}
    assert narrow.aln_is_qspan and narrow.rows_per_read is not None and narrow.h2d_bytes == 13 * t.n_rows + t.n_reads
    eng.run_host(wide, t, p)
    eng.run_host(narrow, t, p)
    assert np.array_equal(wide.out_cluster[:t.n_reads].numpy(), narrow.out_cluster[:t.n_reads].numpy())
    assert np.array_equal(wide.out_n_reads[:t.n_reads].numpy(), narrow.out_n_reads[:t.n_reads].numpy())
    from fslr_b200.engine import DeviceWireTable
    dw = DeviceWireTable(narrow, eng.device, 1)
    dw.upload(narrow, 0, 1)
    eng.run_resident(dw, t, p)
    assert np.array_equal(wide.out_cluster[:t.n_reads].numpy(), dw.out_cluster[:t.n_reads].cpu().numpy())
    assert np.array_equal(wide.out_n_reads[:t.n_reads].numpy(), dw.out_n_reads[:t.n_reads].cpu().numpy())


def test_late_column_upload_and_its_error_path():
    """Host calls with the wire format upload n_alignments behind the other columns on a second stream (tables of >= 2^20 rows;
    resolve_columns / late_columns): same result as the int32 columns, and a call that fails before the column's first reader
    (here: a negative coordinate found by keep_fillings) neither hangs nor disturbs the next call."""
    import pytest
    from fslr_b200 import synth
    from fslr_b200._native import FslrError
    from fslr_b200.engine import PinnedTable, get_engine
    from fslr_b200.table import ClusterParams, ColumnarTable
    t = ColumnarTable.from_synth(synth.make_config("C3", 0.3))
    assert t.n_rows >= (1 << 20)
    p = ClusterParams.from_options(t, cluster_mask=synth.CONFIG_MASK["C3"])
    eng = get_engine(0)
    wide, narrow = PinnedTable(t), PinnedTable(t, compact=True)
    eng.run_host(wide, t, p)
    eng.run_host(narrow, t, p)
    assert np.array_equal(wide.out_cluster[:t.n_reads].numpy(), narrow.out_cluster[:t.n_reads].numpy())
    assert np.array_equal(wide.out_n_reads[:t.n_reads].numpy(), narrow.out_n_reads[:t.n_reads].numpy())
    keep = narrow.cols["rstart"][:40].clone()
    narrow.cols["rstart"][:40] = -100000                                        # (40 rows: some of them are fillings)
    with pytest.raises(FslrError):
        eng.run_host(narrow, t, p)
    narrow.cols["rstart"][:40] = keep
    narrow.out_cluster.zero_()
    eng.run_host(narrow, t, p)
    assert np.array_equal(wide.out_cluster[:t.n_reads].numpy(), narrow.out_cluster[:t.n_reads].numpy())
