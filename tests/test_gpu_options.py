"""Option corners on seeded tables against the CPU oracle: --overlap <= 0 (every same-chromosome filling pair matches:
the ALLMATCH instantiations of k_pair / k_replay), edge_threshold <= 0 and huge, single-element cutoff lists, empty masks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,scale,kw", [
    ("C2", 0.05, dict(overlap=0.0)),
    ("C2", 0.05, dict(overlap=-1.0, edge_threshold=3)),
    ("C5", 0.004, dict(overlap=0.0, edge_threshold=10)),
    ("C2", 0.1, dict(edge_threshold=0)),
    ("C2", 0.1, dict(edge_threshold=10**9)),
    ("C2", 0.1, dict(jaccard_cutoffs="0.2", qlen_diff=0.5, n_alignment_diff=0.9)),
    ("C3", 0.05, dict(cluster_mask="")),
    ("C1", 1.0, dict(overlap=1.0, jaccard_cutoffs="1")),
])
def test_option_corners_match_oracle(name, scale, kw):
    from fslr_b200 import synth
    from fslr_b200.engine import get_engine
    from fslr_b200.table import ClusterParams, ColumnarTable
    from oracle import oracle as orc
    t = ColumnarTable.from_synth(synth.make_config(name, scale))
    opts = dict(cluster_mask=synth.CONFIG_MASK[name])
    opts.update(kw)
    p = ClusterParams.from_options(t, **opts)
    res = get_engine(0).cluster(t, p)
    ocl, onr, ost = orc.oracle_cluster(t, p)
    assert bool(ost["no_clusters"]) == res.no_clusters
    if not res.no_clusters:
        assert np.array_equal(res.cluster, ocl)
        assert np.array_equal(res.n_reads, onr)


def _check(t, p):
    from fslr_b200.engine import get_engine
    from oracle import oracle as orc
    res = get_engine(0).cluster(t, p)
    ocl, onr, ost = orc.oracle_cluster(t, p)
    assert bool(ost["no_clusters"]) == res.no_clusters
    if not res.no_clusters:
        assert np.array_equal(res.cluster, ocl)
        assert np.array_equal(res.n_reads, onr)
    return res


@pytest.mark.parametrize("T", [10, 2])
def test_many_fillings_per_read(T):
    """Reads with up to 12 fillings at scale: the general (> 4 fillings) paths of k_pair and the WALK replay."""
    from fslr_b200 import synth
    from fslr_b200.table import ClusterParams, ColumnarTable
    t = ColumnarTable.from_synth(synth.make_table(60_000, primers=("21q1", "17p6"), seed=77, naln_range=(2, 15), name="long"))
    res = _check(t, ClusterParams.from_options(t, edge_threshold=T))
    assert res.stats["saturating_reads"] > 0


def test_rows_in_arbitrary_order():
    """keep_fillings is defined by first/last OCCURRENCE of a qname (cluster.py:15-24): shuffle the rows of the table."""
    from fslr_b200 import synth
    from fslr_b200.table import ClusterParams, ColumnarTable
    t = ColumnarTable.from_synth(synth.make_config("C2", 0.2))
    perm = np.random.default_rng(11).permutation(t.n_rows)
    for k in ("read_id", "chrom", "rstart", "rend", "aln_size", "qstart", "qend", "n_alignments"):
        setattr(t, k, np.ascontiguousarray(getattr(t, k)[perm]))
    # read ids must stay dense in order of first appearance (the singleton numbering order)
    _, first = np.unique(t.read_id, return_index=True)
    order = np.argsort(first)
    remap = np.empty_like(order); remap[order] = np.arange(order.size)
    t.read_id = remap[t.read_id].astype(np.int32)
    _check(t, ClusterParams.from_options(t))


def test_everything_masked_and_tiny_tables():
    from fslr_b200 import synth
    from fslr_b200.table import ClusterParams, ColumnarTable
    t = ColumnarTable.from_synth(synth.make_config("C1", 0.2))
    p = ClusterParams.from_options(t, cluster_mask=",".join(str(c) for c in t.chrom_names))
    res = _check(t, p)
    assert res.no_clusters and res.stats["n_intervals"] == 0
    for n in (16, 17, 40):
        tt = ColumnarTable.from_synth(synth.make_table(n, seed=n))
        _check(tt, ClusterParams.from_options(tt))


def test_limit_errors():
    """Documented limits surface as error codes, never as wrong answers."""
    from fslr_b200._native import FslrError
    from fslr_b200.engine import get_engine
    from fslr_b200.table import ClusterParams, ColumnarTable
    eng = get_engine(0)

    def one_read(n_rows, naln=None, start0=10_000_000):
        naln = n_rows if naln is None else naln
        return ColumnarTable(read_id=np.zeros(n_rows, np.int32), chrom=np.zeros(n_rows, np.int32),
                             rstart=(start0 + 2000 * np.arange(n_rows)).astype(np.int32),
                             rend=(start0 + 500 + 2000 * np.arange(n_rows)).astype(np.int32), aln_size=np.full(n_rows, 500, np.int32),
                             qstart=(500 * np.arange(n_rows)).astype(np.int32), qend=(500 * np.arange(n_rows) + 500).astype(np.int32),
                             n_alignments=np.full(n_rows, naln, np.int32), n_reads=1, chrom_names=["chr1"],
                             chrom_len=np.array([10**8], np.int64))
    t = one_read(70)                                                            # 68 fillings > 64
    with pytest.raises(FslrError) as e:
        eng.cluster(t, ClusterParams.from_options(t))
    assert e.value.code == -4
    t = one_read(66)                                                            # 64 fillings: allowed, a lone read = singleton
    r = eng.cluster(t, ClusterParams.from_options(t))
    assert r.no_clusters
    t = one_read(5, naln=70000)
    with pytest.raises(FslrError) as e:
        eng.cluster(t, ClusterParams.from_options(t))
    assert e.value.code == -7
    t = one_read(5, start0=-5000)
    with pytest.raises(FslrError) as e:
        eng.cluster(t, ClusterParams.from_options(t))
    assert e.value.code == -7
    t = one_read(5)
    t.read_id = np.full(5, 3, np.int32)                                         # read id >= n_reads
    with pytest.raises(FslrError) as e:
        eng.cluster(t, ClusterParams.from_options(t))
    assert e.value.code == -7


def test_tsv_header_only_and_single_row():
    from fslr_b200 import tsv
    hdr = b"chrom\trstart\trend\tqname\tn_alignments\taln_size\tqstart\tqend\tstrand\tmapq\tqlen\talignment_score\n"
    pb = tsv.read_mappings_bed(hdr, {"chr1": 1000})
    assert pb.n_rows == 0 and pb.n_reads == 0 and pb.chrom_names == []
    pb.close()
    pb = tsv.read_mappings_bed(hdr + b"chr1\t10\t20\tr1\t3\t10\t0\t10\t+\t60\t30\t20", {"chr1": 1000})
    assert (pb.n_rows, pb.n_reads, pb.chrom_names) == (1, 1, ["chr1"])
    assert pb.column("rend").tolist() == [20] and list(pb.qnames()) == ["r1"]
    res = pb.cluster()
    assert res.no_clusters
    pb.close()
