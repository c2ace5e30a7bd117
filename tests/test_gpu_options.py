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
