"""cluster.choose_alignment (cluster.py:237-254): golden vectors from the unmodified reference
(tests/golden/make_golden_choose.py) against the numpy oracle (CPU) and the GPU path through the C ABI."""
import json
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    z = np.load(os.path.join(HERE, "choose_cases.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    for i, m in enumerate(meta):
        yield m["name"], z["%d/qid" % i], z["%d/cluster" % i], z["%d/score" % i], z["%d/kept_rows" % i]


def test_oracle_matches_reference():
    from oracle import oracle as orc
    n = 0
    for name, qid, cl, sc, kept in _cases():
        mask = orc.oracle_choose_alignment(qid, cl, sc)
        assert np.array_equal(np.nonzero(mask)[0], kept), name
        n += 1
    assert n >= 40


@pytest.mark.gpu
def test_gpu_matches_reference():
    import pandas as pd
    from fslr_b200 import cluster as gcluster
    for name, qid, cl, sc, kept in _cases():
        df = pd.DataFrame({"qname": ["q%05d" % q for q in qid], "cluster": cl, "alignment_score": sc})
        out = gcluster.choose_alignment(df)
        assert np.array_equal(out.index.to_numpy(), kept), name
        assert "avg_alignment_score" in df.columns


@pytest.mark.gpu
def test_gpu_matches_oracle_at_scale():
    """C2-sized table: clusters from the GPU clustering step, scores from the generator."""
    import pandas as pd
    from fslr_b200 import synth
    from fslr_b200.cluster import cluster_table
    from fslr_b200.engine import get_engine
    from fslr_b200.table import ColumnarTable
    from oracle import oracle as orc
    t = ColumnarTable.from_synth(synth.make_config("C2"))
    res = cluster_table(t, tie_order="stable")
    rng = np.random.default_rng(5)
    score = rng.integers(0, 50, size=t.n_rows).astype(np.int32)               # many ties between reads of a cluster
    is_rep, rep = get_engine(0).choose_alignment(t.read_id, score, res.cluster, int(res.cluster.max()) + 1)
    mask = orc.oracle_choose_alignment(t.read_id, res.cluster[t.read_id], score)
    assert np.array_equal(is_rep[t.read_id].astype(bool), mask)
    assert np.array_equal(np.sort(rep), np.sort(np.nonzero(is_rep)[0]))
