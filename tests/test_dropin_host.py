"""CPU tests of the host half of the drop-in boundary (fslr_b200/cluster.py): the tie-order permutation against every
permutation the reference produced for the golden fixtures, and the pandas helpers against the reference's own functions
(imported from /root/reference in the build container; those tests skip on the GPU box, where it does not exist)."""
import gzip
import json
import os

import numpy as np
import pandas as pd
import pytest

from tests import golden_io

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dropin_cases():
    with gzip.open(os.path.join(HERE, "dropin_cases.json.gz")) as f:
        return json.loads(f.read())


@pytest.mark.parametrize("fname", ["small_cases.npz", "wide_cases.npz", "config_cases.npz"])
def test_reference_tie_order_reproduces_every_golden_permutation(fname):
    """cluster.reference_tie_order (what the default TIE_ORDER="reference" feeds the GPU) must equal the permutation the
    reference's own sort_values('start') applied when the fixture was generated (cluster.py:114)."""
    from fslr_b200.cluster import reference_tie_order
    n = 0
    for c in golden_io.load(fname):
        if c.order is None:
            continue
        assert np.array_equal(reference_tie_order(c.table), c.order), c.name
        n += 1
    assert n > 0


def test_reference_tie_order_after_filter_false():
    """Same pin on the drop-in fixtures, including tables that went through delete_false first (main.py:229-230)."""
    from fslr_b200.cluster import delete_false, reference_tie_order
    from fslr_b200.table import ColumnarTable
    n_filtered = 0
    for case in _dropin_cases():
        df = pd.DataFrame(case["rows"], columns=case["columns"])
        if case["opts"].get("filter_false"):
            df = delete_false(df)
            n_filtered += 1
        t = ColumnarTable.from_dataframe(df, case["chr_lengths"])
        assert reference_tie_order(t).tolist() == case["expected"]["order"], case["name"]
    assert n_filtered >= 5


def _reference_cluster():
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference checkout not present (GPU box)")
    return rh.import_reference_cluster()


def _messy_frame(seed):
    rng = np.random.default_rng(seed)
    chroms = ["chr1", "chr10", "chr2", "chrX", "chrY", "L1_TALEN", "chr21", "chrUn_KI270742v1", "chr9_alt", "scaffold7"]
    n = 60
    return pd.DataFrame({"chrom": rng.choice(chroms, size=n), "rstart": rng.integers(1, 10**6, size=n),
                         "rend": rng.integers(1, 10**6, size=n),
                         "qname": ["r%d.0.9_0.9.21q1F_%s" % (i // 3, "False" if (i // 3) % 4 == 0 else "21q1R") for i in range(n)]})


@pytest.mark.parametrize("seed", range(6))
def test_rename_chromosomes_equals_the_reference_function(seed):
    """cluster.py:34-43.  Ids of `chrN` names are fixed by the numeric sort; ids of other names follow set iteration order
    in the reference (run dependent), so those are compared as a set of ids and through the round trip back to names."""
    ref = _reference_cluster()
    from fslr_b200 import cluster as mine
    df = _messy_frame(seed)
    lengths = {"chr1": 248_000_000, "chr2": 242_000_000, "chrX": 154_000_000, "notinbed": 5, "L1_TALEN": 8000}
    mask = ["subtelomere", "L1_TALEN", "chr2", "nosuchchrom"]
    rb, rl, rm, rmap = ref.rename_chromosomes(df.copy(), dict(lengths), list(mask))
    mb, ml, mm, mmap = mine.rename_chromosomes(df.copy(), dict(lengths), list(mask))
    assert set(rmap) == set(mmap)
    numeric = [k for k in rmap if k[:3] == "chr" and k[3:].isdigit()]
    assert {k: rmap[k] for k in numeric} == {k: mmap[k] for k in numeric}
    assert sorted(rmap.values()) == sorted(mmap.values()) == list(range(1, len(rmap) + 1))
    inv_r, inv_m = {v: k for k, v in rmap.items()}, {v: k for k, v in mmap.items()}
    assert [inv_r[x] for x in rb["chrom"]] == [inv_m[x] for x in mb["chrom"]] == df["chrom"].tolist()
    assert {inv_r.get(k): v for k, v in rl.items()} == {inv_m.get(k): v for k, v in ml.items()}
    assert [inv_r.get(x, x) for x in rm] == [inv_m.get(x, x) for x in mm]
    # and back (cluster.py:46-49)
    assert ref.chrom_to_str(rb, rmap)["chrom"].tolist() == mine.chrom_to_str(mb, mmap)["chrom"].tolist()


@pytest.mark.parametrize("seed", range(3))
def test_delete_false_equals_the_reference_function(seed):
    ref = _reference_cluster()
    from fslr_b200 import cluster as mine
    df = _messy_frame(seed)
    a, b = ref.delete_false(df.copy()), mine.delete_false(df.copy())
    assert a.index.tolist() == b.index.tolist() and a.equals(b) and len(a) < len(df)
