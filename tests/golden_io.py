"""Loader for the fixtures written by tests/golden/make_golden.py."""
import json
import os

import numpy as np

from fslr_b200 import synth
from fslr_b200.table import ClusterParams, ColumnarTable

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_COLS = ("read_id", "chrom", "rstart", "rend", "aln_size", "qstart", "qend", "n_alignments")


class Case:
    def __init__(self, meta, arrays):
        self.meta, self.arrays, self.name = meta, arrays, meta["name"]
        self._table = None

    @property
    def table(self):
        if self._table is None:
            m, a = self.meta, self.arrays
            if "synth" in m:
                t = ColumnarTable.from_synth(synth.make_config(m["synth"]["config"], m["synth"]["scale"]))
                chk = int(sum(int(np.asarray(getattr(t, k), dtype=np.int64).sum()) * (i + 1)
                              for i, k in enumerate(c for c in _COLS if c != "chrom")))
                assert chk == m["input_checksum"], "synthetic generator drifted from the golden fixture"
                self._table = t
            else:
                self._table = ColumnarTable(**{k: a[k] for k in _COLS}, n_reads=m["n_reads"],
                                            chrom_names=m["chrom_names"],
                                            chrom_len=np.asarray(m["chrom_len"], dtype=np.int64))
        return self._table

    @property
    def filter_false(self):
        return bool(self.meta["opts"].get("filter_false"))

    @property
    def params(self):
        o = {k: v for k, v in self.meta["opts"].items() if k != "filter_false"}
        return ClusterParams.from_options(self.table, **o)

    @property
    def order(self):
        return self.arrays.get("order")

    @property
    def no_clusters(self):
        return self.meta["no_clusters"]

    @property
    def expected(self):
        return self.arrays.get("cluster"), self.arrays.get("n_reads")


def load(fname):
    z = np.load(os.path.join(HERE, fname))
    meta = json.loads(bytes(z["meta"]).decode())
    out = []
    for i, m in enumerate(meta):
        arrs = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("%d/" % i)}
        out.append(Case(m, arrs))
    return out
