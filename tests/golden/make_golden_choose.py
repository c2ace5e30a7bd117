"""Golden vectors for cluster.choose_alignment (cluster.py:237-254) from the UNMODIFIED reference: random tables with
score ties inside clusters, singletons and negative scores.  Run from the repo root: python tests/golden/make_golden_choose.py"""
import json
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh                          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cluster = rh.import_reference_cluster()
    rng = np.random.default_rng(20261020)
    blob, meta = {}, []
    for i in range(40):
        n_reads = int(rng.integers(1, 120))
        n_cl = int(rng.integers(1, max(2, n_reads // 2)))
        rows = []
        for r in range(n_reads):
            k = int(rng.integers(1, 7))
            cl = int(rng.integers(0, n_cl))
            base = int(rng.integers(-5, 40)) if i % 3 else 10               # few distinct values -> ties between reads
            for _ in range(k):
                rows.append(("q%04d" % r, cl, base + int(rng.integers(0, 3)) * (i % 2)))
        df = pd.DataFrame(rows, columns=["qname", "cluster", "alignment_score"])
        df = df.sample(frac=1.0, random_state=int(rng.integers(0, 2**31))).reset_index(drop=True) if i % 4 == 0 else df
        df["cluster"] = df["cluster"].astype(float)                            # main.py:341-342 leaves floats
        out = cluster.choose_alignment(df.copy())
        blob["%d/qid" % i] = pd.factorize(df["qname"])[0].astype(np.int32)
        blob["%d/cluster" % i] = df["cluster"].to_numpy()
        blob["%d/score" % i] = df["alignment_score"].to_numpy().astype(np.int32)
        blob["%d/kept_rows" % i] = out.index.to_numpy().astype(np.int32)
        meta.append({"name": "choose%02d" % i, "n_rows": len(df)})
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "choose_cases.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, len(meta), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
