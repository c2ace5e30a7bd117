"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/fslr_oracle.c (the CPU checker).

May be imported from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only; the
product package (fslr_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfslr_oracle.so")


class _In(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_reads", C.c_int64),
                ("read_id", C.c_void_p), ("chrom", C.c_void_p),
                ("rstart", C.c_void_p), ("rend", C.c_void_p), ("aln_size", C.c_void_p),
                ("qstart", C.c_void_p), ("qend", C.c_void_p), ("n_alignments", C.c_void_p),
                ("n_chrom", C.c_int32), ("chrom_len", C.c_void_p), ("chrom_masked", C.c_void_p),
                ("mask_subtelomere", C.c_int32), ("subtel", C.c_int64), ("order", C.c_void_p),
                ("overlap", C.c_double), ("qlen_diff", C.c_double), ("diff", C.c_double),
                ("cutoffs", C.c_void_p), ("n_cutoffs", C.c_int32), ("edge_threshold", C.c_int64)]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_fillings", "n_data", "n_query_reads", "candidates", "pair_tests",
                                          "edges", "components", "clustered_reads")] + [("no_clusters", C.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "fslr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"] if force else ["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.fslr_oracle_cluster.restype = C.c_int
        _lib.fslr_oracle_cluster.argtypes = [C.POINTER(_In), C.c_void_p, C.c_void_p, C.POINTER(_Stats)]
    return _lib


def oracle_cluster(table, params, order=None):
    """table: fslr_b200.table.ColumnarTable, params: ClusterParams.
    Returns (cluster[int32 R], n_reads[int32 R], stats dict); stats['no_clusters'] mirrors main.py:247-249."""
    lib = _load()
    keep = []

    def p(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    cut = np.asarray(params.jaccard_cutoffs, dtype=np.float64)
    masked = params.chrom_masked if params.chrom_masked is not None else np.zeros(table.n_chrom, np.uint8)
    a = _In(table.n_rows, table.n_reads, p(table.read_id, np.int32), p(table.chrom, np.int32),
            p(table.rstart, np.int32), p(table.rend, np.int32), p(table.aln_size, np.int32),
            p(table.qstart, np.int32), p(table.qend, np.int32), p(table.n_alignments, np.int32),
            table.n_chrom, p(table.chrom_len, np.int64), p(masked, np.uint8),
            int(params.mask_subtelomere), int(params.subtel),
            p(order, np.int64) if order is not None else None,
            params.overlap, params.qlen_diff, params.n_alignment_diff,
            p(cut, np.float64), len(cut), int(params.edge_threshold))
    out_c = np.empty(table.n_reads, dtype=np.int32)
    out_n = np.empty(table.n_reads, dtype=np.int32)
    st = _Stats()
    rc = lib.fslr_oracle_cluster(C.byref(a), out_c.ctypes.data, out_n.ctypes.data, C.byref(st))
    if rc != 0:
        raise ZeroDivisionError("oracle: invalid table (code %d): zero aln_size / qlen2 / n_alignments" % rc)
    stats = {n: int(getattr(st, n)) for n, _ in _Stats._fields_}
    return out_c, out_n, stats


def fillings_start_column(table):
    """`start` column of the frame keep_fillings returns (cluster.py:14-31,111), in bed order —
    what the reference's unstable sort (cluster.py:114) is applied to."""
    rid = np.asarray(table.read_id)
    A = rid.shape[0]
    idx = np.arange(A)
    first = np.full(table.n_reads, A, dtype=np.int64)
    last = np.full(table.n_reads, -1, dtype=np.int64)
    np.minimum.at(first, rid, idx)
    np.maximum.at(last, rid, idx)
    keep = (idx != first[rid]) & (idx != last[rid])
    return np.minimum(table.rstart[keep], table.rend[keep]).astype(np.int64)


def oracle_choose_alignment(qid, cluster_rows, score):
    """numpy restatement of cluster.choose_alignment (cluster.py:237-254): per qname the float64 mean of alignment_score
    (:238-239), per cluster the FIRST row holding the maximum mean (:248, DataFrame.idxmax), kept rows = all rows of the
    selected qnames (:251).  qid: dense read id per row; cluster_rows: cluster value per row.  Returns the kept-row mask."""
    qid = np.asarray(qid)
    n = int(qid.max()) + 1 if qid.size else 0
    avg = np.bincount(qid, weights=np.asarray(score, dtype=np.float64), minlength=n) / np.maximum(np.bincount(qid, minlength=n), 1)
    row_avg = avg[qid]
    selected = set()
    cl = np.asarray(cluster_rows)
    for c in np.unique(cl):
        rows = np.nonzero(cl == c)[0]
        selected.add(int(qid[rows[np.argmax(row_avg[rows])]]))       # argmax: first maximum, like idxmax
    return np.isin(qid, list(selected))
