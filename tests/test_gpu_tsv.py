"""GPU ingest / egress of mappings.bed against pandas: the parsed columns and dense ids equal
ColumnarTable.from_dataframe(pd.read_csv(...)), the clustering result on the device-resident table equals the host path,
and the rendered mappings.cluster.bed is byte-identical to what main.py:334-349 writes with DataFrame.to_csv."""
import io

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _bed_bytes(name, scale):
    from fslr_b200 import synth
    t = synth.make_config(name, scale)
    df = t.to_dataframe()
    df.loc[df.index[::7], "seq"] = "ACGTTGCA" * 11                      # some long trailing fields, some empty ones
    buf = io.StringIO()
    df.to_csv(buf, index=False, sep="\t")
    return buf.getvalue().encode(), t.chr_lengths


@pytest.mark.parametrize("name,scale", [("C1", 1.0), ("C3", 0.02)])
def test_parse_cluster_write(name, scale):
    from fslr_b200 import synth, tsv
    from fslr_b200.engine import get_engine
    from fslr_b200.table import ClusterParams, ColumnarTable
    raw, lens = _bed_bytes(name, scale)
    df = pd.read_csv(io.BytesIO(raw), sep="\t")                          # main.py:209
    ref = ColumnarTable.from_dataframe(df, lens)
    pb = tsv.read_mappings_bed(raw, lens)
    assert (pb.n_rows, pb.n_reads) == (ref.n_rows, ref.n_reads)
    assert pb.chrom_names == [str(c) for c in ref.chrom_names]
    assert np.array_equal(pb.chrom_len, ref.chrom_len)
    for k in ("read_id", "chrom", "rstart", "rend", "aln_size", "qstart", "qend", "n_alignments"):
        assert np.array_equal(pb.column(k), getattr(ref, k)), k
    assert np.array_equal(pb.column("alignment_score"), df["alignment_score"].to_numpy())
    assert list(pb.qnames()) == list(ref.qnames)
    # clustering on the device-resident parse == host-buffer path on the pandas-built table
    mask = synth.CONFIG_MASK[name]
    res = pb.cluster(cluster_mask=mask)
    host = get_engine(0).cluster(ref, ClusterParams.from_options(ref, cluster_mask=mask))
    assert np.array_equal(res.cluster, host.cluster) and np.array_equal(res.n_reads, host.n_reads)
    # egress: byte-identical to the reference's DataFrame.to_csv of the merged frame (floats from the NaN merge)
    df["cluster"] = host.cluster[ref.read_id].astype(np.float64)
    df["n_reads"] = host.n_reads[ref.read_id].astype(np.float64)
    want = io.StringIO()
    df.to_csv(want, index=False, sep="\t")
    assert pb.cluster_bed_bytes().tobytes() == want.getvalue().encode()
    pb.close()


def test_malformed_inputs():
    from fslr_b200 import tsv
    from fslr_b200._native import FslrError
    raw, lens = _bed_bytes("C1", 0.02)
    lines = raw.split(b"\n")
    with pytest.raises(FslrError):                                       # a header without the qname column
        tsv.read_mappings_bed(raw.replace(b"qname", b"name", 1), lens)
    with pytest.raises(FslrError):                                       # a truncated line
        tsv.read_mappings_bed(b"\n".join(lines[:3] + [b"chr1\t5"] + lines[3:]), lens)
    with pytest.raises(FslrError):                                       # a non-numeric coordinate
        tsv.read_mappings_bed(raw.replace(lines[2].split(b"\t")[1], b"12x4", 1), lens)
    no_trailing_newline = tsv.read_mappings_bed(raw.rstrip(b"\n"), lens)
    assert no_trailing_newline.n_rows == len(lines) - 2
    no_trailing_newline.close()


def test_a_second_parse_invalidates_the_first_handle():
    """The library context holds ONE parsed table: an older ParsedBed must refuse to be used (its device columns are freed)
    and closing it must not free the newer table."""
    from fslr_b200 import tsv
    raw1, lens = _bed_bytes("C1", 0.2)
    raw2, _ = _bed_bytes("C1", 0.3)
    pb1 = tsv.read_mappings_bed(raw1, lens)
    pb2 = tsv.read_mappings_bed(raw2, lens)
    with pytest.raises(RuntimeError, match="stale"):
        pb1.column("rstart")
    with pytest.raises(RuntimeError, match="stale"):
        pb1.cluster()
    pb1.close()                                                          # must not touch pb2's table
    res = pb2.cluster()
    assert res.cluster.shape[0] == pb2.n_reads
    assert pb2.cluster_bed_bytes().tobytes().startswith(raw2[:raw2.index(b"\n")])
    pb2.close()
    with pytest.raises(RuntimeError, match="closed"):
        pb2.column("rstart")
