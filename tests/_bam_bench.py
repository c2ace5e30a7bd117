"""Dev script (GPU box): times the BAM -> mappings.bed producer against the oracle restatement of
collect_mapping_info.mapping_info on the same records.  python tests/_bam_bench.py [copies]  -> one JSON line."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fslr_b200 import mapping_info as mi, synth_bam as sb      # noqa: E402
from oracle import mapping_info_oracle as mo                   # noqa: E402

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 10
refs, recs, primers = sb.make_alignments(20000, seed=5, p_unmapped=0.0)
enc = [sb.encode_record(*r) for r in recs]
d = tempfile.mkdtemp()
small = os.path.join(d, "small.bam")
sb.write_bam(small, refs, enc)
t0 = time.perf_counter()
rows = mo.mapping_rows(small, None, primers, "9.9")
txt = mo.mapping_tsv(rows)
t_oracle = time.perf_counter() - t0
# `copies` copies of the read set under distinct names (the first two name characters carry the copy number)
big = []
for c in range(copies):
    tag = b"%02x" % c
    big.extend(e[:36] + tag + e[38:] for e in enc)
path = os.path.join(d, "big.bam")
sb.write_bam(path, refs, big)
n_reads, n_rec = 20000 * copies, len(big)
out = {"reads": n_reads, "records": n_rec, "bam_MB": os.path.getsize(path) / 1e6, "raw_MB": sum(map(len, big)) / 1e6,
       "oracle_reads_per_s": 20000 / t_oracle, "oracle_s_20k_reads": t_oracle}
mi.read_bam_table(small, None, primers).close()                     # warm-up (context, allocator)
for rep in range(2):
    t0 = time.perf_counter(); raw = mi.inflate_bgzf(path, threads=16); t1 = time.perf_counter()
    t = mi.read_bam_table(raw, None, primers); t2 = time.perf_counter()
    bed = t.mappings_bed_bytes("9.9"); t3 = time.perf_counter()
    kernels_ms = t.parse_ms
    t.close()
out.update({"host_inflate": {"inflate_s": t1 - t0, "open_s": t2 - t1, "open_device_ms": kernels_ms, "render_s": t3 - t2,
                             "reads_per_s_from_bgzf": n_reads / (t3 - t0)}, "bed_MB": bed.shape[0] / 1e6})
want = bed.tobytes()
for rep in range(2):
    t0 = time.perf_counter(); t1 = t0
    t = mi.read_bam_table(path, None, primers, device_inflate=True); t2 = time.perf_counter()
    bed = t.mappings_bed_bytes("9.9"); t3 = time.perf_counter()
    kernels_ms, fb = t.parse_ms, int(t.info.reserved)
    t.close()
out.update({"device_inflate": {"read_file_and_open_s": t2 - t0, "open_device_ms": kernels_ms, "render_s": t3 - t2,
                               "host_walk_fallback": fb, "reads_per_s_from_bgzf": n_reads / (t3 - t0), "same_bytes": bool(bed.tobytes() == want)}})
if copies == 1:
    out["identical_to_oracle"] = bool(bed.tobytes() == txt.encode())
print(json.dumps(out))
