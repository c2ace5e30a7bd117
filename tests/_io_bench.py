"""Development measurement (not a test): file -> clusters -> file on the GPU vs pandas, C3-sized table (1M reads)."""
import io
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, ".")
from fslr_b200 import synth, tsv                                        # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
t = synth.make_config("C3", scale)
buf = io.StringIO()
t0 = time.perf_counter(); t.to_dataframe().to_csv(buf, index=False, sep="\t"); raw = buf.getvalue().encode()
print("table: %d rows, %.1f MB of TSV (pandas to_csv %.1f s)" % (t.n_rows, len(raw) / 1e6, time.perf_counter() - t0))
t0 = time.perf_counter(); df = pd.read_csv(io.BytesIO(raw), sep="\t"); t_pd = time.perf_counter() - t0
print("pandas read_csv: %.2f s (%.0f MB/s)" % (t_pd, len(raw) / 1e6 / t_pd))
for rep in range(3):
    t0 = time.perf_counter()
    pb = tsv.read_mappings_bed(raw, t.chr_lengths)
    t1 = time.perf_counter()
    res = pb.cluster(cluster_mask=synth.CONFIG_MASK["C3"])
    t2 = time.perf_counter()
    out = pb.cluster_bed_bytes()
    t3 = time.perf_counter()
    print("gpu: parse %.1f ms wall (%.2f ms on the device after the upload, %.1f GB/s), cluster %.1f ms, render+download %.1f ms, %.1f MB out"
          % (1e3 * (t1 - t0), pb.parse_ms, len(raw) / 1e6 / max(pb.parse_ms, 1e-9), 1e3 * (t2 - t1), 1e3 * (t3 - t2), out.nbytes / 1e6))
    pb.close()
