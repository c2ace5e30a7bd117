#!/usr/bin/env python
"""Benchmark of the read-clustering step (BASELINE.json metric: reads clustered/s + pair tests/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl reference]

A "step" is one pass of the whole clustering step (keep_fillings -> ... -> cluster / n_reads columns) over one
synthetic mappings table.  `value` is measured with the table already resident in HBM; `e2e` goes through the host
buffer C-ABI call (pinned host columns -> H2D -> step -> D2H of both result columns), one call at a time.  For N > 1 the
10M-read table is replicated and the pair space sharded (strong scaling); launch under torchrun.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, kind "port": the reference itself
is pure Python and cannot cluster 10M reads in the time a bench run has) on the SAME full table, one thread.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fslr_b200 import synth                                      # noqa: E402
from fslr_b200.table import ClusterParams, ColumnarTable         # noqa: E402

METRIC, UNIT = "reads_clustered_per_s", "reads/s"
CPU_SAMPLE_READS = 2_000_000

WORKLOADS = {
    "C1": "5k reads, primer 21q1, default cutoffs",
    "C2": "100k reads, 2-6 alignments/read, primers 21q1,17p6",
    "C3": "1M reads, --cluster-mask subtelomere,L1_TALEN",
    "C4": "10M reads, 2-6 alignments/read, primers 21q1,17p6 (40M table rows, 20M fillings)",
    "C5": "2M reads with a 500k-read breakpoint hotspot",
}


def env_int(k, d):
    return int(os.environ.get(k, d))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "MEASURED_PEAKS.json (burst copy bandwidth)"
    return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def host_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    try:
        usable = len(os.sched_getaffinity(0))
    except AttributeError:
        usable = os.cpu_count()
    return {"nproc": os.cpu_count(), "usable_cores": usable, "cpu_model": model,
            "procs": "--procs 1 (ignored: the clustering block main.py:190-352 is single-threaded by construction)"}


def make_workload(config, scale=1.0, genome_scale=1.0):
    t = synth.make_config(config, scale, genome_scale)
    ct = ColumnarTable.from_synth(t)
    params = ClusterParams.from_options(ct, cluster_mask=synth.CONFIG_MASK[config])
    if genome_scale != 1.0:
        params.subtel = max(8000, int(params.subtel * genome_scale))
    return ct, params


def config_dict(config, scale, R, A, F, D, parallelism):
    return {"workload": "%s: %s" % (config, WORKLOADS[config]) + ("" if scale == 1.0 else " x%g" % scale),
            "reads": R, "table_rows": A, "fillings": F, "intervals": D,
            "l2": "inputs larger than L2 (%.2f GB of columns re-read every step)" % (32 * A / 1e9),
            "parallelism": parallelism}


def cpu_port_run(config, n_reads, table=None):
    """One pass of the CPU restatement over `n_reads` reads of the config.  Fewer reads than the config has = a
    density-preserving sample: the genome (and the subtelomere window) is shortened by n_reads / full size, so that every
    filling meets as many others as in the full table — thinning the reads alone would make the CPU look 4x faster per read
    than it is on the real workload.  Returns (seconds, stats, reads)."""
    from oracle import oracle as orc
    kw = dict(synth.CONFIGS[config])
    frac = min(1.0, n_reads / kw["n_reads"])
    if table is None:
        table = make_workload(config, frac, frac)
    ct, params = table
    t0 = time.perf_counter()
    _, _, st = orc.oracle_cluster(ct, params)
    return time.perf_counter() - t0, st, ct.n_reads


def cpu_sample_text(config, n, extra=""):
    full = synth.CONFIGS[config]["n_reads"]
    if n >= full:
        return ("the full %s table (%d reads); whole clustering step%s, oracle/fslr_oracle.c, 1 thread (the reference step is "
                "single-threaded: main.py:190-352 never reads --procs)" % (config, full, extra))
    return ("%d reads of the %s distribution on a genome (and subtelomere window) shortened to %d/%d of its length, so the "
            "filling density and the pair tests per read equal the full table's; whole clustering step%s, "
            "oracle/fslr_oracle.c, 1 thread (the reference step is single-threaded: main.py:190-352 never reads --procs); "
            "on the full-size table the port is slower still per read (working set beyond the CPU caches)"
            % (n, config, n, full, extra))


def python_reference_runs(configs=("C1", "C2")):
    """The UNMODIFIED reference cluster.py (a copy under git-ignored baseline/_ref/, made by __graft_entry__.build() in the
    build container) driven by oracle/ref_harness.run_reference on the configs it finishes in seconds:
    what the Python reference itself does on this host, next to the C port used for the big tables."""
    ref_root = None
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(cand, "fslr", "cluster.py")):
            ref_root = cand
            break
    if ref_root is None:
        return {"unavailable": "no copy of fslr/cluster.py under baseline/_ref (run __graft_entry__.build() where /root/reference exists)"}
    os.environ["FSLR_REFERENCE_ROOT"] = ref_root
    from oracle import ref_harness as rh
    rh.REFERENCE_ROOT = ref_root
    out = {"source": "unmodified fslr/cluster.py from %s, stub pysam / superintervals (oracle/stubs), main.py:209-257,334-342 restated by "
                     "oracle/ref_harness.run_reference; 1 thread" % os.path.relpath(ref_root, ROOT), "runs": []}
    for c in configs:
        t = synth.make_config(c)
        df = t.to_dataframe()
        counters = {}
        t0 = time.perf_counter()
        rh.run_reference(df, t.chr_lengths, cluster_mask=synth.CONFIG_MASK[c], counters=counters)
        s = time.perf_counter() - t0
        out["runs"].append({"config": c, "reads": int(t.n_reads), "seconds": s, "reads_per_s": t.n_reads / s,
                            "pair_tests": counters.get("pair_tests"), "pair_tests_per_s": counters.get("pair_tests", 0) / s})
    return out


class ClockSampler:
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        """t0, t1: wall-clock window (time.time()) of the timed regions; samples outside it are ignored when any fall inside."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        if t0 is not None:
            inside = []
            for r in rows:
                try:
                    ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except Exception:
                    continue
                if t0 - 0.05 <= ts <= t1 + 0.05:
                    inside.append(r)
            if inside:
                rows = inside
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower() == "active":
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            hi = sorted(sm)[len(sm) // 2:]                       # samples under load: the upper half
            out = {"sm_mhz": float(np.median(hi)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def run_reference(args, rank):
    """`--impl reference`: the CPU restatement of the reference's clustering step on the host cores (1 thread — all the step
    can use), on this arm's own config.  Default: the FULL table every step.  The first pass is the warm-up and the clock: if
    K more passes would not fit FSLR_REF_BUDGET_S (default 1500 s) the run falls back to the density-preserving sample and
    says so."""
    if rank != 0:
        return
    full = synth.CONFIGS[args.config]["n_reads"]
    budget = float(os.environ.get("FSLR_REF_BUDGET_S", "1500"))
    n = full if args.cpu_sample_reads <= 0 else min(args.cpu_sample_reads, full)
    if args.gpus > 1 and args.cpu_sample_reads <= 0 and full > CPU_SAMPLE_READS:
        n = CPU_SAMPLE_READS       # the CPU arm does not depend on N: the full table is timed once, by the --gpus 1 run (15 minutes)
    t_gen = time.perf_counter()
    table = make_workload(args.config, min(1.0, n / full), min(1.0, n / full))
    t_gen = time.perf_counter() - t_gen
    s0, st0, _ = cpu_port_run(args.config, n, table)             # warm-up pass (also pages the table in)
    note = ""
    if n == full and (args.steps * s0 + t_gen) > budget and full > CPU_SAMPLE_READS:
        n = CPU_SAMPLE_READS
        note = " (the full table takes %.1f s per pass on this host: %d passes do not fit the %d s budget)" % (s0, args.steps, budget)
        table = make_workload(args.config, n / full, n / full)
        cpu_port_run(args.config, n, table)
    t_tot, reads, tests, st = 0.0, 0, 0, st0
    for _ in range(args.steps):
        s, st, nr = cpu_port_run(args.config, n, table)
        t_tot += s; reads += nr; tests += st["pair_tests"]
    v = reads / t_tot
    ct = table[0]
    cfg = config_dict(args.config, 1.0, full if n == full else ct.n_reads, ct.n_rows, st["n_fillings"], st["n_data"], "host CPU, 1 thread")
    if n != full:
        cfg["sample"] = "%d reads at the full table's filling density per step%s" % (n, note)
    cb = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": cpu_sample_text(args.config, n, " per timed step") + note}
    cb.update(host_info())
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "pair_tests_per_s": tests / t_tot, "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# algorithmic HBM bytes of the HBM-bound stages (F fillings, E recorded pairs, R1 query reads, R reads) and the
# library stages (fslrc_stage_name) whose CUDA-event times they are set against
def hbm_stage_table(st, R):
    F, E, R1 = st["n_fillings"], st["relation_entries"], st["n_query_reads"]
    return [
        ("sort", ["keep_fillings", "data_order_mask", "query_rank_read_lists", "chrom_sort"], 2 * 16 * F,
         "2*16*F: one read + one write of every 16-byte interval record (stages: keep_fillings + the sort by start + query rank + chromosome partition)"),
        ("band_bucket", ["records_bands"], 16 * F + 4 * F, "16*F + 4*F"),
        ("compaction", ["candidates"], 8 * E, "8*E written for E recorded pairs (here: the hit list written by k_hits)"),
        ("union_find", ["union_find"], 8 * E + 4 * R1 + 4 * R1, "8*E read + 4*R1 parent init + 4*R1 final labels"),
        ("numbering", ["numbering"], 8 * R, "8*R"),
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="C4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-sample-reads", type=int, default=None,
                    help="reads of the CPU pass (default: 2M density-preserving sample for the cpu_baseline leg, the FULL table for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true")
    ap.add_argument("--mode", default="shard", choices=["shard", "samples"],
                    help="N > 1: 'shard' = ONE table, pair space sharded over the GPUs (strong scaling, the BASELINE config); "
                         "'samples' = one table per GPU, no collective on the data path (weak scaling: how a run over many samples scales)")
    ap.add_argument("--e2e-depth", type=int, default=2, help="host-buffer calls in flight for the e2e_pipelined figure (1 = skip it)")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        if args.cpu_sample_reads is None:
            args.cpu_sample_reads = 0                          # the full table
        run_reference(args, rank)
        return
    if args.cpu_sample_reads is None:
        args.cpu_sample_reads = CPU_SAMPLE_READS

    import torch
    import torch.distributed as dist
    from fslr_b200.engine import DeviceTable, DeviceWireTable, Engine, HostPipeline, PinnedTable
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    ct, params = make_workload(args.config, args.scale)
    R, A = ct.n_reads, ct.n_rows
    dtab = DeviceTable(ct, eng.device)
    t_prep = time.perf_counter()
    ptab = PinnedTable(ct, compact=True)                      # the table in its wire format: 13.25 B/row (see PinnedTable)
    t_prep = time.perf_counter() - t_prep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    samples = world > 1 and args.mode == "samples"

    def step_resident():
        if world > 1 and not samples:
            return eng.run_sharded(dtab, ct, params, rank, world)
        return eng.run_resident(dtab, ct, params)

    dwire = DeviceWireTable(ptab, eng.device, world) if (world > 1 and not samples) else None

    def step_e2e():
        if samples:
            return eng.run_host(ptab, ct, params)
        if world > 1:
            # every rank uploads 1/world of every wire column over its own PCIe link; in-place NVLink all-gathers complete them
            dwire.upload(ptab, rank, world)
            st = eng.run_sharded(dwire, ct, params, rank, world)
            if rank == 0:                                     # the job's result leaves the box once
                ptab.out_cluster[:R].copy_(dwire.out_cluster[:R], non_blocking=True)
                ptab.out_n_reads[:R].copy_(dwire.out_n_reads[:R], non_blocking=True)
            torch.cuda.synchronize()
            return st
        return eng.run_host(ptab, ct, params)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        e0.record()
        sts = [fn() for _ in range(steps)]
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=eng.device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), sts, eng.launch_count() - l0

    sampler = ClockSampler(local) if rank == 0 else None      # runs across warm-up and every timed region; samples are filtered
    for _ in range(args.warmup):                              # to the timed window below
        step_resident()

    # ---- result check before anything is timed: the sharded run must equal a single-GPU run of the same table, on every rank
    sharded_check = None
    if world > 1 and not samples:
        step_resident()
        sh_c, sh_n = dtab.out_cluster[:R].clone(), dtab.out_n_reads[:R].clone()
        eng.run_resident(dtab, ct, params)                    # this rank alone, no collective
        same = bool(torch.equal(sh_c, dtab.out_cluster[:R]) and torch.equal(sh_n, dtab.out_n_reads[:R]))
        w = torch.arange(1, R + 1, device=eng.device, dtype=torch.int64)
        chk = torch.stack([(sh_c.to(torch.int64) * w).sum(), (sh_n.to(torch.int64) * w).sum(), torch.tensor(int(same), device=eng.device)])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        ok = all(int(c[2]) == 1 and torch.equal(c[:2], allc[0][:2]) for c in allc)
        assert ok, "sharded result differs from the single-GPU result of the same table (rank %d: %s)" % (rank, [c.tolist() for c in allc])
        sharded_check = {"equal_to_single_gpu_run_on_every_rank": True, "checksum": [int(x) for x in allc[0][:2].tolist()]}
        for _ in range(2):                                    # the single-GPU run above resized the scratch pools: settle them again
            step_resident()

    t_first = time.time()
    ms, sts, launches = timed(step_resident, args.steps)
    step_e2e()
    ms_e2e, sts_e2e, _ = timed(step_e2e, args.steps)
    e2e_mode = ("one blocking C-ABI call per step, one table at a time: pinned host columns in their %.2f B/row wire format (chrom uint8, "
                "n_alignments / qstart / qend uint16, rend as int16 span, aln_size = qend - qstart and read_id from run lengths rebuilt on the "
                "device) in, both result columns out" % (ptab.h2d_bytes / max(A, 1))) \
        if (world == 1 or samples) else \
        "every rank uploads 1/%d of the rows (%.2f B/row on the wire), in-place NVLink all-gather of every wire column, sharded step, rank 0 downloads the result" % (world, ptab.h2d_bytes / max(A, 1))
    e2e_serial = {"value": (world if samples else 1) * R * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                  "h2d_bytes_per_step": ptab.h2d_bytes, "d2h_bytes_per_step": ptab.d2h_bytes, "mode": e2e_mode,
                  "host_prep_s_untimed": t_prep}
    e2e_pipelined = None
    if world == 1 and args.e2e_depth > 1:
        # the same call, `e2e_depth` in flight on separate contexts: table k+1 uploads while table k computes (throughput of a
        # service clustering sample after sample, not one table's latency).  Worker threads sleep in their waits.
        pipe = HostPipeline(local, args.e2e_depth, blocking_sync=True)
        ptabs = [PinnedTable(ct, compact=True) for _ in range(args.e2e_depth)]
        wf = []
        for i in range(args.e2e_depth * max(args.warmup, 2)):        # warm-up: same submission pattern as the timed loop
            if i >= len(ptabs):
                wf[i - len(ptabs)].result()
            wf.append(pipe.submit(ptabs[i % len(ptabs)], ct, params))
        for f in wf:
            f.result()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for s_ in pipe.streams:
            s_.wait_event(e0)
        futs = []
        for i in range(args.steps):                       # a buffer set is reused only after its previous step has finished
            if i >= len(ptabs):
                futs[i - len(ptabs)].result()
            futs.append(pipe.submit(ptabs[i % len(ptabs)], ct, params))
        for f in futs:
            f.result()
        for s_ in pipe.streams:
            cur.wait_stream(s_)
        e1.record(cur)
        torch.cuda.synchronize()
        ms_pipe = e0.elapsed_time(e1)
        for pt in ptabs:
            assert np.array_equal(pt.out_cluster[:R].numpy(), ptab.out_cluster[:R].numpy()), "pipelined e2e results disagree"
        pipe.close()
        e2e_pipelined = {"value": R * args.steps / (ms_pipe * 1e-3), "unit": UNIT, "ms_per_step": ms_pipe / args.steps, "depth": args.e2e_depth,
                         "h2d_bytes_per_step": ptabs[0].h2d_bytes, "d2h_bytes_per_step": ptabs[0].d2h_bytes,
                         "mode": "%d C-ABI calls in flight on separate contexts (blocking-sync waits), throughput over a stream of tables" % args.e2e_depth}
        print("[bench] e2e: serial %.2f ms/step, %d in flight %.2f ms/step" % (ms_e2e / args.steps, args.e2e_depth, ms_pipe / args.steps), file=sys.stderr)

    clocks = sampler.stop(t_first, time.time()) if sampler else {}

    # correctness guard inside the bench: both paths agree with each other
    res_a = dtab.out_cluster[:R].cpu().numpy()
    if world == 1:
        assert np.array_equal(res_a, ptab.out_cluster[:R].numpy()), "resident and host paths disagree"

    st = sts[-1]
    stage_ms = {k: float(np.mean([s["stage_ms"][k] for s in sts])) for k in st["stage_ms"]}
    jobs = world if samples else 1                            # tables clustered per step over all ranks
    value = jobs * R * args.steps / (ms * 1e-3)
    peak, peak_kind = load_peaks()
    F, D, Q = st["n_fillings"], st["n_intervals"], st["n_query_reads"]
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram__bytes_read+write per launch from the committed ncu captures
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.config, {})
    # ---- HBM-bound stages against the measured copy bandwidth (SURVEY §8d byte formulas)
    stage_rooflines = []
    for name, stages, nbytes, formula in hbm_stage_table(st, R):
        t_ms = sum(stage_ms[s] for s in stages)
        ach = nbytes / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        tr = traffic.get(name)
        if name == "sort" and tr is not None:
            tr += traffic.get("keep_fillings", 0)
        stage_rooflines.append({"stage": name, "library_stages": stages, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                                "frac": ach / peak, "algorithmic_bytes": nbytes, "formula": formula, "ms": t_ms,
                                "traffic": tr,       # DRAM bytes the stage really moved (ncu, per step): sort passes, permutation gathers
                                "traffic_frac": (tr / (t_ms * 1e-3) / 1e9 / peak) if (tr and t_ms > 0) else None})
    # ---- the pair-test kernel against the integer-issue roofline: 12 + 7*L1*L2 lane-ops per evaluated read pair
    # against the measured dependent-free IADD/LOP/VIMNMX rate of this device
    roofline = None
    if rank == 0:
        ipeak = eng.int_peak()
        Lbar = D / max(Q, 1)
        ops_per = 12 + 7 * Lbar * Lbar
        ops = st["pair_tests"] * ops_per
        t_eval = stage_ms["pair_kernel"] if stage_ms["pair_kernel"] > 0 else stage_ms["pair_heavy"]
        t_stage = stage_ms["candidates"] + stage_ms["pair_kernel"] + stage_ms["pair_heavy"]
        ach = ops / (t_eval * 1e-3) / 1e9 if t_eval > 0 else 0.0
        roofline = {"kernel": "k_eval (pair-test kernel: one lane per candidate pair)" if stage_ms["pair_kernel"] > 0 else "k_pair",
                    "bound": "int_issue", "achieved": ach, "peak": ipeak / 1e9, "unit": "Gop/s (integer lane-ops)",
                    "frac": ach / (ipeak / 1e9) if ipeak > 0 else None,
                    "traffic": traffic.get("pair_kernel"), "ops_per_pair_test": ops_per, "algorithmic_ops_per_launch": ops,
                    "ms_per_launch": t_eval,
                    "peak_source": "measured in this run (fslrc_int_peak: 8 independent add/minmax/xor chains per thread)",
                    "whole_pair_stage": {"kernels": ["k_hits (candidate generation)", "k_eval (pair tests)", "k_pair (heavy reads)"], "ms": t_stage,
                                         "frac": ops / (t_stage * 1e-3) / ipeak if t_stage > 0 and ipeak > 0 else None},
                    "hbm_view": {"bound": "hbm", "achieved": (48 * st["pair_tests"] + 16 * st["pair_tests"]) / (t_eval * 1e-3) / 1e9 if t_eval > 0 else 0.0,
                                 "peak": peak, "unit": "GB/s", "note": "8 B hit + 32 B position records + 8 B result per test at least; gathers of the filling lists hit L2"}}
    par = "1 GPU" if world == 1 else ("one table per GPU on %d GPUs, no data-path collective" % world if samples else
                                      "table replicated (e2e: upload sharded, NVLink all-gather), candidate generation and pair tests sharded by "
                                      "query read over %d GPUs; all-reduce of per-read counters, all-gather of the saturating reads' pairs, "
                                      "all-gather of forests" % world)
    t_pair = stage_ms["candidates"] + stage_ms["pair_kernel"] + stage_ms["pair_heavy"] + stage_ms["replay"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if samples else "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": config_dict(args.config, args.scale, R, A, F, D, par),
            "pair_tests_per_s": st["pair_tests"] / t_pair * 1e3 if t_pair > 0 else None,
            "pair_tests": st["pair_tests"], "band_pairs": st["band_pairs"], "edges": st["edges"], "clusters": st["components"],
            "saturating_reads": st["saturating_reads"], "partner_records": st["partner_records"],
            "stage_ms": stage_ms, "roofline": roofline, "stage_rooflines": stage_rooflines, "clocks": clocks,
            "e2e": {k: e2e_serial[k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "mode")},
            "e2e_serial": e2e_serial, "e2e_pipelined": e2e_pipelined, "sharded_result_check": sharded_check,
            "gpu_launches": launches}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            n = min(args.cpu_sample_reads, R)
            s, ost, nr = cpu_port_run(args.config, n)
            cb = {"value": nr / s, "unit": UNIT, "cores": 1, "kind": "port", "pair_tests_per_s": ost["pair_tests"] / s,
                  "sample": cpu_sample_text(args.config, nr, " once (%.1f s)" % s)}
            cb.update(host_info())
            line["cpu_baseline"] = cb
            if not args.no_python_reference:
                line["cpu_reference_python"] = python_reference_runs()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
