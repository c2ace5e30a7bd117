#!/usr/bin/env python
"""Benchmark of the read-clustering step (BASELINE.json metric: reads clustered/s + pair tests/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl reference]

A "step" is one pass of the whole clustering step (keep_fillings -> ... -> cluster / n_reads columns) over one
synthetic mappings table.  `value` is measured with the table already resident in HBM; `e2e` goes through the host
buffer C-ABI call (pinned host columns -> H2D -> step -> D2H of both result columns) every step.  For N > 1 the
10M-read table is replicated and the pair space sharded (strong scaling); launch under torchrun.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, kind "port": the reference itself
is pure Python and /root/reference does not exist on the GPU box) on a bounded sample of the same workload.
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


def env_int(k, d):
    return int(os.environ.get(k, d))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def make_workload(config, scale=1.0, genome_scale=1.0):
    t = synth.make_config(config, scale, genome_scale)
    ct = ColumnarTable.from_synth(t)
    params = ClusterParams.from_options(ct, cluster_mask=synth.CONFIG_MASK[config])
    if genome_scale != 1.0:
        params.subtel = max(8000, int(params.subtel * genome_scale))
    return ct, params


def cpu_port_run(config, n_reads, table=None):
    """One pass of the CPU restatement over a density-preserving sample of the config: `n_reads` reads on a genome (and a
    subtelomere window) shortened by n_reads / full size, so that every filling meets as many others as in the full table
    — thinning the reads alone would make the CPU look 4x faster per read than it is on the real workload.
    Returns (seconds, stats, reads)."""
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
    return ("%d reads of the %s distribution on a genome (and subtelomere window) shortened to %d/%d of its length, so the "
            "filling density and the pair tests per read equal the full table's; whole clustering step%s, "
            "oracle/fslr_oracle.c, 1 thread (the reference step is single-threaded: main.py:190-352 never reads --procs); "
            "on the full-size table the port is slower still per read (working set beyond the CPU caches)"
            % (n, config, n, full, extra))


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
    if rank != 0:
        return
    n = min(args.cpu_sample_reads, synth.CONFIGS[args.config]["n_reads"])
    frac = min(1.0, n / synth.CONFIGS[args.config]["n_reads"])
    table = make_workload(args.config, frac, frac)
    for _ in range(args.warmup and 1):
        cpu_port_run(args.config, n, table)
    t_tot, reads, tests = 0.0, 0, 0
    for _ in range(args.steps):
        s, st, nr = cpu_port_run(args.config, n, table)
        t_tot += s; reads += nr; tests += st["pair_tests"]
    v = reads / t_tot
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.config, WORKLOADS[args.config]), "sample": "%d reads at the full table's filling density per step" % n},
            "pair_tests_per_s": tests / t_tot,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": cpu_sample_text(args.config, n, " per timed step")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


WORKLOADS = {
    "C1": "5k reads, primer 21q1, default cutoffs",
    "C2": "100k reads, 2-6 alignments/read, primers 21q1,17p6",
    "C3": "1M reads, --cluster-mask subtelomere,L1_TALEN",
    "C4": "10M reads, 2-6 alignments/read, primers 21q1,17p6 (40M table rows, 20M fillings)",
    "C5": "2M reads with a 500k-read breakpoint hotspot",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="C4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-sample-reads", type=int, default=CPU_SAMPLE_READS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="shard", choices=["shard", "samples"],
                    help="N > 1: 'shard' = ONE table, pair space sharded over the GPUs (strong scaling, the BASELINE config); "
                         "'samples' = one table per GPU, no collective on the data path (weak scaling: how a run over many samples scales)")
    ap.add_argument("--e2e-depth", type=int, default=2, help="host-buffer calls in flight for the e2e number (1 = strictly serial)")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from fslr_b200.engine import DeviceTable, Engine, HostPipeline, PinnedTable
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
    ptab = PinnedTable(ct, compact=world > 1 and args.mode == "shard")       # (N = 1: the strictly serial e2e uses int32 columns; the pipelined one compact ones)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    samples = world > 1 and args.mode == "samples"

    def step_resident():
        if world > 1 and not samples:
            return eng.run_sharded(dtab, ct, params, rank, world)
        return eng.run_resident(dtab, ct, params)

    def step_e2e():
        if samples:
            return eng.run_host(ptab, ct, params)
        if world > 1:
            # every rank uploads 1/world of the rows over its own PCIe link; NVLink all-gather rebuilds the columns
            eng.upload_sharded(ptab, dtab, rank, world)
            st = eng.run_sharded(dtab, ct, params, rank, world)
            if rank == 0:                                     # the job's result leaves the box once
                ptab.out_cluster[:R].copy_(dtab.out_cluster[:R], non_blocking=True)
                ptab.out_n_reads[:R].copy_(dtab.out_n_reads[:R], non_blocking=True)
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
    t_first = time.time()
    ms, sts, launches = timed(step_resident, args.steps)
    step_e2e()
    ms_e2e, sts_e2e, _ = timed(step_e2e, args.steps)
    e2e_mode = "one blocking C-ABI call per step (pinned host columns in, host results out)" if (world == 1 or samples) else \
        "every rank uploads 1/%d of the rows (19 B/row on the wire), NVLink all-gather of the columns, sharded step, rank 0 downloads the result" % world
    h2d_bytes = ptab.h2d_bytes
    if world == 1 and args.e2e_depth > 1:
        # the same call, `e2e_depth` in flight: table k+1 uploads while table k computes (fslr_b200.engine.HostPipeline)
        pipe = HostPipeline(local, args.e2e_depth)
        ptabs = [PinnedTable(ct, compact=True) for _ in range(args.e2e_depth)]
        h2d_pipe = ptabs[0].h2d_bytes
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
        for pt in ptabs[1:]:
            assert np.array_equal(pt.out_cluster[:R].numpy(), ptab.out_cluster[:R].numpy()), "pipelined e2e results disagree"
        assert np.array_equal(ptabs[0].out_cluster[:R].numpy(), ptab.out_cluster[:R].numpy()), "pipelined e2e results disagree"
        pipe.close()
        print("[bench] e2e: serial %.2f ms/step, %d in flight with compact columns %.2f ms/step" % (ms_e2e / args.steps, args.e2e_depth, ms_pipe / args.steps), file=sys.stderr)
        if ms_pipe < ms_e2e:
            e2e_serial_ms = ms_e2e / args.steps
            ms_e2e = ms_pipe
            h2d_bytes = h2d_pipe
            e2e_mode = "%d C-ABI calls in flight (upload of table k+1 overlaps the kernels of table k), chrom as uint8, n_alignments " \
                       "as uint16, aln_size (= qend - qstart) and read_id (run lengths) rebuilt on the device: 19 B/row on the wire; one call at a time with int32 columns: %.2f ms/step" % (args.e2e_depth, e2e_serial_ms)

    # throughput over resident tables with two contexts in flight (the latency-bound replay of one table beside the issue-bound
    # kernels of the next): reported beside `value`, which stays the one-table-at-a-time figure
    resident_pipelined = None
    if world == 1 and args.e2e_depth > 1:
        pipe = HostPipeline(local, 2)
        dts = [dtab, DeviceTable(ct, eng.device)]
        wf = []
        for i in range(2 * max(args.warmup, 2)):
            if i >= 2:
                wf[i - 2].result()
            wf.append(pipe.submit_resident(dts[i % 2], ct, params))
        for f in wf:
            f.result()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for s_ in pipe.streams:
            s_.wait_event(e0)
        futs = []
        for i in range(2 * args.steps):
            if i >= 2:
                futs[i - 2].result()
            futs.append(pipe.submit_resident(dts[i % 2], ct, params))
        for f in futs:
            f.result()
        for s_ in pipe.streams:
            cur.wait_stream(s_)
        e1.record(cur)
        torch.cuda.synchronize()
        msp = e0.elapsed_time(e1) / (2 * args.steps)
        assert np.array_equal(dts[1].out_cluster[:R].cpu().numpy(), dtab.out_cluster[:R].cpu().numpy())
        pipe.close()
        del dts
        resident_pipelined = {"depth": 2, "ms_per_step": msp, "value": R / (msp * 1e-3), "unit": UNIT}

    clocks = sampler.stop(t_first, time.time()) if sampler else {}

    # correctness guard inside the bench: both paths agree with each other
    res_a = dtab.out_cluster[:R].cpu().numpy()
    if world == 1:
        assert np.array_equal(res_a, ptab.out_cluster[:R].numpy()), "resident and host paths disagree"

    st = sts[-1]
    stage_ms = {k: float(np.mean([s["stage_ms"][k] for s in sts])) for k in st["stage_ms"]}
    jobs = world if samples else 1                            # tables clustered per step over all ranks
    value = jobs * R * args.steps / (ms * 1e-3)
    e2e = jobs * R * args.steps / (ms_e2e * 1e-3)
    peak, peak_kind = load_peaks()
    F, D, Q = st["n_fillings"], st["n_intervals"], st["n_query_reads"]
    # algorithmic HBM bytes per stage: what one pass over the stage's data must move at least once
    E, NP = st["relation_entries"], st["partner_records"]
    alg_bytes = {
        "keep_fillings": 12 * A + 8 * ct.n_reads + 4 * F,
        "data_order_mask": 2 * 8 * F + 16 * F + 24 * D,
        "query_rank_read_lists": 2 * 8 * D + 8 * D + 32 * Q,
        "chrom_sort": 2 * 8 * D,
        "records_bands": 28 * D + 80 * D,
        "pair_kernel": 16 * D + 32 * D + 16 * Q + 8 * E + 32 * NP,
        "replay": 32 * NP + 16 * st["saturating_reads"] + 8 * st["edges"],
        "union_find": 8 * E + 8 * Q,
        "numbering": 12 * Q + 16 * ct.n_reads,
    }
    top = max((k for k in alg_bytes), key=lambda k: stage_ms[k])
    ach = alg_bytes[top] / (stage_ms[top] * 1e-3) / 1e9 if stage_ms[top] > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram__bytes_read+write per launch from the committed ncu capture
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.config, {}).get(top)
    roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_kind, "algorithmic_bytes_per_launch": alg_bytes[top],
                "ms_per_launch": stage_ms[top],
                "note": "the pair kernel is instruction-issue bound (see int_issue), every other stage HBM bound"}
    # integer-issue view of the pair kernel: 12 + 7*L1*L2 lane-ops per evaluated read pair against the
    # measured dependent-free IADD/LOP/VIMNMX rate of this device
    int_issue = None
    if rank == 0 and stage_ms["pair_kernel"] > 0:
        ipeak = eng.int_peak()
        Lbar = D / max(Q, 1)
        ops = st["pair_tests"] * (12 + 7 * Lbar * Lbar)
        int_issue = {"kernel": "k_pair", "achieved_lane_ops_per_s": ops / (stage_ms["pair_kernel"] * 1e-3), "peak_lane_ops_per_s": ipeak,
                     "frac": ops / (stage_ms["pair_kernel"] * 1e-3) / ipeak, "ops_per_pair_test": 12 + 7 * Lbar * Lbar,
                     "peak_source": "measured in this run (fslrc_int_peak: 8 independent add/minmax/xor chains per thread)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if samples else "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.config, WORKLOADS[args.config]) + ("" if args.scale == 1.0 else " x%g" % args.scale),
                       "reads": R, "table_rows": A, "fillings": F, "intervals": D,
                       "l2": "inputs larger than L2 (%.2f GB of columns re-read every step)" % (32 * A / 1e9),
                       "parallelism": "1 GPU" if world == 1 else ("one table per GPU on %d GPUs, no data-path collective" % world if samples else
                                                                  "table replicated (e2e: upload sharded, NVLink all-gather), pair space sharded by "
                                                                  "query read over %d GPUs, all-reduce + all-gather of forests" % world)},
            "pair_tests_per_s": st["pair_tests"] / (stage_ms["pair_kernel"] + stage_ms["replay"]) * 1e3 if stage_ms["pair_kernel"] > 0 else None,
            "pair_tests": st["pair_tests"], "band_pairs": st["band_pairs"], "edges": st["edges"], "clusters": st["components"],
            "saturating_reads": st["saturating_reads"],
            "partner_records": st["partner_records"], "resident_two_in_flight": resident_pipelined,
            "stage_ms": stage_ms, "roofline": roofline, "int_issue": int_issue, "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": ptab.d2h_bytes, "mode": e2e_mode},
            "gpu_launches": launches}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            n = min(args.cpu_sample_reads, R)
            s, ost, nr = cpu_port_run(args.config, n)
            line["cpu_baseline"] = {"value": nr / s, "unit": UNIT, "cores": 1, "kind": "port",
                                    "pair_tests_per_s": ost["pair_tests"] / s,
                                    "sample": cpu_sample_text(args.config, nr, " once (%.1f s)" % s)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
