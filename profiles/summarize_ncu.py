#!/usr/bin/env python3
"""Per-kernel table and per-stage DRAM traffic of one resident step from an `ncu --set full --page raw --csv` capture.

    python profiles/summarize_ncu.py profiles/r02b_ncu_full_all_kernels_c4.csv [--write-traffic C4]

Prints a markdown table (time, DRAM MB read / written, issue-active %, resident warps %, registers, warp-instructions) and,
with --write-traffic, rewrites the config's entry of profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`)."""
import csv
import json
import os
import sys

STAGE = [  # kernel-name prefix -> bench stage (first match wins)
    ("k_rows_fast", "keep_fillings"), ("k_first_last", "keep_fillings"), ("k_keep", "keep_fillings"), ("k_fill_records", "keep_fillings"),
    ("k_compact_fast", "sort"), ("k_fi_fast", "sort"), ("k_items_fast", "sort"), ("k_tie_delta", "sort"), ("k_firsts_fast", "sort"),
    ("k_assign_fast", "sort"), ("k_apply_delta", "sort"), ("prims::k_rs", "sort"), ("k_mask_flags", "sort"), ("k_build_items", "sort"),
    ("k_records", "band_bucket"), ("k_bands", "band_bucket"), ("k_chrom_bounds", "band_bucket"), ("prims::k_scan_segmax", "band_bucket"),
    ("k_heavy_list", "compaction"), ("k_hits", "compaction"),
    ("k_eval", "pair_kernel"), ("k_pair", "pair_heavy"),
    ("k_light_sat", "saturating_set"), ("k_plinfo", "saturating_set"), ("k_plist", "saturating_set"),
    ("k_run_flags", "replay"), ("k_run_cut", "replay"), ("k_replay", "replay"), ("k_sib", "replay"),
    ("k_union", "union_find"), ("k_iota", "union_find"),
    ("k_flatten", "numbering"), ("k_number", "numbering"), ("k_single_flags", "numbering"),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    h = rows[0]
    col = {c: i for i, c in enumerate(h)}
    units = rows[1]

    def get(r, name):
        return num(r[col[name]]) if name in col else 0.0

    def scale_bytes(name):  # ncu prints bytes in the unit of row 2
        u = units[col[name]].lower() if name in col else "byte"
        return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)

    def scale_time(name):
        u = units[col[name]].lower()
        return {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)

    rd, wr, tm = "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"
    agg, order, stage_bytes = {}, [], {}
    last_stage = "sort"
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        st = next((s for p, s in STAGE if name.startswith(p)), None)
        if st is None:                                   # scans / compactions belong to the stage of the kernel before them
            st = last_stage
        else:
            last_stage = st
        b_r, b_w = get(r, rd) * scale_bytes(rd), get(r, wr) * scale_bytes(wr)
        key = (st, name)
        if key not in agg:
            agg[key] = dict(n=0, us=0.0, rd=0.0, wr=0.0, issue=0.0, warps=0.0, regs=0, inst=0.0)
            order.append(key)
        a = agg[key]
        a["n"] += 1
        a["us"] += get(r, tm) * scale_time(tm)
        a["rd"] += b_r
        a["wr"] += b_w
        a["issue"] = max(a["issue"], get(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"))
        a["warps"] = max(a["warps"], get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"))
        a["regs"] = int(get(r, "launch__registers_per_thread"))
        a["inst"] += get(r, "smsp__inst_executed.sum")
        stage_bytes[st] = stage_bytes.get(st, 0.0) + b_r + b_w
    tot = sum(a["us"] for a in agg.values())
    print("| stage | kernel | launches | us | DRAM rd MB | DRAM wr MB | issue %% | warps %% | regs | M warp-inst |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for key in order:
        a = agg[key]
        print("| %s | `%s` | %d | %.0f | %.0f | %.0f | %.0f | %.0f | %d | %.1f |" % (key[0], key[1][:40], a["n"], a["us"], a["rd"] / 1e6, a["wr"] / 1e6,
                                                                                   a["issue"], a["warps"], a["regs"], a["inst"] / 1e6))
    print("\nstep under ncu: %.2f ms, DRAM traffic %.2f GB" % (tot / 1e3, sum(stage_bytes.values()) / 1e9))
    for s, b in stage_bytes.items():
        print("  %-16s %.3f GB" % (s, b / 1e9))
    if "--write-traffic" in sys.argv:
        cfg = sys.argv[sys.argv.index("--write-traffic") + 1]
        tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
        j = json.load(open(tp)) if os.path.exists(tp) else {}
        ent = {k: int(v) for k, v in stage_bytes.items()}
        ent["_source"] = "profiles/%s (kernel -> stage: profiles/summarize_ncu.py)" % os.path.basename(path)
        j[cfg] = ent
        json.dump(j, open(tp, "w"), indent=1)
        print("wrote", tp)


if __name__ == "__main__":
    main()
