"""Sum the per-launch ncu counters of one direct Fock build into profiles/executed_fp64.json (read by bench.py for `roofline`).

    python tools/executed_from_ncu.py <class_metrics.csv> <key, e.g. "et800|1|1e-16"> [<launches per build>]

The CSV is the `--metrics ... --csv` log of `ncu -k regex:k_shell4 python tools/direct_timing.py child 800` (every class-job launch of the
first build(s)); launches per build defaults to the number of distinct launch IDs divided by the builds seen (4 in direct_timing's child)."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def traffic():
    """--traffic <csv> <key> <kernel substring>: mean DRAM bytes (read + write) and duration per launch of the matching kernels."""
    path, key, pat = sys.argv[2], sys.argv[3], sys.argv[4]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit() and pat in r[4]]
    t = collections.defaultdict(dict)
    for r in rows:
        t[int(r[0])][r[12]] = float(r[14].replace(",", ""))
    n = max(1, len(t))
    dram = sum(v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0) for v in t.values()) / n
    ns = sum(v.get("gpu__time_duration.sum", 0.0) for v in t.values()) / n
    out = os.path.join(ROOT, "profiles", "executed_fp64.json")
    table = json.load(open(out)) if os.path.exists(out) else {}
    table[key] = {"dram_bytes_per_launch": dram, "ncu_us_per_launch": ns / 1e3, "launches": len(t), "kernel": pat, "source": os.path.relpath(os.path.abspath(path), ROOT)}
    json.dump(table, open(out, "w"), indent=1)
    print(json.dumps(table[key], indent=1))


def main():
    if sys.argv[1] == "--traffic":
        return traffic()
    path, key = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    t = collections.defaultdict(dict)
    for r in rows:
        t[int(r[0])][r[12]] = float(r[14].replace(",", ""))
    ids = sorted(t)
    per_build = int(sys.argv[3]) if len(sys.argv) > 3 else len(ids)
    ids = ids[:per_build]
    g = lambda i, m: t[i].get(m, 0.0)
    dfma = sum(g(i, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum") for i in ids)
    dmul = sum(g(i, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum") for i in ids)
    dadd = sum(g(i, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum") for i in ids)
    inst = sum(g(i, "smsp__inst_executed.sum") for i in ids)
    ns = sum(g(i, "gpu__time_duration.sum") for i in ids)
    w = lambda m: sum(g(i, m) * g(i, "gpu__time_duration.sum") for i in ids) / max(ns, 1.0)
    dram = sum(g(i, "dram__bytes_read.sum") + g(i, "dram__bytes_write.sum") for i in ids)
    entry = {"fp64_flops_per_build": 2 * dfma + dmul + dadd, "dfma": dfma, "dmul": dmul, "dadd": dadd, "warp_instructions_per_build": inst,
             "launches_per_build": len(ids), "ncu_kernel_ms_serialised": ns / 1e6,
             "issue_active_pct": w("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "fp64_pipe_active_pct": w("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
             "lanes_per_instruction": w("smsp__thread_inst_executed_per_inst_executed.ratio"),
             "dram_bytes_per_build": dram if dram > 0 else None, "source": os.path.relpath(os.path.abspath(path), ROOT)}
    out = os.path.join(ROOT, "profiles", "executed_fp64.json")
    table = json.load(open(out)) if os.path.exists(out) else {}
    table[key] = entry
    json.dump(table, open(out, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
