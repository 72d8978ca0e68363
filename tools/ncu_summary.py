#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full --import-source on) into the text summary kept under
profiles/: headline metrics per captured launch, instruction mix by SASS opcode with stall samples,
and the hottest source lines.  Runs where there is no GPU (ncu -i only reads the report).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--launch 0] > profiles/rNN_name.txt
"""
import argparse
import collections
import csv
import io
import re
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_summary(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        print(f"== launch {n}: {name}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:78s} {r[i]:>18s} {units[i]}")
        try:
            t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
            unit = units[hdr.index("gpu__time_duration.sum")]
            t_s = t * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(unit, 1e-6)
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")) * scale[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", "")) * scale[units[hdr.index("dram__bytes_write.sum")]]
            print(f"   derived: DRAM traffic {rd + wr:.0f} B per launch, {1e-9 * (rd + wr) / t_s:.1f} GB/s achieved")
        except (ValueError, KeyError):
            pass


def source_summary(rep, top):
    rows = ncu_csv(rep, "source", ["--print-source", "sass"])
    # the file holds one table per launch; take the first
    tables, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            tables.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    if not tables:
        return
    t = tables[0]
    hdr = t["hdr"]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def f(r, name):
        try:
            return float(r[col[name]])
        except (ValueError, KeyError):
            return 0.0
    by_op = collections.defaultdict(lambda: collections.Counter())
    total_samples = sum(f(r, "# Samples") for r in t["rows"])
    total_inst = sum(f(r, "Instructions Executed") for r in t["rows"])
    for r in t["rows"]:
        sass = r[col["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
        op = m.group(2) if m else sass.split()[0]
        op = ".".join(op.split(".")[:3])
        c = by_op[op]
        c["inst"] += f(r, "Instructions Executed")
        c["samples"] += f(r, "# Samples")
        c["wf_shared"] += f(r, "L1 Wavefronts Shared")
        c["wf_shared_ideal"] += f(r, "L1 Wavefronts Shared Ideal")
        c["l2_sectors"] += f(r, "L2 Theoretical Sectors Global")
        c["l1_tags"] += f(r, "L1 Tag Requests Global")
    print(f"\n== SASS opcode mix of '{t['name']}' (first captured launch): {total_inst:.0f} warp instructions, {total_samples:.0f} stall samples")
    print(f"   {'opcode':28s} {'warp inst':>12s} {'%inst':>6s} {'%samples':>8s} {'smem wavefronts':>16s} {'(ideal)':>12s} {'L1 tag req':>12s} {'L2 sectors':>12s}")
    for op, c in sorted(by_op.items(), key=lambda kv: -kv[1]["inst"])[:top]:
        print(f"   {op:28s} {c['inst']:12.0f} {100 * c['inst'] / max(total_inst, 1):6.1f} {100 * c['samples'] / max(total_samples, 1):8.1f} "
              f"{c['wf_shared']:16.0f} {c['wf_shared_ideal']:12.0f} {c['l1_tags']:12.0f} {c['l2_sectors']:12.0f}")
    agg = collections.Counter()
    for r in t["rows"]:
        for s in stall_cols:
            agg[s] += f(r, s)
    tot = sum(agg.values())
    print("\n== stall reasons (all samples): " + ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in agg.most_common(8)))
    print(f"\n== hottest {top} SASS instructions by stall samples")
    for r in sorted(t["rows"], key=lambda r: -f(r, "# Samples"))[:top]:
        main = max(stall_cols, key=lambda s: f(r, s))
        print(f"   {f(r, '# Samples'):8.0f} ({100 * f(r, '# Samples') / max(total_samples, 1):4.1f}%) {main[6:]:12s} {r[col['Source']].strip()[:110]}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=24)
    args = ap.parse_args()
    raw_summary(args.report)
    source_summary(args.report, args.top)
