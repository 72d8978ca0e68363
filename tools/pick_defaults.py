#!/usr/bin/env python
"""Turns probe lines into decisions: reads `also.optin_probes` of a bench.py JSON line (BENCH_rNN.json, or any file
with such a line) and / or .jsonl files written by tools/probe_optin.py / tools/fastbench.py, and prints per
question the candidates that passed their parity check, fastest first, with the gain over today's default.

    python tools/pick_defaults.py BENCH_r01.json [gpurun_out/probe_optin.jsonl gpurun_out/fastbench.jsonl ...]
"""
import collections
import json
import sys


def probes_in(obj, out):
    """Every `optin_probes` list anywhere inside a parsed JSON document (the driver may wrap the bench line)."""
    if isinstance(obj, dict):
        for k, v in obj.items():
            if k == "optin_probes" and isinstance(v, list):
                for p in v:
                    if isinstance(p, dict):
                        out.extend(p["lines"]) if "lines" in p else out.append(p)
            else:
                probes_in(v, out)
    elif isinstance(obj, list):
        for v in obj:
            probes_in(v, out)
    elif isinstance(obj, str) and obj.lstrip().startswith("{") and "optin_probes" in obj:
        try:
            probes_in(json.loads(obj), out)  # a bench line kept as a string
        except ValueError:
            pass


def lines_of(path):
    out = []
    try:
        probes_in(json.load(open(path)), out)  # one JSON document, possibly pretty-printed
        if out:
            return out
    except ValueError:
        pass
    for ln in open(path):
        ln = ln.strip()
        if not ln.startswith("{"):
            continue
        try:
            d = json.loads(ln)
        except ValueError:
            continue
        found = []
        probes_in(d, found)
        if found:
            out.extend(found)
        elif "what" in d or "workload" in d:
            out.append(d)
    return out


def table(title, groups, default_key, fmt):
    print("\n== " + title)
    for g in sorted(groups):
        best = {}
        for key, ms in groups[g]:
            best[key] = min(ms, best.get(key, 1e30))
        base = best.get(default_key)
        ranked = sorted(best.items(), key=lambda kv: kv[1])
        cells = ["%s %.4g%s" % (fmt(k), v, "" if base is None or k == default_key else " (%+.1f%%)" % (100 * (v / base - 1))) for k, v in ranked]
        print("  %-24s %s" % (g, " | ".join(cells)))


def main():
    rows = [r for p in sys.argv[1:] for r in lines_of(p)]
    bad = [r for r in rows if r.get("ok") is False]
    f2d, lag, real2d, cluster, latency, managed, variants = (collections.defaultdict(list) for _ in range(7))
    for r in rows:
        if r.get("ok") is False:
            continue
        w = r.get("what", "")
        if w == "c2c2d" and "fused_mode" in r:
            f2d["%dx%d x%d" % (r["nx"], r["ny"], r["batch"])].append((r["fused_mode"], r["ms"]))
        elif w.startswith("c2c2d_lag"):
            lag["%dx%d x%d" % (r["nx"], r["ny"], r["batch"])].append((int(w[9:]), r["ms"]))
        elif w == "r2c2d":
            real2d["%dx%d x%d" % (r["nx"], r["ny"], r["batch"])].append((r["fused_real"], r["ms"]))
        elif w == "c2c1d_cluster":
            cluster["N=%d dir %+d" % (r["n"], r["direction"])].append((r["cluster_rows"], r["ms"]))
        elif w == "single_call_latency":
            latency["N=%d" % r["n"]].append((r["spin_sync"], r["us_per_call"]))
        elif w == "managed_round_trip":
            managed["%s %dx%d" % (r["kind"], r["nx"], r["ny"])].append((r["prefetch"], r["ms"]))
        elif "workload" in r and "label" in r:
            variants[r["workload"]].append((r["label"], r["ms"]))
    if f2d:
        table("fused 2D C2C: ms per flavour (mufft_cuda_set_fused_2d; 1 = today's per-shape default)", f2d, 1, lambda k: "m%d" % k)
    if lag:
        for g in lag:  # the default lag is what flavour 1 of the same shape ran with
            lag[g] += [(0, ms) for k, ms in f2d.get(g, []) if k == 1]
        table("fused 2D C2C: ms per lag (0 = the launcher's choice)", lag, 0, lambda k: "lag%d" % k)
    if real2d:
        table("2D R2C: ms, 0 = row kernel + column kernel, 1 = fused kernel", real2d, 0, lambda k: "fused=%d" % k)
    if cluster:
        table("rows of 32768 / 65536 points: ms, 0 = two kernels, 1 / 2 = cluster kernel flavours", cluster, 0, lambda k: "cluster=%d" % k)
    if latency:
        table("one transform per fft.h call: us, spin_sync 0 / 1", latency, 0, lambda k: "spin=%d" % k)
    if managed:
        table("host fills, one call, host reads (mufft_alloc memory): ms per round, managed prefetch level", managed, 0, lambda k: "prefetch=%d" % k)
    if variants:
        table("library builds (tools/ab_r2.txt): ms per workload", variants, "base", str)
    if bad:
        print("\n== PARITY FAILED on %d lines:" % len(bad))
        for r in bad:
            print("  ", json.dumps(r)[:240])
    if not rows:
        print("no probe lines found in", sys.argv[1:])


if __name__ == "__main__":
    main()
