#!/usr/bin/env python
"""Static evidence for profiles/: registers, stack (spills), static shared memory and the SASS
mnemonic mix of every kernel in libmufft_b200.so, read with cuobjdump (no GPU needed).

    python tools/sass_summary.py [--lib mufft_b200/libmufft_b200.so] [--filter REGEX] > profiles/rNN_static.txt

The mnemonic columns are STATIC instruction counts (not executed counts): they show which memory
instructions a kernel is built from (LDG/STG width, LDS/STS width, bulk copies UBLKCP / tensor-map
loads UTMALDG, cluster stores, packed FP32x2 adds FADD2) and whether it spills (STL/LDL)."""
import argparse
import collections
import re
import subprocess

COLS = ["LDG.64", "LDG.128", "STG.64", "STG.128", "LDS.64", "LDS.128", "STS.64", "STS.128", "LDL/STL",
        "FADD", "FADD2", "FFMA", "FMUL", "BAR", "MEMBAR", "CCTL", "UBLKCP", "UTMALDG", "SYNCS", "ST.cluster"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def classify(op):
    """SASS opcode (with modifiers) -> column name or None."""
    base = op.split(".")[0]
    if base in ("LDG", "STG", "LDS", "STS"):
        w = "128" if ".128" in op else ("64" if ".64" in op else None)
        return "%s.%s" % (base, w) if w else None
    if base in ("LDL", "STL"):
        return "LDL/STL"
    if base in ("FADD", "FADD2", "FFMA", "FMUL", "MEMBAR", "CCTL", "UBLKCP", "UTMALDG", "SYNCS"):
        return base
    if base == "BAR":
        return "BAR"
    if base == "ST" and ".CLUSTER" in op.upper():
        return "ST.cluster"
    if base == "STAS":
        return "ST.cluster"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default="mufft_b200/libmufft_b200.so")
    ap.add_argument("--filter", default="", help="regex on the demangled kernel name")
    args = ap.parse_args()

    res = subprocess.run(["cuobjdump", "--dump-resource-usage", args.lib], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            usage[name] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            name = None

    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True).stdout
    mix = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            total[cur] += 1
            col = classify(m.group(1))
            if col:
                mix[cur][col] += 1

    names = demangle(sorted(usage))
    rows = []
    for mangled, u in usage.items():
        d = names.get(mangled, mangled)
        d = re.sub(r"^void ", "", d).replace("mufft::", "").replace("(int)", "").replace("(bool)", "")
        d = re.sub(r"\(mcu_\w+\)$", "", d)
        if args.filter and not re.search(args.filter, d):
            continue
        rows.append((d, u, mix[mangled], total[mangled]))
    rows.sort(key=lambda r: r[0])

    print("# %d kernels in %s (sm_100a); static counts from cuobjdump -sass / --dump-resource-usage" % (len(rows), args.lib))
    spilled = [r[0] for r in rows if r[1].get("STACK", 0) or r[1].get("LOCAL", 0) or r[2]["LDL/STL"]]
    print("# kernels with stack / local memory (spills): %d%s" % (len(spilled), "" if not spilled else ": " + "; ".join(spilled)))
    hdr = "%-66s %4s %5s %6s %6s " % ("kernel", "regs", "stack", "smem", "insts") + " ".join("%7s" % c for c in COLS)
    print(hdr)
    for d, u, mx, tot in rows:
        print("%-66s %4d %5d %6d %6d " % (d[:66], u.get("REG", 0), u.get("STACK", 0), u.get("SHARED", 0), tot) +
              " ".join("%7d" % mx[c] for c in COLS))


if __name__ == "__main__":
    main()
