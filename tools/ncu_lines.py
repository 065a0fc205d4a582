#!/usr/bin/env python
"""Attribute the per-instruction counters of an .ncu-rep (source page, SASS view) to CUDA source lines, using the line table
nvdisasm prints for the built library.  Usage: ncu_lines.py report.ncu-rep kernel-substring [min_pct]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "modern-search-engines-project_b200", "csrc", "libmsegpu.so")


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.6
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    seq, cur, on = [], None, False
    for l in dis.split("\n"):
        if l.startswith("//-") and ".text." in l:
            on = kern in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            seq.append((m.group(2).strip(), cur))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, data = rows[1], rows[2:]
    isamp, iex = h.index("# Samples"), h.index("Instructions Executed")
    n = min(len(seq), len(data))
    by, bys, tot, tots = collections.Counter(), collections.Counter(), 0, 0
    for i in range(n):
        e, s = int(data[i][iex]), int(data[i][isamp])
        by[seq[i][1]] += e; bys[seq[i][1]] += s; tot += e; tots += s
    src = {}
    print(f"# {len(seq)} SASS lines (nvdisasm) vs {len(data)} (ncu); {tot} warp instructions, {tots} samples")
    for k, v in sorted(by.items(), key=lambda kv: (kv[0] is None, kv[0])):
        if not k or (v < tot * min_pct / 100 and bys[k] < tots * min_pct / 100):
            continue
        f = os.path.join(ROOT, "modern-search-engines-project_b200", "csrc", k[0])
        if k[0] not in src:
            src[k[0]] = open(f).read().split("\n") if os.path.exists(f) else None
        text = src[k[0]][k[1] - 1].strip()[:95] if src[k[0]] else ""
        print(f"{k[0]}:{k[1]:4d}  inst {100 * v / tot:5.1f}%  samples {100 * bys[k] / tots:5.1f}%  {text}")


if __name__ == "__main__":
    main()
