"""Source-line view of an ncu capture: joins `ncu --page source --csv` (per SASS instruction: executed count, stall samples,
shared-memory wavefronts) with `nvdisasm -g` of the same kernel (per SASS instruction: file and line) and sums per source line.

    python tools/ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top] [demangled-substring]

The cubin is extracted from the in-tree library (cuobjdump -xelf <name>.sm_100a.cubin libmvtools_b200.so); the join is by
instruction index and checks that the opcodes agree, i.e. that the capture was taken with this build's device code."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def sass_rows(report, kernel):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows, hdr, take = [], None, False
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == "Kernel Name":
            take = kernel in r[1]
            hdr = None
            continue
        if not take:
            continue
        if hdr is None:
            hdr = r
            continue
        rows.append(dict(zip(hdr, r)))
    return rows


def disasm_lines(cubin, kernel):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    insts, cur, inside = [], None, False
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            inside = kernel in m.group(1)
            continue
        if line.startswith("\t.section") or line.startswith(".section"):
            inside = False
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            insts.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return insts


def main():
    report, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    rows = sass_rows(report, sys.argv[5] if len(sys.argv) > 5 else kernel)
    insts = disasm_lines(cubin, kernel)
    print("ncu instructions:", len(rows), " nvdisasm instructions:", len(insts))
    n = min(len(rows), len(insts))
    mism = 0
    for i in range(n):
        a = rows[i]["Source"].split()[0].lstrip("@!P0123456789U ") if rows[i]["Source"].split() else ""
        op_n = re.sub(r"^@!?U?P\d+\s+", "", rows[i]["Source"].strip()).split()[0] if rows[i]["Source"].strip() else ""
        op_d = re.sub(r"^@!?U?P\d+\s+", "", insts[i][1]).split()[0] if insts[i][1] else ""
        if op_n.split(".")[0] != op_d.split(".")[0]:
            mism += 1
    print("opcode mismatches:", mism)
    per = defaultdict(lambda: defaultdict(float))
    tot = defaultdict(float)
    keys = ["Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "stall_short_sb", "stall_wait", "stall_long_sb", "stall_math", "stall_branch_resolving", "stall_not_selected", "stall_no_inst", "stall_mio"]
    for i in range(n):
        loc = insts[i][2] or ("?", 0)
        for k in keys:
            try:
                v = float(rows[i].get(k, "0") or 0)
            except ValueError:
                v = 0.0
            per[loc][k] += v
            tot[k] += v
        per[loc]["n_sass"] += 1
    print("totals:", {k: int(v) for k, v in tot.items()})
    order = sorted(per.items(), key=lambda kv: -kv[1]["# Samples"])
    print("%-22s %6s %8s %8s %7s %7s %7s %7s %7s" % ("file:line", "sass", "inst%", "samp%", "wavef%", "ideal%", "shortsb", "wait", "longsb"))
    for loc, d in order[:top]:
        print("%-22s %6d %8.2f %8.2f %7.2f %7.2f %7.0f %7.0f %7.0f" % ("%s:%d" % loc, d["n_sass"], 100 * d["Instructions Executed"] / max(tot["Instructions Executed"], 1), 100 * d["# Samples"] / max(tot["# Samples"], 1),
                                                                 100 * d["L1 Wavefronts Shared"] / max(tot["L1 Wavefronts Shared"], 1), 100 * d["L1 Wavefronts Shared Ideal"] / max(tot["L1 Wavefronts Shared"], 1),
                                                                 d["stall_short_sb"], d["stall_wait"], d["stall_long_sb"]))


if __name__ == "__main__":
    main()
