#!/usr/bin/env python
"""Per-source-line instruction and stall profile of one kernel from an .ncu-rep.

    python tools/ncu_lines.py <report.ncu-rep> <object-or-so> <kernel-substring> [--launch K] [--top N]

ncu's SASS page gives per-instruction counts; nvdisasm -gi gives the (inlined) source line of
every instruction of the same cubin.  Both list the kernel's instructions in address order, so
they are joined by position and aggregated by innermost source line and by function-level line.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_page(report, kernel, launch):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and row and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] and row:
            cur["rows"].append(row)
    blocks = [b for b in blocks if kernel in b["name"].replace("(int)", "").replace("(bool)", "")] or blocks
    return blocks[min(launch, len(blocks) - 1)]


def line_info(obj, kernel_mangled_sub, prefer=None):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    lines = []
    for f in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        active, chain, fresh = False, [("?", 0)], True
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                active = kernel_mangled_sub in m.group(1)
                continue
            if ln.startswith("//---") and ".text." in ln:
                active = False
            if not active:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                if fresh:          # first annotation after an instruction starts a new chain
                    chain, fresh = [], False
                chain.append((os.path.basename(m.group(1)), int(m.group(2))))   # innermost first
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
            if m:
                fresh = True
                cur = chain[0]
                if prefer:
                    for fr in chain:
                        if fr[0] == prefer:
                            cur = fr
                            break
                lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
        if lines:
            break
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("obj")
    ap.add_argument("kernel", help="substring of the demangled name, e.g. 'step_kernel<4, 0, 1>'")
    ap.add_argument("--mangled", default=None, help="substring of the mangled name (default: derived)")
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--file", default=None, help="attribute inlined code to its frame in this file (basename)")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    blk = sass_page(a.report, a.kernel, a.launch)
    hdr = blk["hdr"]
    ci = {n: hdr.index(n) for n in ("Instructions Executed", "Thread Instructions Executed", "# Samples", "Source")}
    mangled = a.mangled
    if mangled is None:
        m = re.search(r"(\w+)<(.*)>", a.kernel)
        mangled = m.group(1) if m else a.kernel
    li = line_info(a.obj, mangled, a.file)
    rows = blk["rows"]
    print(f"# {blk['name']}: {len(rows)} SASS instructions in report, {len(li)} in object", file=sys.stderr)
    n = min(len(rows), len(li))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for k in range(n):
        r = rows[k]
        inst, tinst, samp = int(r[ci["Instructions Executed"]]), int(r[ci["Thread Instructions Executed"]]), int(r[ci["# Samples"]])
        key = li[k][1]
        g = agg[key]
        g[0] += inst
        g[1] += tinst
        g[2] += samp
        g[3] += 1
        tot[0] += inst
        tot[1] += tinst
        tot[2] += samp
    print(f"total warp-inst {tot[0]}  thread-inst {tot[1]}  samples {tot[2]}  lanes/inst {tot[1] / max(tot[0], 1):.1f}")
    print(f"{'file:line':32s} {'sass':>5s} {'warp-inst':>11s} {'%':>6s} {'lanes':>6s} {'samples':>8s} {'%':>6s}")
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print(f"{key[0] + ':' + str(key[1]):32s} {g[3]:5d} {g[0]:11d} {100 * g[0] / max(tot[0], 1):6.2f} "
              f"{g[1] / max(g[0], 1):6.1f} {g[2]:8d} {100 * g[2] / max(tot[2], 1):6.2f}")


if __name__ == "__main__":
    main()
