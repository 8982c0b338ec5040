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
    ap.add_argument("--opcodes", action="store_true", help="dynamic opcode mix instead of the per-line table")
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
    if a.opcodes:
        # dynamic opcode mix (warp-instructions) and how much of it ran at low lane counts
        ops = collections.defaultdict(lambda: [0, 0])
        buckets = collections.defaultdict(int)
        for k in range(n):
            r = rows[k]
            inst, tinst = int(r[ci["Instructions Executed"]]), int(r[ci["Thread Instructions Executed"]])
            m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ci["Source"]].strip())
            op = m.group(1) if m else "?"
            ops[op][0] += inst
            ops[op][1] += tinst
            if inst:
                lanes = tinst / inst
                buckets["<4" if lanes < 4 else "<12" if lanes < 12 else "<20" if lanes < 20 else "<28" if lanes < 28 else ">=28"] += inst
        print("opcode        warp-inst      %  lanes")
        for op, g in sorted(ops.items(), key=lambda kv: -kv[1][0])[: a.top]:
            print(f"{op:12s} {g[0]:10d} {100 * g[0] / max(tot[0], 1):6.2f} {g[1] / max(g[0], 1):6.1f}")
        print("warp-inst by active lanes:", {k: f"{100 * v / max(tot[0], 1):.1f}%" for k, v in sorted(buckets.items())})
        return
    print(f"{'file:line':32s} {'sass':>5s} {'warp-inst':>11s} {'%':>6s} {'lanes':>6s} {'samples':>8s} {'%':>6s}")
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print(f"{key[0] + ':' + str(key[1]):32s} {g[3]:5d} {g[0]:11d} {100 * g[0] / max(tot[0], 1):6.2f} "
              f"{g[1] / max(g[0], 1):6.1f} {g[2]:8d} {100 * g[2] / max(tot[2], 1):6.2f}")


if __name__ == "__main__":
    main()
