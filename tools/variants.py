#!/usr/bin/env python
"""Development aid: build experimental variants of libskyjo_b200.so (extra -D flags, N=4 only by
default) into build/variants/ and, on a GPU box, time each with tools/quick_bench.py.

    python tools/variants.py build  name1:-DFOO=1,-DBAR name2: ...      # here (nvcc cross-compiles)
    python tools/variants.py run [--players 4] [--env K=V ...] name1 name2 ...   # under gpurun
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "build", "variants")


def main():
    cmd, args = sys.argv[1], sys.argv[2:]
    if cmd == "build":
        from skyjo_rl_b200.build import build_variant
        players = (4,)
        for a in args:
            if a.startswith("--players="):
                players = tuple(int(x) for x in a.split("=")[1].split(","))
                continue
            name, _, defs = a.partition(":")
            defines = [d[2:] if d.startswith("-D") else d for d in defs.split(",") if d]
            out = build_variant(os.path.join(VDIR, f"lib_{name}.so"), defines, players)
            print("built", out, defines)
    else:
        extra, names, envs = [], [], {}
        it = iter(args)
        for a in it:
            if a == "--env":
                k, _, v = next(it).partition("=")
                envs[k] = v
            elif a.startswith("--"):
                extra += [a, next(it)]
            else:
                names.append(a)
        for n in names:
            env = dict(os.environ, **envs)
            if n != "base":
                env["SKYJO_LIB"] = os.path.join(VDIR, f"lib_{n}.so")
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench.py"), "--tag",
                                n + "".join(f" {k}={v}" for k, v in envs.items()), *extra],
                               env=env, capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr.strip()[-2000:], flush=True)


if __name__ == "__main__":
    main()
