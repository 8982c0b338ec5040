"""Timeline of the fused policy kernel's first CTA (csrc/skyjo_policy.cu): clock64() marks of the three roles,
recorded by skyjo_policy_trace, printed as cycles per phase averaged over the steady-state tiles.
    python tools/policy_trace.py [--envs 262144]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyjo_rl_b200 import BatchedSkyjoEnv, _lib  # noqa: E402
from skyjo_rl_b200.policy import ActionMaskPolicy, FusedPolicy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 18)
    a = ap.parse_args()
    env = BatchedSkyjoEnv(num_envs=a.envs, num_players=4, seed=0)
    env.reset()
    env.step_random(40)
    torch.manual_seed(0)
    fused = FusedPolicy(ActionMaskPolicy(env.obs_len).to(env.device), env, with_value=False)
    L = _lib.load()
    n = int(L.skyjo_policy_trace_len())
    acts = torch.empty(a.envs, dtype=torch.uint8, device=env.device)
    for _ in range(3):
        fused.sample(0, acts)
    tr = torch.zeros((3, n), dtype=torch.int64, device=env.device)
    _lib.check(L.skyjo_policy_trace(env._h, fused._packed.data_ptr(), acts.data_ptr(), tr.data_ptr(), env._stream()))
    torch.cuda.synchronize()
    t = tr.cpu().numpy()
    H, S, M = (x[x > 0] for x in t)
    tiles = len(H) // 5
    print(f"CTA 0: {tiles} tiles; kernel span {int(max(H.max(), S.max(), M.max()) - min(H.min(), S.min(), M.min()))} cycles")
    Hm = H[:tiles * 5].reshape(tiles, 5)
    Sm = S[:tiles * 5].reshape(tiles, 5)
    Mm = M[:tiles * 18].reshape(tiles, 18)
    sl = slice(2, tiles - 1)
    per_tile = np.diff(Hm[:, 0])[1:-1].mean()
    print(f"steady-state tile period (team H): {per_tile:.0f} cycles")
    d = Hm[sl]
    print("team H  wait D1 %.0f | epilogue 1 %.0f | wait D2 %.0f | epilogue 2 %.0f" % tuple(
        (d[:, k + 1] - d[:, k]).mean() for k in range(4)))
    d = Sm[sl]
    print("team S  wait slot %.0f | stage x(j+1) %.0f | wait logits %.0f | sample %.0f" % tuple(
        (d[:, k + 1] - d[:, k]).mean() for k in range(4)))
    d = Mm[sl]
    print("MMA     wait x(j), MMA3(j-1) %.0f | MMA1 issue -> first h1 chunk seen %.0f" % (
        (d[:, 1] - d[:, 0]).mean(), (d[:, 2] - d[:, 1]).mean()))
    print("        MMA2 chunk marks (cycles after the previous): " + " ".join("%.0f" % (d[:, k] - d[:, k - 1]).mean() for k in range(3, 10)))
    print("        MMA3 chunk marks: " + " ".join("%.0f" % (d[:, k] - d[:, k - 1]).mean() for k in range(10, 18)))
    # offsets of the roles against team H's tile start
    base = Hm[sl, 0]
    print("offsets from H0 (tile start): H1 %.0f H2 %.0f H3 %.0f H4 %.0f | S1 %.0f S2 %.0f S3 %.0f S4 %.0f | M1 %.0f M2 %.0f M9 %.0f M17 %.0f" % (
        *(Hm[sl, k] - base for k in range(1, 5)),) if False else "")
    for name, arr, ks in (("H", Hm, range(5)), ("S", Sm, range(5)), ("M", Mm, (0, 1, 2, 9, 10, 17))):
        print(name, "marks relative to the tile's H0:", " ".join(f"{k}:{(arr[sl, k] - base).mean():.0f}" for k in ks))


if __name__ == "__main__":
    main()
