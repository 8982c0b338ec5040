"""Pure-Python restatement of the product's counter-based RNG protocol ("RNG").

Third, independent implementation (besides csrc/skyjo_rng.cuh and oracle/skyjo_oracle.c),
used by tests/golden/make_golden.py to give the live reference the same in-game reshuffle
rule as the GPU, and by the CPU tests to cross-check the other two.
"""
M32 = 0xFFFFFFFF

PURPOSE_DEAL, PURPOSE_FLIPS, PURPOSE_RESHUFFLE, PURPOSE_POLICY = 1, 2, 3, 4


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return [c0, c1, c2, c3]


def block(seed, env, purpose, a, b):
    ctr = [env & M32, ((env >> 32) & 0xFFFFFF) | (purpose << 24), a & M32, b & M32]
    return philox4x32_10(ctr, [seed & M32, (seed >> 32) & M32])


def bounded(r, n):
    return (r * n) >> 32


def deck(seed, env, episode):
    d = [i // 10 - 2 for i in range(150)]
    blk = None
    for i in range(149, 0, -1):
        k = 149 - i
        if k % 4 == 0:
            blk = block(seed, env, PURPOSE_DEAL, episode, k // 4)
        j = bounded(blk[k % 4], i + 1)
        d[i], d[j] = d[j], d[i]
    return d


def flips(seed, env, episode, num_players):
    out = []
    for p in range(num_players):
        blk = block(seed, env, PURPOSE_FLIPS, episode, p)
        a, b = bounded(blk[0], 12), bounded(blk[1], 11)
        if b >= a:
            b += 1
        out.append((a, b))
    return out


def policy(seed, env, t, mask):
    legal = [i for i, m in enumerate(mask) if m]
    g = t >> 2                      # one Philox block per four lockstep steps, step t takes word t & 3
    blk = block(seed, env, PURPOSE_POLICY, g & M32, g >> 32)
    return legal[bounded(blk[t & 3], len(legal))]


def reshuffle(seed, env, episode, q, pile):
    """New python-list order of `pile` (top = last): e_0 (new discard) is out[-1]."""
    bins = [0] * 15
    for v in pile:
        bins[int(v) + 2] += 1
    n = len(pile)
    out = [0] * n
    for d in range(n):
        remaining = n - d
        blk = block(seed, env, PURPOSE_RESHUFFLE, episode, ((q & 0x7F) << 16) | remaining)
        idx = bounded(blk[0], remaining)
        j = 0
        while idx >= bins[j]:
            idx -= bins[j]
            j += 1
        bins[j] -= 1
        out[n - 1 - d] = j - 2
    return out
