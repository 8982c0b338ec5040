"""Cross-check the three CPU statements of the RNG protocol: pure Python (tests/rng_twin.py),
the C oracle twins, and the Random123 Philox4x32-10 known answers.  CPU-only."""
import numpy as np

import rng_twin
from oracle import oracle as O

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    for ctr, key, out in KAT:
        assert rng_twin.philox4x32_10(ctr, key) == out
        assert [int(x) for x in O.philox(ctr, key)] == out


def test_deck_and_flips_twins_agree():
    for seed, env, ep in [(0, 0, 0), (1, 2, 3), (2**63 + 5, 2**33 + 7, 99), (42, 10**6, 2**31)]:
        d = O.rng_deck(seed, env, ep)
        assert d.tolist() == rng_twin.deck(seed, env, ep)
        assert np.bincount(d + 2).tolist() == [10] * 15
        for N in (1, 4, 12):
            f = O.rng_flips(seed, env, ep, N)
            assert [tuple(x) for x in f.tolist()] == rng_twin.flips(seed, env, ep, N)
            assert np.all(f[:, 0] != f[:, 1]) and f.max() < 12


def test_policy_and_reshuffle_twins_agree():
    rng = np.random.default_rng(3)
    for _ in range(200):
        mask = (rng.random(26) < 0.4).astype(np.int8)
        if mask.sum() == 0:
            mask[rng.integers(26)] = 1
        seed, env, t = int(rng.integers(2**62)), int(rng.integers(2**40)), int(rng.integers(2**40))
        a = O.rng_policy(seed, env, t, mask)
        assert a == rng_twin.policy(seed, env, t, mask) and mask[a] == 1
    for _ in range(50):
        n = int(rng.integers(1, 120))
        pile = rng.integers(-2, 13, n).astype(np.int8)
        seed, env, ep, q = int(rng.integers(2**62)), int(rng.integers(2**40)), int(rng.integers(2**31)), int(rng.integers(300))
        out = O.rng_reshuffle(seed, env, ep, q, pile)
        assert out.tolist() == rng_twin.reshuffle(seed, env, ep, q, pile.tolist())
        assert sorted(out.tolist()) == sorted(pile.tolist())


def test_deal_statistics():
    # slot-value marginals 1/15, flip pairs uniform over the 66 pairs,
    # initial discard top uniform with mean 5.
    n = 3000
    decks = np.stack([O.rng_deck(11, e, 0) for e in range(n)])
    assert abs(decks[:, 0].mean() - 5.0) < 4 * np.sqrt(18.67 / n)
    assert abs(decks[:, 149].mean() - 5.0) < 4 * np.sqrt(18.67 / n)
    counts = np.stack([np.bincount(decks[:, k] + 2, minlength=15) for k in (0, 77, 149)])
    assert np.all(np.abs(counts - n / 15) < 5 * np.sqrt(n / 15))
    pairs = np.zeros((12, 12), dtype=int)
    for e in range(n):
        a, b = O.rng_flips(11, e, 0, 1)[0]
        pairs[min(a, b), max(a, b)] += 1
    iu = np.triu_indices(12, 1)
    assert np.all(np.abs(pairs[iu] - n / 66) < 5 * np.sqrt(n / 66))
