"""Pins the oracle's Philox4x32-10 against the published Random123 known-answer vectors
(Salmon, Moraes, Dror, Shaw, SC'11; Random123 kat_vectors: philox4x32 10 rounds)."""
import numpy as np
import pytest

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


@pytest.mark.parametrize("ctr,key,expect", KAT)
def test_philox_known_answers(oracle, ctr, key, expect):
    from oracle import constv_numpy as onp
    assert [int(x) for x in oracle.philox4x32_10(ctr, key)] == expect
    assert [int(x) for x in onp.philox4x32_10(ctr, key)] == expect


def test_philox_c_matches_python_random_counters(oracle):
    from oracle import constv_numpy as onp
    rng = np.random.default_rng(7)
    for _ in range(200):
        ctr = rng.integers(0, 2**32, 4, dtype=np.uint64).astype(np.uint32)
        key = rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(oracle.philox4x32_10(ctr, key), onp.philox4x32_10(ctr, key))


def test_gaussian_stream_moments_and_keying(oracle):
    n = 200_000
    xi = oracle.fill_noise(n, seed=42, step_index=5)
    assert abs(xi.mean()) < 0.01
    assert abs(xi.std() - 1.0) < 0.01
    assert np.all(np.isfinite(xi))
    assert np.max(np.abs(xi)) < 6.0  # 24-bit uniforms: |xi| <= sqrt(2*25*ln2) = 5.89
    # keyed by (seed, atom, step): each of the three changes the block; same key reproduces it
    a = oracle.gaussian4(42, 17, 5)
    assert np.array_equal(a, xi[17])
    assert not np.array_equal(a, oracle.gaussian4(43, 17, 5))
    assert not np.array_equal(a, oracle.gaussian4(42, 18, 5))
    assert not np.array_equal(a, oracle.gaussian4(42, 17, 6))
    # stream is addressed by ORIGINAL atom index: a permuted slot order permutes the rows
    perm = np.random.default_rng(0).permutation(1000).astype(np.int32)
    assert np.array_equal(oracle.fill_noise(1000, 42, 5, atom_index=perm), xi[perm])
    # and by GLOBAL index: a shard starting at atom 300 sees rows 300...
    assert np.array_equal(oracle.fill_noise(100, 42, 5, global_atom_offset=300), xi[300:400])
