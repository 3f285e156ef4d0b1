"""CPU: the numpy Philox4x32-10 restatement (oracle/philox.py) against the Random123 known-answer vectors, and the
mask helper's statistics.  The CUDA generator (csrc/dropout.cuh) is checked against this restatement bit for bit in
tests/test_gpu_dropout.py."""
import numpy as np

from oracle import philox

# Random123 kat_vectors, philox4x32 10 rounds: counter words, key words, expected output
KAT = [
    ((0x00000000,) * 4, (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_mask_values_and_keep_rate():
    for p in (0.1, 0.5, 0.9):
        m = philox.mask_flat(2022, 7, 3, p, 200_000)
        scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
        assert set(np.unique(m)) <= {np.float32(0.0), scale}
        keep = float((m > 0).mean())
        assert abs(keep - (1 - p)) < 4 * np.sqrt(p * (1 - p) / m.size) + 1e-4
    assert np.all(philox.mask_flat(1, 1, 1, 0.0, 17) == 1.0)


def test_masks_differ_across_steps_sites_and_seeds_and_are_reproducible():
    a = philox.mask_flat(5, 1, 0, 0.5, 4096)
    assert np.array_equal(a, philox.mask_flat(5, 1, 0, 0.5, 4096))
    for other in (philox.mask_flat(5, 2, 0, 0.5, 4096), philox.mask_flat(5, 1, 1, 0.5, 4096), philox.mask_flat(6, 1, 0, 0.5, 4096)):
        assert 0.4 < float((a != other).mean()) < 0.6
    # a window of the stream equals the same elements of the full stream (element-indexed, not sequential)
    assert np.array_equal(philox.mask_flat(5, 1, 0, 0.5, 100, start=1001), philox.mask_flat(5, 1, 0, 0.5, 2000)[1001:1101])
