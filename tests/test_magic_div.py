"""Host-side check of the division-free position decode of the conv kernels (conv_kernel.cuh `div_magic`,
conv_plan.cu `conv_magic`): q = umulhi(n, floor(2^32 / d)), plus one when the remainder is still >= d, must equal
n // d for every position the planner admits (0 <= n < 2^31) and every row / image pitch."""
import numpy as np


def conv_magic(d: int) -> int:
    return 0xFFFFFFFF if d <= 1 else (1 << 32) // d


def div_magic(n: np.ndarray, d: int) -> np.ndarray:
    q = (n.astype(np.uint64) * np.uint64(conv_magic(d))) >> np.uint64(32)
    q = q.astype(np.int64)
    return q + ((n.astype(np.int64) - q * d) >= d)


def test_div_magic_matches_integer_division():
    rng = np.random.default_rng(0)
    edge = np.array([0, 1, 2, 255, 256, (1 << 24) - 1, 1 << 24, (1 << 24) + 1, (1 << 31) - 2, (1 << 31) - 1], dtype=np.int64)
    for d in list(range(1, 1026)) + [4095, 4096, 65535, 65536, (1 << 20) + 7, (1 << 31) - 1]:
        n = np.concatenate([edge, rng.integers(0, 1 << 31, 4000, dtype=np.int64),
                            # multiples of d and their neighbours: where an under-estimated quotient shows
                            np.clip(rng.integers(0, (1 << 31) // d + 1, 2000, dtype=np.int64) * d + rng.integers(-1, 2, 2000), 0, (1 << 31) - 1)])
        assert np.array_equal(div_magic(n, d), n // d), d


def test_div_magic_needs_its_correction_step():
    # the plain multiply-high under-estimates (that is why the kernels test the remainder): e.g. n = d = 3
    n = np.array([3], dtype=np.int64)
    q = (n.astype(np.uint64) * np.uint64(conv_magic(3))) >> np.uint64(32)
    assert int(q[0]) == 0 and int(div_magic(n, 3)[0]) == 1
