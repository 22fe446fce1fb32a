"""CPU: the C restatement against the live, unmodified reference (oracle/_ref/libref.so).
Skipped where the reference could not be built (no /root/reference and no shipped _ref)."""
import numpy as np


def test_leaf_fuzz(oracle, ref):
    rng = np.random.default_rng(7)
    for S in range(1, 17):
        mask = (1 << (4 * S)) - 1
        for _ in range(300):
            v = int.from_bytes(rng.bytes(8), "little") & mask
            assert oracle.leaf_canonical(v, S) == ref.leaf_canonical(v, S)
            assert oracle.mirrored(v, S) == ref.mirrored(v, S)
            assert oracle.inverted(v, S) == ref.inverted(v, S)
            assert oracle.transposed(v, S) == ref.transposed(v, S)


def test_node_fuzz(oracle, ref):
    rng = np.random.default_rng(8)
    null = 0x9FFFFFFF

    def rp():
        if rng.integers(8) == 0:
            return null
        return oracle.pointer(int(rng.integers(5)), int(rng.integers(2)), int(rng.integers(2)), int(rng.integers(4) == 0))

    for _ in range(20000):
        l, r = rp(), rp()
        assert oracle.node_canonical(l, r) == ref.node_canonical(l, r)


def test_random_trees(oracle, ref):
    rng = np.random.default_rng(9)
    for S, n, alphabet in ((12, 5000, 4), (12, 777, 2), (3, 4097, 4), (16, 1000, 16), (7, 2500, 16)):
        codes = np.array([1, 2, 4, 8, 3, 12, 7, 14, 0, 9, 5, 11, 13, 10, 6, 15], dtype=np.uint64)[:alphabet]
        nib = codes[rng.integers(0, alphabet, size=(n, S))]
        leaves = (nib << (4 * np.arange(S, dtype=np.uint64))).sum(axis=1).astype(np.uint64)
        leaves[n // 2:n // 2 + n // 8] = leaves[:n // 8]  # planted repeat
        to, tr = oracle.build(leaves, S), ref.build(leaves, S)
        _, lo = oracle.build_levels(leaves, S)
        _, lr = ref.build_levels(leaves, S)
        assert len(lo) == len(lr) and all(np.array_equal(a, b) for a, b in zip(lo, lr))
        assert to.serialize() == tr.serialize()
        for k in range(to.depth() - 1):
            assert np.array_equal(to.layer(k), tr.layer(k))
        to.sort(); tr.sort()
        assert to.serialize() == tr.serialize()
        assert np.array_equal(to.decode(), leaves) and np.array_equal(tr.decode(), leaves)
        idx = rng.integers(0, n, 500).astype(np.uint64)
        assert np.array_equal(to.random_access(idx), tr.random_access(idx))
