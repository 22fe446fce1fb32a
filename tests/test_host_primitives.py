"""CPU: the __host__ __device__ value functions the kernels run (canonical_node, canonical_leaf, compose in
csrc/common.cuh), called through the C ABI on the host, against the reference's own vectors
(tests/golden/primitives.npz, written by oracle/gen_golden.py from the unmodified reference)."""
import numpy as np


def test_node_canonical_matches_reference_vectors(stb, golden):
    p = golden["prim"]
    for l, r, cl, cr, f in zip(p["node_l"], p["node_r"], p["node_cl"], p["node_cr"], p["node_f"]):
        assert stb.node_canonical(int(l), int(r)) == (int(cl), int(cr), int(f)), (hex(int(l)), hex(int(r)))


def test_leaf_canonical_matches_reference_vectors(stb, golden):
    p = golden["prim"]
    for S, v, c, f in zip(p["leaf_S"], p["leaf_v"], p["leaf_c"], p["leaf_f"]):
        assert stb.leaf_canonical(int(v), int(S)) == (int(c), int(f))


def test_compose_matches_reference_vectors(stb, golden):
    p = golden["prim"]
    for raw, comp in zip(p["ptr_raw"], p["ptr_comp"]):
        got = [stb.pointer_compose(int(raw), m, t) for m in (0, 1) for t in (0, 1)]
        assert got == [int(x) for x in comp]


def test_node_canonical_matches_oracle_on_dense_small_index_space(stb, oracle):
    """Every pair over a tiny index space with all flag combinations and nulls: ties between variants are common."""
    ptrs = [oracle.pointer(i, m, t, inv) for i in range(3) for m in (0, 1) for t in (0, 1) for inv in (0, 1)]
    ptrs.append(stb.NULL)
    ptrs = sorted(set(ptrs))
    for l in ptrs:
        for r in ptrs:
            assert stb.node_canonical(l, r) == oracle.node_canonical(l, r), (hex(l), hex(r))


def test_query_indices_closed_form(stb):
    idx = stb.query_indices(42, 1000, 258333333)
    assert idx.dtype == np.uint64 and idx.max() < 258333333
    # splitmix64(42 + golden) by hand
    x = (42 + 0x9E3779B97F4A7C15) & (2**64 - 1)
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & (2**64 - 1)
    x ^= x >> 31
    assert int(idx[0]) == x % 258333333
