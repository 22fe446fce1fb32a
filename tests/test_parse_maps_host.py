"""CPU: the algorithm behind the device-side deserialize (csrc/serialize.cu), restated in numpy / Python and checked
against the sequential parse of real streams made by the oracle.

A pointer's length is known only from its first byte (src/shared_tree.cpp:25-67), so the start of every pointer
of a layer looks sequential.  The CUDA parser cuts a layer's bytes into 32-byte chunks; seen from a chunk the only
unknown is the offset (0..3) at which its first pointer starts, and every entry offset maps to an exit offset into
the next chunk and a number of pointers started inside.  These maps compose associatively, so a scan places and
numbers every pointer.  This file checks exactly that claim (the kernels themselves are checked on the GPU by
tests/test_gpu_parity.py::test_deserialize_on_device)."""
import numpy as np

CHUNK = 32


def chunk_map(data: bytes, begin: int, first_chunk_start=None):
    """(exit offsets, pointer counts) for the four entry offsets of the chunk data[begin : begin + CHUNK]."""
    exits, counts = [], []
    for e in range(4):
        pos = first_chunk_start if first_chunk_start is not None else e
        c = 0
        while pos < CHUNK:
            b = data[begin + pos] if begin + pos < len(data) else 0
            pos += 1 + (b >> 6)
            c += 1
        exits.append(pos - CHUNK)
        counts.append(c)
    return exits, counts


def compose(a, b):
    """`a` first, then `b` (ParseMap compose in csrc/serialize.cu)."""
    exits = [b[0][a[0][e]] for e in range(4)]
    counts = [a[1][e] + b[1][a[0][e]] for e in range(4)]
    return exits, counts


IDENTITY = ([0, 1, 2, 3], [0, 0, 0, 0])


def sequential_starts(data: bytes, start: int, n_ptrs: int):
    out, pos = [], start
    for _ in range(n_ptrs):
        out.append(pos)
        pos += 1 + (data[pos] >> 6)
    return out, pos


def parallel_starts(data: bytes, start: int, n_ptrs: int):
    """What the three kernels compute: chunk maps, their exclusive scan evaluated at entry 0, then every chunk walked
    from its own entry with its own first pointer number."""
    begin = start & ~(CHUNK - 1)
    start_off = start - begin
    extent = min(len(data) - begin, start_off + 4 * n_ptrs)
    n_chunks = -(-extent // CHUNK)
    maps = [chunk_map(data, begin + c * CHUNK, start_off if c == 0 else None) for c in range(n_chunks)]
    # exclusive scan; the order of composition is arbitrary (associativity): do it as a balanced tree to prove the point
    def reduce(lo, hi):
        if hi - lo == 0:
            return IDENTITY
        if hi - lo == 1:
            return maps[lo]
        mid = (lo + hi) // 2
        return compose(reduce(lo, mid), reduce(mid, hi))
    starts, end = [None] * n_ptrs, None
    for c in range(n_chunks):
        before = reduce(0, c)
        entry, number = before[0][0], before[1][0]
        pos = start_off if c == 0 else entry
        while pos < CHUNK and number < n_ptrs:
            at = begin + c * CHUNK + pos
            starts[number] = at
            pos += 1 + (data[at] >> 6)
            if number == n_ptrs - 1:
                end = begin + c * CHUNK + pos
            number += 1
    return starts, end


def layer_offsets(stream: bytes, leaf_bytes: int):
    """(start, node count) of every node layer of a .dag stream, by the sequential rules of src/shared_tree.cpp:520-538."""
    o = 1 + (stream[0] >> 6)
    n_leaves = int.from_bytes(stream[o:o + 8], "big")
    o += 8 + n_leaves * leaf_bytes
    layers = []
    while o + 8 <= len(stream):
        count = int.from_bytes(stream[o:o + 8], "big")
        o += 8
        layers.append((o, count))
        _, o = sequential_starts(stream, o, 2 * count)
    return layers


def test_chunk_maps_place_every_pointer(oracle):
    rng = np.random.default_rng(4)
    for S, n in ((12, 3000), (5, 2500), (16, 900), (3, 4001)):
        nib = np.array([1, 2, 4, 8], dtype=np.uint64)[rng.integers(0, 4, size=(n, S))]
        leaves = (nib << (4 * np.arange(S, dtype=np.uint64))).sum(axis=1).astype(np.uint64)
        leaves[n // 2:n // 2 + n // 5] = leaves[:n // 5]
        tree = oracle.build(leaves, S)
        tree.sort()
        stream = tree.serialize()
        layers = layer_offsets(stream, (S + 1) // 2)
        assert [c for _, c in layers] == tree.layer_counts()
        for start, count in layers[:6]:  # every alignment of the layer's first byte occurs over the cases
            want, want_end = sequential_starts(stream, start, 2 * count)
            got, got_end = parallel_starts(stream, start, 2 * count)
            assert got == want and got_end == want_end, (S, start, count)


def test_maps_compose_associatively():
    rng = np.random.default_rng(9)
    data = bytes(rng.integers(0, 256, 32 * 40, dtype=np.uint8))
    maps = [chunk_map(data, c * CHUNK) for c in range(40)]
    for _ in range(200):
        i, j, k = sorted(rng.integers(0, 41, 3))
        def fold(lo, hi):
            out = IDENTITY
            for m in maps[lo:hi]:
                out = compose(out, m)
            return out
        assert compose(fold(i, j), fold(j, k)) == fold(i, k)
