"""CPU: the id assignment of a node level deduplicated on chip (csrc/bucket.cu), restated as a small Python model
and checked against the reference's sequential emplace order (src/shared_tree.cpp:662-672: a node's id is the
number of distinct nodes seen before its first occurrence).

What the model pins down is the reasoning the kernels rely on, not the kernels (those are compared with the
oracle on the GPU): (1) positions that repeat their predecessor inside a partition tile make no record and point
at the head of their run; (2) records are deduplicated in ANY grouping (buckets by hash) by taking the smallest
position per key; (3) ids are ranks of the first-occurrence bits in position order; (4) a later occurrence reads
the id of the position it points at, with one more hop when that position is a run's head that is itself a later
occurrence (COLLAPSE_WINDOW in bucket.cuh)."""
import numpy as np

WINDOW = 16  # the model's partition tile (the kernels: 2048 or 4096 positions)


def sequential_ids(keys):
    seen, ids = {}, []
    for k in keys:
        ids.append(seen.setdefault(k, len(seen)))
    return ids


def bucketed_ids(keys, buckets=7, rng=None):
    n = len(keys)
    first = [False] * n
    aux = [None] * n
    records = []
    for p in range(n):
        collapsed = p % WINDOW != 0 and keys[p] == keys[p - 1]
        if collapsed:
            head = p - 1
            while head % WINDOW != 0 and keys[head] == keys[head - 1]:
                head -= 1
            aux[p] = head  # the partition pass writes this itself
        else:
            first[p] = True  # until the dedup kernel says otherwise
            records.append((keys[p], p))
    # records reach their bucket in any order
    if rng is not None:
        rng.shuffle(records)
    by_bucket = {}
    for k, p in records:
        by_bucket.setdefault(hash(k) % buckets, []).append((k, p))
    for recs in by_bucket.values():
        fp = {}
        for k, p in recs:
            fp[k] = min(fp.get(k, p), p)
        for k, p in recs:
            if fp[k] != p:
                first[p] = False
                aux[p] = fp[k]
    rank, ids = 0, [None] * n
    for p in range(n):  # assign: rank of the first-occurrence bits in position order
        if first[p]:
            ids[p] = rank
            rank += 1
    for p in range(n):  # resolve
        if not first[p]:
            q = aux[p]
            if p - q < WINDOW and not first[q]:
                q = aux[q]
            assert first[q], "a later occurrence must end at a first occurrence after at most two hops"
            ids[p] = ids[q]
    return ids


def test_runs_and_repeats_get_the_reference_ids():
    rng = np.random.default_rng(12)
    for case in range(60):
        n = int(rng.integers(1, 400))
        alphabet = int(rng.integers(1, 12))
        keys = []
        while len(keys) < n:
            k = (int(rng.integers(0, alphabet)), int(rng.integers(0, 2)))
            keys.extend([k] * int(rng.choice([1, 1, 1, 2, 5, 40])))  # runs shorter and longer than a tile
        keys = keys[:n]
        assert bucketed_ids(keys, buckets=int(rng.integers(1, 9)), rng=rng) == sequential_ids(keys), case


def test_one_key_everywhere_and_all_distinct():
    assert bucketed_ids([(0, 0)] * 100) == [0] * 100
    keys = [(i, 0) for i in range(100)]
    assert bucketed_ids(keys) == list(range(100))
    # a run that starts as a later occurrence: its head points at the first occurrence, its members at the head
    keys = [(1, 1)] + [(2, 2)] * 3 + [(1, 1)] * 9 + [(3, 3)]
    assert bucketed_ids(keys) == sequential_ids(keys)
