"""CPU: the multi-GPU host logic (genome-compression_b200/dist.py) under gloo, world sizes 1-3,
with the numpy stage twin, checked against the oracle's single-process tree."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_package


def make_leaves(n, S, seed):
    rng = np.random.default_rng(seed)
    codes = np.array([1, 2, 4, 8], dtype=np.uint64)
    nib = codes[rng.integers(0, 2 if n < 64 else 4, size=(n, S))]
    leaves = (nib << (4 * np.arange(S, dtype=np.uint64))).sum(axis=1).astype(np.uint64)
    if n >= 16:  # planted repeat + a reverse-complement-ish mirror block
        leaves[n // 2:n // 2 + n // 4] = leaves[:n // 4]
    return leaves


def _worker(rank, world, port, n, S, cut, seed, result_path):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = load_package()
        from dist_numpy_stages import NumpyStages
        from oracle.pyoracle import Oracle
        from genome_compression_b200.dist import DistBuilder, ShardPlan

        oracle = Oracle()
        leaves = make_leaves(n, S, seed)
        plan = ShardPlan(n, world, cut)
        lo, hi = plan.level_range(rank, 0)
        local = torch.from_numpy(leaves[lo:hi].view(np.int64).copy())
        builder = DistBuilder(NumpyStages(oracle, S), cut=cut)
        if S <= 5 or n < 100:  # text entry point: direct-addressed leaf table + all-reduce(MIN)
            text = b"".join(pkg.leaf_to_str(v, S).encode() for v in leaves[lo:hi])
            tree = builder.build_from_body(torch.frombuffer(bytearray(text or b"A"), dtype=torch.uint8), n * S)
        else:
            tree = builder.build_from_leaves(local, n)
        full = builder.gather(tree)
        if rank == 0:
            want = oracle.build(leaves, S)
            assert len(full.layers_) == want.depth() - 1, (len(full.layers_), want.depth() - 1)
            assert np.array_equal(full.leaves_, want.leaves())
            for k, layer in enumerate(full.layers_):
                assert np.array_equal(layer, want.layer(k)), k
            assert full.root_ == want.root()
            assert tree.layer_totals[0] == want.leaf_count()
            for k, total in enumerate(tree.layer_totals[1:]):
                assert total == want.layer_count(k)
            with open(result_path, "w") as f:
                f.write(f"ok {builder.collectives}")
    finally:
        dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n,S,cut", [(2, 1000, 12, 16), (2, 37, 12, 1), (3, 515, 5, 8), (2, 1, 12, 1), (2, 2, 12, 1), (1, 300, 12, 4)])
def test_sharded_build_equals_oracle(world, n, S, cut):
    with tempfile.TemporaryDirectory() as d:
        result = os.path.join(d, "result")
        mp.spawn(_worker, args=(world, free_port(), n, S, cut, 1234 + n, result), nprocs=world, join=True)
        assert open(result).read().startswith("ok")


def test_shard_plan():
    load_package()
    from genome_compression_b200.dist import ShardPlan
    plan = ShardPlan(258_333_333, 8)
    assert plan.shard == 1 << 25
    assert plan.level_range(7, 0) == (7 << 25, 258_333_333)
    for level in range(plan.sharded_levels()):
        covered = 0
        for r in range(8):
            lo, hi = plan.level_range(r, level)
            assert lo == covered or hi == lo
            covered = max(covered, hi)
            if level:
                plo, phi = plan.level_range(r, level - 1)
                assert hi - lo == -(-(phi - plo) // 2)
        assert covered == plan.level_total(level)
    assert plan.level_total(plan.sharded_levels() - 1) > plan.cut >= plan.level_total(plan.sharded_levels())
    tiny = ShardPlan(1, 4, cut=1)
    assert tiny.sharded_levels() == 1 and tiny.level_range(0, 0) == (0, 1) and tiny.level_range(3, 0) == (1, 1)


def test_default_cut_follows_world(monkeypatch):
    load_package()
    from genome_compression_b200.dist import ShardPlan, default_cut
    monkeypatch.delenv("STB_DIST_CUT_LOG2", raising=False)
    assert [default_cut(w) for w in (1, 2, 4, 8, 16, 64)] == [1 << 25, 1 << 24, 1 << 23, 1 << 22, 1 << 21, 1 << 20]
    # config 4 on 8 ranks: leaf level + 5 node levels sharded, the 4.0 M-position level goes to rank 0
    plan = ShardPlan(258_333_333, 8)
    assert plan.cut == 1 << 22 and plan.sharded_levels() == 6
    monkeypatch.setenv("STB_DIST_CUT_LOG2", "18")
    assert default_cut(8) == 1 << 18 and ShardPlan(258_333_333, 8).sharded_levels() == 10


def test_thread_comm_virtual_ranks(oracle):
    """The in-process communicator used on single-GPU boxes, here with the numpy stages."""
    import threading
    load_package()
    sys.path.insert(0, str(ROOT / "tests"))
    from dist_numpy_stages import NumpyStages
    from genome_compression_b200.dist import DistBuilder, ShardPlan, ThreadComm

    n, S, world, cut = 700, 12, 4, 8
    leaves = make_leaves(n, S, 77)
    shared = ThreadComm.Shared(world)
    plan = ShardPlan(n, world, cut)
    out, errors = [None] * world, []

    def work(rank):
        try:
            lo, hi = plan.level_range(rank, 0)
            builder = DistBuilder(NumpyStages(oracle, S), comm=ThreadComm(shared, rank), cut=cut)
            out[rank] = builder.gather(builder.build_from_leaves(torch.from_numpy(leaves[lo:hi].view(np.int64).copy()), n))
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            shared.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    want = oracle.build(leaves, S)
    assert np.array_equal(out[0].leaves_, want.leaves())
    assert all(np.array_equal(a, want.layer(k)) for k, a in enumerate(out[0].layers_))
    assert out[0].root_ == want.root()


@pytest.mark.parametrize("world,n,S,cut,text", [(4, 700, 12, 8, False), (2, 333, 16, 1, False), (3, 515, 5, 4, True), (2, 1, 12, 1, True),
                                                (3, 2, 12, 1, False)])
def test_peer_exchange_orchestration_virtual_ranks(oracle, world, n, S, cut, text):
    """The peer exchange's host logic (arenas, barriers, deferred counts, pointer hand-over to
    rank 0, buffer reuse across builds) with numpy arenas shared by the virtual ranks."""
    import threading
    pkg = load_package()
    sys.path.insert(0, str(ROOT / "tests"))
    from dist_numpy_stages import NumpyPeerStages
    from genome_compression_b200.dist import DistBuilder, ShardPlan, ThreadComm

    leaves = make_leaves(n, S, 99 + n)
    shared = ThreadComm.Shared(world)
    plan = ShardPlan(n, world, cut)
    out, errors, collectives = [None] * world, [], [0] * world

    def work(rank):
        try:
            lo, hi = plan.level_range(rank, 0)
            builder = DistBuilder(NumpyPeerStages(oracle, S), comm=ThreadComm(shared, rank), cut=cut)
            assert builder.exchange == "peer"
            for _ in range(2):  # the second build reuses arenas, table (new epoch) and cursors
                if text:
                    body = b"".join(pkg.leaf_to_str(v, S).encode() for v in leaves[lo:hi])
                    tree = builder.build_from_body(torch.frombuffer(bytearray(body or b"A"), dtype=torch.uint8), n * S)
                else:
                    tree = builder.build_from_leaves(torch.from_numpy(leaves[lo:hi].view(np.int64).copy()), n)
            out[rank] = (builder.gather(tree), tree.layer_totals)
            collectives[rank] = builder.collectives
            stale = tree
            tree = builder.build_from_leaves(torch.from_numpy(leaves[lo:hi].view(np.int64).copy()), n)
            with pytest.raises(ValueError):
                builder.gather(stale)  # rank 0's top layers were rebuilt by the later build
            builder.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            shared.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    want = oracle.build(leaves, S)
    full, totals = out[0]
    assert np.array_equal(full.leaves_, want.leaves())
    assert len(full.layers_) == want.depth() - 1
    assert all(np.array_equal(a, want.layer(k)) for k, a in enumerate(full.layers_))
    assert full.root_ == want.root()
    assert totals[0] == want.leaf_count() and all(t == want.layer_count(k) for k, t in enumerate(totals[1:]))
    assert len(set(collectives)) == 1  # every rank issued the same collectives


def test_peer_exchange_falls_back_together(oracle):
    """One rank cannot get peer-mapped memory: every rank switches to the collective exchange
    before any stage ran, and the tree is the same."""
    import threading
    load_package()
    sys.path.insert(0, str(ROOT / "tests"))
    from dist_numpy_stages import NumpyPeerStages
    from genome_compression_b200.dist import DistBuilder, ShardPlan, ThreadComm

    class Broken(NumpyPeerStages):
        def peer_alloc(self, shape):
            raise MemoryError("no IPC here")

    n, S, world, cut = 400, 12, 3, 4
    leaves = make_leaves(n, S, 5)
    shared = ThreadComm.Shared(world)
    plan = ShardPlan(n, world, cut)
    out, errors, modes = [None] * world, [], [None] * world

    def work(rank):
        try:
            lo, hi = plan.level_range(rank, 0)
            stages = (Broken if rank == 1 else NumpyPeerStages)(oracle, S)
            builder = DistBuilder(stages, comm=ThreadComm(shared, rank), cut=cut)
            out[rank] = builder.gather(builder.build_from_leaves(torch.from_numpy(leaves[lo:hi].view(np.int64).copy()), n))
            modes[rank] = builder.exchange
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            shared.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    assert modes == ["collective"] * world
    want = oracle.build(leaves, S)
    assert np.array_equal(out[0].leaves_, want.leaves())
    assert all(np.array_equal(a, want.layer(k)) for k, a in enumerate(out[0].layers_))
    assert out[0].root_ == want.root()
