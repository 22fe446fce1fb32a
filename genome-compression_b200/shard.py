"""Sharded multi-GPU shared_tree build: a thin binding over the one-call C ABI of
include/shared_tree_b200_dist.h (stb_shard_*, csrc/shard.cu).

The level loop, the NCCL communicator, the peer-mapped arenas and every kernel live in the
library; this file only creates the ranks and passes pointers.  One process per GPU:

    rank = shard.create_nccl(dna_size=12, device=local_rank)      # torch.distributed carries the 128-byte id
    first, count = rank.range(n_bases_total)                       # which bases this rank builds from
    rank.build_from_body(my_bases, n_bases_total)                  # collective, one call (src/shared_tree.cpp:719-736)
    tree = rank.gather()                                           # collective; a SharedTree on rank 0, None elsewhere

`create_local(world)` makes virtual ranks in one process on one GPU (tests).
"""
from __future__ import annotations

import ctypes as C
import sys
import threading

import numpy as np

_pkg = sys.modules[__name__.rsplit(".", 1)[0]]
lib = _pkg.lib
HOST, DEVICE = _pkg.HOST, _pkg.DEVICE
ID_BYTES = 128

_vp, _u64, _i32, _cp = C.c_void_p, C.c_uint64, C.c_int, C.c_char_p
for _name, _args in {
    "stb_shard_unique_id": [_vp],
    "stb_shard_create": [C.POINTER(_vp), _i32, _i32, _vp, _i32, _i32, _vp],
    "stb_shard_create_local": [C.POINTER(_vp), _i32, _i32, _i32],
    "stb_shard_destroy": [_vp],
    "stb_shard_set_option": [_vp, _cp, _u64],
    "stb_shard_range": [_vp, _u64, C.POINTER(_u64), C.POINTER(_u64)],
    "stb_shard_build_from_body": [_vp, _vp, _u64, _i32],
    "stb_shard_layer_totals": [_vp, C.POINTER(_u64), _u64, C.POINTER(_u64)],
    "stb_shard_gather": [_vp, _vp],
    "stb_shard_profile": [_vp, _i32, C.POINTER(_cp), C.POINTER(C.c_double), C.POINTER(_u64), _u64, C.POINTER(_u64)],
}.items():
    getattr(lib, _name).argtypes = _args
    getattr(lib, _name).restype = _i32
lib.stb_shard_last_error.argtypes = [_vp]
lib.stb_shard_last_error.restype = _cp


class ShardRank:
    """One rank of a sharded build (stb_shard*)."""

    def __init__(self, handle, dna_size: int, device: int, rank: int, world: int):
        self._h, self.dna_size, self.device, self.rank, self.world = C.c_void_p(handle), dna_size, device, rank, world

    def _check(self, st: int):
        if st != 0:
            raise _pkg.StbError(st, lib.stb_shard_last_error(self._h).decode() or lib.stb_status_string(st).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib.stb_shard_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int) -> "ShardRank":
        self._check(lib.stb_shard_set_option(self._h, name.encode(), int(value)))
        return self

    def range(self, n_bases_total: int):
        first, count = C.c_uint64(0), C.c_uint64(0)
        self._check(lib.stb_shard_range(self._h, n_bases_total, C.byref(first), C.byref(count)))
        return int(first.value), int(count.value)

    def build_from_body(self, body_local, n_bases_total: int) -> "ShardRank":
        ptr, n, mem, _keep = _pkg._as_input(body_local, np.uint8)
        assert n >= self.range(n_bases_total)[1], "this rank's slice of the body is shorter than stb_shard_range says"
        self._check(lib.stb_shard_build_from_body(self._h, ptr, n_bases_total, mem))
        return self

    def layer_totals(self):
        out = (C.c_uint64 * 64)()
        count = C.c_uint64(0)
        self._check(lib.stb_shard_layer_totals(self._h, out, 64, C.byref(count)))
        return [int(out[i]) for i in range(count.value)]

    def gather(self, stream: int | None = None):
        """Collective: the complete tree as a SharedTree on rank 0, None on the other ranks."""
        if self.rank == 0:
            tree = _pkg.SharedTree(self.dna_size, device=self.device, stream=stream)
            self._check(lib.stb_shard_gather(self._h, tree._h))
            return tree
        self._check(lib.stb_shard_gather(self._h, None))
        return None

    def profile(self, on: bool = True):
        self._check(lib.stb_shard_profile(self._h, int(on), None, None, None, 0, None))

    def profile_reset(self):
        self._check(lib.stb_shard_profile(self._h, -1, None, None, None, 0, None))

    def profile_read(self) -> dict:
        cap = 64
        names, ms, launches, count = (C.c_char_p * cap)(), (C.c_double * cap)(), (C.c_uint64 * cap)(), C.c_uint64(0)
        self._check(lib.stb_shard_profile(self._h, 1, names, ms, launches, cap, C.byref(count)))
        return {names[i].decode(): {"ms": ms[i], "launches": int(launches[i])} for i in range(min(cap, count.value))}


def create_nccl(dna_size: int = 12, device: int = 0, stream: int | None = None, group=None) -> ShardRank:
    """One rank per process; torch.distributed (already initialised) only carries the communicator id."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ident = (C.c_uint8 * ID_BYTES)()
    if rank == 0:
        st = lib.stb_shard_unique_id(ident)
        if st != 0:
            raise _pkg.StbError(st, "NCCL is not available to libshared_tree_b200.so")
    backend = dist.get_backend(group)
    t = torch.tensor(list(ident), dtype=torch.uint8, device=f"cuda:{device}" if backend == "nccl" else "cpu")
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident = (C.c_uint8 * ID_BYTES)(*t.cpu().tolist())
    h = C.c_void_p()
    st = lib.stb_shard_create(C.byref(h), device, dna_size, C.c_void_p(stream or 0), rank, world, ident)
    if st != 0:
        raise _pkg.StbError(st, "stb_shard_create failed (NCCL communicator / CUDA device)")
    return ShardRank(h.value, dna_size, device, rank, world)


def create_local(world: int, device: int = 0, dna_size: int = 12):
    """`world` virtual ranks of this process on one GPU; drive them with run_local."""
    hs = (C.c_void_p * world)()
    st = lib.stb_shard_create_local(hs, world, device, dna_size)
    if st != 0:
        raise _pkg.StbError(st, "stb_shard_create_local failed")
    return [ShardRank(hs[r], dna_size, device, r, world) for r in range(world)]


def run_local(ranks, fn):
    """Runs fn(rank) for every virtual rank on its own thread (the collectives block until all arrive)."""
    results, errors = [None] * len(ranks), []

    def work(i):
        try:
            results[i] = fn(ranks[i])
        except BaseException as e:  # noqa: BLE001 - reported to the caller below
            errors.append(e)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(ranks))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results
