"""Sharded multi-GPU shared_tree build (BASELINE.json config 4).

One process per GPU; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the
plumbing, the stage kernels of include/shared_tree_b200_dist.h are the work.  Protocol
per level (leaves, then node layers bottom-up):

    every (key, global position) record goes to the key's hash owner -> the owner dedups with
    min global position, tells the source where the key came first, marks first occurrences in
    a bitmap -> all_reduce(bitmap) -> rank index -> ids = rank of the first occurrence's bit

The records travel either through peer-mapped memory — the stage kernels store them straight
into the owner's arena over NVLink and the owner stores the answers back; two small
collectives per level, nothing read by the host (exchange="peer", the default on one node) —
or through all_to_all collectives (exchange="collective": other transports, the CPU tests,
and the joint fallback when peer mapping is unavailable).

Node ids are first-occurrence ranks in GLOBAL position order: identical to the single-GPU
build and to the reference (src/shared_tree.cpp:630-637, :662-672).  Rank g owns a contiguous
power-of-two aligned range of positions at every sharded level (the analogue of the
reference's 2^22-leaf segments, include/shared_tree.h:305-316); when a level is small the
pointer arrays are gathered and rank 0 finishes alone.

`stages` abstracts the device work: `CudaStages` (this file) calls the C ABI; the CPU tests
plug in a numpy twin (tests/dist_numpy_stages.py) to exercise this host logic under gloo.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
from dataclasses import dataclass, field

import torch
import torch.distributed as dist

LEAF, NODE = 0, 1


def ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def default_cut(world: int) -> int:
    """A level with at most this many positions is finished on rank 0: below it the per-level
    collectives cost more than one GPU needs for the whole level.  Measured at 3.1 Gbp
    (profiles/README.md): 2^24 is best on 2 GPUs, 2^22 on 8."""
    env = os.environ.get("STB_DIST_CUT_LOG2")
    return 1 << int(env) if env else max(1 << 20, (1 << 25) // max(1, world))


@dataclass
class ShardPlan:
    """Which positions of which level a rank owns."""
    n_leaves: int
    world: int
    cut: int | None = None  # see default_cut
    shard: int = field(init=False)

    def __post_init__(self):
        if self.cut is None:
            self.cut = default_cut(self.world)
        per = max(1, ceil_div(self.n_leaves, self.world))
        self.shard = 1 << (per - 1).bit_length()  # power of two >= per

    def level_total(self, level: int) -> int:
        """Positions at `level` (0 = leaves, k = node layer k-1)."""
        return ceil_div(self.n_leaves, 1 << level)

    def level_range(self, rank: int, level: int):
        per = self.shard >> level
        total = self.level_total(level)
        return min(total, rank * per), min(total, (rank + 1) * per)

    def sharded_levels(self) -> int:
        """Number of levels (leaf level included) that run distributed."""
        levels = 1
        while (self.shard >> levels) >= 1 and self.level_total(levels) > self.cut and self.level_total(levels) > 1:
            levels += 1
        return levels


@dataclass
class LayerSlice:
    base: int          # id of the first item
    count: int
    items: torch.Tensor  # int64[count] leaves or int32[count, 2] nodes


@dataclass
class DistTree:
    """Result of a sharded build: every rank holds its id-range of each sharded layer;
    rank 0 additionally holds the top layers."""
    dna_size: int
    width: int
    leaves: LayerSlice
    layers: list          # LayerSlice per sharded node layer
    layer_totals: list    # global unique count per sharded layer (leaf level first)
    upper: object         # rank 0: stages' upper-tree object, else None
    root: int | None
    build_id: int = 0     # the builder's build counter: rank 0's top layers live in a handle the next build reuses


class CudaStages:
    """Device work through the C ABI (include/shared_tree_b200_dist.h)."""

    def __init__(self, pkg, dna_size: int, device: int, stream: int | None = None):
        self.pkg = pkg
        self.ctx = pkg.SharedTree(dna_size, device=device, stream=stream)
        self.dna_size = dna_size
        self.device = torch.device("cuda", device)
        self.device_index = device
        self.stream = stream

    def _check(self, st):
        self.ctx._check(st)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)

    def pack_body(self, body, n_leaves, leaves_out):
        self._check(self.pkg.lib.stb_dist_pack_body(self.ctx._h, self._p(body), n_leaves, self._p(leaves_out)))

    def partition(self, kind, items, n_items, gpos0, world, keys, gpos, meta, counts):
        self._check(self.pkg.lib.stb_dist_partition(self.ctx._h, kind, self._p(items), n_items, gpos0, world, self._p(keys),
                                                    self._p(gpos), self._p(meta), self._p(counts)))

    def owner(self, keys, gpos, n, table, cap, answers, bitmap):
        self._check(self.pkg.lib.stb_dist_owner(self.ctx._h, self._p(keys), self._p(gpos), n, self._p(table), cap,
                                                self._p(answers), self._p(bitmap)))

    def rank_index(self, bitmap, n_words, word_prefix, scratch):
        self._check(self.pkg.lib.stb_dist_rank_index(self.ctx._h, self._p(bitmap), n_words, self._p(word_prefix), self._p(scratch)))

    def finish(self, kind, items, n_items, gpos0, bitmap, word_prefix, n_level, meta, answers, pointers, slice_out, base_count):
        self._check(self.pkg.lib.stb_dist_finish(self.ctx._h, kind, self._p(items), n_items, gpos0, self._p(bitmap),
                                                 self._p(word_prefix), n_level, self._p(meta), self._p(answers), self._p(pointers),
                                                 self._p(slice_out), self._p(base_count)))

    def leaf_direct_minpos(self, body, n_local, gpos0, table, tmp) -> int:
        flag = C.c_int(0)
        self._check(self.pkg.lib.stb_dist_leaf_direct_minpos(self.ctx._h, self._p(body), n_local, gpos0, self._p(table), self._p(tmp), C.byref(flag)))
        return int(flag.value)

    def leaf_direct_finish(self, table, n_level, tmp, n_local, bitmap, word_prefix, scratch, ids, pointers, leaves_out):
        self._check(self.pkg.lib.stb_dist_leaf_direct_finish(self.ctx._h, self._p(table), n_level, self._p(tmp), n_local, self._p(bitmap),
                                                             self._p(word_prefix), self._p(scratch), C.c_void_p(ids.data_ptr()),
                                                             self._p(pointers), C.c_void_p(leaves_out.data_ptr())))

    def upper_levels(self, pointers, n, leaf_pointers):
        """pointers: a device tensor or a raw device address."""
        # one handle for every build: its workspace (tables, pointer arrays) is reused
        if getattr(self, "_upper", None) is None:
            self._upper = self.pkg.SharedTree(self.dna_size, device=self.device_index, stream=self.stream)
        tree = self._upper
        ptr = C.c_void_p(pointers) if isinstance(pointers, int) else self._p(pointers)
        tree._check(self.pkg.lib.stb_dist_upper_levels(tree._h, ptr, n, int(leaf_pointers)))
        return tree

    # -- peer exchange (records written straight into the owners' memory) ------------------
    supports_peer = True

    def peer_arena_bytes(self, world, region_cap):
        return int(self.pkg.lib.stb_dist_peer_arena_bytes(world, region_cap))

    def peer_alloc(self, nbytes):
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        self._check(self.pkg.lib.stb_dist_peer_alloc(self.ctx._h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def peer_open(self, handle):
        ptr, buf = C.c_void_p(), (C.c_ubyte * 64)(*handle)
        self._check(self.pkg.lib.stb_dist_peer_open(self.ctx._h, buf, C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr):
        self._check(self.pkg.lib.stb_dist_peer_close(self.ctx._h, C.c_void_p(ptr)))

    def peer_free(self, ptr):
        self._check(self.pkg.lib.stb_dist_peer_free(self.ctx._h, C.c_void_p(ptr)))

    @staticmethod
    def _ptr_array(ptrs):
        return (C.c_void_p * len(ptrs))(*ptrs)

    def peer_scatter(self, kind, items, n_items, gpos0, world, rank, arenas, region_cap, meta):
        self._check(self.pkg.lib.stb_dist_peer_scatter(self.ctx._h, kind, self._p(items), n_items, gpos0, world, rank,
                                                       self._ptr_array(arenas), region_cap, self._p(meta)))

    def peer_owner(self, world, rank, arenas, region_cap, expected, table, table_slots, serial, slot_scratch, planes, bitmap):
        self._check(self.pkg.lib.stb_dist_peer_owner(self.ctx._h, world, rank, self._ptr_array(arenas), region_cap, expected,
                                                     self._p(table), table_slots, serial, self._p(slot_scratch), self._p(planes),
                                                     planes.numel() if planes is not None else 0, self._p(bitmap)))

    def peer_finish(self, kind, items, n_items, gpos0, bitmap, word_prefix, n_level, meta, arena, world, region_cap, pointers,
                    slice_out, base_count):
        self._check(self.pkg.lib.stb_dist_peer_finish(self.ctx._h, kind, self._p(items), n_items, gpos0, self._p(bitmap),
                                                      self._p(word_prefix), n_level, self._p(meta), C.c_void_p(arena), world, region_cap,
                                                      self._p(pointers), self._p(slice_out), self._p(base_count)))

    def peer_payload(self, arena, world, region_cap):
        return int(self.pkg.lib.stb_dist_peer_payload(C.c_void_p(arena), world, region_cap))

    def peer_put(self, dst_ptr, src, nbytes):
        self._check(self.pkg.lib.stb_dist_peer_put(self.ctx._h, C.c_void_p(dst_ptr), self._p(src), nbytes))

    def upper_layers(self, upper):
        """[(count, int32[count,2] device tensor)] of the top layers + root."""
        out = []
        for k in range(upper.depth() - 1):
            cnt = upper.layer_count(k)
            buf = torch.empty((cnt, 2), dtype=torch.int32, device=self.device)
            upper._check(self.pkg.lib.stb_copy_layer(upper._h, k, C.c_void_p(buf.data_ptr()), cnt, self.pkg.DEVICE))
            out.append(buf)
        return out, upper.root()

    def assemble(self, leaves, layers, root, width):
        tree = self.pkg.SharedTree(self.dna_size, device=self.device_index, stream=self.stream)
        counts = (C.c_uint64 * len(layers))(*[int(l.shape[0]) for l in layers])
        ptrs = (C.c_void_p * len(layers))(*[l.data_ptr() if l.numel() else 0 for l in layers])
        tree._check(self.pkg.lib.stb_assemble(tree._h, self._p(leaves), leaves.numel(), len(layers), counts, ptrs, root, width))
        return tree

    def table(self, cap):
        return torch.empty((cap + 1) * 2, dtype=torch.int64, device=self.device)

    def sync(self):
        torch.cuda.current_stream(self.device).synchronize()


class TorchComm:
    """Collectives over a torch.distributed process group (NCCL on GPUs, gloo on CPUs)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def all_to_all(self, send, send_counts, recv_counts):
        recv = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        if self.world == 1:
            recv.copy_(send)
        else:
            dist.all_to_all_single(recv, send, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts), group=self.group)
        return recv

    def count_matrix(self, counts_dev):
        if self.world == 1:
            return [counts_dev.tolist()]
        gathered = torch.empty(self.world * self.world, dtype=counts_dev.dtype, device=counts_dev.device)
        dist.all_gather_into_tensor(gathered, counts_dev, group=self.group)
        return gathered.view(self.world, self.world).tolist()

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_min(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)

    def all_gather_object(self, x):
        if self.world == 1:
            return [x]
        out = [None] * self.world
        dist.all_gather_object(out, x, group=self.group)
        return out

    # -- peer exchange plumbing ----------------------------------------------------------------
    def map_arenas(self, stages, own_ptr, handle):
        """Swap CUDA IPC handles and map every other rank's arena; returns the base pointers in
        rank order (own pointer at [rank])."""
        handles = self.all_gather_object(handle)
        if any(h is None for h in handles):
            return None  # some rank could not allocate: everybody falls back together
        ptrs, opened = [], []
        try:
            for r in range(self.world):
                ptrs.append(own_ptr if r == self.rank else stages.peer_open(handles[r]))
                if r != self.rank:
                    opened.append(ptrs[-1])
        except Exception:  # noqa: BLE001 - no peer access / IPC not permitted in this container
            ptrs = None
        if not all(self.all_gather_object(ptrs is not None)):
            for p in opened:
                stages.peer_close(p)
            return None
        return ptrs

    def unmap_arenas(self, stages, ptrs):
        for r, p in enumerate(ptrs):
            if r != self.rank:
                stages.peer_close(p)

    def stream_barrier(self, token):
        """Every rank's earlier work on its stream (peer stores included) is complete and visible
        once this returns control to the stream: a one-element all-reduce."""
        if self.world > 1:
            dist.all_reduce(token, op=dist.ReduceOp.SUM, group=self.group)


class ThreadComm:
    """Virtual ranks as threads of ONE process sharing one GPU: the same exchanges done by
    slicing each other's tensors.  Lets a single-GPU box run the multi-rank stage kernels
    (tests/test_gpu_dist.py); never used for measurements."""

    class Shared:
        def __init__(self, world):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.box = [None] * world

    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world

    def _exchange(self, item):
        self.sh.box[self.rank] = item
        self.sh.barrier.wait()
        items = list(self.sh.box)
        self.sh.barrier.wait()
        return items

    def all_to_all(self, send, send_counts, recv_counts):
        items = self._exchange((send, list(send_counts)))
        parts = []
        for src, (t, counts) in enumerate(items):
            off = sum(counts[:self.rank])
            assert counts[self.rank] == recv_counts[src]
            parts.append(t[off:off + counts[self.rank]])
        out = torch.cat(parts) if parts else send[:0]
        self.sh.barrier.wait()  # nobody frees a send buffer that is still being read
        return out

    def count_matrix(self, counts_dev):
        return [c.tolist() for c in self._exchange(counts_dev)]

    def all_reduce_sum(self, t):
        items = self._exchange(t)
        total = torch.stack(items).sum(dim=0, dtype=t.dtype)
        self.sh.barrier.wait()
        t.copy_(total)
        self.sh.barrier.wait()

    def all_reduce_min(self, t):
        items = self._exchange(t)
        total = torch.stack(items).min(dim=0).values
        self.sh.barrier.wait()
        t.copy_(total)
        self.sh.barrier.wait()

    def all_gather_object(self, x):
        return self._exchange(x)

    # virtual ranks live in one address space and on one stream: pointers are shared as they
    # are, and a host barrier between enqueues orders the kernels
    def map_arenas(self, stages, own_ptr, handle):
        ptrs = self._exchange(own_ptr if handle is not None else None)
        return None if any(p is None for p in ptrs) else ptrs

    def unmap_arenas(self, stages, ptrs):
        pass

    def stream_barrier(self, token):
        self.sh.barrier.wait()


class PeerUnavailable(RuntimeError):
    """Peer-mapped exchange memory could not be set up on every rank."""


class PeerBuffers:
    """Everything the peer exchange keeps between levels and builds on one rank: the arena that
    the other ranks write into, the owner's table and scratch, the level bitmap and rank index."""

    def __init__(self, stages, comm, region_cap: int, n_level: int):
        st, world, dev = stages, comm.world, stages.device
        self.st, self.comm, self.region_cap, self.n_level = st, comm, region_cap, n_level
        try:
            self.own, handle = st.peer_alloc(st.peer_arena_bytes(world, region_cap))
        except Exception:  # noqa: BLE001
            self.own, handle = 0, None
        self.arenas = comm.map_arenas(st, self.own, handle)
        if self.arenas is None:
            if self.own:
                st.peer_free(self.own)
                self.own = 0
            raise PeerUnavailable("no peer-mapped exchange memory on this node; using the collective exchange")
        self.table_slots = max(1024, 2 * world * region_cap) + 1
        self.table = torch.full((self.table_slots * 2,), -1, dtype=torch.int64, device=dev)  # epoch-tagged: cleared once
        self.serial = 0
        self.slot_scratch = torch.empty(max(1, world * region_cap), dtype=torch.int32, device=dev)
        self.planes = torch.empty(2 * (1 << 28) // 32, dtype=torch.int32, device=dev) if world * region_cap >= (1 << 16) else None
        self.meta = torch.empty(max(1, world * region_cap), dtype=torch.int32, device=dev)
        n_words = ceil_div(n_level, 32)
        self.bitmap = torch.empty(n_words, dtype=torch.int32, device=dev)
        self.word_prefix = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
        self.scratch = torch.empty(ceil_div(n_words, 1024) + 1, dtype=torch.int32, device=dev)
        self.token = torch.zeros(1, dtype=torch.int32, device=dev)

    def fits(self, region_cap: int, n_level: int) -> bool:
        return region_cap <= self.region_cap and n_level <= self.n_level

    def close(self):
        """Local: once this rank's stream has drained, every peer store into its arena has landed
        (each one precedes a collective this rank took part in)."""
        if getattr(self, "own", 0):
            self.st.sync()
            self.comm.unmap_arenas(self.st, self.arenas)
            self.st.peer_free(self.own)
            self.own = 0

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class DistBuilder:
    def __init__(self, stages, comm=None, cut: int | None = None, exchange: str | None = None):
        """exchange: "peer" (records written into the owners' memory by the stage kernels; needs
        peer-mapped device memory: GPUs of one node) or "collective" (all-to-all by the
        communicator).  Default: STB_DIST_EXCHANGE, else "peer" where the stages offer it."""
        self.st = stages
        self.comm = comm or TorchComm()
        self.rank, self.world = self.comm.rank, self.comm.world
        self.device = stages.device
        self.cut = cut if cut is not None else default_cut(self.world)
        if exchange is None:
            exchange = os.environ.get("STB_DIST_EXCHANGE") or ("peer" if getattr(stages, "supports_peer", False) else "collective")
        assert exchange in ("peer", "collective"), exchange
        self.exchange = exchange
        self.peer = None
        self.builds = 0
        self.collectives = 0
        self.trace = bool(int(os.environ.get("STB_DIST_TRACE", "0")))
        self._t0 = time.perf_counter()

    def close(self):
        """Gives the peer-mapped exchange memory back (also happens when the builder is collected)."""
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    # -- collectives ------------------------------------------------------------------
    def _all_to_all(self, send, send_counts, recv_counts):
        self.collectives += 1
        return self.comm.all_to_all(send, send_counts, recv_counts)

    def _count_matrix(self, counts_dev):
        self.collectives += 1
        return self.comm.count_matrix(counts_dev)

    def _all_reduce_sum(self, t):
        self.collectives += 1
        self.comm.all_reduce_sum(t)

    def _gather_rows(self, t, counts, dst=0):
        """Variable-length gather in rank order to `dst` (others get an empty tensor)."""
        send = [t.shape[0] if r == dst else 0 for r in range(self.world)]
        recv = list(counts) if self.rank == dst else [0] * self.world
        return self._all_to_all(t, send, recv)

    def _trace(self, what):
        """STB_DIST_TRACE=1: synchronising wall-clock trace of every phase (rank 0, debugging only)."""
        if not self.trace:
            return
        self.st.sync()
        now = time.perf_counter()
        if self.rank == 0:
            print(f"[dist] {what:28s} {(now - self._t0) * 1e3:8.3f} ms", flush=True)
        self._t0 = time.perf_counter()

    # -- one level --------------------------------------------------------------------
    def _level(self, kind, items, n_items, gpos0, n_level):
        st, world, dev = self.st, self.world, self.device
        self._trace(f"-- level n={n_level}")
        n_pos = n_items if kind == LEAF else ceil_div(n_items, 2)
        keys = torch.empty(n_pos, dtype=torch.int64, device=dev)
        gpos = torch.empty(n_pos, dtype=torch.int32, device=dev)
        meta = torch.empty(n_pos, dtype=torch.int32, device=dev)
        counts = torch.zeros(world, dtype=torch.int32, device=dev)
        st.partition(kind, items, n_items, gpos0, world, keys, gpos, meta, counts)
        self._trace("partition")
        matrix = self._count_matrix(counts)
        send_counts = matrix[self.rank]
        recv_counts = [matrix[src][self.rank] for src in range(world)]
        self._trace("count matrix")
        rkeys = self._all_to_all(keys, send_counts, recv_counts)
        rgpos = self._all_to_all(gpos, send_counts, recv_counts)
        self._trace("all_to_all records")
        m = rkeys.shape[0]
        n_words = ceil_div(n_level, 32)
        bitmap = torch.zeros(n_words, dtype=torch.int32, device=dev)
        answers = torch.empty(m, dtype=torch.int32, device=dev)
        cap = max(1024, 2 * m)
        table = st.table(cap)
        st.owner(rkeys, rgpos, m, table, cap, answers, bitmap)
        self._trace("owner")
        del table, rkeys, rgpos
        back = self._all_to_all(answers, recv_counts, send_counts)
        self._trace("all_to_all answers")
        self._all_reduce_sum(bitmap)  # first-occurrence bits are disjoint across owners: sum == or
        self._trace("all_reduce bitmap")
        word_prefix = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
        scratch = torch.empty(ceil_div(n_words, 1024) + 1, dtype=torch.int32, device=dev)
        st.rank_index(bitmap, n_words, word_prefix, scratch)
        pointers = torch.empty(n_pos, dtype=torch.int32, device=dev)
        slice_out = torch.empty(n_pos if kind == LEAF else (n_pos, 2), dtype=torch.int64 if kind == LEAF else torch.int32, device=dev)
        base_count = torch.zeros(2, dtype=torch.int32, device=dev)
        st.finish(kind, items, n_items, gpos0, bitmap, word_prefix, n_level, meta, back, pointers, slice_out, base_count)
        self._trace("rank index + finish")
        total = word_prefix[n_words:n_words + 1]
        base, count, total = [int(v) & 0xFFFFFFFF for v in torch.cat([base_count, total]).tolist()]
        return pointers, LayerSlice(base, count, slice_out[:count]), total

    # -- one level, peer exchange: nothing comes back to the host ------------------------------
    def _peer_buffers(self, region_cap, n_level):
        if self.peer is None or not self.peer.fits(region_cap, n_level):
            if self.peer is not None:
                self.peer.close()
            self.peer = PeerBuffers(self.st, self.comm, region_cap, n_level)
        return self.peer

    def _level_peer(self, kind, items, n_items, gpos0, n_level, region_cap):
        """region_cap: the largest number of positions a rank holds at this level (same on all ranks)."""
        st, world, dev = self.st, self.world, self.device
        self._trace(f"-- level n={n_level} (peer)")
        pb = self._peer_buffers(region_cap, n_level)
        cap = pb.region_cap
        n_pos = n_items if kind == LEAF else ceil_div(n_items, 2)
        n_words = ceil_div(n_level, 32)
        bitmap = pb.bitmap[:n_words]
        bitmap.zero_()
        st.peer_scatter(kind, items, n_items, gpos0, world, self.rank, pb.arenas, cap, pb.meta)
        self._trace("scatter to owners")
        self.collectives += 1
        self.comm.stream_barrier(pb.token)
        self._trace("barrier")
        pb.serial += 1
        st.peer_owner(world, self.rank, pb.arenas, cap, ceil_div(n_level, world), pb.table, pb.table_slots, pb.serial, pb.slot_scratch,
                      pb.planes, bitmap)
        self._trace("owner")
        self._all_reduce_sum(bitmap)  # first-occurrence bits are disjoint across owners: sum == or; also the second barrier
        self._trace("all_reduce bitmap")
        word_prefix = pb.word_prefix[:n_words + 1]
        st.rank_index(bitmap, n_words, word_prefix, pb.scratch)
        pointers = torch.empty(n_pos, dtype=torch.int32, device=dev)
        slice_out = torch.empty(n_pos if kind == LEAF else (n_pos, 2), dtype=torch.int64 if kind == LEAF else torch.int32, device=dev)
        counts = torch.zeros(3, dtype=torch.int32, device=dev)  # base, count, level total
        st.peer_finish(kind, items, n_items, gpos0, bitmap, word_prefix, n_level, pb.meta, pb.own, world, cap, pointers, slice_out, counts)
        counts[2:3].copy_(word_prefix[n_words:n_words + 1])
        self._trace("rank index + finish")
        return pointers, slice_out, counts

    def _leaf_level_direct(self, body, n_local, gpos0, n_level):
        """ACGT-only leaves, dna_size <= 12: replicated direct table + all-reduce(MIN) instead of
        the record exchange.  Returns None when some rank saw another symbol."""
        st, dev = self.st, self.device
        entries = 1 << (2 * st.dna_size)
        self._trace(f"-- leaf level (direct) n={n_level}")
        table = torch.full((entries,), 0x7F7F7F7F, dtype=torch.int32, device=dev)
        tmp = torch.empty(n_local, dtype=torch.int32, device=dev)
        flag = torch.tensor([st.leaf_direct_minpos(body, n_local, gpos0, table, tmp)], dtype=torch.int32, device=dev)
        self._all_reduce_sum(flag)
        if int(flag.item()):
            return None
        self._trace("leaf minpos")
        self.collectives += 1
        self.comm.all_reduce_min(table)
        self._trace("all_reduce table")
        n_words = ceil_div(n_level, 32)
        bitmap = torch.zeros(n_words, dtype=torch.int32, device=dev)
        word_prefix = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
        scratch = torch.empty(ceil_div(n_words, 1024) + 1, dtype=torch.int32, device=dev)
        ids = torch.empty(entries, dtype=torch.int32, device=dev)
        pointers = torch.empty(n_local, dtype=torch.int32, device=dev)
        leaves_out = torch.empty(max(1, min(n_level, entries)), dtype=torch.int64, device=dev)
        st.leaf_direct_finish(table, n_level, tmp, n_local, bitmap, word_prefix, scratch, ids, pointers, leaves_out)
        total = int(word_prefix[n_words].item()) & 0xFFFFFFFF
        self._trace("leaf ids + resolve")
        # every rank holds the whole (small) leaf table; rank 0's copy is the one that is gathered
        if self.rank == 0:
            return pointers, leaves_out, (0, total, total)
        return pointers, leaves_out[:0], (total, 0, total)

    # -- whole build ------------------------------------------------------------------
    def _run_level(self, kind, items, n_items, gpos0, plan, level):
        """-> (pointers, slice items, (base, count, level total) as ints or as a device tensor)."""
        n_level = plan.level_total(level)
        if self.exchange == "peer":
            try:
                return self._level_peer(kind, items, n_items, gpos0, n_level, max(1, plan.shard >> level))
            except PeerUnavailable as e:  # raised on every rank together, before any stage ran
                if self.rank == 0:
                    print(f"[genome-compression_b200.dist] {e}", file=sys.stderr, flush=True)
                self.exchange, self.peer = "collective", None
        pointers, sl, total = self._level(kind, items, n_items, gpos0, n_level)
        return pointers, sl.items, (sl.base, sl.count, total)

    def build_from_leaves(self, local_leaves, n_leaves_total: int, leaf_level=None) -> DistTree:
        """local_leaves: int64 device tensor with this rank's range of packed leaves
        (ShardPlan.level_range(rank, 0)).  leaf_level: an already finished leaf level."""
        plan = ShardPlan(n_leaves_total, self.world, self.cut)
        lo, hi = plan.level_range(self.rank, 0)
        if leaf_level is None:
            assert local_leaves.numel() == hi - lo, (local_leaves.numel(), lo, hi)
            leaf_level = self._run_level(LEAF, local_leaves, hi - lo, lo, plan, 0)
        pointers = leaf_level[0]
        done = [leaf_level[1:]]
        n_sharded = plan.sharded_levels()
        for level in range(1, n_sharded):
            lo, hi = plan.level_range(self.rank, level)
            prev_lo, prev_hi = plan.level_range(self.rank, level - 1)
            assert hi - lo == ceil_div(prev_hi - prev_lo, 2)
            pointers, items, counts = self._run_level(NODE, pointers, prev_hi - prev_lo, lo, plan, level)
            done.append((items, counts))
        # the last sharded level's pointers go to rank 0, which finishes alone
        last = n_sharded - 1
        per_rank = [plan.level_range(r, last)[1] - plan.level_range(r, last)[0] for r in range(self.world)]
        self._trace("-- upper levels")
        upper, root = None, None
        pb = self.peer if self.exchange == "peer" else None
        if pb is not None and sum(per_rank) * 4 <= self.world * pb.region_cap * 8:
            payload = self.st.peer_payload(pb.arenas[0], self.world, pb.region_cap)
            self.st.peer_put(payload + 4 * sum(per_rank[:self.rank]), pointers, 4 * per_rank[self.rank])
            self.collectives += 1
            self.comm.stream_barrier(pb.token)
            self._trace("pointers to rank 0")
            if self.rank == 0:
                upper = self.st.upper_levels(payload, sum(per_rank), leaf_pointers=(n_sharded == 1))
            self.collectives += 1
            self.comm.stream_barrier(pb.token)  # rank 0 has read its arena: free for the next build
        else:
            gathered = self._gather_rows(pointers, per_rank, dst=0)
            self._trace("pointers to rank 0")
            if self.rank == 0:
                upper = self.st.upper_levels(gathered, gathered.shape[0], leaf_pointers=(n_sharded == 1))
        if self.rank == 0:
            root = upper.root()
        self._trace("upper levels on rank 0")
        # unique counts: one host read for the whole build
        on_device = [c for _, c in done if torch.is_tensor(c)]
        host = torch.stack(on_device).tolist() if on_device else []
        slices, totals = [], []
        for items, c in done:
            base, count, total = [int(v) & 0xFFFFFFFF for v in (host.pop(0) if torch.is_tensor(c) else c)]
            if total >= 1 << 29:
                raise OverflowError("a layer outgrew the 29-bit pointer index (src/shared_tree.cpp:54-67)")
            slices.append(LayerSlice(base, count, items[:count]))
            totals.append(total)
        self.builds += 1
        return DistTree(self.st.dna_size, n_leaves_total, slices[0], slices[1:], totals, upper, root, self.builds)

    def build_from_body(self, local_body, n_bases_total: int) -> DistTree:
        """local_body: uint8 device tensor holding the bases of this rank's leaf range."""
        S = self.st.dna_size
        n_leaves_total = n_bases_total // S
        plan = ShardPlan(n_leaves_total, self.world, self.cut)
        lo, hi = plan.level_range(self.rank, 0)
        if S <= 12:
            done = self._leaf_level_direct(local_body, hi - lo, lo, plan.level_total(0))
            if done is not None:
                return self.build_from_leaves(None, n_leaves_total, leaf_level=done)
        leaves = torch.empty(hi - lo, dtype=torch.int64, device=self.device)
        self.st.pack_body(local_body, hi - lo, leaves)
        return self.build_from_leaves(leaves, n_leaves_total)

    def gather(self, tree: DistTree):
        """Collects every slice on rank 0 and assembles an ordinary SharedTree there
        (None on the other ranks).  Untimed in bench.py: only needed to serialize / verify.
        Must be called before the builder's next build (which reuses rank 0's top-layer handle)."""
        if tree.build_id != self.builds:
            raise ValueError("gather() needs the result of this builder's latest build")
        def gather_slice(sl):
            counts = self.comm.all_gather_object(sl.count)
            return self._gather_rows(sl.items.contiguous(), counts, dst=0)

        leaves = gather_slice(tree.leaves)
        layers = [gather_slice(sl) for sl in tree.layers]
        if self.rank != 0:
            return None
        top, root = self.st.upper_layers(tree.upper)
        return self.st.assemble(leaves, layers + top, root, tree.width)
