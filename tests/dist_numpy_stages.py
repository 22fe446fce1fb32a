"""TEST-ONLY numpy twin of the multi-GPU stage kernels (include/shared_tree_b200_dist.h),
built on the oracle's primitives.  It lets the host-side orchestration in
genome-compression_b200/dist.py (shard plan, splits, collectives, level loop, gather) run
under gloo on CPUs.  The product never imports this."""
from __future__ import annotations

import numpy as np
import torch

IDX = 0x1FFFFFFF
NULL = 0x9FFFFFFF


def _u32(t):
    return t.numpy().view(np.uint32)


def _flags(f):  # oracle: bit0 mirror, bit1 transpose, bit2 invariant -> bits 29..31
    return ((f & 1) << 29) | (((f >> 1) & 1) << 30) | (((f >> 2) & 1) << 31)


def _finish(idx, flags):
    if flags & 0x80000000:
        flags &= ~0x20000000
    return (idx | flags) & 0xFFFFFFFF


class _Upper:
    def __init__(self, layers, root):
        self.layers, self._root = layers, root

    def root(self):
        return self._root


class _Assembled:
    def __init__(self, leaves, layers, root, width):
        self.leaves_, self.layers_, self.root_, self.width_ = leaves, layers, root, width


class NumpyStages:
    def __init__(self, oracle, dna_size):
        self.o = oracle
        self.dna_size = dna_size
        self.device = torch.device("cpu")

    # (key, flags) of every local position
    def _produce(self, kind, items, n_items):
        if kind == 0:
            vals = items.numpy().view(np.uint64)[:n_items]
            out = [self.o.leaf_canonical(int(v), self.dna_size) for v in vals]
            return [(c, _flags(f)) for c, f in out]
        cur = _u32(items)[:n_items]
        res = []
        for i in range((n_items + 1) // 2):
            l = int(cur[2 * i])
            r = int(cur[2 * i + 1]) if 2 * i + 1 < n_items else NULL
            cl, cr, f = self.o.node_canonical(l, r)
            res.append(((cl << 32) | cr, _flags(f)))
        return res

    @staticmethod
    def _owner(key, world):
        return ((key * 0x9E3779B97F4A7C15) >> 40) % world

    def pack_body(self, body, n_leaves, leaves_out):
        text = body.numpy().tobytes()
        S = self.dna_size
        for i in range(n_leaves):
            v, bad = self.o.pack(text[i * S:(i + 1) * S], S)
            assert bad < 0
            leaves_out.numpy().view(np.uint64)[i] = v

    @staticmethod
    def _two_bit(v, S):
        code = 0
        for i in range(S):
            nib = (v >> (4 * i)) & 0xF
            if nib not in (1, 2, 4, 8):
                return None
            code |= {1: 0, 2: 1, 4: 2, 8: 3}[nib] << (2 * i)
        return code

    def leaf_direct_minpos(self, body, n_local, gpos0, table, tmp):
        text, S = body.numpy().tobytes(), self.dna_size
        tab, t = _u32(table), _u32(tmp)
        for i in range(n_local):
            v, bad = self.o.pack(text[i * S:(i + 1) * S], S)
            assert bad < 0
            if self._two_bit(v, S) is None:
                return 1
            canon, f = self.o.leaf_canonical(v, S)
            slot = self._two_bit(canon, S)
            tab[slot] = min(int(tab[slot]), gpos0 + i)
            t[i] = slot | _flags(f)
        return 0

    def leaf_direct_finish(self, table, n_level, tmp, n_local, bitmap, word_prefix, scratch, ids, pointers, leaves_out):
        tab, b = _u32(table), _u32(bitmap)
        used = np.nonzero(tab < 0x7F7F7F7F)[0]
        for s_ in used:
            q = int(tab[s_])
            b[q >> 5] |= np.uint32(1 << (q & 31))
        n_words = (n_level + 31) // 32
        self.rank_index(bitmap, n_words, word_prefix, scratch)
        wp = _u32(word_prefix)
        idv, out = _u32(ids), leaves_out.numpy().view(np.uint64)
        S = self.dna_size
        for s_ in used:
            q = int(tab[s_])
            ident = int(wp[q >> 5]) + bin(int(b[q >> 5]) & ((1 << (q & 31)) - 1)).count("1")
            idv[s_] = ident
            out[ident] = sum((1 << ((int(s_) >> (2 * i)) & 3)) << (4 * i) for i in range(S))
        t, ptr = _u32(tmp), _u32(pointers)
        for i in range(n_local):
            ptr[i] = _finish(int(idv[int(t[i]) & IDX]), int(t[i]) & ~IDX & 0xFFFFFFFF)

    def partition(self, kind, items, n_items, gpos0, world, keys, gpos, meta, counts):
        recs = self._produce(kind, items, n_items)
        order = sorted(range(len(recs)), key=lambda i: self._owner(recs[i][0], world))  # stable
        k, g, m = keys.numpy().view(np.uint64), _u32(gpos), _u32(meta)
        c = np.zeros(world, dtype=np.int32)
        for dst, i in enumerate(order):
            k[dst] = recs[i][0]
            g[dst] = gpos0 + i
            m[dst] = i | recs[i][1]
            c[self._owner(recs[i][0], world)] += 1
        counts.numpy()[:] = c

    def table(self, cap):
        return torch.empty(0, dtype=torch.int64)

    def owner(self, keys, gpos, n, table, cap, answers, bitmap):
        k, g = keys.numpy().view(np.uint64)[:n], _u32(gpos)[:n]
        first = {}
        for key, pos in zip(k.tolist(), g.tolist()):
            if key not in first or pos < first[key]:
                first[key] = pos
        a, b = _u32(answers), _u32(bitmap)
        for j, (key, pos) in enumerate(zip(k.tolist(), g.tolist())):
            q = first[key]
            a[j] = q
            if q == pos:
                b[q >> 5] |= np.uint32(1 << (q & 31))

    def rank_index(self, bitmap, n_words, word_prefix, scratch):
        pop = np.array([bin(int(w)).count("1") for w in _u32(bitmap)[:n_words]], dtype=np.uint64)
        wp = _u32(word_prefix)
        wp[0] = 0
        if n_words:
            wp[1:n_words + 1] = np.cumsum(pop).astype(np.uint32)

    def finish(self, kind, items, n_items, gpos0, bitmap, word_prefix, n_level, meta, answers, pointers, slice_out, base_count):
        b, wp = _u32(bitmap), _u32(word_prefix)
        n_words = (n_level + 31) // 32

        def rank(q):
            if q >= n_level:
                return int(wp[n_words])
            return int(wp[q >> 5]) + bin(int(b[q >> 5]) & ((1 << (q & 31)) - 1)).count("1")

        recs = self._produce(kind, items, n_items)
        n_pos = len(recs)
        base = rank(gpos0)
        bc = _u32(base_count)
        bc[0], bc[1] = base, rank(gpos0 + n_pos) - base
        ptr = _u32(pointers)
        for i, (key, f) in enumerate(recs):
            g = gpos0 + i
            if (int(b[g >> 5]) >> (g & 31)) & 1:
                ident = rank(g)
                if kind == 0:
                    slice_out.numpy().view(np.uint64)[ident - base] = key
                else:
                    s = _u32(slice_out)
                    s[ident - base, 0], s[ident - base, 1] = key >> 32, key & 0xFFFFFFFF
                ptr[i] = _finish(ident, f)
        m, a = _u32(meta), _u32(answers)
        for j in range(n_pos):
            pos, q = int(m[j]) & IDX, int(a[j])
            if q != gpos0 + pos:
                ptr[pos] = _finish(rank(q), int(m[j]) & ~IDX & 0xFFFFFFFF)

    def upper_levels(self, pointers, n, leaf_pointers):
        cur = [int(x) for x in _u32(pointers)[:n]]
        layers = []
        while len(cur) > 1 or (leaf_pointers and not layers):
            ids, nodes, nxt = {}, [], []
            for i in range((len(cur) + 1) // 2):
                l = cur[2 * i]
                r = cur[2 * i + 1] if 2 * i + 1 < len(cur) else NULL
                cl, cr, f = self.o.node_canonical(l, r)
                key = (cl << 32) | cr
                if key not in ids:
                    ids[key] = len(nodes)
                    nodes.append((cl, cr))
                nxt.append(_finish(ids[key], _flags(f)))
            layers.append(np.array(nodes, dtype=np.uint32).reshape(-1, 2))
            cur = nxt
        return _Upper(layers, cur[0])

    def upper_layers(self, upper):
        return [torch.from_numpy(l.view(np.int32).copy()) for l in upper.layers], upper.root()

    def assemble(self, leaves, layers, root, width):
        return _Assembled(leaves.numpy().view(np.uint64).copy(), [_u32(l).reshape(-1, 2).copy() for l in layers], root, width)

    def sync(self):
        pass


class _Arena:
    """Numpy stand-in for one rank's peer-mapped exchange arena (csrc/dist_peer.cu)."""

    def __init__(self, world, cap):
        self.count = np.zeros(world, dtype=np.uint32)            # hdr: records received per source
        self.cursor = np.zeros(world, dtype=np.uint32)           # records sent per owner
        self.ans = np.zeros((world, cap), dtype=np.uint32)       # answers to my records, by owner
        self.gpos = np.zeros((world, cap), dtype=np.uint32)      # received records, by source
        self.keys = np.zeros((world, cap), dtype=np.uint64)
        self.payload = np.zeros(world * cap * 2, dtype=np.uint32)


class NumpyPeerStages(NumpyStages):
    """Adds the peer-exchange stages: 'pointers' are indices into a process-wide arena list, so the
    virtual ranks of ThreadComm (threads of this process) write into each other's arenas the way
    the GPUs do over NVLink."""
    supports_peer = True
    _arenas: list = []
    _lock = __import__("threading").Lock()

    def peer_arena_bytes(self, world, region_cap):
        return (world, region_cap)

    def peer_alloc(self, shape):
        with self._lock:
            NumpyPeerStages._arenas.append(_Arena(*shape))
            return len(NumpyPeerStages._arenas), b""  # 'address' 0 stays null

    def peer_free(self, ptr):
        NumpyPeerStages._arenas[ptr - 1] = None

    def peer_scatter(self, kind, items, n_items, gpos0, world, rank, arenas, region_cap, meta):
        recs = self._produce(kind, items, n_items)
        assert len(recs) <= region_cap
        mine = self._arenas[arenas[rank] - 1]
        mine.cursor[:] = 0
        m = _u32(meta).reshape(world, -1)
        for i, (key, f) in reversed(list(enumerate(recs))):  # any order inside a region is legal
            o = self._owner(key, world)
            k = int(mine.cursor[o])
            dst = self._arenas[arenas[o] - 1]
            dst.keys[rank, k], dst.gpos[rank, k] = key, gpos0 + i
            m[o, k] = i | f
            mine.cursor[o] = k + 1
        for o in range(world):
            self._arenas[arenas[o] - 1].count[rank] = mine.cursor[o]

    def peer_owner(self, world, rank, arenas, region_cap, expected, table, table_slots, serial, slot_scratch, planes, bitmap):
        mine = self._arenas[arenas[rank] - 1]
        first = {}
        for src in range(world):
            for k in range(int(mine.count[src])):
                key, pos = int(mine.keys[src, k]), int(mine.gpos[src, k])
                if key not in first or pos < first[key]:
                    first[key] = pos
        b = _u32(bitmap)
        for src in range(world):
            back = self._arenas[arenas[src] - 1].ans
            for k in range(int(mine.count[src])):
                key, pos = int(mine.keys[src, k]), int(mine.gpos[src, k])
                q = first[key]
                if q == pos:
                    b[q >> 5] |= np.uint32(1 << (q & 31))
                else:
                    back[rank, k] = q  # sparse: only later occurrences are answered

    def peer_finish(self, kind, items, n_items, gpos0, bitmap, word_prefix, n_level, meta, arena, world, region_cap, pointers,
                    slice_out, base_count):
        mine = self._arenas[arena - 1]
        b, wp = _u32(bitmap), _u32(word_prefix)
        n_words = (n_level + 31) // 32

        def rank_of(q):
            if q >= n_level:
                return int(wp[n_words])
            return int(wp[q >> 5]) + bin(int(b[q >> 5]) & ((1 << (q & 31)) - 1)).count("1")

        recs = self._produce(kind, items, n_items)
        base = rank_of(gpos0)
        bc = _u32(base_count)
        bc[0], bc[1] = base, rank_of(gpos0 + len(recs)) - base
        ptr = _u32(pointers)
        for i, (key, f) in enumerate(recs):
            g = gpos0 + i
            if (int(b[g >> 5]) >> (g & 31)) & 1:
                ident = rank_of(g)
                if kind == 0:
                    slice_out.numpy().view(np.uint64)[ident - base] = key
                else:
                    s = _u32(slice_out)
                    s[ident - base, 0], s[ident - base, 1] = key >> 32, key & 0xFFFFFFFF
                ptr[i] = _finish(ident, f)
        m = _u32(meta).reshape(world, -1)
        for o in range(world):
            for k in range(int(mine.cursor[o])):
                pos = int(m[o, k]) & IDX
                g = gpos0 + pos
                if not (int(b[g >> 5]) >> (g & 31)) & 1:
                    ptr[pos] = _finish(rank_of(int(mine.ans[o, k])), int(m[o, k]) & ~IDX & 0xFFFFFFFF)

    def peer_payload(self, arena, world, region_cap):
        return _PayloadRef(arena)

    def peer_put(self, dst, src, nbytes):
        n = nbytes // 4
        self._arenas[dst.arena - 1].payload[dst.offset:dst.offset + n] = _u32(src)[:n]

    def upper_levels(self, pointers, n, leaf_pointers):
        if isinstance(pointers, _PayloadRef):
            pointers = torch.from_numpy(self._arenas[pointers.arena - 1].payload[pointers.offset:pointers.offset + n].view(np.int32).copy())
        return super().upper_levels(pointers, n, leaf_pointers)


class _PayloadRef:
    """payload 'address': arena + word offset; supports the `+ bytes` the builder does."""

    def __init__(self, arena, offset=0):
        self.arena, self.offset = arena, offset

    def __add__(self, nbytes):
        assert nbytes % 4 == 0
        return _PayloadRef(self.arena, self.offset + nbytes // 4)
