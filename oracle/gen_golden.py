#!/usr/bin/env python3
"""Generates tests/golden/ from the UNMODIFIED reference (oracle/_ref/libref.so).

Run in the build container, where /root/reference exists:

    make -C oracle ref && python oracle/gen_golden.py

Everything written is an OUTPUT of executing the reference (or, for
tests/golden/corpus/*.2bit, a re-encoding of the public DNA-corpus text files
the reference ships as test data, needed because /root/reference does not exist
on the GPU box).  No reference source is copied.

Files:
  corpus/<name>.2bit   u64 length + 2-bit codes (a0 c1 g2 t3), 4 per byte, low bits first
  corpus.json          per (file, dna_size): width, leaf/node counts, per-layer counts,
                       pre-/post-sort stream length + sha256, bits/base
  small.json           tiny inputs with complete expected streams (hex), decoded leaves,
                       per-level pointer arrays
  primitives.npz       fuzz vectors for leaf / pointer / node primitives
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Ref, REFERENCE_ROOT  # noqa: E402

GOLD = ROOT / "tests" / "golden"
DATA = REFERENCE_ROOT / "data"
CORPUS = ["chmpxx", "chntxx", "hehcmv", "humdyst", "humghcs", "humhbb", "humhdab", "humprtb", "mpomtcg", "mtpacga", "vaccg"]
HUMHBB_SIZES = [1, 2, 4, 8, 11, 12, 13, 15, 16]


def to_2bit(text: bytes) -> bytes:
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(b"acgt"):
        lut[ch] = i
    codes = lut[np.frombuffer(text, dtype=np.uint8)]
    assert (codes < 4).all(), "corpus file is not pure lowercase acgt"
    pad = (-len(codes)) % 4
    codes = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    packed = codes[:, 0] | (codes[:, 1] << 2) | (codes[:, 2] << 4) | (codes[:, 3] << 6)
    return np.uint64(len(text)).tobytes() + packed.astype(np.uint8).tobytes()


def tree_record(ref: Ref, tree, n_bases_in: int, S: int) -> dict:
    pre = tree.serialize()
    rec = {
        "dna_size": S,
        "width": int(tree.width()),
        "depth": int(tree.depth()),
        "leaves": int(tree.leaf_count()),
        "nodes": int(tree.node_count()),
        "layer_counts": [int(c) for c in tree.layer_counts()],
        "pre_bytes": len(pre),
        "pre_sha256": hashlib.sha256(pre).hexdigest(),
    }
    tree.sort()
    post = tree.serialize()
    assert tree.bytes() == len(post)
    rec["post_bytes"] = len(post)
    rec["post_sha256"] = hashlib.sha256(post).hexdigest()
    rec["bits_per_base"] = 8.0 * len(post) / (rec["width"] * S)
    return rec


def main() -> None:
    ref = Ref()
    (GOLD / "corpus").mkdir(parents=True, exist_ok=True)

    # ---- corpus -----------------------------------------------------------
    corpus = {}
    texts = {}
    for name in CORPUS:
        text = (DATA / name).read_bytes()
        texts[name] = text
        (GOLD / "corpus" / f"{name}.2bit").write_bytes(to_2bit(text))
    merged = b"".join(texts[n] for n in CORPUS)
    assert merged == (DATA / "merged").read_bytes(), "merged != alphabetical concatenation"
    chm = texts["chmpxx"]
    edited = b"".join(chm[i:i + 20] + b"\n" for i in range(0, len(chm), 20))
    ed_ref = (DATA / "edited").read_bytes()
    assert edited.rstrip(b"\n") == ed_ref.rstrip(b"\n"), "edited != chmpxx wrapped at 20 columns"
    edited_trailing_newline = ed_ref.endswith(b"\n")

    for name in CORPUS + ["merged", "edited"]:
        sizes = HUMHBB_SIZES if name == "humhbb" else [12]
        for S in sizes:
            tree = ref.build_file(DATA / name, S)
            rec = tree_record(ref, tree, (DATA / name).stat().st_size, S)
            rec["file_bytes"] = (DATA / name).stat().st_size
            corpus[f"{name}:{S}"] = rec
            print(name, S, rec["width"], rec["post_bytes"], rec["post_sha256"][:12])
    meta = {"corpus_order": CORPUS, "edited_wrap": 20, "edited_trailing_newline": edited_trailing_newline,
            "records": corpus}
    (GOLD / "corpus.json").write_text(json.dumps(meta, indent=1) + "\n")

    # ---- small complete vectors --------------------------------------------
    small = {}

    def add_small(name: str, leaves: np.ndarray, S: int):
        tree = ref.build(leaves, S)
        _, levels = ref.build_levels(leaves, S)
        pre = tree.serialize()
        rec = {"dna_size": S, "input_leaves": [f"{int(v):x}" for v in leaves],
               "pre_hex": pre.hex(), "width": int(tree.width()), "leaf_count": int(tree.leaf_count()),
               "layer_counts": [int(c) for c in tree.layer_counts()],
               "stored_leaves": [f"{int(v):x}" for v in tree.leaves()],
               "layers": [[f"{int(a):08x}" for a in tree.layer(k).reshape(-1)] for k in range(tree.depth() - 1)],
               "levels": [[f"{int(a):08x}" for a in lv] for lv in levels],
               "histograms": [[int(c) for c in tree.histogram(k)] for k in range(tree.depth() - 1)]}
        tree.sort()
        post = tree.serialize()
        rec["post_hex"] = post.hex()
        rec["decoded"] = [f"{int(v):x}" for v in tree.decode()]
        small[name] = rec

    def pk(s: str, S=12) -> int:
        return ref.pack(s.encode(), S)[0]

    A, B = "ACGTTTGACCAT", "GGGACCATTTAC"
    comp = str.maketrans("ACGT", "TGCA")
    hand = [A, B, B[::-1], A[::-1], A.translate(comp), B.translate(comp), B, A, A]
    add_small("hand_9_leaves", np.array([pk(s) for s in hand], dtype=np.uint64), 12)
    rng = np.random.default_rng(20261018)
    for n in (1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 33, 100):
        # tiny alphabet => many duplicates, palindromes and complements
        S = 12
        strs = ["".join(rng.choice(list("AT"), size=2)) * 6 for _ in range(n)]
        add_small(f"at_repeat_{n}", np.array([pk(s) for s in strs], dtype=np.uint64), S)
    for S in (1, 2, 3, 5, 8, 13, 16):
        n = 67
        strs = ["".join(rng.choice(list("ACGT"), size=S)) for _ in range(n)]
        add_small(f"acgt_S{S}_{n}", np.array([pk(s, S) for s in strs], dtype=np.uint64), S)
    for S in (4, 12, 16):
        n = 50
        strs = ["".join(rng.choice(list("ACGTRYKMSWBDHVN-"), size=S)) for _ in range(n)]
        strs += [s[::-1] for s in strs[:10]] + ["-" * S, "S" * S, "N" * S, "-" * S]
        add_small(f"iupac_S{S}", np.array([pk(s, S) for s in strs], dtype=np.uint64), S)
    # the reference's own tests: tests/test.cpp:234-264 ({a,a,t,a,a,t,t,t}) and :266-291
    a = pk("AAAAAAAAAAAA")
    t = pk("TTTTTTTTTTTT")
    add_small("test_tree_transposition", np.array([a, a, t, a, a, t, t, t], dtype=np.uint64), 12)
    (GOLD / "small.json").write_text(json.dumps(small, indent=0) + "\n")

    # ---- FASTA text vectors (ingest semantics, src/fasta_reader.cpp:40-68) ----
    import tempfile, os
    fasta = {}
    cases = {
        "plain_lower": b"acgtacgtacgtacgtacgtacgtac",
        "header_wrapped": b">seq1 description\nACGTACGTAC\nGTACGTACGT\nACGTAC\n",
        "multi_record": b">r1\nACGTRYKMSWBD\nHVN-acgtrykm\n>r2\nswbdhvn-ACGTNN\nNNNNNN\n",
        "no_trailing_newline": b">h\nACGTACGTACGTACGTACGTACGTA",
        "header_blank_header": b">h1\n\n>h2\nACGTACGTACGTACGT\n",
        "blank_blank_data": b"\n\nACGTACGTACGTAAAA\n",
        "trailing_blank_lines": b"ACGTACGTACGTACGTACGTACGT\n\n\n",
    }
    for name, text in cases.items():
        for S in (12, 5):
            with tempfile.NamedTemporaryFile(delete=False) as f:
                f.write(text)
                path = f.name
            leaves = ref.read_genome(path, S)
            os.unlink(path)
            fasta[f"{name}:{S}"] = {"text_hex": text.hex(), "dna_size": S, "leaves": [f"{int(v):x}" for v in leaves]}
    (GOLD / "fasta.json").write_text(json.dumps(fasta, indent=0) + "\n")

    # ---- primitive fuzz vectors ---------------------------------------------
    recs = {"leaf_S": [], "leaf_v": [], "leaf_t": [], "leaf_m": [], "leaf_i": [], "leaf_c": [], "leaf_f": []}
    for S in range(1, 17):
        mask = (1 << (4 * S)) - 1
        for j in range(400):
            mode = j % 3
            if mode == 0:
                v = int.from_bytes(rng.bytes(8), "little")
            elif mode == 1:
                v = sum(int(rng.choice([1, 2, 4, 8])) << (4 * i) for i in range(S))
            else:
                half = [int(rng.choice([1, 2, 4, 8, 0, 9, 6, 15])) for _ in range((S + 1) // 2)]
                full = half + half[::-1][S % 2:]
                v = sum(c << (4 * i) for i, c in enumerate(full[:S]))
            v &= mask
            c, f = ref.leaf_canonical(v, S)
            recs["leaf_S"].append(S); recs["leaf_v"].append(v)
            recs["leaf_t"].append(ref.transposed(v, S)); recs["leaf_m"].append(ref.mirrored(v, S))
            recs["leaf_i"].append(ref.inverted(v, S)); recs["leaf_c"].append(c); recs["leaf_f"].append(f)
    NULL = ref.lib.ref_pointer_null()
    assert NULL == 0x9FFFFFFF

    def rp():
        k = rng.integers(10)
        if k == 0:
            return NULL
        inv = int(rng.integers(4) == 0)
        idx = int(rng.integers(4)) if rng.integers(2) else int(rng.integers(1 << 28))
        return ref.pointer(idx, int(rng.integers(2)), int(rng.integers(2)), inv)

    node_l, node_r, node_cl, node_cr, node_f = [], [], [], [], []
    for _ in range(20000):
        l, r = rp(), rp()
        cl, cr, f = ref.node_canonical(l, r)
        node_l.append(l); node_r.append(r); node_cl.append(cl); node_cr.append(cr); node_f.append(f)
    ptr_raw, ptr_ser, ptr_comp = [], [], []
    edge = [0, 15, 16, 4111, 4112, 1052687, 1052688, (1 << 28) + 1052687, 0x1FFFFFFE]
    for j in range(4000):
        idx = edge[j] if j < len(edge) else int(rng.choice([rng.integers(16), rng.integers(4112), rng.integers(1052688), rng.integers(1 << 28)]))
        p = ref.pointer(idx, int(rng.integers(2)), int(rng.integers(2)), int(rng.integers(2)))
        ptr_raw.append(p)
        s = ref.pointer_serialize(p)
        ptr_ser.append(int.from_bytes(s + b"\0" * (4 - len(s)), "big") | 0)  # left-aligned
        ptr_comp.append([ref.compose(p, m, t) for m in (0, 1) for t in (0, 1)])
    ptr_raw.append(NULL); s = ref.pointer_serialize(NULL)
    ptr_ser.append(int.from_bytes(s, "big")); ptr_comp.append([ref.compose(NULL, m, t) for m in (0, 1) for t in (0, 1)])
    np.savez_compressed(
        GOLD / "primitives.npz",
        **{k: np.array(v, dtype=np.uint64) for k, v in recs.items()},
        node_l=np.array(node_l, dtype=np.uint32), node_r=np.array(node_r, dtype=np.uint32),
        node_cl=np.array(node_cl, dtype=np.uint32), node_cr=np.array(node_cr, dtype=np.uint32),
        node_f=np.array(node_f, dtype=np.uint8),
        ptr_raw=np.array(ptr_raw, dtype=np.uint32), ptr_ser=np.array(ptr_ser, dtype=np.uint32),
        ptr_comp=np.array(ptr_comp, dtype=np.uint32))
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
