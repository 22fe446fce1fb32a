"""CPU: the C restatement (oracle/oracle.c) against golden values produced by the
unmodified reference (tests/golden/, generator: oracle/gen_golden.py)."""
import hashlib

import numpy as np
import pytest

from conftest import CORPUS_CASES, corpus_text


def hexs(a):
    return [f"{int(v):x}" for v in a]


def test_leaf_primitives(oracle, golden):
    p = golden["prim"]
    for S, v, t, m, i, c, f in zip(p["leaf_S"], p["leaf_v"], p["leaf_t"], p["leaf_m"], p["leaf_i"], p["leaf_c"], p["leaf_f"]):
        S, v = int(S), int(v)
        assert oracle.transposed(v, S) == int(t)
        assert oracle.mirrored(v, S) == int(m)
        assert oracle.inverted(v, S) == int(i)
        assert oracle.leaf_canonical(v, S) == (int(c), int(f))


def test_node_and_pointer_primitives(oracle, golden):
    p = golden["prim"]
    for l, r, cl, cr, f in zip(p["node_l"], p["node_r"], p["node_cl"], p["node_cr"], p["node_f"]):
        assert oracle.node_canonical(int(l), int(r)) == (int(cl), int(cr), int(f))
    for raw, ser, comp in zip(p["ptr_raw"], p["ptr_ser"], p["ptr_comp"]):
        raw = int(raw)
        s = oracle.pointer_serialize(raw)
        assert int.from_bytes(s + b"\0" * (4 - len(s)), "big") == int(ser)
        got = [oracle.compose(raw, m, t) for m in (0, 1) for t in (0, 1)]
        assert got == [int(x) for x in comp]


def test_fasta_ingest(oracle, golden):
    for key, rec in golden["fasta"].items():
        text = bytes.fromhex(rec["text_hex"])
        leaves = oracle.fasta_to_leaves(text, rec["dna_size"])
        assert hexs(leaves) == rec["leaves"], key


def test_unknown_symbol_and_tail(oracle):
    with pytest.raises(ValueError, match="unknown symbol 88"):
        oracle.fasta_to_leaves(b"ACGTACGTACGx", 12)
    # an invalid symbol inside the dropped tail is never converted (fasta_reader.cpp:59-61)
    assert len(oracle.fasta_to_leaves(b"ACGTACGTACGTACx", 12)) == 1
    with pytest.raises(ValueError, match="unknown symbol 13"):
        oracle.fasta_to_leaves(b"ACGTACGTACG\r\nACGT", 12)


def test_small_vectors(oracle, golden):
    for name, rec in golden["small"].items():
        S = rec["dna_size"]
        leaves = np.array([int(x, 16) for x in rec["input_leaves"]], dtype=np.uint64)
        tree, levels = oracle.build_levels(leaves, S)
        assert [[f"{int(a):08x}" for a in lv] for lv in levels] == rec["levels"], name
        assert tree.serialize().hex() == rec["pre_hex"], name
        assert hexs(tree.leaves()) == rec["stored_leaves"], name
        for k, want in enumerate(rec["layers"]):
            assert [f"{int(a):08x}" for a in tree.layer(k).reshape(-1)] == want, (name, k)
            assert [int(c) for c in tree.histogram(k)] == rec["histograms"][k], (name, k)
        assert tree.width() == rec["width"] and tree.leaf_count() == rec["leaf_count"]
        tree.sort()
        assert tree.serialize().hex() == rec["post_hex"], name
        assert hexs(tree.decode()) == rec["decoded"] == rec["input_leaves"], name
        back = oracle.deserialize(bytes.fromhex(rec["post_hex"]), S)
        assert hexs(back.decode()) == rec["input_leaves"], name
        assert back.width() == rec["width"]
        idx = np.arange(len(leaves), dtype=np.uint64)
        assert np.array_equal(back.random_access(idx), leaves)


def test_hand_checked_stream(oracle, golden):
    # SURVEY Appendix B, hand-checkable 9-leaf vector (80 bytes)
    want = ("00" "0000000000000002" "124888412218" "218881221444" "0000000000000003" "1001" "0011" "cfffffff00"
            "0000000000000003" "0010" "0031" "cfffffff02" "0000000000000002" "0021" "cfffffff02"
            "0000000000000001" "0011")
    assert golden["small"]["hand_9_leaves"]["pre_hex"] == want


@pytest.mark.parametrize("name,S", CORPUS_CASES)
def test_corpus(oracle, golden, name, S):
    rec = golden["corpus"]["records"][f"{name}:{S}"]
    text = corpus_text(name)
    assert len(text) == rec["file_bytes"]
    leaves = oracle.fasta_to_leaves(text, S)
    tree = oracle.build(leaves, S)
    assert tree.width() == rec["width"] == len(leaves)
    assert tree.depth() == rec["depth"]
    assert tree.leaf_count() == rec["leaves"] and tree.node_count() == rec["nodes"]
    assert tree.layer_counts() == rec["layer_counts"]
    pre = tree.serialize()
    assert len(pre) == rec["pre_bytes"] and hashlib.sha256(pre).hexdigest() == rec["pre_sha256"]
    tree.sort()
    post = tree.serialize()
    assert tree.bytes() == len(post) == rec["post_bytes"]
    assert hashlib.sha256(post).hexdigest() == rec["post_sha256"]
    assert np.array_equal(tree.decode(), leaves)
