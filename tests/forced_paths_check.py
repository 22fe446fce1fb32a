"""Forces the large-input code paths onto inputs small enough for the oracle, through the
handle's options (stb_set_option), given as name=value arguments:
  bucket_min=1                                  on-chip (bucketed) deduplication of the node levels
  bucket_min=1 bucket_levels=9                  ... of every large-path node level, with the exact singleton filter in front
  bucket_min=1 bucket_cap=16                    ... with final buckets that outgrow their regions: the exact-size second pass + chunked dedup
  bucket_min=1 bucket_slack_permille=0 bucket_headroom=0   ... with first-pass buckets that overflow: the hash-table fallback
  coop_max=0                                    no cooperative middle launch: every level as separate kernels
  stream_chunk_log2=12 stream_min_chunks=2      streaming (chunked) build from host memory"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import corpus_text, load_package  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def main():
    import torch
    stb, oracle = load_package(), Oracle()
    pkg = stb
    options = dict(a.split("=") for a in sys.argv[1:])

    class Tree(stb.SharedTree):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            for name, value in options.items():
                self.set_option(name, int(value))
    stb = type("Pkg", (), {"SharedTree": Tree, "synth_genome": staticmethod(stb.synth_genome)})
    for name, S in (("merged", 12), ("humhbb", 5), ("vaccg", 16), ("chmpxx", 1)):
        text = corpus_text(name)
        leaves = oracle.fasta_to_leaves(text, S)
        got, want = stb.SharedTree(S).build_from_fasta(text), oracle.build(leaves, S)
        assert got.layer_counts() == want.layer_counts(), (name, S)
        assert got.serialize() == want.serialize(), (name, S)
        got.sort(); want.sort()
        assert got.serialize() == want.serialize(), (name, S)
        assert np.array_equal(got.decode(), leaves)
        if S <= 12:  # bare body in HOST memory: the entry point that streams large inputs
            host = stb.SharedTree(S).build_from_body(text)
            want2 = oracle.build(leaves, S)
            assert host.layer_counts() == want2.layer_counts(), (name, S, "host body")
            assert host.serialize() == want2.serialize(), (name, S, "host body")
            assert np.array_equal(host.decode(), leaves)
    # FASTA text in HOST memory: headers, wrapped lines, blank lines, several records (streams chunk by chunk
    # with the line automaton's state carried across the chunks when the stream options say so)
    def wrap(b, w):
        return b"\n".join(b[i:i + w] for i in range(0, len(b), w))
    for S in (12, 5):
        fasta = (b">rec one\n" + wrap(corpus_text("merged"), 60) + b"\n>rec two | x\n" + wrap(corpus_text("vaccg").upper(), 71) +
                 b"\n>three\n" + wrap(corpus_text("hehcmv"), 4097) + b"\n>four\n" + corpus_text("humhbb") + b"\n")
        lv = oracle.fasta_to_leaves(fasta, S)
        got = stb.SharedTree(S).build_from_fasta(fasta)
        want = oracle.build(lv, S)
        assert got.layer_counts() == want.layer_counts(), ("fasta host", S)
        assert got.serialize() == want.serialize(), ("fasta host", S)
        assert np.array_equal(got.decode(), lv)
    n = 7_000_000
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    stb.synth_genome(buf, n, seed=11, repeat_permille=500)
    leaves = oracle.fasta_to_leaves(buf.cpu().numpy().tobytes(), 12)
    got, want = stb.SharedTree(12).build_from_body(buf), oracle.build(leaves, 12)
    assert got.serialize() == want.serialize()
    tree = stb.SharedTree(12)
    for _ in range(2):  # twice on one handle: the workspace (tables, epoch tags) is reused
        host = tree.build_from_body(buf.cpu().numpy().tobytes())
        assert host.serialize() == want.serialize(), "host body, synthetic"
    host.sort(); want.sort()
    assert host.serialize() == want.serialize()
    # one handle, host-streaming and device builds of DIFFERENT inputs interleaved: the two node
    # tables keep separate epoch counters, a reset of one must not revive tags of the other
    tree = stb.SharedTree(12)
    bufs, wants = [], []
    for seed in (21, 22):
        b = torch.empty(n, dtype=torch.uint8, device="cuda")
        stb.synth_genome(b, n, seed=seed, repeat_permille=500)
        bufs.append(b)
        wants.append(oracle.build(oracle.fasta_to_leaves(b.cpu().numpy().tobytes(), 12), 12).serialize())
    small = bufs[0][:400_000].contiguous()
    want_small = oracle.build(oracle.fasta_to_leaves(small.cpu().numpy().tobytes(), 12), 12).serialize()
    for which in (0, 0, None, 1, 0, None, 1):
        if which is None:
            assert tree.build_from_body(small).serialize() == want_small, "device build between streaming builds"
        else:
            assert tree.build_from_body(bufs[which].cpu().numpy().tobytes()).serialize() == wants[which], f"interleaved host build {which}"
    # an IUPAC symbol in the middle: the streaming path must fall back, not mis-build
    text = bytearray(buf[:1_200_000].cpu().numpy().tobytes())
    text[600_001] = ord("N")
    lv = oracle.fasta_to_leaves(bytes(text), 12)
    assert stb.SharedTree(12).build_from_body(bytes(text)).serialize() == oracle.build(lv, 12).serialize()
    # the "real genome" variant: N runs (leaves outside ACGT go to the side table next to the direct one; runs
    # of equal nodes collapse before the on-chip dedup) and soft-masked lower case, from device and host memory
    masked = buf.clone()
    pkg.synth_mask(masked, seed=11)
    mtext = masked.cpu().numpy().tobytes()
    assert mtext.count(b"N") + mtext.count(b"n") > 1000 and any(c in mtext for c in b"acgt")
    lv = oracle.fasta_to_leaves(mtext, 12)
    want = oracle.build(lv, 12)
    for entry, arg in (("device", masked), ("host", mtext)):
        got = stb.SharedTree(12).build_from_body(arg)
        assert got.layer_counts() == want.layer_counts(), ("masked", entry)
        assert got.serialize() == want.serialize(), ("masked", entry)
        assert np.array_equal(got.decode(), lv), ("masked", entry)
    # a side table that fills up (512 usable slots, thousands of distinct leaves with IUPAC codes): the hashed leaf level takes over
    tiny = stb.SharedTree(12).set_option("side_table_slots", 1)
    iupac = bytearray(mtext[:2_400_000])
    spots = np.random.default_rng(3).integers(0, len(iupac), 6000)
    for at, code in zip(spots, b"RYKMSWBDHVN" * 600):
        iupac[at] = code
    iupac = bytes(iupac)
    want_iupac = oracle.build(oracle.fasta_to_leaves(iupac, 12), 12).serialize()
    assert tiny.build_from_body(torch.frombuffer(bytearray(iupac), dtype=torch.uint8).cuda()).serialize() == want_iupac, "side table overflow"
    assert tiny.build_from_body(iupac).serialize() == want_iupac, "side table overflow, host"
    assert stb.SharedTree(12).build_from_body(iupac).serialize() == want_iupac, "IUPAC leaves in the side table, host"
    # long runs of one letter and of short periods between random stretches (every run longer than a partition tile)
    rnd = buf[:600_000].cpu().numpy().tobytes()
    runs = rnd[:240_000] + b"A" * 480_000 + rnd[240_000:360_000] + b"ACGTTGCA" * 90_000 + b"N" * 300_000 + rnd[360_000:] + b"T" * 120_007
    for S in (12, 4):
        lvS = oracle.fasta_to_leaves(runs, S)
        wantS = oracle.build(lvS, S)
        for entry, arg in (("device", torch.frombuffer(bytearray(runs), dtype=torch.uint8).cuda()), ("host", runs)):
            got = stb.SharedTree(S).build_from_body(arg)
            assert got.layer_counts() == wantS.layer_counts(), ("runs", S, entry)
            assert got.serialize() == wantS.serialize(), ("runs", S, entry)
    rng = np.random.default_rng(5)
    codes = np.array([1, 2, 4, 8, 3, 12, 7, 14, 0, 9, 5, 11, 13, 10, 6, 15], dtype=np.uint64)
    nib = codes[rng.integers(0, 16, size=(40000, 12))]
    lv = (nib << (4 * np.arange(12, dtype=np.uint64))).sum(axis=1).astype(np.uint64)
    lv[20000:30000] = lv[:10000]
    assert stb.SharedTree(12).build_from_leaves(lv).serialize() == oracle.build(lv, 12).serialize()
    print("forced_paths_check ok", flush=True)


if __name__ == "__main__":
    main()
