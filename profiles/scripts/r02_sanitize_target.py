"""Target of the compute-sanitizer runs (memcheck / racecheck): every kernel family once, on inputs of a
few Mbp (the tools slow kernels down by one to two orders of magnitude).  Results are checked against the oracle."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import load_package  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

pkg, oracle = load_package(), Oracle()
n = 1_500_000
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
pkg.synth_genome(buf, n, seed=5, repeat_permille=500)
pkg.synth_mask(buf, seed=5)
text = buf.cpu().numpy().tobytes()
text = text[:700_000] + b"A" * 60_000 + text[700_000:]
leaves = oracle.fasta_to_leaves(text, 12)
want = oracle.build(leaves, 12)
pre = want.serialize()
want.sort()
post = want.serialize()
dev = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
configs = [
    {},                                                                            # default: hash-table node levels, cooperative middle
    {"bucket_min": 1, "bucket_levels": 4, "coop_max": 0},                          # on-chip dedup of four levels
    {"bucket_min": 1, "bucket_cap": 16, "bucket_levels": 2},                       # final buckets outgrow their regions: exact pass + chunked dedup
    {"bucket_min": 1, "bucket_levels": 2, "bucket_slack_permille": 0, "bucket_headroom": 0},  # first-pass overflow: table fallback
    {"side_table_slots": 1},                                                       # hashed leaf level
]
for opts in configs:
    t = pkg.SharedTree(12)
    for k, v in opts.items():
        t.set_option(k, v)
    t.build_from_body(dev)
    assert t.serialize() == pre, opts
    t.sort()
    assert t.serialize() == post, opts
# streaming host builds (bare body and FASTA text), decode, random access, deserialize on the device
t = pkg.SharedTree(12).set_option("stream_chunk_log2", 13).set_option("stream_min_chunks", 2)
assert t.build_from_body(text).serialize() == pre
fasta = b">x\n" + b"\n".join(text[i:i + 60] for i in range(0, len(text), 60)) + b"\n"
assert t.build_from_fasta(fasta).serialize() == pre
t.sort()
assert np.array_equal(t.decode(), leaves)
assert t.decode_ascii().upper() == text[: len(leaves) * 12].upper()
idx = np.random.default_rng(1).integers(0, len(leaves), 20000).astype(np.uint64)
assert np.array_equal(t.random_access(idx), leaves[idx])
back = pkg.SharedTree(12).deserialize(post)
assert back.serialize() == post and np.array_equal(back.decode(), leaves)
# two virtual ranks of the sharded build (ACGT only)
from genome_compression_b200 import shard  # noqa: E402
acgt = torch.empty(n, dtype=torch.uint8, device="cuda")
pkg.synth_genome(acgt, n, seed=6, repeat_permille=500)
atext = acgt.cpu().numpy().tobytes()
awant = oracle.build(oracle.fasta_to_leaves(atext, 12), 12).serialize()
ranks = shard.create_local(2, device=0, dna_size=12)


def work(rank):
    first, count = rank.range(n)
    rank.set_option("cut", 2048).build_from_body(atext[first:first + count], n)
    return rank.gather()


assert shard.run_local(ranks, work)[0].serialize() == awant
print("sanitize target ok")
