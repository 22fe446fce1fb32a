"""GPU: the sharded build behind one C-ABI call per rank (csrc/shard.cu, include/shared_tree_b200_dist.h).
Virtual ranks (threads of this process sharing one GPU; the library's own local communicator) always run;
the NCCL runs need >= 2 GPUs.  Everything is compared with the single-GPU stream and the oracle."""
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, corpus_text

pytestmark = pytest.mark.gpu


def sharded_tree(stb, text: bytes, S: int, world: int, cut: int, options=()):
    from genome_compression_b200 import shard
    ranks = shard.create_local(world, device=0, dna_size=S)
    n_bases = len(text)
    for r in ranks:
        r.set_option("cut", cut)
        for name, value in options:
            r.set_option(name, value)

    def work(rank):
        first, count = rank.range(n_bases)
        rank.build_from_body(text[first:first + count], n_bases)
        return rank.gather()
    out = shard.run_local(ranks, work)
    totals = ranks[0].layer_totals()
    for r in ranks:
        r.close()
    return out[0], totals


@pytest.mark.parametrize("world,cut", [(1, 1 << 16), (2, 64), (4, 1), (8, 1024), (2, 1 << 30)])
def test_virtual_ranks_match_single_gpu_and_oracle(stb, oracle, world, cut):
    text = corpus_text("merged")
    leaves = oracle.fasta_to_leaves(text, 12)
    full, totals = sharded_tree(stb, text, 12, world, cut)
    single = stb.SharedTree(12).build_from_body(text)
    assert full.layer_counts() == single.layer_counts()
    assert totals[0] == single.leaf_count() and totals[1:] == single.layer_counts()[:len(totals) - 1]
    assert full.serialize() == single.serialize() == oracle.build(leaves, 12).serialize()
    full.sort()
    single.sort()
    assert full.serialize() == single.serialize()
    assert np.array_equal(full.decode(), leaves)


@pytest.mark.parametrize("name,S,world", [("humhbb", 5, 4), ("humhbb", 1, 2), ("vaccg", 12, 8), ("chmpxx", 8, 2)])
def test_virtual_ranks_other_leaf_sizes(stb, oracle, name, S, world):
    text = corpus_text(name)
    full, _ = sharded_tree(stb, text, S, world, cut=16)
    want = oracle.build(oracle.fasta_to_leaves(text, S), S)
    assert full.serialize() == want.serialize()


def test_virtual_ranks_tiny_inputs(stb, oracle):
    for n_leaves, world in ((1, 2), (3, 2), (5, 4), (17, 8)):
        text = (corpus_text("humdyst")[: n_leaves * 12 + 5])
        full, _ = sharded_tree(stb, text, 12, world, cut=1)
        assert full.serialize() == oracle.build(oracle.fasta_to_leaves(text, 12), 12).serialize(), (n_leaves, world)


@pytest.mark.parametrize("world,options", [(4, ()), (2, (("child_filter", 0),)), (8, (("bucket_cap", 256),))])
def test_virtual_ranks_synthetic(stb, world, options):
    import torch
    n_bases = 48_000_000
    buf = torch.empty(n_bases, dtype=torch.uint8, device="cuda")
    stb.synth_genome(buf, n_bases, seed=3, repeat_permille=500)
    single = stb.SharedTree(12).build_from_body(buf)
    full, totals = sharded_tree(stb, buf.cpu().numpy().tobytes(), 12, world, cut=1 << 14, options=options)
    assert full.layer_counts() == single.layer_counts()
    assert full.serialize() == single.serialize()


def test_symbols_outside_acgt_stop_every_rank(stb):
    text = bytearray(corpus_text("humhbb").upper())
    text[40000] = ord("N")
    with pytest.raises(stb.StbError) as e:
        sharded_tree(stb, bytes(text), 12, 4, cut=16)
    assert "ACGT" in str(e.value)
    text[40000] = ord("!")
    with pytest.raises(stb.StbError) as e:
        sharded_tree(stb, bytes(text), 12, 4, cut=16)
    assert e.value.name in ("STB_ERR_UNKNOWN_SYMBOL", "STB_ERR_INVALID_ARG")  # the ACGT fast path does not tell the two apart


def test_cabi_two_ranks_from_cpp(stb, tmp_path):
    """A C++ program drives two ranks through the C ABI alone (host buffers, std::thread) and compares the gathered
    stream with the single-GPU one."""
    exe = tmp_path / "shard_cabi_test"
    cmd = ["g++", "-std=c++17", "-O1", "-o", str(exe), str(ROOT / "tests" / "shard_cabi_test.cpp"), f"-I{ROOT / 'include'}",
           f"-L{ROOT / 'genome-compression_b200'}", "-lshared_tree_b200", f"-Wl,-rpath,{ROOT / 'genome-compression_b200'}", "-pthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "shard_cabi_test ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_nccl_ranks(stb):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(ROOT / "tests" / "shard_gpu_check.py"), "60000000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "shard_gpu_check ok" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
