"""GPU: the drop-in proof.  The reference's OWN tests/test.cpp and compress.cpp, compiled
unmodified against genome-compression_b200/host/include (oracle/Makefile target `dropin`),
run on the CUDA library; plus this repo's host driver (host/compress.cpp)."""
import hashlib
import subprocess

import numpy as np
import pytest

from conftest import ROOT, corpus_text

pytestmark = pytest.mark.gpu

REF_TESTS = ROOT / "oracle" / "_ref" / "ref_tests_on_b200"
REF_COMPRESS = ROOT / "oracle" / "_ref" / "ref_compress_on_b200"
OWN_COMPRESS = ROOT / "genome-compression_b200" / "host" / "compress_b200"


@pytest.fixture(scope="module")
def workdir(tmp_path_factory, golden):
    d = tmp_path_factory.mktemp("dropin")
    (d / "data").mkdir()
    for name in golden["corpus"]["corpus_order"] + ["merged", "edited"]:
        (d / "data" / name).write_bytes(corpus_text(name))
    return d


def run(cmd, cwd):
    return subprocess.run([str(c) for c in cmd], cwd=cwd, capture_output=True, text=True, timeout=600)


def test_reference_test_binary_passes_on_the_cuda_library(workdir):
    if not REF_TESTS.exists():
        pytest.skip("oracle/_ref/ref_tests_on_b200 was not built (needs /root/reference at build time)")
    res = run([REF_TESTS], workdir)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("Finished without errors") == 10, res.stdout + res.stderr


@pytest.mark.parametrize("name", ["humhbb", "chmpxx", "edited", "merged", "vaccg"])
def test_reference_compress_binary_writes_the_golden_dag(workdir, golden, name):
    if not REF_COMPRESS.exists():
        pytest.skip("oracle/_ref/ref_compress_on_b200 was not built (needs /root/reference at build time)")
    rec = golden["corpus"]["records"][f"{name}:12"]
    out = workdir / f"{name}.ref.dag"
    res = run([REF_COMPRESS, "--statistics", f"--output={out}", f"data/{name}"], workdir)
    assert res.returncode == 0, res.stdout + res.stderr
    cols = res.stdout.strip().split(",")
    assert int(cols[0]) == 12 and int(cols[1]) == rec["width"] and int(cols[4]) == rec["post_bytes"]
    assert hashlib.sha256(out.read_bytes()).hexdigest() == rec["post_sha256"]


def test_reference_compress_binary_error_exit(workdir):
    if not REF_COMPRESS.exists():
        pytest.skip("not built")
    bad = workdir / "bad.fa"
    bad.write_bytes(b">x\nACGTACGTACGTACGTACGTAXGT\n")
    res = run([REF_COMPRESS, "--no-save", bad], workdir)
    assert res.returncode == 1
    assert "Encountered unknown symbol: 88 (ASCII code 88)" in res.stderr


@pytest.mark.parametrize("name,S", [("humhbb", 12), ("humhbb", 5), ("humhbb", 16), ("merged", 12), ("edited", 12)])
def test_own_driver(workdir, golden, oracle, name, S):
    assert OWN_COMPRESS.exists(), "genome-compression_b200/host/compress_b200 missing: run __graft_entry__.build()"
    rec = golden["corpus"]["records"].get(f"{name}:{S}")
    out = workdir / f"{name}.{S}.dag"
    hist = workdir / f"{name}.{S}.hist.csv"
    res = run([OWN_COMPRESS, "--statistics", f"--dna-size={S}", f"--output={out}", f"--histogram={hist}", f"data/{name}"], workdir)
    assert res.returncode == 0, res.stdout + res.stderr
    text = corpus_text(name)
    leaves = oracle.fasta_to_leaves(text, S)
    want = oracle.build(leaves, S)
    # histogram CSV (src/shared_tree.cpp:332-345): built BEFORE save but AFTER sort in compress.cpp
    want.sort()
    stream = want.serialize()
    assert out.read_bytes() == stream
    if rec:
        assert hashlib.sha256(stream).hexdigest() == rec["post_sha256"]
    blocks = hist.read_text().split("\n\n")
    for k in range(want.depth() - 1):
        got = [int(x) for x in blocks[k].replace("\n", "").split(",") if x]
        assert got == sorted((int(c) for c in want.histogram(k)), reverse=True), k
    # decompress: .dag -> text
    txt = workdir / f"{name}.{S}.txt"
    res = run([OWN_COMPRESS, "--decompress", f"--dna-size={S}", f"--output={txt}", out], workdir)
    assert res.returncode == 0, res.stdout + res.stderr
    body = oracle.fasta_body(text).upper()
    assert txt.read_bytes() == body[: len(leaves) * S]
