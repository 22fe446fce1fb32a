"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on CPU."""
from __future__ import annotations

import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_package():
    """Imports genome-compression_b200/ (the directory name is not an identifier)."""
    name = "genome_compression_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg = ROOT / "genome-compression_b200"
    spec = importlib.util.spec_from_file_location(name, pkg / "__init__.py", submodule_search_locations=[str(pkg)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def stb():
    pkg = ROOT / "genome-compression_b200"
    if not (pkg / "libshared_tree_b200.so").exists():
        spec = importlib.util.spec_from_file_location("_stb_build", pkg / "_build.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Ref, build_ref
    if build_ref() is None:
        pytest.skip("reference library not available (no /root/reference and no oracle/_ref)")
    return Ref()


def corpus_text(name: str) -> bytes:
    """Lower-case acgt text of a DNA-corpus file, rebuilt from tests/golden/corpus/*.2bit."""
    meta = json.loads((GOLD / "corpus.json").read_text())
    if name == "merged":
        return b"".join(corpus_text(n) for n in meta["corpus_order"])
    if name == "edited":
        chm = corpus_text("chmpxx")
        w = meta["edited_wrap"]
        text = b"\n".join(chm[i:i + w] for i in range(0, len(chm), w))
        return text + (b"\n" if meta["edited_trailing_newline"] else b"")
    raw = (GOLD / "corpus" / f"{name}.2bit").read_bytes()
    n = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
    packed = np.frombuffer(raw[8:], dtype=np.uint8)
    codes = np.stack([(packed >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)[:n]
    return np.frombuffer(b"acgt", dtype=np.uint8)[codes].tobytes()


@pytest.fixture(scope="session")
def golden():
    return {
        "corpus": json.loads((GOLD / "corpus.json").read_text()),
        "small": json.loads((GOLD / "small.json").read_text()),
        "fasta": json.loads((GOLD / "fasta.json").read_text()),
        "prim": np.load(GOLD / "primitives.npz"),
    }


CORPUS_CASES = [(n, 12) for n in ("chmpxx", "chntxx", "hehcmv", "humdyst", "humghcs", "humhbb", "humhdab", "humprtb",
                                  "mpomtcg", "mtpacga", "vaccg", "merged", "edited")] + \
               [("humhbb", s) for s in (1, 2, 4, 8, 11, 13, 15, 16)]
