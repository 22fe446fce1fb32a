"""CPU: the C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import re

from conftest import ROOT


def declared_symbols():
    text = (ROOT / "include" / "shared_tree_b200.h").read_text() + (ROOT / "include" / "shared_tree_b200_dist.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_whole_capability_set():
    names = declared_symbols()
    for needed in ("stb_create", "stb_destroy", "stb_clone", "stb_build_from_fasta", "stb_build_from_leaves",
                   "stb_depth", "stb_width", "stb_leaf_count", "stb_node_count", "stb_layer_count", "stb_sort_tree",
                   "stb_bytes", "stb_serialize", "stb_deserialize", "stb_copy_leaves", "stb_copy_layer",
                   "stb_histogram", "stb_decode_leaves", "stb_decode_ascii", "stb_random_access",
                   "stb_status_string", "stb_last_error"):
        assert needed in names


def test_library_exports_every_declared_symbol(stb):
    lib = ctypes.CDLL(str(stb.LIB_PATH))
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_gpu_means_loud_failure_not_fallback(stb):
    import torch
    if torch.cuda.is_available():
        return  # covered by the gpu tests
    try:
        stb.SharedTree(12)
    except stb.StbError as e:
        assert e.name == "STB_ERR_CUDA"
    else:
        raise AssertionError("SharedTree() must not succeed without a CUDA device")


def test_status_strings(stb):
    assert stb.lib.stb_status_string(0) == b"ok"
    assert b"unknown" in stb.lib.stb_status_string(3)


def test_product_never_touches_the_oracle():
    pkg = ROOT / "genome-compression_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cpp")):
        if "build" in path.parts:
            continue
        text = path.read_text()
        assert "pyoracle" not in text and "liboracle" not in text and "libref" not in text, path
