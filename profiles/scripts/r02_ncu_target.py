"""Target of the round-2 `ncu --set full` captures: everything is warmed up first, then ONE stage runs inside
cudaProfilerStart/Stop (ncu --profile-from-start off).   usage: r02_ncu_target.py build|sort|serialize|decode|deserialize [bases]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
stage = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3_100_000_000
text = torch.empty(n, dtype=torch.uint8, device="cuda")
pkg.synth_genome(text, n, seed=42, repeat_permille=500)
tree = pkg.SharedTree(12)
for _ in range(2):
    tree.build_from_body(text)
torch.cuda.synchronize()
prof = torch.cuda.profiler
if stage == "all":
    # every stage once inside one profiled region; ncu picks the first invocation of every kernel function
    # (--kernel-id ::regex:.*:1): the leaf level, the first node levels, the largest layer of sort / serialize / parse
    tree.sort()
    nb = tree.bytes()
    dag = torch.empty(nb + 64, dtype=torch.uint8, device="cuda")
    tree.serialize_into(dag)
    host = dag[:nb].cpu().numpy()
    w = tree.width()
    out = torch.empty(w * 12, dtype=torch.uint8, device="cuda")
    tree.decode_ascii(out=out)
    idx = torch.from_numpy(pkg.query_indices(42, 10_000_000, w).astype("int64")).cuda()
    got = torch.empty(10_000_000, dtype=torch.int64, device="cuda")
    tree.random_access(idx, out=got)
    back = pkg.SharedTree(12)
    back.deserialize(host)
    tree.build_from_body(text)
    torch.cuda.synchronize()
    prof.start()
    tree.build_from_body(text)
    tree.sort()
    tree.bytes()
    tree.serialize_into(dag)
    tree.decode_ascii(out=out)
    tree.random_access(idx, out=got)
    back.deserialize(host)
    torch.cuda.synchronize()
    prof.stop()
elif stage == "build":
    prof.start()
    tree.build_from_body(text)
    torch.cuda.synchronize()
    prof.stop()
elif stage == "sort":
    tree.sort()
    tree.build_from_body(text)
    prof.start()
    tree.sort()
    torch.cuda.synchronize()
    prof.stop()
else:
    tree.sort()
    nb = tree.bytes()
    dag = torch.empty(nb + 64, dtype=torch.uint8, device="cuda")
    tree.serialize_into(dag)
    if stage == "serialize":
        tree.build_from_body(text)
        tree.sort()
        prof.start()
        tree.bytes()
        tree.serialize_into(dag)
        torch.cuda.synchronize()
        prof.stop()
    elif stage == "decode":
        w = tree.width()
        out = torch.empty(w * 12, dtype=torch.uint8, device="cuda")
        tree.decode_ascii(out=out)
        idx = torch.from_numpy(pkg.query_indices(42, 10_000_000, w).astype("int64")).cuda()
        got = torch.empty(10_000_000, dtype=torch.int64, device="cuda")
        tree.random_access(idx, out=got)
        prof.start()
        tree.decode_ascii(out=out)
        tree.random_access(idx, out=got)
        torch.cuda.synchronize()
        prof.stop()
    elif stage == "deserialize":
        host = dag[:nb].cpu().numpy()
        back = pkg.SharedTree(12)
        back.deserialize(host)
        prof.start()
        back.deserialize(host)
        torch.cuda.synchronize()
        prof.stop()
print("done", stage)
