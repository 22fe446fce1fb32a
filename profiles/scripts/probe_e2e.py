"""Wall-clocks every call of the end-to-end compress / decompress paths (host text -> .dag bytes on the
host and back), several repetitions on one handle, to find where the time outside the kernels goes."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
n_bases = int(sys.argv[1]) if len(sys.argv) > 1 else 3_100_000_000
text = torch.empty(n_bases, dtype=torch.uint8, device="cuda")
pkg.synth_genome(text, n_bases, seed=42, repeat_permille=500)
host = torch.empty(n_bases, dtype=torch.uint8, pin_memory=True)
host.copy_(text)
torch.cuda.synchronize()
tree = pkg.SharedTree(12)


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, out


dag_host = None
for rep in range(4):
    b, _ = wall(lambda: tree.build_from_body(host))
    s, _ = wall(tree.sort)
    p, nb = wall(tree.bytes)
    if dag_host is None:
        dag_host = torch.empty(nb + 16, dtype=torch.uint8, pin_memory=True)
    w, _ = wall(lambda: tree.serialize_to_host(dag_host))
    print(f"compress #{rep}: build(host) {b:.2f}  sort {s:.2f}  bytes {p:.2f}  serialize_to_host {w:.2f}  total {b + s + p + w:.2f} ms", flush=True)
for rep in range(3):
    b, _ = wall(lambda: tree.build_from_body(text))
    s, _ = wall(tree.sort)
    print(f"device build #{rep}: build {b:.2f}  sort {s:.2f}", flush=True)
n0 = tree.width()
text_host = torch.empty(n0 * 12, dtype=torch.uint8, pin_memory=True)
back = pkg.SharedTree(12)
view = dag_host[:nb].numpy()
for rep in range(3):
    d, _ = wall(lambda: back.deserialize(view))
    a, _ = wall(lambda: back.decode_ascii_to_host(text_host))
    print(f"decompress #{rep}: deserialize(host bytes) {d:.2f}  decode_ascii_to_host {a:.2f} ms", flush=True)
print("roundtrip equal:", bool(torch.equal(text_host, host[: n0 * 12])))
print("free/total GB", [round(x / 1e9, 1) for x in torch.cuda.mem_get_info()])
