"""Times sort_tree / decode repeatedly on one handle, wall (CUDA events) against the sum of the
kernel classes, to separate kernel time from allocation time."""
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
tree = pkg.SharedTree(12)
stream = torch.cuda.current_stream()


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record(stream)
    fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b), (time.perf_counter() - t0) * 1e3


for rep in range(4):
    tree.build_from_body(text)
    tree.profile(True)
    tree.profile_reset()
    ev, wall = timed(tree.sort)
    prof = tree.profile_read()
    tree.profile(False)
    ksum = sum(r["ms"] for r in prof.values())
    print(f"sort #{rep}: events {ev:.2f} ms, wall {wall:.2f} ms, kernels {ksum:.2f} ms:",
          ", ".join(f"{k} {v['ms']:.2f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:6]), flush=True)
n0 = tree.width()
out = torch.empty(n0 * 12, dtype=torch.uint8, device="cuda")
for rep in range(3):
    tree.profile(True)
    tree.profile_reset()
    ev, wall = timed(lambda: tree.decode_ascii(out=out))
    prof = tree.profile_read()
    tree.profile(False)
    ksum = sum(r["ms"] for r in prof.values())
    print(f"decode_ascii #{rep}: events {ev:.2f} ms, wall {wall:.2f} ms, kernels {ksum:.2f} ms:",
          ", ".join(f"{k} {v['ms']:.2f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:5]), flush=True)
print("free/total GB", [round(x / 1e9, 1) for x in torch.cuda.mem_get_info()])
