"""Per-level kernel classes of one single-GPU build (option profile_levels): where the 9.8 ms go."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_100_000_000
variant = sys.argv[2] if len(sys.argv) > 2 else "plain"
text = torch.empty(n, dtype=torch.uint8, device="cuda")
pkg.synth_genome(text, n, seed=42, repeat_permille=500)
if variant == "nruns":
    pkg.synth_mask(text, seed=42)
tree = pkg.SharedTree(12).set_option("profile_levels", 1)
for a in sys.argv[3:]:
    k, v = a.split("=")
    tree.set_option(k, int(v))
for _ in range(3):
    tree.build_from_body(text)
tree.profile(True)
tree.profile_reset()
reps = 5
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(reps):
    tree.build_from_body(text)
ev1.record()
torch.cuda.synchronize()
prof = tree.profile_read()
print(f"{variant} {n} bases: {ev0.elapsed_time(ev1) / reps:.3f} ms per build; layers {tree.layer_counts()[:8]}")
levels = {}
for name, rec in prof.items():
    base, _, lvl = name.partition("@L")
    levels.setdefault(int(lvl) if lvl else -1, {})[base] = rec["ms"] / reps
for lvl in sorted(levels):
    row = levels[lvl]
    print(f"L{lvl:>2}: {sum(row.values()):7.3f} ms  " + "  ".join(f"{k} {v:.3f}" for k, v in row.items()))
