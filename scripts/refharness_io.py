"""Glue between the committed golden rollout and the C# RefHarness (ppo-bipedalwalker_b200/csharp/RefHarness): pins the oracle
to the REAL reference arithmetic wherever a .NET SDK exists (there is none in the B200 image).

  python scripts/refharness_io.py export  [golden_actions.txt]          # tests/golden/physics_rollout.npz -> harness input
  dotnet run -c Release -p:ReferenceRoot=/path/to/PPO-BipedalWalker --project ppo-bipedalwalker_b200/csharp/RefHarness \\
         -- golden_actions.txt refharness_states.txt                      # the reference's own physics classes replay it
  python scripts/refharness_io.py compare [refharness_states.txt]        # bit-for-bit against the golden states

`compare` exits 0 only if all 90 x 8 records (92 floats each) are bit-identical; otherwise it names the first differing
env-step / walker / field with both values, which is exactly what is needed to correct SURVEY.md Appendix C's assumptions.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "physics_rollout.npz")
FIELDS = ([f"{b}.v{i}.{c}" for b, n in (("LLL", 6), ("LLU", 6), ("Body", 5), ("RLL", 6), ("RLU", 6)) for i in range(n) for c in "xy"]
          + [f"{b}.centroid.{c}" for b in ("LLL", "LLU", "Body", "RLL", "RLU") for c in "xy"]
          + [f"{b}.velocity.{c}" for b in ("LLL", "LLU", "Body", "RLL", "RLU") for c in "xy"]
          + [f"{b}.omega" for b in ("LLL", "LLU", "Body", "RLL", "RLU")] + [f"{b}.angle" for b in ("LLL", "LLU", "Body", "RLL", "RLU")]
          + [f"joint{k}.torque" for k in range(4)])
assert len(FIELDS) == 92


def export(path):
    g = np.load(GOLDEN)
    actions, done = g["actions"], g["done"]
    steps, n, _ = actions.shape
    with open(path, "w") as fh:
        fh.write(f"{n} {steps} 50\n")
        fh.write(" ".join(str(x) for x in g["floors"]) + "\n")
        for t in range(steps):
            for e in range(n):
                fh.write(" ".join(f"{w:08x}" for w in actions[t, e].view(np.uint32)) + f" {int(done[t, e])}\n")
    print(f"wrote {path}: {steps} env-steps x {n} walkers")


def parse_states(path, steps, n):
    rows = [ln.split() for ln in open(path).read().strip().splitlines()]
    if len(rows) != steps * n or any(len(r) != 92 for r in rows):
        raise SystemExit(f"{path}: expected {steps * n} lines of 92 hex words")
    return np.array([[int(w, 16) for w in r] for r in rows], np.uint32).reshape(steps, n, 92)


def compare(path):
    g = np.load(GOLDEN)
    want = np.ascontiguousarray(g["states"], np.float32).view(np.uint32)
    steps, n, _ = want.shape
    got = parse_states(path, steps, n)
    bad = np.argwhere(want != got)
    if len(bad) == 0:
        print(f"RefHarness output is bit-identical to the golden rollout ({steps} env-steps x {n} walkers x 92 floats): the oracle is pinned")
        return 0
    t, e, f = (int(x) for x in bad[0])
    print(f"{len(bad)} words differ; first at env-step {t}, walker {e} (floor {g['floors'][e]}), field {FIELDS[f]}: "
          f"reference 0x{got[t, e, f]:08x} ({got[t, e, f:f + 1].view(np.float32)[0]!r}) vs oracle 0x{want[t, e, f]:08x} "
          f"({want[t, e, f:f + 1].view(np.float32)[0]!r})")
    return 1


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else ""
    if cmd == "export":
        export(sys.argv[2] if len(sys.argv) > 2 else "golden_actions.txt")
    elif cmd == "compare":
        sys.exit(compare(sys.argv[2] if len(sys.argv) > 2 else "refharness_states.txt"))
    else:
        raise SystemExit(__doc__)
