#!/usr/bin/env python
"""Prints the headline numbers of a bench.py JSON line (the file may carry an NCCL banner before the line)."""
import json
import sys

raw = open(sys.argv[1]).read()
lines = [ln for ln in raw.splitlines() if ln.startswith("{")]
if not lines:
    print("no JSON line in", sys.argv[1])
    sys.exit(0)
d = json.loads(lines[0])
print("c2" if "c5" in d or "maxcut" in d["config"]["workload"] else d["config"]["workload"], "n_gpus", d["n_gpus"], "step ms",
      round(d["value"], 2), {k: round(v, 2) for k, v in d.get("phase_ms", {}).items()})
if "shard_assembly_ms" in d:
    print("  shard", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["shard_assembly_ms"].items() if k != "note"})
print("  frac", round(d["roofline"]["frac"], 3), "solve", d.get("solve_ms"), d.get("solve_iterations"), d.get("solved"),
      "e2e", round(d["e2e"]["value"], 2))
for key in ("c5", "c3"):
    b = d.get(key)
    if not b:
        continue
    if "error" in b:
        print(" ", key, "ERROR", b["error"])
        continue
    print(" ", key, "step ms", round(b["value"], 2), {k: round(v, 2) for k, v in b.get("phase_ms", {}).items()},
          "solve_ms", b.get("solve_ms"), "programs/s", b.get("programs_per_s"))
    if "shard_assembly_ms" in b:
        print("    shard", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in b["shard_assembly_ms"].items() if k != "note"})
