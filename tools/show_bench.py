"""Prints the headline numbers of a bench.py JSON line read from stdin (development helper)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d["n_gpus"], "GPU", round(d["ms_per_step"], 2), "ms/step", round(d["value"] / 1e3, 2), "TFLOP/s", "checksum", d["checksum"],
      {k: round(v, 1) for k, v in d["stages_ms"].items()}, "chunk", d["config"]["level_chunk"])
