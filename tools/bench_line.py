#!/usr/bin/env python
"""Condense a bench.py JSON line (stdin) to one readable line: tag value ms/step stage times."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print(tag, "rows/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
      {k: round(v, 2) for k, v in d.get("stage_ms_per_segment", {}).items()})
