#!/usr/bin/env python
"""Reads one bench.py JSON line on stdin and prints the headline fields (used in gpurun one-liners)."""
import json
import sys

d = json.loads(sys.stdin.read())
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], d.get("step_breakdown_ms"), d["e2e"]["value"], d.get("clocks"))
