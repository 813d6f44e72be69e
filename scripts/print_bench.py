import sys,json
d=json.loads(sys.stdin.read()); print(d["n_gpus"], round(d["value"]), d["ms_per_step"], d["step_breakdown_ms"], d["e2e"]["value"], d["clocks"])
