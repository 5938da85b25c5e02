"""Print the headline numbers of bench.py JSON lines: python scripts/bench_value.py file [file ...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        line = [l for l in open(f) if l.startswith("{")][-1]
        d = json.loads(line)
        e2e = d.get("e2e") or {}
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(e2e.get("value", 0), 1),
              "launches", d.get("gpu_launches"), "clk", (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as ex:  # noqa
        print(f, "unreadable:", ex)
