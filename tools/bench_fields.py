"""Print the headline fields of one or more bench.py output files (last JSON line of each)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([ln for ln in open(f) if ln.startswith("{")][-1])
    except Exception as ex:  # noqa: BLE001
        print(f, "unreadable:", ex)
        continue
    e2e = d.get("e2e", {})
    print(f"{f}: value {d.get('value', 0):.1f} {d.get('unit')} | e2e {e2e.get('value', 0):.1f} ({e2e.get('ms_per_step', 0):.3f} ms/step) | "
          f"reduce {d.get('reduce', {}).get('value', 0):.0f} GB/s, reduce e2e {d.get('reduce', {}).get('e2e', {}).get('value', 0):.1f} GB/s | "
          f"roofline {d.get('roofline', {}).get('frac', 0):.3f}")
