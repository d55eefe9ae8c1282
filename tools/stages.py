#!/usr/bin/env python
"""Print the per-stage table of one or more bench.py JSON lines (files given as arguments)."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    st = d.get("stages", {})
    print(f"{f}: {d['value']:.1f} {d['unit']}  {d['ms_per_step']:.4f} ms/frame  e2e {d['e2e']['value']:.1f}  | " +
          "  ".join(f"{k} {v['ms']:.4f}" for k, v in st.items()))
