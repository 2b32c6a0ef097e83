"""Print the key numbers of bench.py JSON lines: python tools/bench_brief.py file..."""
import json, sys
for f in sys.argv[1:]:
    for ln in open(f):
        ln = ln.strip()
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        k = d.get("kernel_time_share", {})
        r = d.get("roofline", {})
        print("%s: ms/step %.2f  e2e %.2f  value %.3g q/s | prof: nn %.2f upd %.2f list %.2f walk %.2f | top %s frac %.3f | launches %s rmse %.17g" % (
            f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], k.get("nn_ms", 0), k.get("update_ms", 0), k.get("list_ms", 0),
            k.get("walk_ms", 0), r.get("kernel"), r.get("frac", 0), d.get("gpu_launches"), d.get("best_rmse", 0)))
