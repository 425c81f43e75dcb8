"""Print the interesting fields of a bench.py JSON line read from stdin (helper for gpurun one-liners)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d.get("roofline") or {}
e = d.get("e2e") or {}
f = d.get("fused_step") or {}
print(tag, "value %.4g rays/s  %.3f ms/step  step_frac %s  bwd %s us (%s)  fwd %s us (%s)  e2e %.4g  e2e_deferred %s  fused_step %s/%s us  clocks %s" % (
    d["value"], d["ms_per_step"], (r.get("step") or {}).get("frac"), r.get("us_per_launch"), r.get("frac"),
    (r.get("fwd_kernel") or {}).get("us_per_launch"), (r.get("fwd_kernel") or {}).get("frac"), e.get("value", float("nan")),
    (e.get("deferred_grads") or {}).get("value"), f.get("us"), f.get("unfused_us"), (d.get("clocks") or {}).get("sm_mhz")))
