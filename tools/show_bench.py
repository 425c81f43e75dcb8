"""Print the interesting fields of a bench.py JSON line (helper for gpurun one-liners).

    python bench.py | python tools/show_bench.py [tag]        # from a pipe
    python tools/show_bench.py gpurun_out/bench.json [tag]    # from a file

Never waits on a terminal: with no file argument and an interactive stdin it exits with a usage message (a bare call once
sat in a gpurun command line until the call's time limit ran out)."""
import json
import os
import sys

args = sys.argv[1:]
if args and os.path.isfile(args[0]):
    text, args = open(args[0]).read(), args[1:]
elif sys.stdin.isatty():
    raise SystemExit(__doc__)
else:
    text = sys.stdin.read()
tag = args[0] if args else ""
d = json.loads(text.strip().splitlines()[-1])
r = d.get("roofline") or {}
e = d.get("e2e") or {}
f = d.get("fused_step") or {}
print(tag, "value %.4g rays/s  %.3f ms/step  step_frac %s  bwd %s us (%s)  fwd %s us (%s)  e2e %.4g  e2e_deferred %s  fused_step %s/%s us  clocks %s" % (
    d["value"], d["ms_per_step"], (r.get("step") or {}).get("frac"), r.get("us_per_launch"), r.get("frac"),
    (r.get("fwd_kernel") or {}).get("us_per_launch"), (r.get("fwd_kernel") or {}).get("frac"), e.get("value", float("nan")),
    (e.get("deferred_grads") or {}).get("value"), f.get("us"), f.get("unfused_us"), (d.get("clocks") or {}).get("sm_mhz")))
