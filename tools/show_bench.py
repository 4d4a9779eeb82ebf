"""Print the main numbers of a bench.py JSON line.  usage: python tools/show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def g(o, *ks):
    for k in ks:
        o = o.get(k) if isinstance(o, dict) else None
        if o is None:
            return None
    return o
print("n_gpus", d.get("n_gpus"), "steps", d.get("steps"))
print("shuffle value %.0f (single %.0f, lanes %.0f) e2e %.0f  frac %.3f whole_step %.3f" % (
    d["value"], g(d, "single_stream", "value"), g(d, "multi_lane", "value"), g(d, "e2e", "value"), g(d, "roofline", "frac"),
    g(d, "roofline", "whole_step", "frac_of_peak")))
if "fixed" in d:
    f = d["fixed"]
    print("fixed   value %.0f (single %.0f, lanes %.0f) e2e %.0f  whole_step %.3f" % (
        f["value"], g(f, "single_stream", "value"), g(f, "multi_lane", "value"), g(f, "e2e", "value"), g(f, "roofline", "whole_step", "frac_of_peak")))
bv = d.get("batch_verify")
if bv:
    print("batch_verify corrupted %.0f/s (%.2f ms) all-valid %.0f/s (%.2f ms) ratio %.2f ok=%s" % (
        bv["value"], bv["ms_per_step"], bv["all_valid"]["value"], bv["all_valid"]["ms_per_step"], bv["corrupted_over_all_valid"], bv["decisions_match_expected"]))
if d.get("large_deck"):
    print("large deck prove %.2f ms verify %.2f ms" % (d["large_deck"]["prove_ms"], d["large_deck"]["verify_ms"]))
m = d.get("msm")
if m:
    print("msm value %.0f (single call %.0f) e2e %.0f accumulate frac %.3f whole %.3f / single %.3f" % (
        m["value"], g(m, "single_call", "value"), g(m, "e2e", "value"), g(m, "roofline", "frac"), g(m, "roofline", "whole_msm", "frac_of_peak"),
        g(m, "roofline", "whole_msm", "frac_of_peak_single_call")))
