"""print the interesting parts of a bench.py JSON line"""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.4g  e2e %.4g  ms/step %.4f  n_gpus %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["n_gpus"]))
r = d["roofline"]
print("k_step_a %.4f ms  frac %.3f | sweep frac %.3f" % (r["avg_launch_ms"], r["frac"], r["sweep"]["frac"]))
print("kernels_ms", {k[:24]: round(v, 4) for k, v in r["sweep"]["kernels_ms"].items()}, "stream step", round(r["sweep"]["stream_form_step_ms"], 4))
print("query", round(r["codebook_query"]["ms"], 4), "frac", round(r["codebook_query"]["frac"], 3))
print("converged", d.get("converged_cloud"))
print("cpu_baseline", d.get("cpu_baseline"))
print("graph", d.get("graph"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
print("stats", d.get("engine_stats"))
print("filter", d.get("filter"))
print("gemm", d.get("codebook_gemm"), "tcn", d.get("tcn_forward_ms"))
if d.get("per_rank"):
    for k in d["per_rank"]:
        print(k)
print("sharded", d.get("sharded_check"))
