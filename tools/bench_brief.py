# Prints the main numbers of a bench.py JSON line (development helper).
import json, sys
d = json.load(open(sys.argv[1]))
o = d["roofline"]["other_kernels"]
print("value %.2f G/s  step %.3f ms  zero %.3f inject %.3f x %.3f y %.3f  frac %.3f  e2e %.2f G/s  fp32 %.3f ms" % (
    d["value"] / 1e9, d["ms_per_step"], o["ms_zero_fill"], o["ms_minmax_inject"], o["ms_sweep_x"], o["ms_sweep_y"],
    d["roofline"]["frac"], d["e2e"]["value"] / 1e9, d["fp32_path"]["ms_per_step"]))
