"""Print the headline numbers of a multi-GPU bench line (development helper): multi_print.py gpurun_out/r02_bench_n8.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), d["clocks"])
e = d["e2e"]
print("  e2e", round(e["value"]), round(e["ms_per_step"], 1), "d2h GB/s", round(e["d2h_GBps_all_gpus"], 1), "threads", e.get("host_threads"))
s = d["strong"]
print("  strong", round(s["value"]), round(s["ms_per_step"], 3), [round(x, 2) for x in s["per_rank_device_ms"]])
print("  strong e2e", round(s["e2e"]["value"]), round(s["e2e"]["ms_per_step"], 1), round(s["e2e"]["d2h_GBps_all_gpus"], 1), s["gather"])
