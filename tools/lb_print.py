"""Print the headline numbers of bench lines under gpurun_out/ (development helper): lb_print.py [names...]"""
import json, sys
names = sys.argv[1:] or ["lb_speed1", "lb_mixed", "lb_mixed128"]
for f in names:
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e)
        continue
    r = d["details"]["region_dedup"]
    e = d.get("e2e") or {}
    print(f, round(d["ms_per_step"], 3), "regions only", r["ms_per_step_regions_only"], "without", r["ms_per_step_without"],
          "e2e", e.get("ms_per_step"), e.get("d2h_GBps_all_gpus"), "packed", (e.get("plan_to_pcm") or {}).get("rank0_packed_ms_per_step"),
          "spans", (e.get("plan_to_pcm") or {}).get("ms_per_step"))
