import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    if "name" in d:
        print("%-28s n=%8d pts=%6d step %8.3f kernel %8.3f other %7.3f resample %7.3f e2e_p50 %8.3f" % (
            d["name"], d["particles"], d["points"], d["step_ms"], d["kernel_ms"], d["step_ms"] - d["kernel_ms"], d["resample_ms"], d.get("e2e_p50_ms", 0)))
