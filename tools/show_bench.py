"""Print the headline fields of bench.py JSON lines:  python tools/show_bench.py 'gpurun_out/*.json'"""
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1])):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"], 3), round(d["verify_step_us"]), round(d["draft_step_us"]), d["verify_breakdown_ms"], d["draft_breakdown_ms"])
    except Exception as e:
        print(f, "ERR", e)
