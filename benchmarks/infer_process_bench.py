"""`loom_infer` (the process loom.runner starts) on the C2 files: wall clock and where it goes.  -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
model, rows = bench.load_workload("C2", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
r = bench.measure_loom_infer(model, rows)
r = bench.measure_loom_infer(model, rows)     # second run: files in the page cache, driver warm
print(json.dumps(r))
