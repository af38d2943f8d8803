"""The closed_form record of bench.py alone (development: a one-minute run)."""
import json
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench

ctx = bench.Ctx()
print(json.dumps(bench.closed_form_records(ctx, int(sys.argv[1]) if len(sys.argv) > 1 else 5)))
