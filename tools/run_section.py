#!/usr/bin/env python
"""Run one section of bench.py alone (a short target for ncu): bpr_c2 | bpr_large | eval_c2 | eval_large"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
what = sys.argv[1]
kind, shape = what.split("_")
if kind == "bpr":
    print(json.dumps(bench.run_bpr_ours(shape, epochs=2, warmup=1)))
else:
    print(json.dumps(bench.run_eval_ours(shape, reps=2)))
