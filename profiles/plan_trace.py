"""Stages of tron_plan_create and of the cold legacy span on cfg2 (TRON_PLAN_TRACE): python profiles/plan_trace.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TRON_PLAN_TRACE"] = "1"
import torch, tron_b200 as t
from bench import WORKLOADS
dims, flags, _ = WORKLOADS["cfg2"]
torch.cuda.init(); torch.zeros(1, device="cuda")
for rep in range(2):
    t0 = time.perf_counter(); p = t.Plan(t.make_config(dims, device=0, **flags)); t1 = time.perf_counter()
    print("plan create %.2f ms" % ((t1 - t0) * 1e3), file=sys.stderr)
    t0 = time.perf_counter(); p.close(); t1 = time.perf_counter()
    print("plan destroy %.2f ms" % ((t1 - t0) * 1e3), file=sys.stderr)
