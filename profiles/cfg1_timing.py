"""BASELINE cfg1 (Shepp-Logan 256^2 forward + adjoint, one coil, one slice) as bench.py measures it: back-to-back
launches, CUDA events.  usage: python profiles/cfg1_timing.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import tron_b200 as t
import bench
r = bench.measure_cfg1(torch, t, 0, steps=50, warmup=10)
print(json.dumps({k: r[k] for k in ("ms_per_step", "forward_ms", "adjoint_ms")}))
