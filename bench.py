#!/usr/bin/env python
"""bench.py -- radial k-space samples gridded per second (BASELINE.json metric).

Workload (config.workload): BASELINE config 2, the one configuration the
reference publishes a time for -- whole-body golden-angle adjoint, input
[6 coils, 1, 512 readout, 20271 spokes, 1], flags `-u 0.4 -d 21 -a -G`
-> 956 sliding-window slices (204 spokes, slide 21) of 256x256, coil RSS.
Synthetic N(0,1) data of that shape (the reference's data file is a git-LFS
pointer).  One "step" = the whole 956-slice job.

  value  : coil-samples/s with the acquisition already resident in HBM
           (tron_recon_device, CUDA events on the launching stream)
  e2e    : the same job through the host-buffer C-ABI call (tron_recon_host):
           pinned host input -> H2D -> kernels -> D2H, wall clock
  roofline: the gridding kernel alone (tron_grid_device), algorithmic bytes
           8*nc*(nro*npe1work + nxos^2) per slice over CUDA-event time
  cpu_baseline: OpenMP C gridding operator (oracle/, "port") on the host cores,
           bounded sample of the same slices
  --impl reference: the UNMODIFIED reference (oracle/_ref, tron.cu compiled in
           place for sm_100 + cuFFT) through its own recon_radial2d on pinned
           host buffers on the same GPU; rank 0 only.

Multi-GPU (torchrun, one rank per GPU): every rank reconstructs its own
acquisition of the same shape (slices/frames are independent, no data-path
collective) -> weak scaling; value = total samples / max-over-ranks time.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (dims, flags, description)
    "cfg2": ([6, 1, 512, 20271, 1], dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=21),
             "BASELINE cfg2 whole-body golden-angle adjoint: dims [6,1,512,20271,1], -u 0.4 -d 21 -a -G, "
             "956 slices of 256x256, coil RSS"),
    "cfg3": ([32, 1, 512, 804 * 32, 1], dict(adjoint=True, golden=True, undersamp=1.5703125, prof_slide=804),
             "BASELINE cfg3 per-GPU shard: dims [32,1,512,25728,1], -a -G -u 1.5703125 -d 804, 32 slices of 256x256"),
    "cfg4": ([16, 1, 256, 128 + 21 * 249, 1], dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=21),
             "BASELINE cfg4 per-GPU shard: dims [16,1,256,5357,1], -u 0.5 -d 21 -a -G, 250 frames of 128x128"),
    # SURVEY 8(f) rows N4 / N3 on the cfg2 acquisition (not driver-run; profiles/ holds their lines)
    "cfg2_walsh": ([6, 1, 512, 20271, 1], dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=21,
                                               coil_combine=1, walsh_npatch=1),
                   "cfg2 acquisition, -a -G -u 0.4 -d 21 -w 1: adaptive (Walsh) coil combine instead of RSS"),
    "cfg2_cgnr3": ([6, 1, 512, 20271, 1], dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=21, niter=3),
                   "cfg2 acquisition, -a -G -u 0.4 -d 21 -i 3: three CGNR iterations per slice, coil RSS"),
    "small": ([6, 1, 512, 204 + 21 * 15, 1], dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=21),
              "16-slice stretch of cfg2 (smoke)"),
}
PUBLISHED_SAMPLES_PER_S = 182.7e6      # 599.1 M coil-samples / 3.28 s (hardware not stated)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 6] or \
               [r for (_, r) in self.rows if len(r) >= 6]
        if not rows:
            return None
        sm = sorted(float(r[0]) for r in rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}


def bind_to_gpu_numa(local):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are
    allocated (first touch places them on that node): with several ranks per box the host<->device copies of
    the end-to-end leg otherwise cross sockets.  Best effort: returns the node or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind_to_gpu_numa(local)
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_input(torch, n_elems, rank):
    """complex64 N(0,1) acquisition, generated on the device, returned as (device f32 tensor, pinned host copy)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(20261017 + 2 + 1000 * rank)
    d = torch.randn(n_elems * 2, dtype=torch.float32, device="cuda", generator=g)
    h = torch.empty(n_elems * 2, dtype=torch.float32, pin_memory=True)
    h.copy_(d)
    torch.cuda.synchronize()
    return d, h


def cpu_baseline(dims, flags, geom, budget_s=15.0):
    """OpenMP C gridding operator (precompensate + gridradial2d restated) on a bounded number of slices."""
    from oracle.oracle import Oracle
    from util import synth_complex
    o = Oracle()
    nc, nro, npe, n = geom["nc"], geom["nro"], geom["npe1work"], geom["nxos"]
    s = synth_complex((npe, nro, nc), stream=2)
    t0 = time.perf_counter()
    o.grid(o.precompensate(s, nc, nro, npe), n, nc, nro, npe, W=flags.get("kernwidth", 2.0), skip=0,
           golden=flags.get("golden", False))
    t1 = time.perf_counter() - t0
    reps = int(max(1, min(63, budget_s / max(t1, 1e-3) - 1)))
    t0 = time.perf_counter()
    for z in range(reps):
        o.grid(o.precompensate(s, nc, nro, npe), n, nc, nro, npe, W=flags.get("kernwidth", 2.0),
               skip=(z + 1) * geom["prof_slide"], golden=flags.get("golden", False))
    dt = time.perf_counter() - t0 + t1
    nsl = reps + 1
    return {"value": nc * nro * npe * nsl / dt, "unit": "samples/s", "cores": o.num_threads(), "kind": "port",
            "sample": "%d of %d slices, density compensation + gridding operator only (OpenMP C, oracle/), %.1f s"
                      % (nsl, geom["nz"], dt)}


def run_ours(args):
    import torch
    import tron_b200 as t
    from tron_b200 import build
    build.build()
    rank, world, local = dist_setup(args.gpus)
    dims, flags, desc = WORKLOADS[args.workload]
    cfg = t.make_config(dims, device=local, **flags)
    plan = t.Plan(cfg)
    g = plan.geom.as_dict()
    nsamp = g["nc"] * g["nro"] * g["npe1work"] * g["nz"]            # coil-samples gridded per step
    in_elems, out_elems = g["shard_in_elems"], g["shard_out_elems"]
    d_in, h_in = make_input(torch, in_elems, rank)
    d_out = torch.zeros(out_elems * 2, dtype=torch.float32, device="cuda")
    h_out = torch.zeros(out_elems * 2, dtype=torch.float32, pin_memory=True)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- device-resident throughput
    for _ in range(args.warmup):
        plan.recon_device(d_out.data_ptr(), d_in.data_ptr(), stream)
    barrier(world)
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        plan.recon_device(d_out.data_ptr(), d_in.data_ptr(), stream)
    e1.record()
    barrier(world)
    wall1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = plan.last_launches() * args.steps
    clocks = sampler.stop(wall0, wall1) if sampler else None
    ms_per_step = ms / args.steps
    value = nsamp * world / (ms_per_step * 1e-3)
    checksum = float(d_out[::4097].double().abs().sum().item())

    # ---- end to end through the host-buffer C-ABI call
    for _ in range(max(1, args.warmup // 2)):
        plan.recon_host_ptr(h_out.data_ptr(), h_in.data_ptr())
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.recon_host_ptr(h_out.data_ptr(), h_in.data_ptr())
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps, world)
    barrier(world)
    e2e_ok = bool(torch.equal(h_out[::4097], d_out[::4097].cpu()))

    # ---- roofline of the dominant kernel (gridding), timed alone on this stream
    n, nc = g["nxos"], g["nc"]
    B = min(256, g["nz"])                                          # the launch length the device pipeline uses
    d_grid = torch.empty(B * nc * n * n * 2, dtype=torch.float32, device="cuda")
    nlaunch = 0
    for z0 in range(0, min(g["nz"], 4 * B), B):                     # warm-up
        plan.grid_device(d_grid.data_ptr(), d_in.data_ptr(), z0, min(B, g["nz"] - z0), stream)
    torch.cuda.synchronize()
    e0.record()
    for z0 in range(0, g["nz"] - B + 1, B):
        plan.grid_device(d_grid.data_ptr(), d_in.data_ptr(), z0, B, stream)
        nlaunch += 1
    e1.record()
    torch.cuda.synchronize()
    grid_ms = e0.elapsed_time(e1) / max(nlaunch, 1)
    bytes_per_slice = 8 * nc * (g["nro"] * g["npe1work"] + n * n)
    achieved = bytes_per_slice * B / (grid_ms * 1e-3) / 1e9
    peak, peak_src = peak_hbm()
    traffic = None
    tj = os.path.join(ROOT, "profiles", "grid_traffic.json")
    if os.path.isfile(tj):
        try:
            per_slice = json.load(open(tj)).get(args.workload + "_bytes_per_slice")
            traffic = per_slice * B if per_slice else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "grid_gather_kernel (tron_grid_device, %d slices/launch)" % B,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_slice * B,
                "ms_per_launch": grid_ms, "share_of_step": grid_ms * (g["nz"] / B) / ms_per_step}

    out = None
    if rank == 0:
        cpu = cpu_baseline(dims, flags, g) if world == 1 and not args.no_cpu else None
        out = {"metric": "radial k-space samples gridded/sec", "value": value, "unit": "samples/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak",
               "vs_baseline": value / PUBLISHED_SAMPLES_PER_S if args.workload == "cfg2" else None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": desc, "name": args.workload, "slices_per_gpu": g["nz"],
                          "coil_samples_per_step_per_gpu": nsamp, "parallelism": "slices x%d (no collective)" % world,
                          "l2": "inputs %.0f MB + outputs %.0f MB per GPU, larger than the 126 MB L2"
                                % (in_elems * 8 / 1e6, out_elems * 8 / 1e6),
                          "vs_baseline_note": "published 3.28 s is the reference's end-to-end span on unnamed hardware"},
               "images_per_s": g["nz"] * world / (ms_per_step * 1e-3),
               "e2e": {"value": nsamp * world / e2e_s, "unit": "samples/s", "ms_per_step": e2e_s * 1e3,
                       "images_per_s": g["nz"] * world / e2e_s,
                       "h2d_bytes_per_step": in_elems * 8, "d2h_bytes_per_step": out_elems * 8,
                       "matches_device_path": e2e_ok},
               "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "checksum": checksum}
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    plan.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_cfg5(args):
    """BASELINE cfg5: fp16-storage forward + adjoint NUFFT pair, 1024 matrix, 2x grid, kernel width 6,
    64 coils sharded over the ranks; the coil root-sum-of-squares is one NCCL reduce of nx*ny floats.
    Strong scaling: the 64 coils are split over the ranks (nc_local = 64 / N)."""
    import torch
    import tron_b200 as t
    from tron_b200 import build
    build.build()
    rank, world, local = dist_setup(args.gpus)
    nc, nx = 64, 1024
    ncl = nc // world
    fw = t.Plan(t.make_config([ncl, 1, nx, nx, 1], adjoint=False, kernwidth=6.0, half_in=True, half_out=True, device=local))
    gf = fw.geom.as_dict()
    ad = t.Plan(t.make_config([ncl, 1, gf["nro"], gf["npe1work"], 1], adjoint=True, kernwidth=6.0, half_in=True,
                              sos_partial=(ncl > 1), device=local))
    ga = ad.geom.as_dict()
    gen = torch.Generator(device="cuda"); gen.manual_seed(20261017 + 5 + 1000 * rank)
    d_img = torch.randn(gf["in_elems"] * 2, device="cuda", generator=gen).to(torch.float16)
    d_smp = torch.zeros(gf["out_elems"] * 2, device="cuda", dtype=torch.float16)
    d_sos = torch.zeros(ga["nx"] * ga["ny"] * (1 if ncl > 1 else 2), device="cuda", dtype=torch.float32)
    h_img = torch.empty(gf["in_elems"] * 2, dtype=torch.float16, pin_memory=True); h_img.copy_(d_img)
    h_out = torch.empty(ga["nx"] * ga["ny"], dtype=torch.float32, pin_memory=True)
    stream = torch.cuda.current_stream().cuda_stream

    def step(from_host):
        if from_host:
            d_img.copy_(h_img, non_blocking=True)
        fw.recon_device(d_smp.data_ptr(), d_img.data_ptr(), stream)
        ad.recon_device(d_sos.data_ptr(), d_smp.data_ptr(), stream)
        if world > 1:
            import torch.distributed as dist
            dist.reduce(d_sos, dst=0, op=dist.ReduceOp.SUM)          # the one collective of the coil-sharded path
        img = torch.sqrt(d_sos) if ncl > 1 else d_sos
        if from_host:
            h_out.copy_(img[: h_out.numel()], non_blocking=True)
        return img

    for _ in range(args.warmup):
        step(False)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(False)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world) / args.steps
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps, world)
    nsamp = 2 * nc * gf["nro"] * gf["npe1work"]                     # forward + adjoint coil-samples
    if rank == 0:
        print(json.dumps({"metric": "radial k-space samples gridded/sec", "value": nsamp / (ms * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fp16 storage)",
                          "data": "synthetic",
                          "config": {"workload": "BASELINE cfg5 fp16-storage forward+adjoint pair: 1024 matrix, 2x grid, "
                                                 "kernel width 6, 64 coils coil-sharded, NCCL reduce of the partial sum of squares",
                                     "name": "cfg5", "coils_per_gpu": ncl},
                          "e2e": {"value": nsamp / e2e_s, "unit": "samples/s", "ms_per_step": e2e_s * 1e3,
                                  "h2d_bytes_per_step": gf["in_elems"] * 4, "d2h_bytes_per_step": ga["nx"] * ga["ny"] * 4},
                          "gpu_launches": (fw.last_launches() + ad.last_launches()) * args.steps}), flush=True)
    fw.close(); ad.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


def run_reference(args):
    """The unmodified reference (CUDA + cuFFT) on the same GPU, through its own recon_radial2d."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    dims, flags, desc = WORKLOADS[args.workload]
    try:
        from oracle.oracle import RefLib
        ref = RefLib()
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built: %s" % e}), flush=True)
        return None
    geom = ref.configure(dims, True, golden=flags.get("golden", False), gridos=flags.get("gridos", 2.0),
                         kernwidth=flags.get("kernwidth", 2.0), undersamp=flags.get("undersamp", 1.0),
                         prof_slide=flags.get("prof_slide", 0))
    if geom["nc"] > ref.maxchan:
        print(json.dumps({"impl": "reference", "unavailable": "workload has %d channels, the stock reference "
                          "supports %d (tron.h:51)" % (geom["nc"], ref.maxchan)}), flush=True)
        return None
    L = ref.lib
    n_in = int(np.prod(dims))
    rng = np.random.Generator(np.random.Philox(key=20261017 + 2))
    p_in = L.tronref_host_alloc(n_in * 8)
    p_out = L.tronref_host_alloc(ref.out_elems * 8)
    h_in = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_float)), shape=(n_in * 2,))
    h_in[:] = rng.standard_normal(n_in * 2, dtype=np.float32)
    times = []
    sampler = None
    wall0 = time.time()
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            sampler = ClockSampler(0)
            wall0 = time.time()
        times.append(L.tronref_recon(C.c_void_p(p_out), C.c_void_p(p_in)))
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if sampler else None
    sec = float(np.mean(times[args.warmup:]))
    nsamp = geom["nc"] * geom["nro"] * geom["npe1work"] * geom["nz"]
    value = nsamp / sec
    window_bytes = geom["nc"] * geom["nro"] * geom["npe1work"] * 8
    out = {"impl": "reference", "metric": "radial k-space samples gridded/sec", "value": value, "unit": "samples/s",
           "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc, "name": args.workload,
                      "how": "oracle/_ref/libtronref.so = /root/reference/src/tron.cu compiled in place "
                             "(nvcc -O3 --use_fast_math, sm_100, cuFFT), recon_radial2d on pinned host buffers, "
                             "wall clock incl. its per-call tron_init/tron_shutdown"},
           "images_per_s": geom["nz"] / sec,
           "cpu_baseline": {"value": value, "unit": "samples/s", "cores": 0, "kind": "reference",
                            "sample": "full workload on the GPU: the reference has no CPU path (SURVEY 8c)"},
           "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": window_bytes * geom["nz"],
                   "d2h_bytes_per_step": ref.out_elems * 8},
           "clocks": clocks}
    print(json.dumps(out), flush=True)
    L.tronref_host_free(C.c_void_p(p_in)); L.tronref_host_free(C.c_void_p(p_out))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg5":
        run_cfg5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
