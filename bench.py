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
           pinned host input -> H2D -> kernels -> D2H, wall clock; with the rel-L2
           distance of its output from the device path's and the cost of the two
           copies alone (copy_floor_ms) -- the step cannot beat its PCIe traffic
  roofline: the gridding kernel alone (tron_grid_device), algorithmic bytes
           8*nc*(nro*npe1work + nxos^2) per slice over CUDA-event time
  parity : slices of this run's output against the unmodified reference run on
           the same input windows (oracle/_ref as checker, untimed)
  e2e_cold: the reference's own span (tron.cu:726-786: init + buffers + recon +
           shutdown) through the legacy recon_radial2d symbol
  other_configs (N = 1): cfg1, cfg3 shard, cfg4 shard, cfg5 pair -- value, e2e,
           roofline and the reference's time for the same job (<= 6-coil chunks,
           sampled and scaled)
  shards (N > 1): one cfg3 shard (32 slices, 32 coils) and one cfg4 shard (250 frames, 16 coils) per GPU
  strong / coil_sharded (N > 1): ONE cfg2 acquisition sharded by slice; the cfg5
           pair with 64/N coils per GPU and the library's single ncclReduce
  cpu_baseline: OpenMP C gridding operator (oracle/, "port") on the host cores,
           bounded sample of the same slices
  --impl reference: the UNMODIFIED reference (oracle/_ref, tron.cu compiled in
           place for sm_100 + cuFFT) through its own recon_radial2d on pinned
           host buffers on the same GPU; rank 0 only.
  --lean : the headline line only (kernel experiments).

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


class _StdoutToStderr:
    """NCCL announces its version on stdout at the first communicator; stdout carries the JSON line only."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


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
        with _StdoutToStderr():                      # NCCL's version banner belongs on stderr
            dist.barrier()
            torch.cuda.synchronize()
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


def make_input(torch, n_elems, rank, pinned=True):
    """complex64 N(0,1) acquisition, generated on the device, returned as (device f32 tensor, pinned host copy)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(20261017 + 2 + 1000 * rank)
    d = torch.randn(n_elems * 2, dtype=torch.float32, device="cuda", generator=g)
    if not pinned:
        return d, None
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


def rel_l2_t(a, b):
    """relative L2 distance of two torch tensors (float64 accumulation)."""
    d = (a.double() - b.double()).norm().item()
    return d / max(b.double().norm().item(), 1e-300)


def parity_vs_reference(torch, d_out, d_in, dims, flags, g, slices):
    """Checker, outside every timed region: slices of THIS run's output against the unmodified reference
    (oracle/_ref) run on the same windows of the same input.  Slice z of a sliding-window job is the
    reference's single-slice job on spokes [z*slide, z*slide + npe1work) with skip_angles = z*slide
    (absolute golden-angle index, tron.cu:630,738)."""
    nc, nro, win, slide = g["nc"], g["nro"], g["npe1work"], g["prof_slide"]
    against = "oracle/_ref (unmodified tron.cu + cuFFT on this GPU), same input windows"
    try:
        from oracle.oracle import RefLib
        ref = RefLib()
        if nc > ref.maxchan:           # tron.h:51 MAXCHAN 6: the same sources compiled with -DMAXCHAN=64 (oracle/build.py)
            ref = RefLib(widened=True)
            against = "oracle/_ref built with -DMAXCHAN=64 (tron.h:51; otherwise unmodified tron.cu + cuFFT), same input windows"
    except Exception as e:
        return {"unavailable": "oracle/_ref not built: %s" % e}
    if nc > ref.maxchan:
        return {"unavailable": "nc = %d > MAXCHAN of the reference builds" % nc}
    npix = g["nx"] * g["ny"]
    worst, per = 0.0, {}
    for z in slices:
        w = d_in[2 * nc * nro * slide * z: 2 * nc * nro * (slide * z + win)].cpu().numpy().view(np.complex64)
        ref.configure([nc, 1, nro, win, 1], True, golden=flags.get("golden", False), gridos=flags.get("gridos", 2.0),
                      kernwidth=flags.get("kernwidth", 2.0), undersamp=float(win + 0.5) / nro, prof_slide=0,
                      skip_angles=flags.get("skip_angles", 0) + z * slide)
        assert ref.geom["npe1work"] == win and ref.geom["nz"] == 1
        want = torch.from_numpy(ref.recon(w).view(np.float32))
        got = d_out[2 * npix * z: 2 * npix * (z + 1)].cpu()
        per[str(z)] = rel_l2_t(got, want)
        worst = max(worst, per[str(z)])
    return {"rel_l2": worst, "tolerance": 1e-5, "ok": bool(worst <= 1e-5), "slices": per, "against": against}


def copy_floor_ms(torch, h_in, d_in, h_out, d_out, reps, world):
    """What the two PCIe copies of the end-to-end step cost on their own: the step's H2D and D2H bytes from the
    same pinned buffers, concurrently on two streams, wall clock, max over ranks."""
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    def once():
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)
    once()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    return max_over_ranks((time.perf_counter() - t0) / reps, world) * 1e3


def grid_roofline(torch, plan, g, d_in, half_in, B, kernel_name, workload):
    """The gridding kernel alone (tron_grid_device), CUDA events on the launching stream, against the HBM peak."""
    n, nc = g["nxos"], g["nc"]
    stream = torch.cuda.current_stream().cuda_stream
    d_grid = torch.empty(B * nc * n * n * 2, dtype=torch.float32, device="cuda")
    starts = list(range(0, g["nz"] - B + 1, B)) or [0]
    for z0 in starts[:4]:
        plan.grid_device(d_grid.data_ptr(), d_in.data_ptr(), z0, B, stream)
    torch.cuda.synchronize()
    reps = max(1, 8 // len(starts))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for z0 in starts:
            plan.grid_device(d_grid.data_ptr(), d_in.data_ptr(), z0, B, stream)
    e1.record()
    torch.cuda.synchronize()
    grid_ms = e0.elapsed_time(e1) / (reps * len(starts))
    del d_grid
    bytes_per_slice = (4 if half_in else 8) * nc * g["nro"] * g["npe1work"] + 8 * nc * n * n
    achieved = bytes_per_slice * B / (grid_ms * 1e-3) / 1e9
    peak, peak_src = peak_hbm()
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, "profiles", "grid_traffic.json")
    if os.path.isfile(tj):
        try:
            j = json.load(open(tj))
            per_slice = j.get(workload + "_bytes_per_slice")
            traffic = per_slice * B if per_slice else None
            traffic_src = j.get(workload + "_source") if per_slice else None
        except Exception:
            traffic = None
    return {"bound": "hbm", "kernel": "%s (tron_grid_device, %d slices/launch)" % (kernel_name, B),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_per_slice * B, "ms_per_launch": grid_ms}, grid_ms


def time_device(torch, fn, steps, warmup, world=1):
    for _ in range(warmup):
        fn()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world) / steps


def time_wall(torch, fn, steps, warmup, world=1):
    for _ in range(warmup):
        fn()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    dt = max_over_ranks((time.perf_counter() - t0) / steps, world)
    barrier(world)
    return dt * 1e3


def reference_chunks_ms(dims, flags, slices_timed, slices_total, adjoint=True):
    """the stock reference holds at most MAXCHAN = 6 channels (tron.h:51), so a many-coil job is
    timed in <= 6-coil chunks (its recon_radial2d span on pinned host buffers, per-call init included) on
    `slices_timed` slices and scaled to `slices_total`.  Returns (ms, description) or (None, why)."""
    try:
        from oracle.oracle import RefLib
        ref = RefLib()
    except Exception as e:
        return None, "oracle/_ref not built: %s" % e
    nc = dims[0]
    rng = np.random.Generator(np.random.Philox(key=20261017 + 77))
    total, chunks = 0.0, []
    c0 = 0
    while c0 < nc:
        c = min(ref.maxchan if adjoint else nc, nc - c0)      # degridradial2d has no channel limit (tron.cu:540-577)
        d = [c] + list(dims[1:])
        h = rng.standard_normal(int(np.prod(d)) * 2, dtype=np.float32).view(np.complex64)
        ref.configure(d, adjoint, golden=flags.get("golden", False), gridos=flags.get("gridos", 2.0),
                      kernwidth=flags.get("kernwidth", 2.0), undersamp=flags.get("undersamp", 1.0),
                      prof_slide=flags.get("prof_slide", 0))
        ref.recon(h, return_seconds=True)                      # warm-up (cuFFT plan caches, clocks)
        _, sec = ref.recon(h, return_seconds=True)
        total += sec
        chunks.append(c)
        c0 += c
    scale = slices_total / float(slices_timed)
    return total * 1e3 * scale, "oracle/_ref in coil chunks %s on %d of %d slices, scaled x%.1f" % (
        chunks, slices_timed, slices_total, scale)


def measure_adjoint_config(torch, t, name, local, steps=3, warmup=2, with_reference=True, ref_slices=2):
    """One of the other BASELINE shapes on this GPU: device-resident value, host-buffer e2e, gridding roofline and
    the reference's time for the same job (chunked, sampled)."""
    dims, flags, desc = WORKLOADS[name]
    plan = t.Plan(t.make_config(dims, device=local, **flags))
    g = plan.geom.as_dict()
    nsamp = g["nc"] * g["nro"] * g["npe1work"] * g["nz"]
    in_elems, out_elems = g["shard_in_elems"], g["shard_out_elems"]
    d_in, h_in = make_input(torch, in_elems, 7)
    d_out = torch.zeros(out_elems * 2, dtype=torch.float32, device="cuda")
    h_out = torch.zeros(out_elems * 2, dtype=torch.float32, pin_memory=True)
    stream = torch.cuda.current_stream().cuda_stream
    ms = time_device(torch, lambda: plan.recon_device(d_out.data_ptr(), d_in.data_ptr(), stream), steps, warmup)
    launches = plan.last_launches()
    e2e_ms = time_wall(torch, lambda: plan.recon_host_ptr(h_out.data_ptr(), h_in.data_ptr()), steps, 1)
    B = min(plan.batch_slices(), g["nz"])
    kern = "grid_scatter_kernel" if g["nc"] in (2, 4, 6, 16) else ("grid_wide_kernel" if g["nc"] >= 32 else "grid_gather_kernel")
    roof, grid_ms = grid_roofline(torch, plan, g, d_in, False, B, kern, name)
    roof["share_of_step"] = grid_ms * (g["nz"] / B) / ms
    out = {"workload": desc, "value": nsamp / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
           "images_per_s": g["nz"] / (ms * 1e-3), "gpu_launches_per_step": launches,
           "e2e": {"value": nsamp / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": in_elems * 8, "d2h_bytes_per_step": out_elems * 8,
                   "rel_l2_vs_device": rel_l2_t(h_out, d_out.cpu())},
           "roofline": roof}
    try:                                        # checker, untimed: first and last slice of the device-resident output
        plan.recon_device(d_out.data_ptr(), d_in.data_ptr(), stream)
        torch.cuda.synchronize()
        out["parity"] = parity_vs_reference(torch, d_out, d_in, dims, flags, g, sorted({0, g["nz"] - 1}))
    except Exception as e:
        out["parity"] = {"error": "%s: %s" % (type(e).__name__, e)}
    plan.close()
    del d_in, d_out, h_in, h_out
    torch.cuda.empty_cache()
    if with_reference:
        k = min(ref_slices, g["nz"])
        rd = list(dims)
        rd[3] = g["npe1work"] + g["prof_slide"] * (k - 1)
        ref_ms, how = reference_chunks_ms(rd, flags, k, g["nz"])
        out["reference_ms"] = ref_ms
        out["reference_how"] = how
        if ref_ms:
            out["speedup_vs_reference_e2e"] = ref_ms / e2e_ms
    return out


def shard_leg(torch, t, name, rank, world, local, steps=5, warmup=3):
    """One per-GPU shard of BASELINE cfg3 / cfg4 on every rank at once (device resident; the N = 1 line carries the
    end-to-end leg and the reference's time for the same shard)."""
    dims, flags, desc = WORKLOADS[name]
    plan = t.Plan(t.make_config(dims, device=local, **flags))
    g = plan.geom.as_dict()
    nsamp = g["nc"] * g["nro"] * g["npe1work"] * g["nz"]
    d_in, _ = make_input(torch, g["shard_in_elems"], 7 + rank, pinned=False)
    d_out = torch.zeros(g["shard_out_elems"] * 2, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ms = time_device(torch, lambda: plan.recon_device(d_out.data_ptr(), d_in.data_ptr(), stream), steps, warmup, world)
    B = min(plan.batch_slices(), g["nz"])
    kern = "grid_scatter_kernel" if g["nc"] in (2, 4, 6, 16) else ("grid_wide_kernel" if g["nc"] >= 32 else "grid_gather_kernel")
    roof, grid_ms = grid_roofline(torch, plan, g, d_in, False, B, kern, name)
    plan.close()
    del d_in, d_out
    return {"workload": desc + " -- one shard per GPU", "scaling": "weak", "value": nsamp * world / (ms * 1e-3),
            "unit": "samples/s", "ms_per_step": ms, "images_per_s": g["nz"] * world / (ms * 1e-3),
            "roofline_rank0": {k: roof[k] for k in ("kernel", "achieved", "peak", "unit", "frac", "ms_per_launch")}}


def measure_cfg1(torch, t, local, steps=20, warmup=5):
    """BASELINE cfg1: Shepp-Logan 256^2 forward (RUNME1 flags = defaults) then adjoint of the result, one GPU.
    One slice: launch-latency bound (the HBM time of the pair is ~1.3 us)."""
    from util import shepp_logan
    ph = np.ascontiguousarray(shepp_logan(256).ravel())
    fw = t.Plan(t.make_config([1, 1, 256, 256, 1], adjoint=False, device=local))
    gf = fw.geom.as_dict()
    ad = t.Plan(t.make_config([1, 1, gf["nro"], gf["npe1work"], 1], adjoint=True, device=local))
    ga = ad.geom.as_dict()
    d_img = torch.from_numpy(ph.view(np.float32).copy()).cuda()
    d_smp = torch.zeros(gf["out_elems"] * 2, dtype=torch.float32, device="cuda")
    d_rec = torch.zeros(ga["out_elems"] * 2, dtype=torch.float32, device="cuda")
    h_img = torch.from_numpy(ph.view(np.float32).copy()).pin_memory()
    h_smp = torch.zeros(gf["out_elems"] * 2, dtype=torch.float32, pin_memory=True)
    h_rec = torch.zeros(ga["out_elems"] * 2, dtype=torch.float32, pin_memory=True)
    stream = torch.cuda.current_stream().cuda_stream

    def pair():
        fw.recon_device(d_smp.data_ptr(), d_img.data_ptr(), stream)
        ad.recon_device(d_rec.data_ptr(), d_smp.data_ptr(), stream)

    def pair_host():
        fw.recon_host_ptr(h_smp.data_ptr(), h_img.data_ptr())
        ad.recon_host_ptr(h_rec.data_ptr(), h_smp.data_ptr())

    ms = time_device(torch, pair, steps, warmup)
    ms_fw = time_device(torch, lambda: fw.recon_device(d_smp.data_ptr(), d_img.data_ptr(), stream), steps, 2)
    e2e_ms = time_wall(torch, pair_host, steps, 2)
    nsamp = 2 * gf["nro"] * gf["npe1work"]
    peak, peak_src = peak_hbm()
    n = gf["nxos"]
    b_interp = 8 * (gf["nro"] * gf["npe1work"] + n * n)            # SURVEY 8d, one channel, either direction
    out = {"workload": "BASELINE cfg1: Shepp-Logan 256^2 forward (defaults, linear angles) + adjoint of the result, 1 coil",
           "value": nsamp / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "forward_ms": ms_fw,
           "adjoint_ms": ms - ms_fw, "gpu_launches_per_step": fw.last_launches() + ad.last_launches(),
           "e2e": {"value": nsamp / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": int(ph.nbytes + gf["out_elems"] * 8),
                   "d2h_bytes_per_step": int(gf["out_elems"] * 8 + ga["out_elems"] * 8)},
           "roofline": {"bound": "hbm", "kernel": "forward + adjoint pair, all six launches (one slice: latency bound)",
                        "achieved": (2 * b_interp + 4 * 8 * n * n) / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": (2 * b_interp + 4 * 8 * n * n) / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": 2 * b_interp + 4 * 8 * n * n}}
    try:
        from oracle.oracle import RefLib
        ref = RefLib()
        ref.configure([1, 1, 256, 256, 1], False)
        ref.recon(ph)
        want_s, s1 = ref.recon(ph, return_seconds=True)
        ref.configure([1, 1, 512, 512, 1], True)
        ref.recon(want_s)
        want_i, s2 = ref.recon(want_s, return_seconds=True)
        out["reference_ms"] = (s1 + s2) * 1e3
        out["reference_how"] = "oracle/_ref, recon_radial2d span of each direction on pinned host buffers"
        out["speedup_vs_reference_e2e"] = out["reference_ms"] / e2e_ms
        out["parity"] = {"forward_rel_l2": rel_l2_t(h_smp, torch.from_numpy(want_s.view(np.float32))),
                         "pair_rel_l2": rel_l2_t(h_rec, torch.from_numpy(want_i.view(np.float32)))}
    except Exception as e:
        out["reference_ms"] = None
        out["reference_how"] = "unavailable: %s" % e
    fw.close(); ad.close()
    return out


class Cfg5Pair:
    """BASELINE cfg5: fp16-storage forward + adjoint pair, 1024 matrix, 2x grid, kernel width 6, 64 coils split
    over `world` ranks; the adjoint ends in a partial sum of squares per rank and ONE ncclReduce (library call,
    tron_coil_reduce) + sqrt on rank 0."""

    def __init__(self, torch, t, rank, world, local, comm):
        self.torch, self.t, self.rank, self.world, self.comm = torch, t, rank, world, comm
        nc, nx = 64, 1024
        self.nc, self.ncl = nc, nc // world
        ncl = self.ncl
        self.fw = t.Plan(t.make_config([ncl, 1, nx, nx, 1], adjoint=False, kernwidth=6.0, half_in=True, half_out=True, device=local))
        self.gf = gf = self.fw.geom.as_dict()
        self.ad = t.Plan(t.make_config([ncl, 1, gf["nro"], gf["npe1work"], 1], adjoint=True, kernwidth=6.0, half_in=True,
                                       sos_partial=True, device=local))
        self.ga = ga = self.ad.geom.as_dict()
        gen = torch.Generator(device="cuda"); gen.manual_seed(20261017 + 5 + 1000 * rank)
        self.npix = ga["nx"] * ga["ny"]
        self.d_img = torch.randn(gf["in_elems"] * 2, device="cuda", generator=gen).to(torch.float16)
        self.d_smp = torch.zeros(gf["out_elems"] * 2, device="cuda", dtype=torch.float16)
        self.d_sos = torch.zeros(self.npix, device="cuda", dtype=torch.float32)
        self.d_out = torch.zeros(self.npix * 2, device="cuda", dtype=torch.float32)
        self.h_img = torch.empty(gf["in_elems"] * 2, dtype=torch.float16, pin_memory=True); self.h_img.copy_(self.d_img)
        self.h_out = torch.empty(self.npix * 2, dtype=torch.float32, pin_memory=True)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.nsamp = 2 * nc * gf["nro"] * gf["npe1work"]                # forward + adjoint coil-samples, all ranks

    def step(self, from_host=False, reduce=True):
        if from_host:
            self.d_img.copy_(self.h_img, non_blocking=True)
        self.fw.recon_device(self.d_smp.data_ptr(), self.d_img.data_ptr(), self.stream)
        self.ad.recon_device(self.d_sos.data_ptr(), self.d_smp.data_ptr(), self.stream)
        if reduce:
            self.comm.coil_reduce(self.d_out.data_ptr(), self.d_sos.data_ptr(), self.npix, root=0, stream=self.stream)
        if from_host and self.rank == 0:
            self.h_out.copy_(self.d_out, non_blocking=True)

    def reduce_only(self):
        self.comm.coil_reduce(self.d_out.data_ptr(), self.d_sos.data_ptr(), self.npix, root=0, stream=self.stream)

    def measure(self, steps, warmup):
        torch, world = self.torch, self.world
        ms = time_device(torch, lambda: self.step(False), steps, warmup, world)
        ms_fw = time_device(torch, lambda: self.fw.recon_device(self.d_smp.data_ptr(), self.d_img.data_ptr(), self.stream), steps, 1, world)
        ms_red = time_device(torch, self.reduce_only, max(steps, 5), 2, world)
        e2e_ms = time_wall(torch, lambda: self.step(True), steps, 1, world)
        gf, ga, ncl = self.gf, self.ga, self.ncl
        n = gf["nxos"]
        # interpolation bytes of either direction with fp16 samples, f32 grid, per rank
        b_dir = 4 * ncl * gf["nro"] * gf["npe1work"] + 8 * ncl * n * n
        peak, peak_src = peak_hbm()
        return {"value": self.nsamp / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "forward_ms": ms_fw,
                "adjoint_plus_reduce_ms": ms - ms_fw, "reduce_ms": ms_red, "reduce_share_of_step": ms_red / ms,
                "reduce_bytes": self.npix * 4, "coils_per_gpu": ncl, "n_gpus": world,
                "gpu_launches_per_step": self.fw.last_launches() + self.ad.last_launches() + 1 + (1 if world > 1 else 0),
                "e2e": {"value": self.nsamp / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": gf["in_elems"] * 4, "d2h_bytes_per_step": self.npix * 8},
                "roofline": self.fp32_roofline(ms, b_dir, peak, peak_src)}

    def fp32_roofline(self, ms, b_dir, hbm_peak, hbm_src):
        """cfg5 is FP32 bound (36 flop/B at W = 6).  Useful arithmetic: one complex x real FMA = 4 flop per
        tap and channel; the forward transform has (2W)^2 = 144 taps per sample, the adjoint 0.943 of that (the
        reference's annulus drops the corners of the square support).  Peak = the FFMA2 rate measured on this pool's
        B200 by profiles/ffma2_bench.cu (profiles/fp32_peak.json), else 148 SMs x 128 lanes x 2 x 1.965 GHz."""
        gf, ncl = self.gf, self.ncl
        nsamp = gf["nro"] * gf["npe1work"]
        taps = 144.0
        flops = 4.0 * ncl * nsamp * taps * (1.0 + 0.943)
        peak, src = 148 * 128 * 2 * 1.965e9 / 1e12, "nominal (148 SMs x 128 FP32 lanes x 2 x 1.965 GHz)"
        pj = os.path.join(ROOT, "profiles", "fp32_peak.json")
        if os.path.isfile(pj):
            try:
                j = json.load(open(pj))
                peak, src = float(j["ffma2_tflops"]), j.get("source", "profiles/fp32_peak.json")
            except Exception:
                pass
        ach = flops / (ms * 1e-3) / 1e12
        return {"bound": "fp32", "kernel": "degrid_wide + grid_wide: useful interpolation flops of both directions over the WHOLE pair's time (FFT passes included)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": src,
                "algorithmic_flops_per_step": flops,
                "hbm": {"achieved": 2 * b_dir / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": 2 * b_dir / (ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "algorithmic_bytes_per_step": 2 * b_dir}}

    def image(self):
        self.step(False)
        self.torch.cuda.synchronize()
        return self.d_out.clone()

    def close(self):
        self.fw.close(); self.ad.close()


def make_comm(torch, t, rank, world, local):
    """The library's NCCL communicator; torch.distributed only carries the 128-byte id (plumbing)."""
    with _StdoutToStderr():
        if world == 1:
            return t.Comm(t.comm_unique_id(), 0, 1, local)
        import torch.distributed as dist
        buf = torch.zeros(t.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(t.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, src=0)
        c = t.Comm(bytes(buf.cpu().numpy().tobytes()), rank, world, local)
        torch.cuda.synchronize()
        return c


def cfg5_pipe_utilisation():
    """FP32-pipe utilisation of the two cfg5 interpolation kernels, from the committed ncu capture."""
    pj = os.path.join(ROOT, "profiles", "cfg5_pipe_util.json")
    if os.path.isfile(pj):
        try:
            return json.load(open(pj))
        except Exception:
            return None
    return None


def strong_leg(torch, t, args, rank, world, local):
    """ONE cfg2 acquisition, its 956 slices split over the ranks with slice_begin/slice_end: each rank
    holds its window of the spokes (halo included) and writes its slab; no collective."""
    dims, flags, desc = WORKLOADS["cfg2"]
    nz = t.geometry(t.make_config(dims, **flags)).nz
    lo, hi = t.shard_slices(nz, rank, world)
    plan = t.Plan(t.make_config(dims, device=local, slices=(lo, hi), **flags))
    g = plan.geom.as_dict()
    d_in, h_in = make_input(torch, g["shard_in_elems"], 0)              # every rank: its stretch of a same-seeded stream
    d_out = torch.zeros(g["shard_out_elems"] * 2, dtype=torch.float32, device="cuda")
    h_out = torch.zeros(g["shard_out_elems"] * 2, dtype=torch.float32, pin_memory=True)
    stream = torch.cuda.current_stream().cuda_stream
    ms = time_device(torch, lambda: plan.recon_device(d_out.data_ptr(), d_in.data_ptr(), stream), args.steps, args.warmup, world)
    e2e_ms = time_wall(torch, lambda: plan.recon_host_ptr(h_out.data_ptr(), h_in.data_ptr()), args.steps, 1, world)
    nsamp = g["nc"] * g["nro"] * g["npe1work"] * nz
    plan.close()
    return {"workload": "ONE cfg2 acquisition (956 slices) sharded by slice over %d GPUs, no collective" % world,
            "scaling": "strong", "value": nsamp / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
            "slices_per_gpu": hi - lo,
            "e2e": {"value": nsamp / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": g["shard_in_elems"] * 8, "d2h_bytes_per_step": g["shard_out_elems"] * 8}}


def run_ours(args):
    import torch
    import tron_b200 as t
    from tron_b200 import build
    build.build()
    rank, world, local = dist_setup(args.gpus)
    dims, flags, desc = WORKLOADS[args.workload]
    cfg = t.make_config(dims, device=local, **flags)
    g = t.geometry(cfg).as_dict()                                   # host-only (tron_geometry_compute)
    nsamp = g["nc"] * g["nro"] * g["npe1work"] * g["nz"]            # coil-samples gridded per step
    in_elems, out_elems = g["shard_in_elems"], g["shard_out_elems"]
    d_in, h_in = make_input(torch, in_elems, rank)
    h_out = torch.zeros(out_elems * 2, dtype=torch.float32, pin_memory=True)

    # ---- cold span, measured FIRST: what the reference's recon_radial2d brackets (tron.cu:726-786: init, buffers,
    # recon, shutdown) through the same-named legacy symbol = plan create + recon + destroy per call.  First because
    # the span is mostly cudaMalloc / cudaFree, whose cost grows with what the process already holds (with this
    # bench's 19 GB of work buffers and torch's cache alive: 50 instead of 23 ms) and with the reference's library
    # mapped by the parity checker further down (its own CUDA runtime, cuFFT, cuBLAS: ~5x, profiles/r02_cold_span.txt)
    cold_leg = None
    if world == 1 and not args.lean and args.workload == "cfg2":
        L = t.load_library()
        assert L.tron_set_config(C.byref(cfg)) == 0
        cold = []
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            L.recon_radial2d(C.c_void_p(h_out.data_ptr()), C.c_void_p(h_in.data_ptr()))
            cold.append((time.perf_counter() - t0) * 1e3)
        med = float(np.median(cold[1:]))
        cold_leg = {"ms_per_step": med, "value": nsamp / (med * 1e-3), "unit": "samples/s", "runs_ms": cold,
                    "how": "legacy recon_radial2d(h_out, h_in): tron_plan_create + tron_recon_host + tron_plan_destroy per "
                           "call, CUDA context already up (as in the reference arm); median of the last 3 of 4 calls (the "
                           "first loads the kernels, like the reference arm's warm-up step)"}

    t_plan0 = time.perf_counter()
    plan = t.Plan(cfg)
    plan_create_ms = (time.perf_counter() - t_plan0) * 1e3
    g = plan.geom.as_dict()
    d_out = torch.zeros(out_elems * 2, dtype=torch.float32, device="cuda")
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
    e2e_s = time_wall(torch, lambda: plan.recon_host_ptr(h_out.data_ptr(), h_in.data_ptr()),
                      args.steps, max(1, args.warmup // 2), world) * 1e-3
    e2e_rel = rel_l2_t(h_out, d_out.cpu())
    nz_local = g["slice_end"] - g["slice_begin"]
    per_slice_rel = float(((h_out.view(nz_local, -1).double() - d_out.cpu().view(nz_local, -1).double()).norm(dim=1)
                           / d_out.cpu().view(nz_local, -1).double().norm(dim=1).clamp_min(1e-300)).max().item())
    floor_ms = copy_floor_ms(torch, h_in, d_in, h_out, d_out, args.steps, world)
    plan.recon_device(d_out.data_ptr(), d_in.data_ptr(), stream)     # (the floor probe overwrote d_in/h_out with themselves: no-op)
    torch.cuda.synchronize()

    # ---- roofline of the dominant kernel (gridding), timed alone on this stream
    B = min(plan.batch_slices(), g["nz"])                          # the launch length the device pipeline uses
    kern = "grid_scatter_kernel" if g["nc"] in (2, 4, 6, 16) else ("grid_wide_kernel" if g["nc"] >= 32 else "grid_gather_kernel")
    roofline, grid_ms = grid_roofline(torch, plan, g, d_in, False, B, kern, args.workload)
    roofline["share_of_step"] = grid_ms * (g["nz"] / B) / ms_per_step

    extras = {}
    if not args.lean and args.workload == "cfg2":
        if world == 1:
            if cold_leg is not None:
                cold_leg["plan_create_ms_first"] = plan_create_ms
                extras["e2e_cold"] = cold_leg
            # parity of THIS run's output against the unmodified reference (checker, untimed)
            extras["parity"] = parity_vs_reference(torch, d_out, d_in, dims, flags, g, [0, 1, 31, 32, 477, 954, 955])
        if world == 1:
            # the same job with fp16 storage at the boundary (north-star item 4): complex-half samples in,
            # complex-half images out, f32 arithmetic -- half the PCIe bytes in both directions
            try:
                p16 = t.Plan(t.make_config(dims, device=local, half_in=True, half_out=True, **flags))
                h_in16 = torch.empty(in_elems * 2, dtype=torch.float16, pin_memory=True)
                h_in16.copy_(h_in)
                h_out16 = torch.zeros(out_elems * 2, dtype=torch.float16, pin_memory=True)
                ms16 = time_wall(torch, lambda: p16.recon_host_ptr(h_out16.data_ptr(), h_in16.data_ptr()), args.steps, 2)
                extras["e2e_fp16_storage"] = {
                    "value": nsamp / (ms16 * 1e-3), "unit": "samples/s", "ms_per_step": ms16,
                    "h2d_bytes_per_step": in_elems * 4, "d2h_bytes_per_step": out_elems * 4,
                    "rel_l2_vs_f32_path": rel_l2_t(h_out16.float(), h_out), "tolerance": 2e-3,
                    "how": "tron_recon_host with half_in + half_out (complex-half RA payloads), f32 arithmetic"}
                p16.close()
                del h_in16, h_out16
            except Exception as e:
                extras["e2e_fp16_storage"] = {"error": "%s: %s" % (type(e).__name__, e)}
        plan.close(); plan = None
        del d_in, d_out, h_in, h_out
        torch.cuda.empty_cache()
        if world == 1:
            other = {}
            for name, fn in (("cfg1", lambda: measure_cfg1(torch, t, local)),
                             ("cfg3", lambda: measure_adjoint_config(torch, t, "cfg3", local, ref_slices=2)),
                             ("cfg4", lambda: measure_adjoint_config(torch, t, "cfg4", local, ref_slices=16))):
                try:
                    other[name] = fn()
                except Exception as e:           # an extra must never cost the headline line
                    other[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.empty_cache()
            try:
                comm = make_comm(torch, t, rank, world, local)
                pair = Cfg5Pair(torch, t, rank, world, local, comm)
                c5 = pair.measure(3, 2)
                c5["workload"] = "BASELINE cfg5: fp16-storage forward+adjoint pair, 1024 matrix, 2x grid, kernel width 6, 64 coils on one GPU"
                c5["fp32_pipe"] = cfg5_pipe_utilisation()
                pair.close(); comm.close()
                fms, fhow = reference_chunks_ms([6, 1, 1024, 1024, 1], dict(kernwidth=6.0), 6, 64, adjoint=False)
                ams, ahow = reference_chunks_ms([6, 1, 2048, 2048, 1], dict(kernwidth=6.0), 6, 64, adjoint=True)
                if fms and ams:
                    c5["reference_ms"] = fms + ams
                    c5["reference_how"] = "one 6-coil chunk per direction through oracle/_ref (f32, its recon_radial2d span), scaled x64/6"
                    c5["speedup_vs_reference_e2e"] = (fms + ams) / c5["e2e"]["ms_per_step"]
                other["cfg5"] = c5
            except Exception as e:
                other["cfg5"] = {"error": "%s: %s" % (type(e).__name__, e)}
            extras["other_configs"] = other
        else:
            try:
                extras["strong"] = strong_leg(torch, t, args, rank, world, local)
            except Exception as e:
                extras["strong"] = {"error": "%s: %s" % (type(e).__name__, e)}
            torch.cuda.empty_cache()
            # cfg3 / cfg4 are DEFINED as per-GPU shards of a larger job (256 slices / 2000 frames over 8 GPUs): every rank
            # runs its shard, no collective; value = all ranks' samples over the slowest rank's time
            shards = {}
            for name in ("cfg3", "cfg4"):
                try:
                    shards[name] = shard_leg(torch, t, name, rank, world, local)
                except Exception as e:
                    shards[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.empty_cache()
            extras["shards"] = shards
            try:
                comm = make_comm(torch, t, rank, world, local)
                pair = Cfg5Pair(torch, t, rank, world, local, comm)
                cs = pair.measure(args.steps, args.warmup)
                cs["workload"] = ("BASELINE cfg5 pair, 64 coils sharded %d per GPU, partial sums of squares combined by "
                                  "one ncclReduce (tron_coil_reduce) + sqrt on rank 0" % pair.ncl)
                cs["scaling"] = "strong"
                pair.close(); comm.close()
                extras["coil_sharded"] = cs
            except Exception as e:
                extras["coil_sharded"] = {"error": "%s: %s" % (type(e).__name__, e)}

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
                       "rel_l2_vs_device": e2e_rel, "max_slice_rel_l2_vs_device": per_slice_rel,
                       "copy_floor_ms": floor_ms, "frac_of_copy_floor": floor_ms / (e2e_s * 1e3),
                       "copy_floor_how": "the step's H2D and D2H bytes alone, concurrently on two streams from the same "
                                         "pinned buffers, max over ranks"},
               "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "checksum": checksum}
        out.update(extras)
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if plan is not None:
        plan.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_cfg5(args):
    """BASELINE cfg5 as the main line (--workload cfg5): see Cfg5Pair."""
    import torch
    import tron_b200 as t
    from tron_b200 import build
    build.build()
    rank, world, local = dist_setup(args.gpus)
    comm = make_comm(torch, t, rank, world, local)
    pair = Cfg5Pair(torch, t, rank, world, local, comm)
    m = pair.measure(args.steps, args.warmup)
    if rank == 0:
        line = {"metric": "radial k-space samples gridded/sec", "value": m["value"], "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fp16 storage)",
                "data": "synthetic",
                "config": {"workload": "BASELINE cfg5 fp16-storage forward+adjoint pair: 1024 matrix, 2x grid, "
                                       "kernel width 6, 64 coils coil-sharded, one ncclReduce of the partial sum of squares "
                                       "(tron_coil_reduce)", "name": "cfg5", "coils_per_gpu": pair.ncl},
                "e2e": m["e2e"], "gpu_launches": m["gpu_launches_per_step"] * args.steps, "roofline": m["roofline"],
                "coil_sharded": m}
        print(json.dumps(line), flush=True)
    pair.close(); comm.close()
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
        # the stock reference holds at most MAXCHAN = 6 channels per call (tron.h:51): the job is
        # run in <= 6-coil chunks and the chunk times are summed (a bounded sample of slices, scaled).
        k = min(geom["nz"], 4)
        rd = list(dims)
        rd[3] = geom["npe1work"] + geom["prof_slide"] * (k - 1)
        ms, how = reference_chunks_ms(rd, flags, k, geom["nz"])
        nsamp = geom["nc"] * geom["nro"] * geom["npe1work"] * geom["nz"]
        value = nsamp / (ms * 1e-3)
        print(json.dumps({"impl": "reference", "metric": "radial k-space samples gridded/sec", "value": value,
                          "unit": "samples/s", "n_gpus": 1, "steps": 1, "warmup": 1, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": desc, "name": args.workload, "how": how},
                          "cpu_baseline": {"value": value, "unit": "samples/s", "cores": 0, "kind": "reference", "sample": how},
                          "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
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
    ap.add_argument("--lean", action="store_true", help="headline line only: no parity / cold / other-config / strong / coil-sharded legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg5":
        run_cfg5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
