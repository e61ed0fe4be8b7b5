#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on B200: frames/s and Gde/s of
rectify -> SGBM (MODE_HH, left + right matcher) -> WLS -> depth -> Steger centre line -> 3D
at 1280x720, 128 disparities, block 9 (BASELINE config 3), frames sharded frame-wise over the GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over a batch of `--frames` synthetic frames per GPU (weak
scaling).  Prints ONE JSON line on rank 0:
  value / ms_per_step : inputs resident in HBM, CUDA-event timed (l3d_pipeline_run_dev), max over ranks
  e2e                 : the same batch through the host-buffer C-ABI call (l3d_pipeline_run_host):
                        pinned H2D of every frame + D2H of depth maps and point clouds inside the timed region
  roofline            : one SGBM matcher run (cost volume + aggregation + WTA kernels) timed alone with CUDA
                        events on its stream, algorithmic bytes / duration vs the measured HBM peak
  cpu_baseline        : the reference's CPU path (cv2.StereoSGBM etc. through oracle/ref_ops.py) on a bounded
                        sample of the same frames, frame-parallel over the host cores (rank 0, N=1 only)
`--impl reference` times only that CPU path (all host threads, bounded sample per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# one hardware work queue per stream (default 8): the frame pipeline drives 14 lane streams + 4 aggregation streams
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, D, BS, MODE = 1280, 720, 128, 9, 1  # BASELINE config 3 (MODE_HH = 1)
MAX_POINTS = 20000
WORKLOAD = "c3: 1280x720 per eye, 128 disparities, block 9, SGBM MODE_HH (left+right matcher) + WLS(8000,1.5) " \
           "+ ImprovedSteger sigma 3 + reconstruct_from_depth"


def sgbm_algorithmic_bytes():
    """B_sgbm = (2 + 4*npasses)*N + 4*W*H bytes per matcher run; N = width1*H*D, npasses = 2 (HH)."""
    n = (W - D) * H * D
    return (2 + 4 * 2) * n + 4 * W * H


def make_frames(n_distinct, nframes, seed0):
    from laser_3d_reconstruction_b200 import synth
    base = [synth.stereo_pair(W, H, D, seed0 + s) for s in range(n_distinct)]
    L = np.stack([base[i % n_distinct][0] for i in range(nframes)])
    R = np.stack([base[i % n_distinct][1] for i in range(nframes)])
    return L, R


# ----------------------------------------------------------------------------------------------
# CPU reference leg (oracle/ref_ops.py on the real cv2: TEST INFRASTRUCTURE used as a timed baseline)
# ----------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(maps, K, Q):
    import cv2
    cv2.setNumThreads(1)  # MODE_HH is single-threaded inside OpenCV anyway: parallelise over frames
    _CPU.update(maps=maps, K=K, Q=Q)


def _cpu_frame(pair):
    from laser_3d_reconstruction_b200 import synth
    from oracle import ref_ops
    left, right = pair
    rect, depth = ref_ops.depth_path(left, right, _CPU["maps"], D, BS, MODE, _CPU["Q"], use_wls=True)
    pts = ref_ops.improved_steger_extract(rect, loop=True)  # per-pixel np.linalg.eig, as the reference does
    xyz = ref_ops.ReconstructorRef(_CPU["K"], synth.LASER_PLANE, False).reconstruct_from_depth(pts, depth)
    return int(len(xyz))


class CpuPath:
    def __init__(self, max_workers=32):
        import multiprocessing as mp
        from laser_3d_reconstruction_b200 import synth
        try:
            ncpu = len(os.sched_getaffinity(0))
        except AttributeError:
            ncpu = os.cpu_count() or 1
        self.cores = max(1, min(ncpu, max_workers))  # each worker holds ~0.5 GB of OpenCV cost volumes
        K, Q = synth.camera_model(W, H)
        maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init, initargs=(maps, K, Q))

    def run(self, pairs):
        t0 = time.perf_counter()
        counts = self.pool.map(_cpu_frame, pairs, chunksize=1)
        return time.perf_counter() - t0, counts

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path, all host threads."""
    if rank != 0:
        return
    from laser_3d_reconstruction_b200 import synth
    cpu = CpuPath()
    nfr = cpu.cores  # one frame per worker and step: a bounded sample of the workload
    base = [synth.stereo_pair(W, H, D, s) for s in range(min(nfr, 8))]
    pairs = [base[i % len(base)] for i in range(nfr)]
    for _ in range(args.warmup):
        cpu.run(pairs[:max(1, min(nfr, cpu.cores))])
    t = 0.0
    for _ in range(args.steps):
        dt, counts = cpu.run(pairs)
        t += dt
    cpu.close()
    fps = nfr * args.steps / t
    sample = "%d frames/step (one per worker), %d steps" % (nfr, args.steps)
    out = {
        "impl": "reference", "metric": "frames_per_s", "value": fps, "unit": "frames/s", "gde_per_s": fps * W * H * D / 1e9,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": nfr, "sharding": "frame-parallel over host cores"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cpu.cores, "kind": "port", "sample": sample,
                         "note": "cv2 %s StereoSGBM/remap/cvtColor (the binary the reference calls) driven by "
                                 "oracle/ref_ops.py; WLS = oracle C restatement (cv2.ximgproc absent); Steger = "
                                 "per-pixel np.linalg.eig loop as in the reference" % __import__("cv2").__version__},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [s for (t, s) in self.samples if t0 <= t <= t1] or [s for (_, s) in self.samples[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from laser_3d_reconstruction_b200 import _native as N
    from laser_3d_reconstruction_b200 import pipeline, sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nfr = args.frames
    K, Q = synth.camera_model(W, H)
    maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
    # frames are sharded r::world from one global list of nfr*world frames (weak scaling)
    mine = sharding.shard_frames(nfr * world, rank, world)
    n_distinct = min(args.distinct, nfr)
    L, R = make_frames(n_distinct, nfr, seed0=1000 * rank)
    ctx = N.Context(local_rank)
    cfg = pipeline.make_pipeline_config(W, H, D, BS, MODE, Q, K, extractor=N.STEGER_IMPROVED, lanes=args.lanes,
                                        max_points=MAX_POINTS)
    fp = pipeline.FramePipeline(cfg, maps=maps, ctx=ctx)
    dL, dR = fp.upload(L), fp.upload(R)
    pL = pipeline.pinned_empty(L.shape, np.uint8)
    pR = pipeline.pinned_empty(R.shape, np.uint8)
    pL[:] = L
    pR[:] = R
    depth_h = pipeline.pinned_empty((nfr, H, W), np.float32)
    xyz_h = pipeline.pinned_empty((nfr, MAX_POINTS, 3), np.float64)

    def gather(counts):
        """NCCL gather of the per-frame point clouds to rank 0 (the only exchange, off the hot path)."""
        if world == 1:
            return None
        table = torch.empty((max(int(np.sum(counts)), 1), 4), dtype=torch.float64, device=dev)
        rows = fp.pack_points_dev(mine, table.data_ptr())  # device-side pack; nothing goes through the host
        return sharding.gather_point_clouds(table[:rows], device=dev, to_host=False)  # stays on rank 0's GPU

    def step_dev():
        counts = fp.run_dev(dL, dR, nfr)
        gather(counts)
        return counts, fp.last_ms

    def step_host():
        counts = fp.run_host(pL, pR, depth_h, xyz_h)
        return counts, fp.last_ms

    # ---- device-resident arm -----------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        counts, _ = step_dev()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = fp.launches
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        counts, ms = step_dev()
        dev_ms += ms
    barrier()
    t1 = time.perf_counter()
    wall_ms = (t1 - t0) * 1e3
    launches = fp.launches - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    # a step's duration = CUDA-event time of the batch (first enqueue -> last lane done) + the gather's wall time
    # for N > 1; the wall clock over the K steps bounds it from above, so use the wall clock when gathering.
    el = torch.tensor([wall_ms if world > 1 else dev_ms, wall_ms], dtype=torch.float64, device=dev)
    nl = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    elapsed_ms, wall_max = float(el[0]), float(el[1])
    launches = int(nl[0])  # kernels of libl3d.so launched inside the timed region, summed over ranks

    # ---- end-to-end arm (host buffers through the C ABI) ----------------------------------------
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    e2e_ms = 0.0
    for _ in range(args.steps):
        counts_h, ms = step_host()
        e2e_ms += ms
    barrier()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    ee = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ee, op=dist.ReduceOp.MAX)
    e2e_elapsed = float(ee[0])
    assert list(counts_h) == list(counts), "host-buffer and device-resident runs disagree"
    h2d = int(L.nbytes + R.nbytes)
    d2h = int(depth_h.nbytes + xyz_h.nbytes + 8 * nfr)

    total_frames = nfr * world * args.steps
    fps = total_frames / (elapsed_ms / 1e3)
    e2e_fps = total_frames / (e2e_elapsed / 1e3)
    out = {
        "metric": "frames_per_s", "value": fps, "unit": "frames/s", "gde_per_s": fps * W * H * D / 1e9,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
        "wall_ms_per_step": wall_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": nfr, "distinct_frames": n_distinct, "lanes": args.lanes,
                   "sharding": "frame-wise r::world, no collective on the hot path; NCCL gather of point clouds to rank 0"
                               " per step when N>1",
                   "l2": "inputs %.0f MB/step plus 425 MB of C/S cost volumes per matcher run exceed the 126 MB L2"
                         % ((L.nbytes + R.nbytes) / 1e6),
                   "gde_definition": "W*H*D per frame, counted once although two matchers run (SURVEY 8d)"},
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_elapsed / args.steps},
        "gpu_launches": int(launches),
        "points_per_frame": float(np.mean(counts)),
    }
    if clocks is not None:
        out["clocks"] = clocks

    # ---- roofline leg: the SGBM kernels timed alone (lanes chained, nothing overlaps them) ----------
    if rank == 0:
        cfg1 = pipeline.make_pipeline_config(W, H, D, BS, MODE, Q, K, extractor=N.STEGER_IMPROVED, lanes=args.lanes,
                                             max_points=MAX_POINTS)
        fp1 = pipeline.FramePipeline(cfg1, maps=maps, ctx=ctx)
        nr = min(nfr, 14)
        for _ in range(2):
            fp1.run_dev(dL, dR, nr)
        fp1.set_timing(True)
        fp1.run_dev(dL, dR, nr)
        names = ["sgbm_cost", "sgbm_scan_k0", "sgbm_vgroup_down", "sgbm_vgroup_up", "sgbm_wta", "wls"] + \
                ["sgbm_scan_k%d" % kk for kk in range(1, 8)]
        groups = {}
        for g in names:
            t, k = fp1.kernel_time(g)
            if k:
                groups[g] = {"ms_total": t, "timed_regions": k}
        fp1.set_timing(False)
        runs = 2 * nr  # left + right matcher per frame
        sgbm_ms = sum(v["ms_total"] for g, v in groups.items() if g.startswith("sgbm_")) / runs
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = sgbm_algorithmic_bytes() / (sgbm_ms * 1e-3) / 1e9
        traffic = None  # DRAM bytes of the same kernels from the committed ncu --set full capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["sgbm_run_dram_bytes"]
        except (OSError, ValueError, KeyError):
            pass
        out["roofline"] = {
            "bound": "hbm", "kernel": "sgbm matcher run = cost volume + horizontal paths + cluster-fused previous-row "
                                      "paths (both passes) + WTA",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
            "algorithmic_bytes_per_run": sgbm_algorithmic_bytes(), "ms_per_run": sgbm_ms, "traffic": traffic,
            "how": "CUDA events on the launching streams around every kernel group; lanes chained so each timed kernel "
                   "runs alone; %d frames = %d matcher runs; the cluster-fused aggregation launches carry all runs of a "
                   "lane set at once and their time is divided by the runs they process" % (nr, runs),
            "groups_ms_per_run": {g: v["ms_total"] / runs for g, v in groups.items() if g.startswith("sgbm_")},
            "wls_ms_per_frame": groups.get("wls", {"ms_total": 0.0})["ms_total"] / nr,
        }
        fp1.close()

    # ---- CPU baseline leg (rank 0, N = 1 only): bounded sample ----------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = CpuPath()
        ns = 8 * cpu.cores  # ~10-30 s of CPU work
        pairs = [(L[i % nfr], R[i % nfr]) for i in range(ns)]
        dt, ccounts = cpu.run(pairs)
        cpu.close()
        out["cpu_baseline"] = {"value": ns / dt, "unit": "frames/s", "cores": cpu.cores, "kind": "port",
                               "sample": "%d frames of the same workload, frame-parallel over %d worker processes, %.1f s"
                                         % (ns, cpu.cores, dt),
                               "points_match_gpu": bool(abs(ccounts[0] - int(counts[0])) <= max(2, 0.001 * ccounts[0]))}
    if rank == 0:
        print(json.dumps(out), flush=True)
    fp.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=112, help="frames per step per GPU (a multiple of the 7-frame lane set)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic frames rendered per rank")
    ap.add_argument("--lanes", type=int, default=28, help="frames in flight per GPU (streams): four lane sets of 7 frames")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one process per GPU
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
