#!/usr/bin/env python3
"""bench.py -- headline benchmark: SPH particles/sec projected to map (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repository's CUDA path
  python bench.py --impl reference [...]                          the reference's own CPU path (oracle/_ref)

Workload (BASELINE.json configs[1]): synthetic 256^3 = 16.8 M gas particles (recipe S1 of SURVEY.md 8(d): jittered
lattice in a unit periodic box, seed 12345), 2048 x 2048 projected mass and temperature-weighted maps (two weight
fields, one pass), M4 cubic spline (the reference kernel), smoothing lengths h = d_48 (the SPH-realistic choice: a
particle covers ~4000 pixels, so the accumulation is FP32-issue-bound, not HBM-bound -- SURVEY 7.2 H1).
A "step" is one full pass of the hot path over the particle set: bin -> scan -> emit -> sort -> tile accumulate
(-> NCCL reduce of the partial maps when N > 1).  N > 1 is weak scaling: every rank owns its own 256^3 shard
and all shards are deposited onto the same map.

One JSON line is printed by rank 0 (keys documented in the task contract): value = device-resident throughput,
e2e = the same through the public host-buffer API (H2D + D2H inside the timed region), roofline = the dominant
kernel against the measured HBM peak, cpu_baseline = the oracle's OpenMP restatement on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SPH particles/sec projected to map"
UNIT = "particles/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=256, help="lattice size per rank (particles = n^3)")
    ap.add_argument("--npix", type=int, default=2048)
    ap.add_argument("--k", type=int, default=48)
    ap.add_argument("--h-scale", type=float, default=1.0, help="multiplies h = d_k (footprint sweep)")
    ap.add_argument("--h-mode", default="auto", choices=["auto", "uniform", "knn"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="full footprint sweep (h scale 1/64 .. 1) instead of the three default regimes")
    ap.add_argument("--no-regimes", action="store_true", help="skip the footprint regimes entirely")
    return ap.parse_args()


def hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """number of samples read so far (brackets the timed region)"""
        return len(self.lines)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        # samples taken between the two marks (the timed region); if the region was shorter than nvidia-smi's period, the
        # samples of the identical warm-up steps just before it
        window = self.lines[first:(last if last is not None else len(self.lines)) + 1]
        if not window:
            window = self.lines[max(0, first - 4):first + 1]
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(args, rank):
    from astro_sph_tools_b200 import synthetic
    pos, rng = synthetic.s1_positions(args.n, 1.0, 12345 + rank)
    N = pos.shape[0]
    mass = np.full(N, 1.0 / N)
    T = 10.0 ** rng.uniform(4.0, 7.0, N)
    return pos, mass, mass * T


def smoothing_lengths(args, pos_d, pos):
    """h = d_k: the k-th neighbour distance, self included (io/SWIFT/_SnapshotSWIFT.py:62-83 convention), periodic box.
    Uses this package's CUDA k-NN when it is available, otherwise the constant lattice estimate (documented in config)."""
    from astro_sph_tools_b200 import synthetic
    mode = args.h_mode
    if mode in ("auto", "knn"):
        try:
            from astro_sph_tools_b200.tools.smoothing import compute_smoothing_lengths_device
            h = compute_smoothing_lengths_device(pos_d, args.k, box_size=1.0)
            return h * args.h_scale, f"d_{args.k} from the CUDA k-NN (periodic)"
        except (ImportError, NotImplementedError):
            if mode == "knn":
                raise
    import torch
    h = torch.full((pos.shape[0],), synthetic.s1_h_lattice_estimate(args.n, args.k) * args.h_scale, dtype=torch.float64,
                   device=pos_d.device)
    return h, f"constant lattice estimate of d_{args.k} = (3k/4pi)^(1/3) L/n"


def cpu_baseline(args, pos, h, props, target_s=12.0):
    """oracle port (OpenMP float64 scatter restatement) on a bounded sub-box sample of the same workload"""
    import oracle
    nthreads = oracle.max_threads()

    def sample(frac):
        w = 1.0 * frac
        sel = (pos[:, 0] < w) & (pos[:, 1] < w)
        npx = max(int(round(args.npix * frac)), 1)
        return np.ascontiguousarray(pos[sel]), h[sel], np.ascontiguousarray(np.stack([q[sel] for q in props])), npx, w

    def run(frac):
        p, hh, pr, npx, w = sample(frac)
        t0 = time.time()
        oracle.project2d(p, hh, pr, (npx, npx), 2, 0.0, w, 0.0, w, kernel="cubic_spline_3d", nthreads=0)
        return len(hh), time.time() - t0, npx

    n0, t0, _ = run(1.0 / 16)                                        # calibration
    rate = n0 / max(t0, 1e-3)
    frac = min(1.0, max(1.0 / 16, np.sqrt(rate * target_s / len(h))))
    frac = max(1, int(frac * 16)) / 16.0
    n1, t1, npx = run(frac)
    return {"value": n1 / t1, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"sub-box [0,{frac:g})^2 of the workload: {n1} particles -> {npx}^2 window, both weight fields, {t1:.2f} s, "
                      f"oracle/sph_oracle.c OpenMP float64 scatter"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref, serial by construction,
    tools/projections/_projector.py:111) on a bounded sub-box sample of the same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    pos, m, mT = make_inputs(args, 0)
    from astro_sph_tools_b200 import synthetic
    h = np.full(len(m), synthetic.s1_h_lattice_estimate(args.n, args.k) * args.h_scale)
    frac = 1.0 / 32
    sel = (pos[:, 0] < frac) & (pos[:, 1] < frac)
    p, hh, a, b = np.ascontiguousarray(pos[sel]), h[sel], m[sel], mT[sel]
    npx = max(int(round(args.npix * frac)), 1)
    kind = "reference"
    try:
        mod, Axes = oracle.reference_module()
        def step():
            mod.create_image(p, hh, a, (npx, npx), 32, Axes.Z, 0.0, frac, 0.0, frac)
            mod.create_image(p, hh, b, (npx, npx), 32, Axes.Z, 0.0, frac, 0.0, frac)
        cores = 1
    except Exception:
        kind = "port"
        cores = oracle.max_threads()
        pr = np.stack([a, b])
        def step():
            oracle.project2d(p, hh, pr, (npx, npx), 2, 0.0, frac, 0.0, frac)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / args.steps
    val = len(hh) / dt
    sample = (f"sub-box [0,1/32)^2 of the workload: {len(hh)} particles -> {npx}^2 window at the workload's pixel scale, "
              f"mass and mass*T maps (two create_image calls), {dt:.2f} s per step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"S1 {args.n}^3 particles -> {args.npix}^2 mass + T-weighted maps, cubic spline, h=d_{args.k} (lattice estimate)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from astro_sph_tools_b200 import CoordinateAxes, distributed as astd
    from astro_sph_tools_b200.tools.projections import Projector2D, quartic_spline_kernel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    pos, m, mT = make_inputs(args, rank)
    N = pos.shape[0]
    pos_d = torch.from_numpy(pos).to(dev)
    h_d, h_note = smoothing_lengths(args, pos_d, pos)
    props_d = [torch.from_numpy(m).to(dev), torch.from_numpy(mT).to(dev)]
    size = (args.npix, args.npix)
    bounds = (0.0, 1.0, 0.0, 1.0)
    eng = Projector2D(device=dev)
    out = torch.empty((2,) + size, dtype=torch.float64, device=dev)

    def step():
        astd.project_sharded(eng, pos_d, h_d, props_d, size, CoordinateAxes.Z, bounds, "cubic_spline_3d", out=out)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    launches_per_step = eng.last_stats["n_launches"]
    if rank == 0:                     # nvidia-smi needs a few hundred ms to deliver its first sample: keep the GPU under the
        t_wait = time.time()          # same load (extra untimed warm-up steps) until it has
        while sampler.proc is not None and sampler.mark() == 0 and time.time() - t_wait < 3.0:
            if world == 1:
                step()
                torch.cuda.synchronize()
            else:                     # step() holds a collective: rank 0 must not run it alone
                time.sleep(0.02)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    m0 = sampler.mark()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    m1 = sampler.mark()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop(m0, m1) if rank == 0 else None
    ms_per_step = float(ms.item()) / args.steps
    value = world * N / (ms_per_step * 1e-3)

    # ---- per-stage device times (CUDA events recorded inside the library on the launching stream)
    stage_names = ["bin+direct", "scan", "emit", "sort", "tile_ranges", "tile_accumulate", "memset", "total"]
    acc = np.zeros(8)
    reps = 3
    for _ in range(reps):
        eng.project(pos_d, h_d, props_d, size, CoordinateAxes.Z, bounds, "cubic_spline_3d", out=out, timing=True)
        acc += np.array(eng.last_stats["stage_ms"])
    stage_ms = acc / reps
    stats = eng.last_stats
    dom = int(np.argmax(stage_ms[:6]))
    peak, peak_src = hbm_peak()
    alg_bytes = N * (24 + 8 + 8 * 2) + 2 * size[0] * size[1] * 8          # N*48 + 2*nx*ny*8
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(stage_names[dom])
    except Exception:
        pass
    achieved = alg_bytes / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": stage_names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                "kernel_ms": float(stage_ms[dom]),
                "note": "h = d_48 footprints (~4000 pixel updates per particle) make this stage FP32-issue/MUFU-bound; see fp32 key"}
    # FP32 issue roofline of the accumulate stage: useful kernel evaluations (pixel, particle) pairs inside the support
    px_per_particle = float((np.pi * (2.0 * h_d.double().mean().item() * args.npix) ** 2))
    evals = N * px_per_particle
    # rowcol_accum_kernel<cubic, 2 props>: 10.6 SASS instructions per evaluated pixel in the full-shape loop, 7.6 in the
    # outer-annulus loop, one MUFU.SQRT each.  Two ceilings for USEFUL updates (pixels inside the support): the issue
    # ceiling at the full-loop instruction count, and the MUFU ceiling (16 sqrt per clock per SM, measured by
    # benchmarks/micro/ffma2_probe.cu: 8 cycles per warp-wide MUFU.SQRT) -- every update needs exactly one square root.
    sm_hz = 148 * 1.965e9
    fp32 = {"pixel_updates_per_particle": px_per_particle, "updates_per_s": evals / (stage_ms[5] * 1e-3) if stage_ms[5] > 0 else None,
            "lane_instr_per_evaluated_pixel_sass": {"full": 10.6, "outer_annulus": 7.6},
            "issue_peak_lane_instr_per_s": sm_hz * 4 * 32, "mufu_peak_sqrt_per_s": sm_hz * 16}
    if fp32["updates_per_s"]:
        fp32["useful_issue_frac_at_max_clock"] = fp32["updates_per_s"] * 10.6 / fp32["issue_peak_lane_instr_per_s"]
        fp32["useful_mufu_frac_at_max_clock"] = fp32["updates_per_s"] / fp32["mufu_peak_sqrt_per_s"]

    # ---- end to end through the public host-buffer API (pinned host arrays in, numpy map out)
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
        h_host = pin(h_d.cpu().numpy())
        pos_p, m_p, mT_p = pin(pos), pin(m), pin(mT)

        def e2e_step():
            return astd.create_images_sharded(pos_p, h_host, [m_p, mT_p], size, 32, CoordinateAxes.Z, *bounds,
                                              kernel_func=quartic_spline_kernel)
        res = None
        for _ in range(3):            # warm-up holds the previous result like the timed loop does, so both pinned result
            res = e2e_step()          # blocks of torch's caching host allocator exist before the clock starts
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = e2e_step()
        sync_all()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * N * args.steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(world * N * (24 + 8 + 16)), "d2h_bytes_per_step": int(2 * size[0] * size[1] * 8),
               "ms_per_step": float(dt.item()) / args.steps * 1e3}

    # footprint regimes: the same particle set with h scaled down; shows where the path is HBM-bound (sub-pixel supports,
    # every particle deposited by the binning kernel) and where it is FP32-issue-bound (SPH-realistic supports)
    sweep = None
    if not args.no_regimes and rank == 0 and world == 1:
        sweep = []
        for sc in ((1 / 64, 1 / 32, 1 / 16, 1 / 8, 1 / 4, 1 / 2, 1.0) if args.sweep else (1 / 64, 1 / 8, 1.0)):
            hs = h_d * sc
            for _ in range(2):
                eng.project(pos_d, hs, props_d, size, CoordinateAxes.Z, bounds, out=out)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                eng.project(pos_d, hs, props_d, size, CoordinateAxes.Z, bounds, out=out)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 3
            eng.project(pos_d, hs, props_d, size, CoordinateAxes.Z, bounds, out=out, timing=True)
            sweep.append({"h_scale": sc, "stage_ms": [round(x, 4) for x in eng.last_stats["stage_ms"]], "support_radius_px": float(2 * hs.mean().item() * args.npix), "ms": t,
                          "particles_per_s": N / (t * 1e-3), "hbm_frac": alg_bytes / (t * 1e-3) / 1e9 / peak,
                          "pairs": eng.last_stats["n_pairs"]})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, pos, h_d.cpu().numpy(), [m, mT])

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 weights, f64 index work and accumulation",
            "data": "synthetic",
            "config": {"workload": f"S1 {args.n}^3 = {N} particles per GPU -> {args.npix}^2 projected mass + T-weighted maps (one pass), "
                                   f"cubic spline (reference kernel), axis Z", "h": h_note, "h_scale": args.h_scale,
                       "pairs_per_step": stats["n_pairs"], "rounds": stats["n_rounds"],
                       "l2": f"inputs {N * 48 / 1e6:.0f} MB per step exceed the 126 MB L2, no explicit flush",
                       "parallelism": f"particles sharded by index over {world} GPU(s), NCCL sum-reduce of the map" if world > 1 else "1 GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roofline, "fp32": fp32, "stages_ms": dict(zip(stage_names, [float(x) for x in stage_ms])),
            "cpu_baseline": cpu,
        }
        if sweep:
            line["regimes"] = sweep
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
