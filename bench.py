#!/usr/bin/env python
"""bench.py -- EDaGe-PP hot path on B200 (BASELINE.json metric: collision-checked segments/sec and valid
paths/sec vs host CPU).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                   (CPU arm: the reference's path on the host cores)

One PASS = the hot path over one batch of `--maps` fresh synthetic maps per GPU (config 2: 10 000 maps, R=224, O=50
candidate circles, 1024 candidate segments/map, clearance 1 -> 4.48 px):
  generate_maps (placement rejection + label transform + obstacle draws + clearance verdict + bit raster)
  -> verdict_fused (A11 f64 + A12 f32 on ONE read of the f64 segments, bit-packed verdicts)
  -> dda_gridcheck on the same array vs the bit-packed maps -> warp-scan compaction of the free segments and of the
  valid maps -> GMM samples.
One STEP = `--passes` passes (default 32: the timed region of the driver's 20-step run is > 0.5 s).  The verdict
kernels, the DDA and the sampler run on separate streams and pass i+1's generator overlaps pass i's verdicts; the
per-kernel times of the `roofline` / `kernels` objects come from a second, sequential, event-bracketed run.
`value` is device-resident throughput (inputs already in HBM); `e2e` runs the same passes through the host-buffer C
ABI with pinned HOST inputs/outputs, copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

CLEAR_UNITS = 1.0                 # MapGenerate ctor default clearance (map units) -> 4.48 px at R=224, M=50
R, MAP_SIZE, O, SEGS_PER_MAP, OBST_SIZE = 224, 50.0, 50, 1024, 5.0
GMM_PER_MAP = 1000                # 10 M samples at 10 k maps
N_BANK, REPS = 100, 10
SEED = 0x5050_4E45_54
SIGMA = 15.0                      # e = s + N(0, 15^2)


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz") or 1965.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def config_dict(maps, passes, world):
    """Identical for both arms (the driver compares them)."""
    return {"workload": "config2: %d maps/GPU/pass, R=224, O=50 candidate circles, %d segments/map x {A11 f64, A12 f32, DDA}, "
                        "%d GMM samples/map, clearance 1 (4.48 px), bank of %d target paths; %d passes per step"
                        % (maps, SEGS_PER_MAP, GMM_PER_MAP, N_BANK, passes),
            "maps_per_gpu_per_pass": maps, "passes_per_step": passes,
            "segments_per_step": 3 * maps * SEGS_PER_MAP * passes * world, "parallelism": "map-sharded x%d" % world,
            "l2": "inputs larger than L2 (segments 328 MB + labels 162 MB + bitmaps 63 MB per pass); no flush needed"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Samples before this call (start-up, warm-up) are dropped: only the timed region is reported."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines[max(self.first - 1, 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU arm
def _cpu_inputs(sample_maps, seed=123):
    from ppnet_b200.synthetic import synthetic_bank, synthetic_segments
    rng = np.random.default_rng(seed)
    bank = synthetic_bank(8, seed=1)
    segs = synthetic_segments(sample_maps, SEGS_PER_MAP, sigma=SIGMA, seed=9)
    seg_map = np.repeat(np.arange(sample_maps, dtype=np.int32), SEGS_PER_MAP)
    obs = np.zeros([sample_maps, O, 3])
    obs[..., 0] = rng.uniform(0, R, (sample_maps, O))
    obs[..., 1] = rng.uniform(0, R, (sample_maps, O))
    obs[..., 2] = rng.uniform(0, 22.4, (sample_maps, O))
    cnt = rng.integers(20, O + 1, sample_maps).astype(np.int32)
    pp = bank["pathpt"][rng.integers(0, 8, sample_maps)]
    cand = np.stack([rng.uniform(0, 50, (sample_maps, O)), rng.uniform(0, 50, (sample_maps, O)),
                     rng.uniform(0, 5, (sample_maps, O))], axis=2)
    return dict(segs=segs, xy32=np.ascontiguousarray(segs[:, [1, 0, 3, 2]].astype(np.float32)), seg_map=seg_map, obs=obs, cnt=cnt,
                pp=pp, cand=cand)


def cpu_port_pass(inp, threads):
    """One pass of the path over the sample on the plain-C oracle port, all `threads` host threads: clearance filter +
    raster of the sample's maps, then A11 + A12 + DDA verdicts of their segments.  -> (seconds, verdicts, positives)."""
    from oracle import c_oracle
    clear_px = CLEAR_UNITS / MAP_SIZE * R
    t0 = time.perf_counter()
    c_oracle.clearance_filter(inp["pp"], inp["cand"], MAP_SIZE, R, CLEAR_UNITS, threads=threads)
    bits = c_oracle.raster_circles_bits(inp["obs"], inp["cnt"], R, clear_px / 2, threads=threads)
    v64 = c_oracle.segcheck_f64(inp["segs"], inp["seg_map"], inp["obs"], inp["cnt"], clear_px, threads=threads)
    c_oracle.segcheck_f32(inp["xy32"], inp["seg_map"], inp["obs"], inp["cnt"], clear_px, threads=threads, want_steer=False)
    c_oracle.dda_gridcheck(bits, R, inp["xy32"], inp["seg_map"], threads=threads)
    return time.perf_counter() - t0, 3 * len(inp["segs"]), float(v64.mean())


def _py_port_chunk(args):
    """Worker of the Python-port timing: the line-by-line numpy restatement, one segment at a time (as the reference runs)."""
    lo, hi, segs, seg_map, obs, cnt, clear_px = args
    from oracle import ppnet_oracle as orc
    t0 = time.perf_counter()
    for i in range(lo, hi):
        m = seg_map[i]
        ob = obs[m, :cnt[m]].tolist()
        orc.segcheck_edage_f64(segs[i, :2], segs[i, 2:], ob, clear_px)
        orc.segcheck_mpnet_f32(segs[i, [1, 0]].astype(np.float32), segs[i, [3, 2]].astype(np.float32), ob, clear_px)
    return time.perf_counter() - t0


_REF = {}


def _ref_chunk(args):
    """Worker of the REAL-reference timing (only where the reference tree exists): process_map.collision_check_circle_edge
    (EDaGe-PP/process_map.py:383-425) and the AST-lifted neuralplanner.collision_check_circle_edge (:43-69)."""
    import contextlib
    import io
    import torch
    lo, hi, segs, seg_map, obs, cnt, clear_px = args
    from oracle import ref_loader
    if "pm" not in _REF:
        _REF["pm"] = ref_loader.load_edage()["process_map"]
    pm = _REF["pm"]
    obc = [obs[m, :cnt[m]].tolist() for m in range(len(cnt))]
    ns = ref_loader.load_mpnet_checker(obc, clear_px)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):                      # the reference prints every collision
        for i in range(lo, hi):
            m = int(seg_map[i])
            pm.collision_check_circle_edge(torch.tensor(segs[i, :2]), torch.tensor(segs[i, 2:]), obc[m], clear_px)
            ns["collision_check_circle_edge"](torch.tensor(segs[i, [1, 0]], dtype=torch.float32),
                                              torch.tensor(segs[i, [3, 2]], dtype=torch.float32), m)
    return time.perf_counter() - t0


def python_legs(threads, n_seg=2000):
    """The reference-style scalar Python path, 1 core and all cores (multiprocessing, one process per core): the REAL
    reference functions where the reference tree is found ($PPNET_REF -> /root/reference -> baseline/_ref), otherwise the
    line-by-line Python restatement in oracle/ (BASELINE.md section 3)."""
    import multiprocessing as mp
    from oracle import ref_loader
    inp = _cpu_inputs(max(8, n_seg // 256), seed=321)
    segs, seg_map = inp["segs"][::SEGS_PER_MAP // 256][:n_seg], inp["seg_map"][::SEGS_PER_MAP // 256][:n_seg]
    clear_px = CLEAR_UNITS / MAP_SIZE * R
    ref = ref_loader.find_reference()
    fn = _ref_chunk if ref else _py_port_chunk
    out = {"kind": "reference" if ref else "port", "segments": int(2 * len(segs)),
           "what": ("process_map.collision_check_circle_edge (f64) + neuralplanner.collision_check_circle_edge (f32), the reference's own "
                    "functions, plot_obstacles not involved" if ref else
                    "oracle/ppnet_oracle.py line-by-line restatement of the same two functions (no reference tree on this box)")}
    n1 = min(len(segs), 400)
    t1 = fn((0, n1, segs, seg_map, inp["obs"], inp["cnt"], clear_px))
    out["one_core_segments_per_s"] = 2 * n1 / t1
    cuts = [len(segs) * t // threads for t in range(threads + 1)]
    try:
        ctx = mp.get_context("spawn")
        t0 = time.perf_counter()
        with ctx.Pool(threads) as pool:
            pool.map(fn, [(cuts[t], cuts[t + 1], segs, seg_map, inp["obs"], inp["cnt"], clear_px) for t in range(threads)])
        wall = time.perf_counter() - t0
        t0 = time.perf_counter()
        with ctx.Pool(threads) as pool:                                  # start-up cost of the pool (imports), subtracted
            pool.map(fn, [(0, 0, segs, seg_map, inp["obs"], inp["cnt"], clear_px)] * threads)
        startup = time.perf_counter() - t0
        out["all_cores_segments_per_s"] = 2 * len(segs) / max(wall - startup, 1e-9)
        out["all_cores"] = threads
    except Exception as e:                                               # no fork/spawn in this sandbox: report 1 core only
        out["all_cores_error"] = type(e).__name__
    return out


def reference_config1():
    """BASELINE.md section 3 item 1, only where the reference tree exists: MapGenerate(path_num=10, resolution=224, map_size=50,
    obstacles_num=20, clearance=3).generate(map_num=100), seed 0, plot_obstacles patched to a white tensor (excluded)."""
    import contextlib
    import io
    import tempfile
    import torch
    from oracle import ref_loader
    if ref_loader.find_reference() is None:
        return None
    mods = ref_loader.load_edage()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            np.random.seed(0)
            torch.manual_seed(0)
            with contextlib.redirect_stdout(io.StringIO()):
                t0 = time.perf_counter()
                mg = mods["MapGenerate"].MapGenerate(path_num=10, resolution=224, map_size=50, obstacles_num=20, clearance=3)
                t_paths = time.perf_counter() - t0
                t0 = time.perf_counter()
                mg.generate(map_num=100, folder_path=td + "/", round_index=0)
                t_maps = time.perf_counter() - t0
                # a 100-map sample of the clearance filter through generate_map_randomly
                # (MapGenerate.py:126-151) on the labels just produced
                t0 = time.perf_counter()
                for q in range(100):
                    lab = mg.MapLabel[q]
                    mg.generate_map_randomly(path_point=lab[4], init=lab[3][0], end=lab[3][10], length=1.0, path_obstacles=[], index=q)
                t_a14 = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
    return {"paths_s": t_paths, "maps_s": t_maps, "maps": 100, "valid_paths_per_s": 100 / t_maps, "cores": 1,
            "a14_generate_map_randomly_100_maps_s": t_a14, "a14_obstacle_verdicts_per_s": 100 * 20 / t_a14,
            "excluded": "plot_obstacles (matplotlib is not installed): patched to a white tensor"}


def cpu_baseline(sample_maps, threads, with_python=True):
    inp = _cpu_inputs(sample_maps)
    cpu_port_pass(_cpu_inputs(32), threads)
    dt, n, pos = cpu_port_pass(inp, threads)
    cb = {"value": n / dt, "unit": "segments/s", "cores": threads, "kind": "port",
          "sample": "%d maps x %d segments x {A11 f64, A12 f32, DDA} + clearance filter + raster of those maps, plain-C "
                    "oracle port on %d threads (%.2f s)" % (sample_maps, SEGS_PER_MAP, threads, dt),
          "valid_paths_per_s": sample_maps / dt, "positives": pos}
    if with_python:
        cb["python"] = python_legs(threads)
    return cb


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores, same workload definition as
    the GPU arm (config identical); each step = one pass over a bounded sample of `--cpu-sample-maps` maps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    inp = _cpu_inputs(args.cpu_sample_maps)
    for _ in range(max(args.warmup, 1)):
        cpu_port_pass(_cpu_inputs(64), threads)
    t0 = time.perf_counter()
    res = [cpu_port_pass(inp, threads) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    n = res[0][1]
    v = n * args.steps / dt
    py = python_legs(threads)                                             # outside the step loop
    c1 = reference_config1() if not args.no_config1 else None
    cb = {"value": v, "unit": "segments/s", "cores": threads, "kind": "port",
          "sample": "each step: %d maps x %d segments x {A11 f64, A12 f32, DDA} + clearance filter + raster of those maps, "
                    "plain-C oracle port (oracle/oracle_c.c) on %d threads" % (args.cpu_sample_maps, SEGS_PER_MAP, threads),
          "valid_paths_per_s": args.cpu_sample_maps * args.steps / dt, "positives": res[0][2], "python": py,
          "reference_config1": c1,
          "note": "value = the C port (the fastest faithful CPU implementation of the path, so the GPU/CPU ratio is a lower bound); "
                  "the reference itself is scalar Python: see `python` (kind = reference where the reference tree exists)"}
    line = {"impl": "reference", "metric": "collision-checked segments/sec", "value": v, "unit": "segments/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32+u32",
            "data": "synthetic", "valid_paths_per_s": args.cpu_sample_maps * args.steps / dt,
            "config": config_dict(args.maps, args.passes, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's threads to the CPUs next to its GPU (sysfs local_cpulist) so that the pinned host buffers it
    allocates afterwards are first-touched on the GPU's NUMA node: the e2e path is host<->device copies."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            txt = f.read().strip()
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": bdf, "numa_node": node, "cpus": len(cpus)}
    except Exception as e:                                   # containers without sysfs access: leave the affinity alone
        return {"unbound": type(e).__name__}


# ------------------------------------------------------------------------------------------------- secondary configs
def secondary_configs(ops, torch, quick=True):
    """BASELINE.json configs 3 (MPNet checker on batched predicted paths) and 5 (dense 1024^2 maps, long segments) as
    device-time microbenchmarks (CUDA events, 3 warm-ups): parity cases of tests/, reported as secondary keys."""
    def timed(fn, n=10, w=3):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    out = {}
    rng = np.random.default_rng(0)
    C = CLEAR_UNITS / MAP_SIZE * R
    # ---- config 3: 4000 MPNet problems, ragged paths of 4..64 f32 waypoints, <= 50 circles
    P = 4000
    obs = np.zeros([P, 50, 3])
    obs[..., 0] = rng.uniform(0, 224, (P, 50))
    obs[..., 1] = rng.uniform(0, 224, (P, 50))
    obs[..., 2] = rng.uniform(0, 9, (P, 50))
    cnt = rng.integers(0, 51, P).astype(np.int32)
    lens = rng.integers(4, 65, P)
    wps = []
    for L in lens:
        a, b = rng.uniform(5, 219, 2), rng.uniform(5, 219, 2)
        t = np.linspace(0, 1, L)[:, None]
        wps.append((a + t * (b - a) + rng.normal(0, 10, (L, 2))).astype(np.float32))
    wp = torch.from_numpy(np.concatenate(wps)).cuda()
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).cuda()
    pm = torch.arange(P, dtype=torch.int32).cuda()
    dobs, dcnt = torch.from_numpy(obs).cuda(), torch.from_numpy(cnt).cuda()
    n_edges = int(lens.sum() - P)
    ms = timed(lambda: ops.path_feasible_f32(wp, off, pm, dobs, dcnt, C))
    feas, _ = ops.path_feasible_f32(wp, off, pm, dobs, dcnt, C)
    out["config3_feasibility_check"] = {"ms": ms, "paths_per_s": P / ms * 1e3, "edges_per_s": n_edges / ms * 1e3,
                                        "feasible_frac": float(feas.float().mean().item())}
    ms = timed(lambda: ops.lvc_f32(wp, off, pm, dobs, dcnt, C))
    out["config3_lvc"] = {"ms": ms, "paths_per_s": P / ms * 1e3}
    segs = torch.from_numpy(np.concatenate([np.concatenate([w[:-1], w[1:]], axis=1) for w in wps])).cuda()
    soff = torch.from_numpy(np.concatenate([[0], np.cumsum(lens - 1)]).astype(np.int64)).cuda()
    ms = timed(lambda: ops.segcheck_mpnet_f32(segs, dobs, dcnt, C, seg_off=soff))
    out["config3_steerTo_flat"] = {"ms": ms, "segments_per_s": n_edges / ms * 1e3}
    # ---- config 5: dense 1024^2 maps (400 small circles), long segments: about half of them blocked, the free ones walk
    #      their whole length (hundreds of cells) through the 131 KB shared-memory bitmap
    Rr, S, Oo, M, SPM = 1024, 40, 400, 128 if quick else 256, 4096
    MS, CL, OS = 200.0, 0.4, 0.8
    c_px = CL / MS * Rr
    bank = ops.path_synthesize(0, 64, seg_num=S, clearance=CL, map_size=MS, resolution=Rr, seed=9, hmax=128, pomax=64).to_bank()
    gen = ops.generate_maps(bank, 0, M, 4, Oo, Rr, MS, OS, CL, seed=9, raster_inflate=c_px / 2, max_tries=1 << 16)
    ms = timed(lambda: ops.generate_maps(bank, 0, M, 4, Oo, Rr, MS, OS, CL, seed=9, raster_inflate=c_px / 2, max_tries=1 << 16, out=gen), n=5)
    out["config5_generate_maps"] = {"ms": ms, "maps_per_s": M / ms * 1e3, "valid": int(gen.valid.sum().item()), "maps": M,
                                    "avg_circles": float(gen.obs_cnt.float().mean().item())}
    s = rng.uniform(0, Rr, (M * SPM, 2))
    ang = rng.uniform(0, 2 * np.pi, M * SPM)
    ln = rng.uniform(64, 448, M * SPM)
    e = s + np.stack([np.cos(ang), np.sin(ang)], axis=1) * ln[:, None]
    seg64 = torch.from_numpy(np.concatenate([s, e], axis=1)).cuda()
    o = {}
    ms = timed(lambda: ops.verdict_fused(seg64, gen.obs, gen.obs_cnt, c_px, bound=float(Rr), want=("bits64", "bits32"), out=o), n=5)
    pos = float(ops.unpack_bits(o["bits64"], M * SPM).float().mean().item())
    out["config5_verdict_fused"] = {"ms": ms, "segments_per_s": 2 * M * SPM / ms * 1e3, "positives": pos}
    d = {}
    ms = timed(lambda: ops.dda_gridcheck_rc64(gen.bits, Rr, seg64, want=("bits", "first"), out=d), n=5)
    fh = d["first"]
    walked = torch.where(fh >= 0, fh, torch.from_numpy(np.maximum(np.abs(np.rint(e[:, 0]) - np.rint(s[:, 0])),
                                                                  np.abs(np.rint(e[:, 1]) - np.rint(s[:, 1]))).astype(np.int32)).cuda())
    out["config5_dda_gridcheck"] = {"ms": ms, "segments_per_s": M * SPM / ms * 1e3,
                                    "positives": float((fh >= 0).float().mean().item()),
                                    "mean_cells_walked": float(walked.float().mean().item()),
                                    "bitmap_bytes": Rr * Rr // 8}
    return out


# ------------------------------------------------------------------------------------------------- config 4
def config4(torch, dist, ops, sharding, rank, world, dev, total_maps, batch=10000, n_bank=1000):
    """BASELINE config 4: `total_maps` maps of the full generator (placement + labels + obstacle sets + bit raster),
    STRONG-scaled: the global index range [0, total) is split contiguously over the ranks (sharding.shard_range, index
    rule MapGenerate.py:68), every rank works through its shard in batches, and the only exchange is the all-gather of
    the per-rank counters and of a 64-bit content digest of everything produced (labels, obstacle sets, bitmaps).  The
    combined digest must be the same for every world size: N ranks emit the same bytes as one."""
    clear_px = CLEAR_UNITS / MAP_SIZE * R
    tb = time.perf_counter()
    paths = ops.path_synthesize(0, n_bank, seg_num=10, poly_order=4, clearance=CLEAR_UNITS, map_size=MAP_SIZE, resolution=R,
                                seed=SEED, hmax=64, pomax=24, device=dev)        # replicated: every rank draws the same bank
    bank = paths.to_bank()
    torch.cuda.synchronize()
    bank_s = time.perf_counter() - tb
    bank_ok = int(paths.hull_cnt.max().item()) <= 64 and int(paths.status.max().item()) == 0
    first, count = sharding.shard_range(total_maps, rank, world)
    counters = torch.zeros([4], dtype=torch.int64, device=dev)
    acc = torch.zeros([1], dtype=torch.int64, device=dev)
    gen = None

    def run(timed):
        nonlocal gen
        done = 0
        while done < count:
            nb = min(batch, count - done)
            if gen is None or gen.n_maps != nb:
                gen = ops.generate_maps(bank, first + done, nb, REPS, O, R, MAP_SIZE, OBST_SIZE, CLEAR_UNITS, SEED,
                                        counters=counters, raster_inflate=clear_px / 2)
            else:
                ops.generate_maps(bank, first + done, nb, REPS, O, R, MAP_SIZE, OBST_SIZE, CLEAR_UNITS, SEED, out=gen,
                                  raster_inflate=clear_px / 2)
            if not timed:
                ops.digest_maps(gen, acc)
            done += nb

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run(False)                                   # pass 1 (untimed): warm-up + the digest of every byte produced
    sync()
    cnt_digest = counters.clone()
    counters.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    a.record()
    run(True)                                    # pass 2 (timed): generation only
    b.record()
    sync()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    _, totals, offs = sharding.gather_counts(counters)
    combined, _ = sharding.combine_digests(acc)
    assert torch.equal(cnt_digest, counters), "the two passes over the same global range disagree"
    return {"total_maps": int(total_maps), "scaling": "strong", "world": world, "seconds": ms * 1e-3,
            "maps_per_s": total_maps / (ms * 1e-3), "valid_paths_per_s": totals["valid_paths"] / (ms * 1e-3),
            "bank": "%d target paths, replicated per rank, %.2f s untimed, capacities ok: %s" % (n_bank, bank_s, bank_ok),
            "counts_all_gathered": totals, "digest": "0x%016x" % combined,
            "digest_of": "angle, translation, SegPoint, PathPoint, accepted obstacle sets, obs_cnt, bit-packed maps, valid flags "
                         "of all %d maps; summed over ranks mod 2^64 after an all-gather" % total_maps,
            "shard_of_rank0": [int(first), int(count)]}


def run_config4(args):
    import torch
    import torch.distributed as dist
    from ppnet_b200 import ops, sharding
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c4 = config4(torch, dist, ops, sharding, rank, world, dev, args.total_maps)
    if rank == 0:
        emit({"metric": "config4: strong-scaled dataset generation", "value": c4["maps_per_s"], "unit": "maps/s", "n_gpus": world,
              "higher_is_better": True, "scaling": "strong", "config4": c4})
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from ppnet_b200 import _lib, host, ops, sharding
    from ppnet_b200.synthetic import synthetic_segments

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    numa_all = [numa]
    if world > 1:                                            # every rank's PCIe function, NUMA node and CPU set
        numa_all = [None] * world
        dist.all_gather_object(numa_all, numa)
    M, P = args.maps, args.passes
    n_seg = M * SEGS_PER_MAP
    n_words = (n_seg + 31) // 32
    clear_px = CLEAR_UNITS / MAP_SIZE * R

    # ---- inputs (untimed): target-path bank, candidate segments (host pinned + device), GMM parameters
    tb0 = time.perf_counter()
    paths = ops.path_synthesize(0, N_BANK, seg_num=10, poly_order=4, clearance=CLEAR_UNITS, map_size=MAP_SIZE, resolution=R,
                                seed=SEED, hmax=64, pomax=24, device=dev)
    bank = paths.to_bank()
    torch.cuda.synchronize()
    bank_ms = 1e3 * (time.perf_counter() - tb0)
    assert int(paths.hull_cnt.max().item()) <= 64 and int(paths.status.max().item()) == 0, "target-path bank overflowed its capacities"
    bk = {k: getattr(bank, k).cpu().numpy() for k in ("pathpt", "segpt", "hull", "hull_cnt", "obs", "obs_cnt")}
    segs_h = torch.from_numpy(synthetic_segments(M, SEGS_PER_MAP, sigma=SIGMA, seed=100 + rank)).pin_memory()
    segs = segs_h.to(dev)
    g_mean, g_std, g_w = ops.gmm_params(SEED, 10, 2, 70.0, 5.0, device=dev)
    n_gmm = GMM_PER_MAP * M
    counters = torch.zeros([4], dtype=torch.int64, device=dev)
    # two-deep buffers: pass i+1's generator runs while pass i's verdict kernels read pass i's maps
    gens = [ops.generate_maps(bank, rank * M, M, REPS, O, R, MAP_SIZE, OBST_SIZE, CLEAR_UNITS, SEED, counters=counters,
                              raster_inflate=clear_px / 2) for _ in range(2)]
    vout = [dict(bits64=torch.empty(n_words, dtype=torch.int32, device=dev), bits32=torch.empty(n_words, dtype=torch.int32, device=dev))
            for _ in range(2)]
    dout = [dict(bits=torch.empty(n_words, dtype=torch.int32, device=dev)) for _ in range(2)]
    cbuf = [ops.compact_bits(vout[b]["bits64"], vout[b]["bits32"], dout[b]["bits"], n=n_seg) for b in range(2)]
    gmm_out = [torch.empty([n_gmm, 2], dtype=torch.float32, device=dev) for _ in range(2)]
    free_total = torch.zeros([1], dtype=torch.int64, device=dev)
    s_ver, s_dda, s_gmm = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    ev_gen = [torch.cuda.Event() for _ in range(2)]
    ev_dda = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_gmm = [torch.cuda.Event() for _ in range(2)]
    pass_no = [0]

    def one_pass(step):
        """Software-pipelined pass: main = generator; s_ver = fused verdicts + compaction; s_dda = DDA; s_gmm = sampler."""
        i = pass_no[0]
        pass_no[0] += 1
        b = i & 1
        map0, _ = sharding.step_range(step * P + (i % P), rank, world, M)     # every pass generates NEW maps
        if i >= 2:
            main.wait_event(ev_done[b])                                       # buffers b were last read by pass i - 2
            main.wait_event(ev_gmm[b])
        ops.generate_maps(bank, map0, M, REPS, O, R, MAP_SIZE, OBST_SIZE, CLEAR_UNITS, SEED, out=gens[b], raster_inflate=clear_px / 2)
        ev_gen[b].record(main)
        with torch.cuda.stream(s_gmm):
            ops.gmm_sample(SEED, map0 * GMM_PER_MAP, n_gmm, g_mean, g_std, g_w, out=gmm_out[b])
            ev_gmm[b].record(s_gmm)
        with torch.cuda.stream(s_dda):
            s_dda.wait_event(ev_gen[b])
            ops.dda_gridcheck_rc64(gens[b].bits, R, segs, want=("bits",), out=dout[b])
            ev_dda[b].record(s_dda)
        with torch.cuda.stream(s_ver):
            s_ver.wait_event(ev_gen[b])
            ops.verdict_fused(segs, gens[b].obs, gens[b].obs_cnt, clear_px, want=("bits64", "bits32"), out=vout[b])
            s_ver.wait_event(ev_dda[b])
            ops.compact_bits(vout[b]["bits64"], vout[b]["bits32"], dout[b]["bits"], n=n_seg, out=cbuf[b])
            free_total.add_(cbuf[b][1])
            ev_done[b].record(s_ver)

    def join():
        for b in range(2):
            main.wait_event(ev_done[b])
            main.wait_event(ev_gmm[b])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler starts BEFORE the warm-up: nvidia-smi attaching to the driver stalls kernel launches for
    # several ms, which must not land inside the timed region; its 50 ms samples then cover warm-up + timed steps
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.6)
    for it in range(args.warmup):
        for _ in range(P):
            one_pass(it)
    join()
    barrier()
    counters.zero_()
    free_total.zero_()
    launches0 = _lib.launch_count()
    sampler.mark()
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(main)
    for it in range(args.steps):
        for _ in range(P):
            one_pass(args.warmup + it)
    join()
    t_end.record(main)
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    ms_total = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    # the path's only collective: one all_gather of int64[4] per rank
    _, totals, _ = sharding.gather_counts(counters)
    tot = np.asarray([totals[n] for n in sharding.COUNTER_NAMES])
    free_all = float(free_total.item())
    if world > 1:
        t = torch.tensor([free_all], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        free_all = float(t.item())
    ms_step = ms_total / args.steps
    ms_pass = ms_step / P
    maps_done, valid, acc_obs, tries = (int(x) for x in tot)
    seg_per_step = 3 * n_seg * P * world
    value = seg_per_step / (ms_step * 1e-3)
    valid_per_s = valid / (ms_total * 1e-3)

    # ---- per-kernel times: the same pass, sequential on one stream, CUDA events between the launches (untimed above)
    names = ["generate_maps", "verdict_fused", "dda_gridcheck", "compact_survivors", "gmm_sample"]
    seq_n = 20
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(seq_n)]
    for q in range(seq_n + 2):
        e = evs[q - 2] if q >= 2 else None
        map0, _ = sharding.step_range(10_000 + q, rank, world, M)
        if e: e[0].record()
        ops.generate_maps(bank, map0, M, REPS, O, R, MAP_SIZE, OBST_SIZE, CLEAR_UNITS, SEED, out=gens[0], raster_inflate=clear_px / 2)
        if e: e[1].record()
        ops.verdict_fused(segs, gens[0].obs, gens[0].obs_cnt, clear_px, want=("bits64", "bits32"), out=vout[0])
        if e: e[2].record()
        ops.dda_gridcheck_rc64(gens[0].bits, R, segs, want=("bits",), out=dout[0])
        if e: e[3].record()
        ops.compact_bits(vout[0]["bits64"], vout[0]["bits32"], dout[0]["bits"], n=n_seg, out=cbuf[0])
        if e: e[4].record()
        ops.gmm_sample(SEED, map0 * GMM_PER_MAP, n_gmm, g_mean, g_std, g_w, out=gmm_out[0])
        if e: e[5].record()
    torch.cuda.synchronize()
    k_ms = np.asarray([[e[i].elapsed_time(e[i + 1]) for i in range(5)] for e in evs]).mean(axis=0)
    n_free = int(cbuf[0][1].item())

    # ---- roofline of the dominant kernel (algorithmic bytes per launch / its mean launch duration)
    avg_cnt = float(gens[0].obs_cnt.double().mean().item())
    bits_b = R * ((R + 31) // 32) * 4
    alg = {
        "generate_maps": M * (16 * bank.np + 16 * bank.nseg1 + 8 + 8 + 24 * avg_cnt + bits_b + 13),
        # one 32-B read per segment feeds both flavours; two verdict bits out; circles read once per map
        "verdict_fused": n_seg * (32 + 0.25) + M * 24 * avg_cnt,
        "dda_gridcheck": M * bits_b + n_seg * (32 + 0.125),
        "compact_survivors": 3 * 4 * n_words + 4 * n_free,
        "gmm_sample": n_gmm * 8,
    }
    # the figures SURVEY 8(d) quotes per unit (A11 33 B/seg + A12 17 B/seg as separate launches), for comparison
    alg_survey = {"verdict_fused": n_seg * (33 + 17) + 2 * M * 24 * avg_cnt, "dda_gridcheck": M * bits_b + n_seg * 17}
    peak, sm_mhz, peak_src = peaks()
    issue_peak = 148 * 4 * sm_mhz * 1e6                      # warp instructions per second the 592 schedulers can issue
    warp_inst = {}
    tp = os.path.join(REPO, "profiles", "traffic.json")     # per-launch counters from the committed ncu --set full capture
    traffic = {}
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f)
    kernels = {}
    for i, n in enumerate(names):
        k = {"ms": float(k_ms[i]), "share": float(k_ms[i] / k_ms.sum()), "alg_bytes": float(alg[n]),
             "achieved_gbs": float(alg[n] / (k_ms[i] * 1e-3) / 1e9), "frac": float(alg[n] / (k_ms[i] * 1e-3) / 1e9 / peak)}
        if n in alg_survey:
            k["frac_survey_bytes"] = float(alg_survey[n] / (k_ms[i] * 1e-3) / 1e9 / peak)
        wi = (traffic.get(n) or {}).get("warp_instructions_per_launch")
        if wi:                                               # issue-slot fraction: how close the kernel is to the OTHER ceiling
            k["issue_slot_frac"] = float(wi / (issue_peak * k_ms[i] * 1e-3))
            warp_inst[n] = wi
        f64 = (traffic.get(n) or {}).get("fp64_pipe_pct")
        if f64 is not None:                                  # how busy the FP64 pipe is (ncu, same capture) -- not the bound
            k["fp64_pipe_pct"] = f64
        kernels[n] = k
    dom = names[int(np.argmax(k_ms))]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": (traffic.get(dom) or {}).get("dram_bytes_per_launch"),
                "peak_source": peak_src,
                "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full",
                "issue_slot_frac": kernels[dom].get("issue_slot_frac"),
                "note": "instruction-issue bound, not HBM bound: issue_slot_frac = warp instructions (ncu) / (148 SMs x 4 schedulers x "
                        "clock x time); DESIGN.md 4.2 has the instruction lower bound per segment"}

    # ---- e2e through the host-buffer C ABI: pinned host inputs -> host outputs, copies inside the timed region
    e2e = e2e_gen = None
    if not args.no_e2e:
        # Two passes in flight, each on its own context (own streams + arena) and its own set of pinned host buffers: the
        # calls are synchronous, so a caller that wants pass i+1's uploads to ride under pass i's downloads issues them
        # from two threads (contexts are thread-safe when distinct; ctypes drops the GIL).  The GMM sampler of each pass
        # runs from a third / fourth context.
        from concurrent.futures import ThreadPoolExecutor
        DEPTH = args.e2e_depth
        ctxs = [host.HostContext(local) for _ in range(DEPTH)]
        ctxs_gmm = [host.HostContext(local) for _ in range(DEPTH)]
        pool = ThreadPoolExecutor(max_workers=2 * DEPTH)
        hbank = host.HostBank(bk["pathpt"], bk["segpt"], bk["hull"], bk["hull_cnt"], bk["obs"], bk["obs_cnt"], device=local)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
        houts = [dict(angle=pin([M], torch.float64), trans=pin([M, 2], torch.int32),
                      segpt=pin([M, bank.nseg1, 2], torch.float64), pathpt=pin([M, bank.np, 2], torch.float64),
                      obs=pin([M, O + bank.pomax, 3], torch.float64), obs_cnt=pin([M], torch.int32),
                      rand_cnt=pin([M], torch.int32), bits=pin([M, R, (R + 31) // 32], torch.int32),
                      tries=pin([M], torch.int32), valid=pin([M], torch.uint8), counters=np.zeros(4, dtype=np.uint64))
                 for _ in range(DEPTH)]
        hws = [[pin([n_words], torch.int32).view(np.uint32) for _ in range(3)] for _ in range(DEPTH)]
        hvalid_idxs, hvalid_cnts = [pin([M], torch.int32) for _ in range(DEPTH)], [np.zeros(1, dtype=np.int64) for _ in range(DEPTH)]
        hgmms = [pin([n_gmm, 2], torch.float32) for _ in range(DEPTH)]
        s64 = segs_h.numpy()
        gm, gs, gw = g_mean.cpu().numpy(), g_std.cpu().numpy(), g_w.cpu().numpy()
        outs = [dict(vbits_f64=hws[d][0], vbits_f32=hws[d][1], vbits_dda=hws[d][2], valid_idx=hvalid_idxs[d],
                     valid_count=hvalid_cnts[d]) for d in range(DEPTH)]
        checks_up = [dict(segs_rc_f64=s64, clearance_px=clear_px, **outs[d]) for d in range(DEPTH)]     # ONE upload: 32 B / segment
        checks_gen = [dict(propose_sigma=SIGMA, segs_per_map=SEGS_PER_MAP, clearance_px=clear_px, **outs[d]) for d in range(DEPTH)]

        def e2e_pass(idx, checks, d):
            # one host call: upload this pass's candidate segments, generate the maps, run the verdict kernels against
            # them, download labels / obstacle sets / bitmaps / bit-packed verdicts / valid list (copies overlap kernels)
            map0, _ = sharding.step_range(idx, rank, world, M)
            f1 = pool.submit(ctxs_gmm[d].gmm_sample, SEED, map0 * GMM_PER_MAP, n_gmm, gm, gs, gw, out=hgmms[d])
            f2 = pool.submit(host.generate_maps_host, ctxs[d], hbank, map0, M, REPS, O, houts[d], R, MAP_SIZE, OBST_SIZE, CLEAR_UNITS,
                             SEED, raster_inflate=clear_px / 2, checks=checks[d])
            return map0, (f1, f2)

        def bytes_moved():
            tot = [0, 0]
            for c in ctxs + ctxs_gmm:
                a, b2 = c.bytes_moved()
                tot[0] += a
                tot[1] += b2
            return tot

        def e2e_run(checks, base):
            e2e_steps = max(2, min(args.steps, 3))
            inflight = []

            def submit(idx, q):
                while len(inflight) >= DEPTH:                    # slot q % DEPTH is free again once pass q - DEPTH returned
                    for f in inflight.pop(0)[1]:
                        f.result()
                inflight.append(e2e_pass(idx, checks, q % DEPTH))

            def drain():
                while inflight:
                    for f in inflight.pop(0)[1]:
                        f.result()

            for q in range(2 * DEPTH):
                submit(base + q, q)
            drain()
            b0 = bytes_moved()
            barrier()
            t0 = time.perf_counter()
            n_pass = e2e_steps * P
            for q in range(n_pass):
                submit(base + 10 + q, q)
            last = inflight[-1][0]
            drain()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            b1 = bytes_moved()
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return {"value": seg_per_step * e2e_steps / dt, "unit": "segments/s",
                    "h2d_bytes_per_step": (b1[0] - b0[0]) // e2e_steps, "d2h_bytes_per_step": (b1[1] - b0[1]) // e2e_steps,
                    "h2d_bytes_per_pass": (b1[0] - b0[0]) // n_pass, "d2h_bytes_per_pass": (b1[1] - b0[1]) // n_pass,
                    "ms_per_step": 1e3 * dt / e2e_steps, "ms_per_pass": 1e3 * dt / n_pass, "steps": e2e_steps,
                    "passes_in_flight": DEPTH, "valid_paths_per_s": M * P * world * e2e_steps / dt}, last, (n_pass - 1) % DEPTH

        # the region is short (0.7 s) and a shared box has hiccups: run it e2e_repeats times, report the MEDIAN run and list all
        def e2e_median(checks, base):
            runs = [e2e_run(checks, base + 1000 * r) for r in range(max(1, args.e2e_repeats))]
            order = sorted(range(len(runs)), key=lambda r: runs[r][0]["ms_per_pass"])
            pick = order[len(order) // 2]
            runs[pick][0]["repeats_ms_per_pass"] = [round(r[0]["ms_per_pass"], 4) for r in runs]
            runs[pick][0]["repeats_note"] = "median of %d runs of the timed region" % len(runs)
            # the equality check below needs the LAST run's outputs: they are what the host buffers hold now
            return runs[pick][0], runs[-1][1], runs[-1][2]

        e2e, last_map0, ld = e2e_median(checks_up, 20_000)
        hw, hout, hvalid_idx, hvalid_cnt = hws[ld], houts[ld], hvalid_idxs[ld], hvalid_cnts[ld]
        # the host results of the last timed pass equal the device-resident path on the same global map range
        chk = ops.generate_maps(bank, last_map0, M, REPS, O, R, MAP_SIZE, OBST_SIZE, CLEAR_UNITS, SEED, raster_inflate=clear_px / 2)
        cv = ops.verdict_fused(segs, chk.obs, chk.obs_cnt, clear_px, want=("bits64", "bits32"))
        cd = ops.dda_gridcheck_rc64(chk.bits, R, segs, want=("bits",))
        nv = int(hvalid_cnt[0])
        same = (np.array_equal(hw[0], cv["bits64"].cpu().numpy().view(np.uint32)) and
                np.array_equal(hw[1], cv["bits32"].cpu().numpy().view(np.uint32)) and
                np.array_equal(hw[2], cd["bits"].cpu().numpy().view(np.uint32)) and
                np.array_equal(hout["pathpt"], chk.pathpt.cpu().numpy()) and
                np.array_equal(hout["bits"], chk.bits.cpu().numpy()) and
                np.array_equal(hout["obs_cnt"], chk.obs_cnt.cpu().numpy()) and
                np.array_equal(hvalid_idx[:nv], np.nonzero(chk.valid.cpu().numpy() == 1)[0]))
        if not same:
            raise SystemExit("bench.py: e2e outputs differ from the device-resident path")
        e2e.update({"check": "host outputs of the last timed pass == device-resident path (verdict words, labels, bitmaps, valid list)",
                    "timer": "host wall clock around the whole run of synchronous host-API calls (each call synchronises before returning; "
                             "two passes in flight from two threads)",
                    "api": "ppnet_generate_and_check_host, one-array mode (f64 segments uploaded once, float32 flavours derived on the "
                           "device, verdicts bit-packed, valid maps compacted) + ppnet_gmm_sample_host (second context, concurrent); "
                           "pinned host buffers",
                    "cpu_affinity": numa_all})
        e2e_gen, _, _ = e2e_median(checks_gen, 40_000)

        # the ceiling of this box for these bytes: raw pinned copies of the same volume per pass, both directions at once,
        # every rank at the same time (no kernels, no API of ours) -- at N > 1 the ranks share the host's memory system
        def raw_copy_ms(up_bytes, down_bytes, reps=3, chunk=32 << 20):
            hi = torch.empty(max(up_bytes, 1), dtype=torch.uint8).pin_memory()
            ho = torch.empty(max(down_bytes, 1), dtype=torch.uint8).pin_memory()
            di = torch.empty(max(up_bytes, 1), dtype=torch.uint8, device=dev)
            do = torch.empty(max(down_bytes, 1), dtype=torch.uint8, device=dev)
            su, sd = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

            def once():
                for o in range(0, max(up_bytes, down_bytes), chunk):
                    if o < up_bytes:
                        with torch.cuda.stream(su):
                            di[o:o + chunk].copy_(hi[o:o + chunk], non_blocking=True)
                    if o < down_bytes:
                        with torch.cuda.stream(sd):
                            ho[o:o + chunk].copy_(do[o:o + chunk], non_blocking=True)
            once()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                once()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return 1e3 * dt

        for e in (e2e, e2e_gen):
            raw = raw_copy_ms(int(e["h2d_bytes_per_pass"]), int(e["d2h_bytes_per_pass"]))
            e["raw_copy_ms_per_pass"] = raw
            e["frac_of_raw_copy_ceiling"] = raw / e["ms_per_pass"]
            e["raw_copy_note"] = ("raw pinned cudaMemcpyAsync of the same bytes per pass, both directions at once on every rank "
                                  "simultaneously: what this box's PCIe + host memory system delivers at this N")
        e2e_gen["api"] = ("same call with the device-side segment source (propose_sigma): generator-mode callers upload nothing; "
                          "not the contract's e2e (no host->device input copy), reported beside it")

    c4 = None
    if not args.no_config4:
        c4 = config4(torch, dist, ops, sharding, rank, world, dev, args.total_maps)

    cb = sec = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_baseline(args.cpu_sample_maps, os.cpu_count() or 1)
    if rank == 0 and world == 1 and not args.no_secondary:
        sec = secondary_configs(ops, torch)

    if rank == 0:
        cfg = config_dict(M, P, world)
        line = {"metric": "collision-checked segments/sec", "value": value, "unit": "segments/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_pass": ms_pass,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32+u32", "data": "synthetic",
                "valid_paths_per_s": valid_per_s, "free_segments_per_s": free_all / (ms_total * 1e-3),
                "config": cfg,
                "run_info": {"bank": "%d target paths synthesised on the device (A1-A9) in %.1f ms, untimed" % (N_BANK, bank_ms),
                             "placement_tries_per_map": tries / max(maps_done, 1),
                             "accepted_random_obstacles_per_map": acc_obs / max(maps_done, 1),
                             "free_fraction": free_all / max(n_seg * P * args.steps * world, 1),
                             "streams": "generator | fused verdicts + compaction | DDA | sampler; two-deep buffers",
                             "sequential_pass_ms": float(k_ms.sum()), "overlapped_pass_ms": ms_pass},
                "roofline": roofline, "kernels": kernels, "cpu_baseline": cb, "e2e": e2e, "e2e_generator_mode": e2e_gen,
                "config4": c4, "secondary_configs": sec, "gpu_launches": int(launches), "clocks": clocks,
                "counts_all_gathered": {"maps": maps_done, "valid_paths": valid, "accepted_obstacles": acc_obs,
                                        "placement_tries": tries}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries (NCCL prints its version banner) write to fd 1
    behind Python's back: point fd 1 at stderr for the whole run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--maps", type=int, default=10000, help="maps per GPU per pass")
    ap.add_argument("--passes", type=int, default=32, help="passes (fresh map batches) per step")
    ap.add_argument("--cpu-sample-maps", type=int, default=4096, help="maps in the CPU arm's bounded sample (per step / for cpu_baseline)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-depth", type=int, default=2, help="host-API passes in flight (one thread + context each)")
    ap.add_argument("--e2e-repeats", type=int, default=3, help="runs of the e2e timed region; the median is reported, all are listed")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-config1", action="store_true", help="reference arm: skip the 45 s config-1 run of the real reference")
    ap.add_argument("--config4", action="store_true", help="only the strong-scaled 1 M-map dataset generation (BASELINE config 4)")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--total-maps", type=int, default=1_000_000)
    args = ap.parse_args()
    _claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config4:
        run_config4(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
