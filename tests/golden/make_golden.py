"""Generate the golden fixtures in this directory by running the REAL PPNet reference code.

Run in the build container only (needs /root/reference or $PPNET_REF):

    cd /tmp && python /root/repo/tests/golden/make_golden.py [--only NAME]

Every array in every .npz was computed by the unmodified reference functions (loaded through
oracle/ref_loader.py: matplotlib/imgviz stubbed, plot_obstacles patched out, MPNet checker
AST-lifted).  The oracle (oracle/ppnet_oracle.py) is used here only to *construct sharp inputs*
(radii placed within a few ulp of the decision threshold); the recorded outputs are the
reference's.  Container: numpy 2.3.5 / OpenBLAS 0.3.30 (SkylakeX kernels unless
OPENBLAS_CORETYPE says otherwise -- recorded as `dot_mode` in the f64 fixture), scipy 1.18.1,
torch 2.11.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ppnet_oracle as orc  # noqa: E402
from oracle.ref_loader import load_edage, load_mpnet_checker  # noqa: E402

CLEAR = 1 / 50 * 224      # 4.48, what every caller passes


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def detect_dot_mode():
    """Which ddot does this numpy use?  fused (SkylakeX) or un-fused."""
    rng = np.random.default_rng(123)
    a = rng.standard_normal((2000, 2))
    b = rng.standard_normal((2000, 2))
    fused = sum(float(np.dot(a[i], b[i])) == orc.fma64(a[i, 1], b[i, 1], a[i, 0] * b[i, 0])
                for i in range(2000))
    unf = sum(float(np.dot(a[i], b[i])) == a[i, 0] * b[i, 0] + a[i, 1] * b[i, 1]
              for i in range(2000))
    if fused == 2000:
        return orc.DOT_FUSED_SKX
    if unf == 2000:
        return orc.DOT_UNFUSED
    raise RuntimeError("np.dot is neither model: fused %d unfused %d" % (fused, unf))


# ---------------------------------------------------------------------------------------------
def make_maps(rng, n_maps, omax=50):
    """Ragged obstacle sets like MapGenerate produces: [x, y, r] in pixels."""
    obs = np.zeros([n_maps, omax, 3])
    cnt = np.zeros(n_maps, dtype=np.int32)
    for m in range(n_maps):
        k = int(rng.integers(0, omax + 1)) if m >= 2 else (0 if m == 0 else omax)
        cnt[m] = k
        obs[m, :k, 0] = rng.uniform(0, 224, k)
        obs[m, :k, 1] = rng.uniform(0, 224, k)
        obs[m, :k, 2] = rng.uniform(0, 22.4 if m % 2 == 0 else 5.0, k)
    return obs, cnt


def sharp_cases(rng, n, flavour):
    """Segments + a single circle whose radius sits within +-3 ulp of the decision threshold of
    either the edge test (|dis| vs r + c/2) or the vertex test (|e-o| vs r + c/2)."""
    segs, circ = [], []
    ftype = np.float32 if flavour == "f32" else np.float64
    while len(segs) < n:
        s = rng.uniform(5, 219, 2)
        e = s + rng.normal(0, 40, 2)
        if not (0 <= e[0] <= 224 and 0 <= e[1] <= 224):
            continue
        s = s.astype(ftype)
        e = e.astype(ftype)
        t = rng.uniform(0.1, 0.9)
        nrm = np.array([e[1] - s[1], -(e[0] - s[0])], dtype=np.float64)
        nrm /= np.linalg.norm(nrm)
        kind = rng.integers(0, 2)
        if kind == 0:      # edge-sharp: centre off the segment interior
            off = rng.uniform(3, 20) * (1 if rng.random() < 0.5 else -1)
            o = s.astype(np.float64) + t * (e.astype(np.float64) - s) + off * nrm
            target = abs(off)
        else:              # vertex-sharp: centre near e
            ang = rng.uniform(0, 2 * np.pi)
            dist = rng.uniform(3, 20)
            o = e.astype(np.float64) + dist * np.array([np.cos(ang), np.sin(ang)])
            target = dist
        o = o.astype(np.float32).astype(np.float64)      # centre goes through f32 either way
        # the reference's own threshold quantity, from the oracle, then nudge the radius by ulps
        if flavour == "f32":
            d = e - o.astype(np.float32)
            meas = float(np.sqrt(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1]))) if kind else None
        size = target - CLEAR / 2
        if size <= 0:
            continue
        k = int(rng.integers(-3, 4))
        size = float(size)
        for _ in range(abs(k)):
            size = float(np.nextafter(size, np.inf if k > 0 else -np.inf))
        segs.append([s[0], s[1], e[0], e[1]])
        circ.append([float(o[0]), float(o[1]), size])
    return np.asarray(segs, dtype=ftype), np.asarray(circ, dtype=np.float64)


def sharpen_radius(seg, circ, flavour, rng):
    """Replace circ radius so that thr lands within a few ulp of the *reference-computed*
    comparison quantity (using the oracle's bit-exact intermediate)."""
    if flavour == "f32":
        f = np.float32
        s0, s1, e0, e1 = (f(v) for v in seg)
        o0, o1 = f(circ[0]), f(circ[1])
        if rng.random() < 0.5:
            v0, v1 = f(e0 - o0), f(e1 - o1)
            q = float(np.sqrt(f(f(v0 * v0) + f(v1 * v1))))
        else:
            d0, d1 = f(e0 - s0), f(e1 - s1)
            L = np.sqrt(f(f(d0 * d0) + f(d1 * d1)))
            n0, n1 = f(d1 / L), f(f(-d0) / L)
            q = float(abs(f(f(n0 * f(o0 - s0)) + f(n1 * f(o1 - s1)))))
        # thr = f32(size + c/2); choose size so thr is q nudged by k f32-ulps
        k = int(rng.integers(-2, 3))
        tq = np.float32(q)
        for _ in range(abs(k)):
            tq = np.nextafter(tq, np.float32(np.inf if k > 0 else -np.inf))
        size = float(tq) - CLEAR / 2
    else:
        s0, s1, e0, e1 = (np.float64(v) for v in seg)
        o0, o1 = np.float64(np.float32(circ[0])), np.float64(np.float32(circ[1]))
        mode = detect_dot_mode()
        if rng.random() < 0.5:
            v0, v1 = e0 - o0, e1 - o1
            q = float(np.sqrt(v0 * v0 + v1 * v1))
        else:
            d0, d1 = e0 - s0, e1 - s1
            L = np.sqrt(orc.dot2_f64(d0, d1, d0, d1, mode))
            n0, n1 = d1 / L, (-d0) / L
            q = float(abs(orc.dot2_f64(n0, n1, o0 - s0, o1 - s1, mode)))
        k = int(rng.integers(-2, 3))
        tq = q
        for _ in range(abs(k)):
            tq = float(np.nextafter(tq, np.inf if k > 0 else -np.inf))
        size = tq - CLEAR / 2          # thr = size + c/2 lands within an ulp or two of tq
    return size if size > 0 else circ[2]


KAT = [   # name, s(x,y), e(x,y), circle(x,y,r)
    ("edge_hit", (10, 10), (100, 10), (50, 14, 2)),
    ("edge_miss", (10, 10), (100, 10), (50, 14.3, 2)),
    ("vertex_hit_e", (10, 10), (100, 10), (103, 10, 1)),
    ("start_inside_only", (10, 10), (100, 10), (8, 10, 3)),
    ("beyond_end_proj", (10, 10), (100, 10), (110, 10, 1)),
    ("on_line_center", (10, 10), (100, 10), (50, 10, 1)),
    ("oob_xneg", (-1, 10), (100, 10), (500, 500, 1)),
    ("oob_y225", (10, 225), (100, 10), (500, 500, 1)),
    ("x230", (230, 10), (100, 10), (500, 500, 1)),
    ("degenerate", (60, 60), (60, 60), (200, 200, 1)),
    ("degenerate_in", (60, 60), (60, 60), (61, 60, 2)),
]


def gen_segcheck(flavour, out_name):
    """A11 (f64, process_map.py:383-425) or A12 (f32, neuralplanner.py:43-69)."""
    rng = np.random.default_rng(20261018 if flavour == "f64" else 20261019)
    ftype = np.float32 if flavour == "f32" else np.float64
    n_maps, per_map = 48, 48
    obs, cnt = make_maps(rng, n_maps)
    segs, seg_map = [], []
    for m in range(n_maps):
        for i in range(per_map):
            s = rng.uniform(0, 224, 2)
            r = rng.random()
            if r < 0.55:
                e = s + rng.normal(0, 15, 2)
            elif r < 0.8:
                e = rng.uniform(0, 224, 2)
            elif r < 0.85:
                e = s.copy()                       # degenerate
            elif r < 0.95:
                e = rng.uniform(-10, 234, 2)       # bounds quirks
                s = rng.uniform(-10, 234, 2)
            else:
                e = s + rng.normal(0, 1e-3, 2)     # tiny
            segs.append([s[0], s[1], e[0], e[1]])
            seg_map.append(m)
    segs = np.asarray(segs).astype(ftype)
    seg_map = np.asarray(seg_map, dtype=np.int32)
    # sharpen: for ~1/3 of segments, re-tune one circle's radius to an ulp-neighbourhood
    for i in range(len(segs)):
        m = seg_map[i]
        if cnt[m] and rng.random() < 0.35:
            j = int(rng.integers(0, cnt[m]))
            new_r = sharpen_radius(segs[i], obs[m, j], flavour, rng)
            if new_r < 25:                      # keep the maps realistic (no map-covering disks)
                obs[m, j, 2] = new_r
    # extra single-circle maps: sharp + KAT
    sh_segs, sh_circ = sharp_cases(rng, 700, flavour)
    for i in range(len(sh_segs)):
        sh_circ[i, 2] = sharpen_radius(sh_segs[i], sh_circ[i], flavour, rng)
    kat_segs = np.asarray([[k[1][0], k[1][1], k[2][0], k[2][1]] for k in KAT]).astype(ftype)
    kat_circ = np.asarray([k[3] for k in KAT], dtype=np.float64)
    ex_segs = np.concatenate([sh_segs, kat_segs])
    ex_circ = np.concatenate([sh_circ, kat_circ])
    n_ex = len(ex_segs)
    obs2 = np.zeros([n_ex, obs.shape[1], 3])
    obs2[:, 0, :] = ex_circ
    cnt2 = np.ones(n_ex, dtype=np.int32)
    all_obs = np.concatenate([obs, obs2])
    all_cnt = np.concatenate([cnt, cnt2])
    all_segs = np.concatenate([segs, ex_segs])
    all_map = np.concatenate([seg_map, np.arange(n_ex, dtype=np.int32) + n_maps])

    # ---- run the real reference --------------------------------------------------------
    verdict = np.zeros(len(all_segs), dtype=np.uint8)
    if flavour == "f64":
        pm = load_edage()["process_map"]
        for i, sg in enumerate(all_segs):
            m = all_map[i]
            ol = [[float(v) for v in all_obs[m, j]] for j in range(all_cnt[m])]
            # the reference takes (row, col) points and swaps; our stored segs are (x, y)
            s_rc = torch.tensor([sg[1], sg[0]], dtype=torch.float64)
            e_rc = torch.tensor([sg[3], sg[2]], dtype=torch.float64)
            with quiet():
                verdict[i] = bool(pm.collision_check_circle_edge(s_rc, e_rc, ol, CLEAR))
        np.savez_compressed(os.path.join(HERE, out_name), segs_xy=all_segs, seg_map=all_map,
                            obs=all_obs, obs_cnt=all_cnt, clearance=CLEAR, verdict=verdict,
                            dot_mode=detect_dot_mode(), n_kat=len(KAT),
                            kat_names=np.asarray([k[0] for k in KAT]))
    else:
        obc = [[[float(v) for v in all_obs[m, j]] for j in range(all_cnt[m])]
               for m in range(len(all_cnt))]
        ns = load_mpnet_checker(obc)
        steer = np.zeros(len(all_segs), dtype=np.uint8)
        for i, sg in enumerate(all_segs):
            s = torch.from_numpy(np.asarray(sg[0:2], dtype=np.float32))
            e = torch.from_numpy(np.asarray(sg[2:4], dtype=np.float32))
            verdict[i] = bool(ns["collision_check_circle_edge"](s, e, int(all_map[i])))
            steer[i] = ns["steerTo"](s, e, int(all_map[i]))
        # paths: feasibility_check + lvc on the multi-circle maps
        paths, path_map, feas, lvc_out = [], [], [], []
        for m in range(n_maps):
            for rep in range(2):
                L = int(rng.integers(2, 40))
                a, b = rng.uniform(10, 214, 2), rng.uniform(10, 214, 2)
                tt = np.linspace(0, 1, L)[:, None]
                wp = (a + tt * (b - a) + rng.normal(0, 10 if rep else 2, (L, 2))).astype(np.float32)
                if rep and L > 4:
                    wp[L // 2] = wp[L // 2 - 1]     # duplicate waypoint: steerTo dist == 0 branch
                pl = [torch.from_numpy(wp[k].copy()) for k in range(L)]
                feas.append(ns["feasibility_check"](pl, m))
                out = ns["lvc"](pl, m)
                lvc_out.append(np.stack([o.numpy() for o in out]))
                paths.append(wp)
                path_map.append(m)
        np.savez_compressed(
            os.path.join(HERE, out_name), segs_xy=all_segs, seg_map=all_map, obs=all_obs,
            obs_cnt=all_cnt, clearance=CLEAR, verdict=verdict, steer=steer, n_kat=len(KAT),
            kat_names=np.asarray([k[0] for k in KAT]),
            path_pts=np.concatenate(paths), path_off=np.cumsum([0] + [len(p) for p in paths]),
            path_map=np.asarray(path_map, dtype=np.int32), feasible=np.asarray(feas, dtype=np.uint8),
            lvc_pts=np.concatenate(lvc_out), lvc_off=np.cumsum([0] + [len(p) for p in lvc_out]))
    print(out_name, "segments", len(all_segs), "positives", int(verdict.sum()))


# ---------------------------------------------------------------------------------------------
def build_paths(n_paths, seed, clearance, seg_num=10):
    """Run the real Path pipeline under a fixed seed, capturing the random draws by replay."""
    mods = load_edage()
    PathMod = mods["Path"]
    recs = []
    np.random.seed(seed)
    torch.manual_seed(seed)
    tries = 0
    while len(recs) < n_paths:
        tries += 1
        state = np.random.get_state()
        path = PathMod.Path(seg_num=seg_num, poly_order=4, dim=2, clearance=clearance, is_straight=False)
        with quiet():
            path.generate(show_now=False)
        after = np.random.get_state()
        # replay the draws Path.generate consumed: per segment random(1) [straight?], random(1000), random(1)
        np.random.set_state(state)
        draws_straight, draws_y, draws_end = [], [], []
        for i in range(seg_num):
            draws_straight.append(np.random.random(1)[0])
            draws_y.append(np.random.random(1000))
            draws_end.append(np.random.random(1)[0])
        assert np.array_equal(np.random.get_state()[1], after[1])
        np.random.set_state(after)
        captured = {}
        orig_norm = path.space_normalization

        def spy(space, point_trans=False, check_free=False, _o=orig_norm, _c=captured, _p=path):
            _c["space_raw"] = (space[0].cpu().numpy() * 255).round().astype(np.uint8)
            _c["hull_raw"] = np.asarray(_p.ConvexHull).copy()
            _c["PathPoint_raw"] = np.asarray(_p.PathPoint).copy()
            _c["SegPoint_raw"] = np.asarray(_p.SegPoint).copy()
            _c["BoundaryPoint_raw"] = np.asarray(_p.BoundaryPoint).copy()
            return _o(space, point_trans=point_trans, check_free=check_free)

        path.space_normalization = spy
        with quiet():
            path.draw_boundary(show_now=False)
            captured["up"] = np.asarray(path.Boundary.upboundary.point).copy()
            captured["up_dir"] = np.asarray(path.Boundary.upboundary.direction).copy()
            captured["down"] = np.asarray(path.Boundary.downboundary.point).copy()
            captured["init"] = np.asarray(path.Boundary.initboundary).copy()
            captured["end"] = np.asarray(path.Boundary.endboundary).copy()
            # A8 / A9 capture: the isles search_isle returns and the torch.rand(1) values set_obstacles consumes
            orig_isle, orig_rand = path.search_isle, torch.rand
            rand_log = []

            def spy_isle(*a, _o=orig_isle, _c=captured, **k):
                isles = _o(*a, **k)
                _c["isles"] = [np.asarray(b, dtype=np.float64).copy() for b in isles]
                return isles

            def spy_rand(*a, _o=orig_rand, **k):
                v = _o(*a, **k)
                rand_log.append(float(v.reshape(-1)[0]))
                return v

            path.search_isle = spy_isle
            torch.rand = spy_rand
            try:
                ok = path.path_obstacles(resolution=224, map_size=50, map_offset=112)
            finally:
                torch.rand = orig_rand
            captured["rand"] = np.asarray(rand_log, dtype=np.float64)
        if not ok:
            continue
        recs.append(dict(path=path, straight=np.asarray(draws_straight), y=np.asarray(draws_y),
                         end=np.asarray(draws_end), cap=captured))
    return recs


def gen_paths():
    """A1-A8: PathSeg.random, Path.generate, draw_boundary, path_space (corridor paint, hull,
    normalisation) under seeds, c=1 and c=3."""
    out = {}
    k = 0
    for clearance, seed, n in ((1, 11, 3), (3, 12, 2)):
        for rec in build_paths(n, seed, clearance):
            p, cap = rec["path"], rec["cap"]
            pre = "p%d_" % k
            out[pre + "clearance"] = clearance
            out[pre + "draw_straight"] = rec["straight"]
            out[pre + "draw_y"] = rec["y"]
            out[pre + "draw_end"] = rec["end"]
            out[pre + "is_straight"] = np.asarray([s.is_straight for s in p.PathSeg])
            out[pre + "Poly"] = np.asarray([s.Poly for s in p.PathSeg])
            out[pre + "EndPoint"] = np.asarray([float(np.reshape(s.EndPoint, -1)[0]) for s in p.PathSeg])
            out[pre + "SegLength"] = np.asarray([float(np.reshape(s.Length, -1)[0]) for s in p.PathSeg])
            out[pre + "GradSt"] = np.asarray([float(np.reshape(s.GradSt, -1)[0]) for s in p.PathSeg])
            out[pre + "GradEnd"] = np.asarray([float(np.reshape(s.GradEnd, -1)[0]) for s in p.PathSeg])
            out[pre + "SegRotation"] = np.asarray([float(s.Rotation) for s in p.PathSeg])
            out[pre + "SegTranslation"] = np.asarray([np.reshape(s.Translation, 2) for s in p.PathSeg])
            out[pre + "SegPoint_raw"] = cap["SegPoint_raw"]
            out[pre + "PathPoint_raw"] = cap["PathPoint_raw"]
            out[pre + "BoundaryPoint_raw"] = cap["BoundaryPoint_raw"]
            out[pre + "Length"] = float(p.Length)
            for kk in ("up", "up_dir", "down", "init", "end", "space_raw", "hull_raw"):
                out[pre + kk] = cap[kk]
            out[pre + "Rotation"] = float(p.Rotation)
            out[pre + "Translation"] = np.asarray([float(t) for t in p.Translation])
            out[pre + "ConvexHull"] = np.asarray(p.ConvexHull)
            out[pre + "SegPointImage"] = np.asarray(p.SegPointImage)
            out[pre + "PathPoint"] = np.asarray(p.PathPoint)
            out[pre + "BoundaryPoint"] = np.asarray(p.BoundaryPoint)
            out[pre + "Space"] = (p.Space[0].cpu().numpy() * 255).round().astype(np.uint8)
            obs = [[float(o[0]), float(o[1]), float(o[2])] for o in p.obstacles]
            out[pre + "obstacles"] = np.asarray(obs, dtype=np.float64).reshape(-1, 3)
            isles = cap.get("isles", [])
            out[pre + "isle_off"] = np.cumsum([0] + [len(b) for b in isles]).astype(np.int64)
            out[pre + "isle_pts"] = (np.concatenate(isles, axis=0) if isles else np.zeros([0, 2])).astype(np.float64)
            out[pre + "obst_rand"] = cap["rand"]
            k += 1
    out["n_paths"] = k
    np.savez_compressed(os.path.join(HERE, "paths.npz"), **out)
    print("paths.npz", k, "paths")


# ---------------------------------------------------------------------------------------------
def gen_grid():
    """A4 coord_euclidean2image and A5 free_space_bydirection on the real class."""
    PathMod = load_edage()["Path"]
    rng = np.random.default_rng(77)
    p = PathMod.Path(seg_num=1, clearance=1)
    p.Resolution, p.MapSize = 224, 50
    pts = rng.uniform(-60, 60, (4000, 2))
    step = 50 / 224
    # exact ties: (k + 0.5) * step - style values and small integers
    ties = np.asarray([[(k + 0.5) * step, (k - 0.5) * step] for k in range(-40, 40)])
    pts = np.concatenate([pts, ties, np.asarray([[0.0, -0.0], [1e-300, -1e-300]])])
    idx224 = p.coord_euclidean2image(pts, 224)
    idx112 = p.coord_euclidean2image(pts, 112.0)
    p2 = PathMod.Path(seg_num=1, clearance=1)
    p2.Resolution, p2.MapSize = 1024, 50
    idx1024 = p2.coord_euclidean2image(pts, 1024)
    # half-integer pixel-space ties through step 1 (used by the DDA endpoint snap)
    p3 = PathMod.Path(seg_num=1, clearance=1)
    p3.Resolution, p3.MapSize = 1, 1
    halves = np.asarray([[k + 0.5, -(k + 0.5)] for k in range(-8, 9)], dtype=np.float64)
    idx_half = p3.coord_euclidean2image(halves, 0)
    # rays
    n_rays = 600
    x0 = rng.uniform(-52, 52, (n_rays, 2))
    ang = rng.uniform(0, 2 * np.pi, n_rays)
    step_len = 1 / 224 * 50
    dirs = step_len * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    step_num = rng.choice([0.8 * 1 / step_len, 0.8 * 3 / step_len, 2.5, 3.5, 30.0], n_rays)
    cells, offs = [], [0]
    for i in range(n_rays):
        space = torch.zeros([448, 448])
        sp = p.free_space_bydirection(space, x0[i], dirs[i], step_num[i], mapoffset=224)
        nz = torch.nonzero(sp).numpy()
        # order is not recoverable from the painted image; store as sorted set
        nz = nz[np.lexsort((nz[:, 1], nz[:, 0]))]
        cells.append(nz)
        offs.append(offs[-1] + len(nz))
    np.savez_compressed(os.path.join(HERE, "grid.npz"), pts=pts, idx224=idx224, idx112=idx112,
                        idx1024=idx1024, halves=halves, idx_half=idx_half, ray_x0=x0, ray_dir=dirs,
                        ray_step_num=step_num, ray_cells=np.concatenate(cells),
                        ray_off=np.asarray(offs))
    print("grid.npz", len(pts), "points", n_rays, "rays", offs[-1], "cells")


# ---------------------------------------------------------------------------------------------
def gen_mapgen():
    """A10, A13, A14 through the real MapGenerate.generate (P=3 target paths, 9 maps, O=50, c=1
    and P=2, 4 maps, O=20, c=3), capturing draws by replay of the numpy stream."""
    mods = load_edage()
    MG = mods["MapGenerate"]
    out = {}
    gi = 0
    cwd = os.getcwd()
    for (P, O, c, seed) in ((3, 50, 1, 5), (2, 20, 3, 6)):
        tmp = tempfile.mkdtemp(prefix="ppnet_golden_")
        os.chdir(tmp)
        np.random.seed(seed)
        torch.manual_seed(seed)
        with quiet():
            mg = MG.MapGenerate(path_num=P, resolution=224, map_size=50, obstacles_num=O, clearance=c)
        MG.cnt = 0
        # spy boundary_check + generate_map_randomly to record each call's inputs / outputs
        calls = []
        for tp in mg.PathGroup.TargetPaths:
            orig = tp.boundary_check

            def spy_bc(angle, translation, _o=orig, _tp=tp):
                rst, hull = _o(angle, translation)
                calls.append(("bc", mg.PathGroup.TargetPaths.index(_tp), float(np.reshape(angle, -1)[0]),
                              [int(translation[0]), int(translation[1])], bool(rst), np.asarray(hull).copy()))
                return rst, hull

            tp.boundary_check = spy_bc
        orig_gmr = mg.generate_map_randomly

        def spy_gmr(path_point, init, end, length, path_obstacles, index, _o=orig_gmr):
            st = np.random.get_state()
            cand = np.stack([np.random.random(O) * 50, np.random.random(O) * 50,
                             np.random.random(O) * mg.ObstacleSize], axis=1)
            np.random.set_state(st)
            calls.append(("gmr", index, np.asarray(path_point).copy(), cand,
                          np.asarray(path_obstacles, dtype=np.float64).reshape(-1, 3)))
            return _o(path_point=path_point, init=init, end=end, length=length,
                      path_obstacles=path_obstacles, index=index)

        mg.generate_map_randomly = spy_gmr
        with quiet():
            mg.generate(map_num=P * P, folder_path=os.path.join(tmp, "out"), round_index=0)
        problems = [json.loads(l) for l in open(os.path.join(tmp, "unsolved_problems.txt"))]
        os.chdir(cwd)
        pre = "g%d_" % gi
        out[pre + "P"], out[pre + "O"], out[pre + "clearance"] = P, O, c
        for j, tp in enumerate(mg.PathGroup.TargetPaths):
            out[pre + "tp%d_hull" % j] = np.asarray(tp.ConvexHull)
            out[pre + "tp%d_SegPointImage" % j] = np.asarray(tp.SegPointImage)
            out[pre + "tp%d_PathPoint" % j] = np.asarray(tp.PathPoint)
            out[pre + "tp%d_Length" % j] = float(tp.Length)
            obs = [[float(o[0]), float(o[1]), float(o[2])] for o in tp.obstacles]
            out[pre + "tp%d_obstacles" % j] = np.asarray(obs, dtype=np.float64).reshape(-1, 3)
        bcs = [cl for cl in calls if cl[0] == "bc"]
        out[pre + "bc_path"] = np.asarray([b[1] for b in bcs], dtype=np.int32)
        out[pre + "bc_angle_arg"] = np.asarray([b[2] for b in bcs])          # = -angle drawn
        out[pre + "bc_trans_arg"] = np.asarray([b[3] for b in bcs], dtype=np.int64)   # = [t1, t0]
        out[pre + "bc_ok"] = np.asarray([b[4] for b in bcs], dtype=np.uint8)
        out[pre + "bc_hull_off"] = np.cumsum([0] + [len(b[5]) for b in bcs])
        out[pre + "bc_hull_out"] = np.concatenate([b[5] for b in bcs])
        gm = [cl for cl in calls if cl[0] == "gmr"]
        out[pre + "map_index"] = np.asarray([g[1] for g in gm], dtype=np.int64)
        out[pre + "map_pathpoint"] = np.asarray([g[2] for g in gm])
        out[pre + "map_cand"] = np.asarray([g[3] for g in gm])
        out[pre + "map_pathobs_off"] = np.cumsum([0] + [len(g[4]) for g in gm])
        out[pre + "map_pathobs"] = np.concatenate([g[4] for g in gm]) if gm else np.zeros([0, 3])
        # MapLabel: [label, angle, translation, segpoint, pathpoint]
        out[pre + "label_angle"] = np.asarray([float(np.reshape(l[1], -1)[0]) for l in mg.MapLabel])
        out[pre + "label_translation"] = np.asarray([[int(l[2][0]), int(l[2][1])] for l in mg.MapLabel])
        out[pre + "label_segpoint"] = np.asarray([l[3] for l in mg.MapLabel])
        out[pre + "label_pathpoint"] = np.asarray([l[4] for l in mg.MapLabel])
        # problems: accepted random obstacles come first, then the path obstacles
        out[pre + "prob_obs_off"] = np.cumsum([0] + [len(p["Obstacles"]) for p in problems])
        out[pre + "prob_obs"] = np.concatenate(
            [np.asarray(p["Obstacles"], dtype=np.float64).reshape(-1, 3) for p in problems])
        out[pre + "prob_init"] = np.asarray([p["Init"] for p in problems])
        out[pre + "prob_end"] = np.asarray([p["End"] for p in problems])
        out[pre + "prob_length"] = np.asarray([p["Length"] for p in problems])
        out[pre + "prob_index"] = np.asarray([p["Index"] for p in problems])
        gi += 1
    out["n_groups"] = gi
    np.savez_compressed(os.path.join(HERE, "mapgen.npz"), **out)
    print("mapgen.npz", gi, "groups")


# ---------------------------------------------------------------------------------------------
def gen_misc():
    """A16 add_init_end_single, convehull.py fixture, PathSeg seeded smoke value."""
    mods = load_edage()
    pm = mods["process_map"]
    rng = np.random.default_rng(5)
    imgs_in, imgs_out, inits, ends = [], [], [], []
    for i in range(6):
        base = ((np.add.outer(np.arange(224), np.arange(224)) % 7) / 7).astype(np.float32)
        img = torch.from_numpy(np.stack([base, base * 0.5, 1 - base]).copy())   # low entropy
        init = rng.uniform(-2, 226, 2) if i < 4 else np.asarray([0.5, 223.5])
        end = rng.uniform(-2, 226, 2) if i < 4 else np.asarray([2.5, 1.5])
        imgs_in.append(img.numpy().copy())
        o = pm.add_init_end_single(img, init, end)
        imgs_out.append(o.numpy().copy())
        inits.append(init)
        ends.append(end)
    from scipy.spatial import ConvexHull
    np.random.seed(0)
    pts = np.random.rand(30, 2)                 # convehull.py:5-7
    hv = ConvexHull(pts).vertices
    np.random.seed(0)
    ps = mods["PathSeg"].PathSeg(4, 2)
    ps.random()
    np.savez_compressed(os.path.join(HERE, "misc.npz"), img_in=np.asarray(imgs_in),
                        img_out=np.asarray(imgs_out), init=np.asarray(inits), end=np.asarray(ends),
                        hull_pts=pts, hull_vertices=hv, seg0_poly=ps.Poly,
                        seg0_end=float(ps.EndPoint[0]), seg0_len=float(np.reshape(ps.Length, -1)[0]))
    print("misc.npz")


# ---------------------------------------------------------------------------------------------
def gen_masks():
    """N1-N3 ("next" rows): corridor placement (MapGenerate.py:102-106), the label-mask rasterisers
    process_map.generate_gen_path / generate_seg_space (:148-191) and extract_path (:293-365), all through the real
    reference functions."""
    from PIL import Image
    mods = load_edage()
    MG, pm = mods["MapGenerate"], mods["process_map"]
    out = {}
    P, O, c, seed = 2, 20, 1, 21
    tmp = tempfile.mkdtemp(prefix="ppnet_golden_")
    cwd = os.getcwd()
    os.chdir(tmp)
    np.random.seed(seed)
    torch.manual_seed(seed)
    with quiet():
        mg = MG.MapGenerate(path_num=P, resolution=224, map_size=50, obstacles_num=O, clearance=c)
    MG.cnt = 0
    placed = []
    orig_affine = MG.T.functional.affine

    def spy_affine(img, *a, **k):
        r = orig_affine(img, *a, **k)
        placed.append((r[0].cpu().numpy() * 255).round().astype(np.uint8))
        return r

    MG.T.functional.affine = spy_affine
    try:
        with quiet():
            mg.generate(map_num=P * P, folder_path=os.path.join(tmp, "out"), round_index=0)
    finally:
        MG.T.functional.affine = orig_affine
    n = len(mg.MapLabel)
    assert len(placed) == n
    spaces = [(tp.Space[0].cpu().numpy() * 255).round().astype(np.uint8) for tp in mg.PathGroup.TargetPaths]
    out["n1_space"] = np.asarray(spaces)                                   # [P,224,224] Path.Space per target path
    out["n1_path"] = np.asarray([(g // P) % P for g in range(n)], dtype=np.int32)
    out["n1_angle"] = np.asarray([float(np.reshape(l[1], -1)[0]) for l in mg.MapLabel])
    out["n1_translation"] = np.asarray([[int(l[2][0]), int(l[2][1])] for l in mg.MapLabel], dtype=np.int64)
    out["n1_placed"] = np.asarray(placed)                                  # [n,224,224] path_space after rotate + affine
    pathpoints = [np.asarray(l[4]) for l in mg.MapLabel]
    segpoints = [np.asarray(l[3]) for l in mg.MapLabel]
    out["n2_pathpoint"] = np.asarray(pathpoints)
    out["n2_segpoint"] = np.asarray(segpoints)
    # N2: generate_gen_path writes {root}/{index}.png ('L'); generate_seg_space writes palette PNGs of the placed corridor
    root_p, root_s = os.path.join(tmp, "mask_path"), os.path.join(tmp, "mask_space")
    pm.generate_gen_path(pathpoints, 0, root=root_p)
    out["n2_gen_path"] = np.asarray([np.asarray(Image.open(os.path.join(root_p, "%d.png" % i))) for i in range(n)])
    pm.imgviz.label_colormap = lambda: (np.arange(768) % 251).astype(np.uint8).reshape(256, 3)
    pil_spaces = [Image.fromarray(np.stack([sp] * 3, axis=2)) for sp in spaces]
    rot = [float(a) for a in out["n1_angle"]]
    tr = [[int(t[0]), int(t[1])] for t in out["n1_translation"]]
    with quiet():
        pm.generate_seg_space(pil_spaces, pathpoints, rot, tr, 0, root=root_s)
    out["n2_seg_space"] = np.asarray([np.asarray(Image.open(os.path.join(root_s, "%d.png" % i))) for i in range(n)])
    # N3: extract_path on heat-maps made from the label path (3x3 dilation, values fading with distance)
    pm.time.time = lambda: 0.0                                              # neutralise the 1 s wall-clock timeout
    pm.time_synchronized = lambda: 0.0
    ds = 2
    k = 0
    for i in range(n):
        for variant in range(2):
            heat = np.zeros([224, 224], dtype=np.float64)
            for q in pathpoints[i]:
                r0, c0 = int(np.round(q[0])), int(np.round(q[1]))
                for dj in range(-2, 3):
                    for dk in range(-2, 3):
                        if 0 <= r0 + dj < 224 and 0 <= c0 + dk < 224:
                            v = 255 - 40 * max(abs(dj), abs(dk)) - (0 if variant == 0 else (r0 * 7 + c0 * 3) % 23)
                            heat[r0 + dj, c0 + dk] = max(heat[r0 + dj, c0 + dk], v)
            if variant == 1:
                heat[100:110, :] = 0                                       # a gap: the walk must fail or detour
            img = Image.fromarray(heat.astype(np.uint8), mode="L")
            small = img.resize((int(img.size[0] / ds), int(img.size[1] / ds)), Image.BILINEAR)
            small = MG.T.ToTensor()(small).squeeze().numpy()
            with quiet():
                ok, path = pm.extract_path(img, init_state=segpoints[i][0], end_state=segpoints[i][10], down_sample_rate=ds)
            out["n3_%d_mask" % k] = small.astype(np.float32)
            out["n3_%d_init" % k] = np.asarray(segpoints[i][0], dtype=np.float64)
            out["n3_%d_end" % k] = np.asarray(segpoints[i][10], dtype=np.float64)
            out["n3_%d_ok" % k] = bool(ok)
            out["n3_%d_path" % k] = path.numpy().astype(np.float64) if ok else np.zeros([0, 2])
            k += 1
    out["n3_count"], out["n3_ds"] = k, ds
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **out)
    print("masks.npz", n, "maps", k, "extract_path cases, ok:", [bool(out["n3_%d_ok" % j]) for j in range(k)])


# ---------------------------------------------------------------------------------------------
def gen_planner_masks():
    """N4: gerated_by_planners.generated_by_planners (EDaGe-PP/gerated_by_planners.py:56-161) run unmodified on a
    synthetic solved-problems file (OMPL itself is not available): mask_space / mask_path PNGs per solution."""
    import importlib
    from PIL import Image
    from oracle.ref_loader import find_reference
    mods = load_edage()
    sys.path.insert(0, os.path.join(find_reference(), "EDaGe-PP"))
    try:
        gp = importlib.import_module("gerated_by_planners")
    finally:
        sys.path.pop(0)
    gp.imgviz.label_colormap = lambda: (np.arange(768) % 251).astype(np.uint8).reshape(256, 3)
    rng = np.random.default_rng(31)
    sols = []
    for L in (2, 3, 5):
        a, b = rng.uniform(20, 60, 2), rng.uniform(150, 205, 2)
        t = np.linspace(0, 1, L)[:, None]
        w = a + t * (b - a) + np.concatenate([[[0, 0]], rng.normal(0, 12, (L - 2, 2)), [[0, 0]]]) if L > 2 else a + t * (b - a)
        sols.append(w)
    sols.append(np.asarray([[3.2, 100.7], [120.4, 221.9], [222.6, 40.1]]))          # touches the borders
    tmp = tempfile.mkdtemp(prefix="ppnet_golden_")
    cwd = os.getcwd()
    os.chdir(tmp)
    with open("solved.txt", "w") as f:
        for w in sols:
            prob = {"Obstacles": [[50.0, 60.0, 8.0]],
                    "Solution": [{"Planner": "BITstar", "Waypoint": [[float(x), float(y)] for x, y in w], "Time": 1.0}]}
            f.write(json.dumps(prob) + "\n")
    with quiet():
        gp.generated_by_planners("solved.txt")
    out = {"n": len(sols), "wp": np.concatenate(sols), "off": np.cumsum([0] + [len(w) for w in sols]).astype(np.int64)}
    out["mask_space"] = np.asarray([np.asarray(Image.open("data_BITstar/mask_space/%d.png" % i)) for i in range(len(sols))])
    out["mask_path"] = np.asarray([np.asarray(Image.open("data_BITstar/mask_path/%d.png" % i)) for i in range(len(sols))])
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "planner_masks.npz"), **out)
    print("planner_masks.npz", len(sols), "solutions; space px", [(m != 0).sum() for m in out["mask_space"]],
          "path px", [(m != 0).sum() for m in out["mask_path"]], "values", np.unique(out["mask_space"]), np.unique(out["mask_path"]))


def gen_extract_image():
    """N3's driver, process_map.extract_path_image (:452-506), run through the REAL reference on a four-map dataset the
    real MapGenerate wrote: heat-maps made from the label paths (one with a gap: extraction fails; one problem gets an
    extra obstacle on its path: collision), `NUM_PER_FOLDER` set to the dataset's size, `plot_solution` (matplotlib) and
    the 1 s wall-clock timeout neutralised.  Fixture: the inputs (masks, segment points, problems) and the exact text of
    solved_problems.txt."""
    from PIL import Image
    import json
    mods = load_edage()
    MG, pm = mods["MapGenerate"], mods["process_map"]
    P, O, c, seed = 2, 20, 1, 33
    tmp = tempfile.mkdtemp(prefix="ppnet_golden_xi_")
    cwd = os.getcwd()
    os.chdir(tmp)
    np.random.seed(seed)
    torch.manual_seed(seed)
    with quiet():
        mg = MG.MapGenerate(path_num=P, resolution=224, map_size=50, obstacles_num=O, clearance=c)
    MG.cnt = 0
    folder = os.path.join(tmp, "original_data", "0")
    os.makedirs(os.path.join(tmp, "original_data"))                         # generate() makes the folder, GMM/ and data/ itself
    with quiet():
        mg.generate(map_num=P * P, folder_path=folder, round_index=0)
    n = len(mg.MapLabel)
    torch.save(mg.MapLabel, os.path.join(folder, "data", "MapLabel"))
    problems = [json.loads(l) for l in open(os.path.join(tmp, "unsolved_problems.txt"))]
    assert len(problems) == n and [p["Index"] for p in problems] == list(range(n))
    segpoints = [np.asarray(l[3], dtype=np.float64) for l in mg.MapLabel]
    pathpoints = [np.asarray(l[4], dtype=np.float64) for l in mg.MapLabel]
    # problem 1 gets an obstacle on the middle of its path: its extracted path must be rejected
    mid = pathpoints[1][len(pathpoints[1]) // 2]
    problems[1]["Obstacles"].append([float(mid[1]), float(mid[0]), 6.0])
    utxt = os.path.join(tmp, "unsolved_edit.txt")
    with open(utxt, "w") as f:
        for p in problems:
            f.write(json.dumps(p) + "\n")
    mask_root = os.path.join(tmp, "masks")
    os.makedirs(mask_root)
    heats = []
    for i in range(n):
        heat = np.zeros([224, 224], dtype=np.float64)
        for q in pathpoints[i]:
            r0, c0 = int(np.round(q[0])), int(np.round(q[1]))
            for dj in range(-2, 3):
                for dk in range(-2, 3):
                    if 0 <= r0 + dj < 224 and 0 <= c0 + dk < 224:
                        v = 255 - 40 * max(abs(dj), abs(dk)) - (r0 * 7 + c0 * 3) % 23
                        heat[r0 + dj, c0 + dk] = max(heat[r0 + dj, c0 + dk], v)
        if i == 2:
            r_mid = int(np.round(pathpoints[i][len(pathpoints[i]) // 2][0]))
            heat[max(0, r_mid - 6):r_mid + 6, :] = 0                            # a gap: extraction fails
        heats.append(heat.astype(np.uint8))
        Image.fromarray(heats[-1], mode="L").save(os.path.join(mask_root, "%d.png" % i))
    pm.NUM_PER_FOLDER = n
    pm.time.time = lambda: 0.0
    pm.time_synchronized = lambda: 0.0
    pm.plot_solution = lambda **k: None
    result_root = os.path.join(tmp, "result")
    real_load = torch.load                  # the reference predates torch.load's weights_only=True default (torch 2.6)
    torch.load = lambda *a, **k: real_load(*a, **dict(k, weights_only=False))
    try:
        with quiet():
            pm.extract_path_image(mask_root, os.path.join(tmp, "original_data"), result_root, utxt, c / 50 * 224)
    finally:
        torch.load = real_load
    solved = open(os.path.join(result_root, "solved_problems.txt")).read() if os.path.exists(os.path.join(result_root, "solved_problems.txt")) else ""
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "extract_image.npz"), masks=np.asarray(heats), segpoint=np.asarray(segpoints),
                        unsolved=np.frombuffer(open(utxt, "rb").read(), dtype=np.uint8),
                        solved=np.frombuffer(solved.encode(), dtype=np.uint8), clearance=c / 50 * 224)
    print("extract_image.npz", n, "maps, solved lines:", solved.count("\n"),
          "indices:", [json.loads(l)["Index"] for l in solved.splitlines()])


GENS = {
    # run a second time under OPENBLAS_CORETYPE=HASWELL (set before numpy loads) to pin the un-fused ddot model:
    #   OPENBLAS_CORETYPE=HASWELL python make_golden.py --only segcheck_f64   ->  segcheck_f64_unfused.npz
    "segcheck_f64": lambda: gen_segcheck("f64", "segcheck_f64.npz" if detect_dot_mode() == orc.DOT_FUSED_SKX
                                         else "segcheck_f64_unfused.npz"),
    "segcheck_f32": lambda: gen_segcheck("f32", "segcheck_f32.npz"),
    "grid": gen_grid,
    "paths": gen_paths,
    "mapgen": gen_mapgen,
    "misc": gen_misc,
    "masks": gen_masks,
    "planner_masks": gen_planner_masks,
    "extract_image": gen_extract_image,
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    for name, fn in GENS.items():
        if args.only in (None, name):
            fn()
