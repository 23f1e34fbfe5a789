"""GPU parity tests for target-path synthesis (A1-A3, A5-A9): ppnet_path_synthesize through the C ABI against the
reference-generated golden fixture (parity mode: the reference's own draws) and against the oracle pipeline
(Philox mode).  Floating-point outputs: 1e-5 relative as BASELINE.json states (np.polyfit is LAPACK least
squares, not bit-stable); integer outputs (cells, hull, isles, painted corridor) exact."""
import numpy as np
import pytest
import torch

from oracle import philox
from oracle import ppnet_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "these tests need a B200"
    from ppnet_b200 import ops as _ops
    return _ops


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def close(got, want, what, rtol=RTOL, scale=None):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    s = np.abs(want).max() if scale is None else scale
    err = np.abs(got - want).max() if got.size else 0.0
    assert err <= rtol * max(s, 1e-300), (what, err, s)


def _golden_group(g, ks):
    S, HM, MR = 10, 64, 64
    n = len(ks)
    y = np.stack([g["p%d_draw_y" % k] for k in ks])
    ue = np.stack([g["p%d_draw_end" % k] for k in ks])
    st = np.stack([g["p%d_is_straight" % k] for k in ks]).astype(np.uint8)
    hull = np.zeros([n, HM, 2], dtype=np.int32)
    hcnt = np.zeros(n, dtype=np.int32)
    rnd = np.zeros([n, MR], dtype=np.float32)
    rcnt = np.zeros(n, dtype=np.int32)
    for i, k in enumerate(ks):
        h = g["p%d_hull_raw" % k]
        hull[i, :len(h)] = h.astype(np.int32)
        hcnt[i] = len(h)
        r = g["p%d_obst_rand" % k]
        rnd[i, :len(r)] = r.astype(np.float32)
        rcnt[i] = len(r)
    return S, HM, y, ue, st, hull, hcnt, rnd, rcnt


@pytest.mark.parametrize("use_ref_hull", [True, False])
def test_path_synthesis_parity_mode_vs_reference(ops, golden, use_ref_hull):
    g = golden("paths")
    for c, ks in ((1, [0, 1, 2]), (3, [3, 4])):
        S, HM, y, ue, st, hull, hcnt, rnd, rcnt = _golden_group(g, ks)
        kw = dict(in_hull=dev(hull), in_hull_cnt=dev(hcnt)) if use_ref_hull else {}
        out = ops.path_synthesize(0, len(ks), seg_num=S, poly_order=4, clearance=float(c), map_size=50.0, resolution=224,
                                  hmax=HM, pomax=32, want_space=True, in_y=dev(y), in_uend=dev(ue), in_straight=dev(st),
                                  in_obst_rand=dev(rnd), in_obst_rand_cnt=dev(rcnt), **kw)
        torch.cuda.synchronize()
        o = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in vars(out).items()}
        for i, k in enumerate(ks):
            pre = "p%d_" % k
            # A1
            close(o["poly"][i], g[pre + "Poly"], "Poly")
            close(o["endpoint"][i], g[pre + "EndPoint"], "EndPoint")
            assert np.array_equal(o["is_straight"][i].astype(bool), g[pre + "is_straight"])
            close(o["seg_length"][i], g[pre + "SegLength"], "SegLength")
            close(o["grad_st"][i], g[pre + "GradSt"], "GradSt")
            close(o["grad_end"][i], g[pre + "GradEnd"], "GradEnd")
            # A2
            close(o["seg_rot"][i], g[pre + "SegRotation"], "SegRotation")
            close(o["seg_trans"][i], g[pre + "SegTranslation"], "SegTranslation")
            close(o["segpoint_raw"][i], g[pre + "SegPoint_raw"], "SegPoint")
            close(o["pathpoint_raw"][i], g[pre + "PathPoint_raw"], "PathPoint_raw")
            close(o["length"][i], float(g[pre + "Length"]), "Length")
            # A3
            close(o["up"][i], g[pre + "up"], "up")
            close(o["up_dir"][i], g[pre + "up_dir"], "up_dir")
            close(o["down"][i], g[pre + "down"], "down")
            close(o["cap_init"][i], g[pre + "init"], "init")
            close(o["cap_end"][i], g[pre + "end"], "end")
            close(o["boundary_raw"][i], g[pre + "BoundaryPoint_raw"], "BoundaryPoint_raw")
            # A5: the painted corridor, bit-exact set of cells
            assert np.array_equal(o["space_raw"][i], g[pre + "space_raw"]), "corridor cells differ"
            # A6: same vertex set as scipy/Qhull (cyclic order may start elsewhere)
            H = int(o["hull_cnt"][i])
            want_h = g[pre + "hull_raw"].astype(np.int64)
            got_h = o["hull_raw"][i, :H]
            assert H == len(want_h)
            assert set(map(tuple, got_h.tolist())) == set(map(tuple, want_h.tolist()))
            # A7
            close(o["rotation"][i], float(g[pre + "Rotation"]), "Rotation")
            close(o["translation"][i], g[pre + "Translation"], "Translation", scale=224.0)
            close(o["segpoint_img"][i], g[pre + "SegPointImage"], "SegPointImage", scale=224.0)
            close(o["pathpoint"][i], g[pre + "PathPoint"], "PathPoint", scale=224.0)
            close(o["boundary"][i], g[pre + "BoundaryPoint"], "BoundaryPoint", scale=224.0)
            # A7 mask part: Path.Space after torchvision rotate + affine + crop, pixel for pixel
            assert np.array_equal(o["space"][i], g[pre + "Space"]), "normalised corridor differs"
            if use_ref_hull:
                close(o["hull"][i, :H], g[pre + "ConvexHull"], "ConvexHull", scale=224.0)
                # A8: the same isles in the same order (the order follows the hull's vertex order)
                off = g[pre + "isle_off"]
                assert int(o["isle_cnt"][i]) == len(off) - 1
                for t in range(len(off) - 1):
                    lo, hi = o["isle"][i, t]
                    assert np.allclose(o["pathpoint"][i, lo:hi], g[pre + "isle_pts"][off[t]:off[t + 1]], rtol=0, atol=1e-6)
                # A9: same obstacles from the same torch.rand values
                want_o = g[pre + "obstacles"]
                assert int(o["obs_cnt"][i]) == len(want_o)
                assert int(o["obst_rand_used"][i]) == len(g[pre + "obst_rand"])
                close(o["obs"][i, :len(want_o)], want_o, "obstacles", scale=224.0)
                assert int(o["status"][i]) == 0
            else:
                # own hull order: same isle SET
                off = g[pre + "isle_off"]
                got = sorted(tuple(x) for x in o["isle"][i, :int(o["isle_cnt"][i])].tolist())
                want = sorted(tuple(x) for x in orc.search_isle(g[pre + "PathPoint"], g[pre + "ConvexHull"], c, 50, 224))
                assert got == want


def _oracle_path(seed, gid, S, c, hmax=64):
    forced, straight, ys, ue = philox.path_draws(seed, gid, S)
    draws = philox.path_obst_draws(seed, gid, 400)
    return forced, orc.synthesize_path(ys, ue, straight, c, 50.0, 224, draws, path_straight=forced)


def test_path_synthesis_philox_mode_vs_oracle_and_shard_invariance(ops):
    seed, S, c, n = 20261018, 10, 1.0, 6
    out = ops.path_synthesize(100, n, seg_num=S, clearance=c, seed=seed, want_space=True)
    a = ops.path_synthesize(100, 2, seg_num=S, clearance=c, seed=seed)
    b = ops.path_synthesize(102, 4, seg_num=S, clearance=c, seed=seed)
    torch.cuda.synchronize()
    for name in ("pathpoint", "hull", "obs", "obs_cnt", "isle", "segpoint_img", "length"):
        whole = getattr(out, name).cpu().numpy()
        parts = np.concatenate([getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy()])
        assert np.array_equal(whole, parts), name           # bit-identical for any sharding of the path ids
    o = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in vars(out).items()}
    for i in range(n):
        forced, w = _oracle_path(seed, 100 + i, S, c)
        assert bool(o["path_straight"][i]) == forced
        close(o["poly"][i], np.asarray([s["Poly"] for s in w["segs"]]), "Poly")
        close(o["pathpoint_raw"][i], w["chain"]["PathPoint"], "PathPoint_raw")
        close(o["boundary_raw"][i], w["bnd"]["BoundaryPoint"], "BoundaryPoint_raw")
        close(o["ray_x0"][i], w["ray_x0"], "ray_x0")
        close(o["ray_dir"][i], w["ray_dir"], "ray_dir", scale=1.0)
        assert np.array_equal(o["cells"][i], w["cells"])
        H = int(o["hull_cnt"][i])
        assert np.array_equal(o["hull_raw"][i, :H], np.asarray(w["hull_raw"]))      # both CCW from the smallest vertex
        space = orc.corridor_paint(w["ray_x0"], w["ray_dir"], w["step_num"], 50.0, 224)
        assert np.array_equal(o["space_raw"][i], space)
        close(o["pathpoint"][i], w["norm"]["PathPoint"], "PathPoint", scale=224.0)
        close(o["hull"][i, :H], w["norm"]["ConvexHull"], "ConvexHull", scale=224.0)
        assert [tuple(x) for x in o["isle"][i, :int(o["isle_cnt"][i])].tolist()] == [tuple(x) for x in w["isles"]]
        assert int(o["obs_cnt"][i]) == len(w["obstacles"])
        assert int(o["obst_rand_used"][i]) == w["used"]
        close(o["obs"][i, :len(w["obstacles"])], w["obstacles"], "obstacles", scale=224.0)


def test_synthesized_bank_feeds_the_generator(ops):
    """PathGroup.generate -> MapGenerate.generate on the device end to end: every emitted map keeps the A9 / A14
    clearance invariant and its hull inside the map."""
    c, R, M, O = 1.0, 224, 50.0, 50
    paths = ops.path_synthesize(0, 16, clearance=c, resolution=R, map_size=M, seed=7)
    assert int(paths.status.max().item()) & 3 == 0
    bank = paths.to_bank()
    gen = ops.generate_maps(bank, 0, 160, 10, O, R, M, 5.0, c, seed=7)
    torch.cuda.synchronize()
    assert int(gen.valid.sum().item()) == 160
    pp = gen.pathpt.cpu().numpy()
    obs = gen.obs.cpu().numpy()
    cnt = gen.obs_cnt.cpu().numpy()
    rc = gen.rand_cnt.cpu().numpy()
    c_px = c / M * R
    for m in range(0, 160, 7):
        odd = pp[m, 1::2]
        for k in range(cnt[m]):
            x, y, r = obs[m, k]
            d = np.sqrt((odd[:, 0] - y) ** 2 + (odd[:, 1] - x) ** 2).min()
            assert d > r + c_px - (1e-4 if k >= rc[m] else 0.0), (m, k, d, r)
        assert pp[m].min() > -1e-9 and pp[m].max() < R + 1e-9
