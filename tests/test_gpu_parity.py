"""GPU parity tests: the CUDA path (through the C ABI) against the reference-generated golden
fixtures and against the oracle on seeded inputs.  Bit-exact for verdicts / indices / cells."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import ppnet_oracle as orc

pytestmark = pytest.mark.gpu

CLEAR = 1 / 50 * 224


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


def _csr(seg_map, n_maps):
    """Sort segments by map -> (order, seg_off)."""
    order = np.argsort(seg_map, kind="stable")
    counts = np.bincount(seg_map, minlength=n_maps)
    return order, np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


def _rc(segs_xy):
    return np.stack([segs_xy[:, 1], segs_xy[:, 0], segs_xy[:, 3], segs_xy[:, 2]], axis=1)


# ------------------------------------------------------------------ A11
@pytest.mark.parametrize("fixture", ["segcheck_f64", "segcheck_f64_unfused"])
def test_segcheck_f64_golden(ops, golden, fixture):
    g = golden(fixture)                                   # reference outputs under the fused / un-fused OpenBLAS ddot
    order, off = _csr(g["seg_map"], len(g["obs_cnt"]))
    pts = _rc(g["segs_xy"])[order]
    v = ops.segcheck_edage_f64(dev(pts), dev(g["obs"]), dev(g["obs_cnt"]), float(g["clearance"]),
                               seg_off=dev(off), dot_mode=int(g["dot_mode"]))
    assert np.array_equal(v.cpu().numpy(), g["verdict"][order])
    # the other ddot model must differ somewhere (sharp cases) and must equal the oracle in that mode
    other = 1 - int(g["dot_mode"])
    v2 = ops.segcheck_edage_f64(dev(pts), dev(g["obs"]), dev(g["obs_cnt"]), float(g["clearance"]),
                                seg_off=dev(off), dot_mode=other).cpu().numpy()
    want2 = c_oracle.segcheck_f64(pts, g["seg_map"][order], g["obs"], g["obs_cnt"], float(g["clearance"]),
                                  dot_mode=other)
    assert np.array_equal(v2, want2)
    assert (v2 != g["verdict"][order]).any()


def _config2_inputs(rng, n_maps, segs_per_map, dtype, omax=50, sharp_frac=0.0):
    obs = np.zeros([n_maps, omax, 3])
    obs[..., 0] = rng.uniform(0, 224, (n_maps, omax))
    obs[..., 1] = rng.uniform(0, 224, (n_maps, omax))
    obs[..., 2] = rng.uniform(0, 22.4, (n_maps, omax))
    cnt = rng.integers(0, omax + 1, n_maps).astype(np.int32)
    s = rng.uniform(0, 224, (n_maps * segs_per_map, 2))
    e = s + rng.normal(0, 15, s.shape)
    segs = np.concatenate([s, e], axis=1).astype(dtype)
    return segs, obs, cnt


@pytest.mark.parametrize("dot_mode", [0, 1])
def test_segcheck_f64_vs_oracle_seeded(ops, dot_mode):
    rng = np.random.default_rng(42 + dot_mode)
    n_maps, spm = 300, 1024
    segs, obs, cnt = _config2_inputs(rng, n_maps, spm, np.float64)
    seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), spm)
    v = ops.segcheck_edage_f64(dev(segs), dev(obs), dev(cnt), CLEAR, dot_mode=dot_mode).cpu().numpy()
    want = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, CLEAR, dot_mode=dot_mode, threads=8)
    assert np.array_equal(v, want)
    assert 0.05 < v.mean() < 0.6


def test_segcheck_f64_edge_cases(ops):
    # empty maps, maps without circles, ragged CSR with empty rows, NaN / inf coordinates
    obs = np.zeros([4, 3, 3])
    obs[1, 0] = [50, 14, 2]
    obs[3, :, :] = [[50, 14, 2], [np.nan, 3, 1], [10, 10, np.inf]]
    cnt = np.asarray([0, 1, 0, 3], dtype=np.int32)
    pts = np.asarray([[10, 10, 10, 100], [10, 10, 10, 100], [np.nan, 10, 10, 100], [10, 10, 10, np.inf],
                      [10, 10, 10, 100], [60, 60, 60, 60]], dtype=np.float64)
    off = np.asarray([0, 1, 4, 4, 6], dtype=np.int64)
    seg_map = np.asarray([0, 1, 1, 1, 3, 3], dtype=np.int32)
    v = ops.segcheck_edage_f64(dev(pts), dev(obs), dev(cnt), CLEAR, seg_off=dev(off)).cpu().numpy()
    want = c_oracle.segcheck_f64(pts, seg_map, obs, cnt, CLEAR)
    assert np.array_equal(v, want)
    for i in range(len(pts)):
        m = seg_map[i]
        assert bool(v[i]) == orc.segcheck_edage_f64(pts[i, :2], pts[i, 2:], obs[m, :cnt[m]].tolist(), CLEAR)
    # zero segments
    z = ops.segcheck_edage_f64(torch.zeros([0, 4], dtype=torch.float64, device="cuda"), dev(obs), dev(cnt), CLEAR,
                               seg_off=dev(np.zeros(5, dtype=np.int64)))
    assert z.numel() == 0


# ------------------------------------------------------------------ A12
def test_segcheck_f32_golden(ops, golden):
    g = golden("segcheck_f32")
    order, off = _csr(g["seg_map"], len(g["obs_cnt"]))
    v, st = ops.segcheck_mpnet_f32(dev(g["segs_xy"][order]), dev(g["obs"]), dev(g["obs_cnt"]),
                                   float(g["clearance"]), seg_off=dev(off), want_steer=True)
    assert np.array_equal(v.cpu().numpy(), g["verdict"][order])
    assert np.array_equal(st.cpu().numpy(), g["steer"][order])


def test_segcheck_f32_vs_oracle_seeded(ops):
    rng = np.random.default_rng(7)
    n_maps, spm = 300, 1024
    segs, obs, cnt = _config2_inputs(rng, n_maps, spm, np.float32)
    segs[::97, 2:] = segs[::97, :2]                      # dist == 0 -> steerTo == 1 whatever the verdict
    seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), spm)
    v, st = ops.segcheck_mpnet_f32(dev(segs), dev(obs), dev(cnt), CLEAR, want_steer=True)
    want, want_st = c_oracle.segcheck_f32(segs, seg_map, obs, cnt, CLEAR, threads=8)
    assert np.array_equal(v.cpu().numpy(), want)
    assert np.array_equal(st.cpu().numpy(), want_st)


def _margin_cases(rng, n, ftype):
    """One circle per segment, placed so that the decision sits at a *relative* distance 10^U(-17,-2) from a
    threshold of the fast path: |dis| vs thr, foot vs either end, |e-o| vs thr; plus degenerate / extreme
    operands (zero and tiny segments, centre on an endpoint, tiny and huge radii, huge and non-finite coords)."""
    scale = rng.choice([1.0, 1.0, 1.0, 1e-3, 1e3, 1e6], n)
    s = rng.uniform(5, 219, (n, 2))
    e = s + rng.normal(0, 30, (n, 2))
    d = e - s
    L = np.linalg.norm(d, axis=1)
    nrm = np.stack([d[:, 1], -d[:, 0]], axis=1) / L[:, None]
    thr = rng.uniform(2.3, 25, n)
    pert = 10.0 ** rng.uniform(-17, -2, n) * rng.choice([-1, 1], n)
    kind = rng.integers(0, 8, n)
    t = rng.uniform(0.05, 0.95, n) * L                 # foot position along the segment
    off = thr * rng.uniform(0.05, 0.95, n)             # lateral offset
    # 0: |dis| ~ thr (foot inside)   1: foot ~ s   2: foot ~ e   3: |e-o| ~ thr   4: foot ~ s and |dis| ~ thr
    off = np.where((kind == 0) | (kind == 4), thr * (1 + pert), off)
    t = np.where((kind == 1) | (kind == 4), L * pert, t)
    t = np.where(kind == 2, L * (1 + pert), t)
    side = rng.choice([-1.0, 1.0], n)
    o = s + d / L[:, None] * t[:, None] + nrm * (off * side)[:, None]
    ang = rng.uniform(0, 2 * np.pi, n)
    ov = e + np.stack([np.cos(ang), np.sin(ang)], axis=1) * (thr * (1 + pert))[:, None]
    o = np.where((kind == 3)[:, None], ov, o)
    # 5: centre exactly on s / e / the line   6, 7: random generic
    on = rng.integers(0, 3, n)
    o5 = np.where((on == 0)[:, None], s, np.where((on == 1)[:, None], e, s + d * rng.uniform(0, 1, (n, 1))))
    o = np.where((kind == 5)[:, None], o5, o)
    o = np.where((kind >= 6)[:, None], rng.uniform(0, 224, (n, 2)), o)
    segs = np.concatenate([s, e], axis=1) * scale[:, None]
    circ = np.concatenate([o * scale[:, None], ((thr - CLEAR / 2) * scale)[:, None]], axis=1)
    # degenerate / extreme rows
    idx = rng.permutation(n)[:n // 10]
    for i in idx:
        c = rng.integers(0, 8)
        if c == 0: segs[i, 2:] = segs[i, :2]                                   # zero length
        elif c == 1: segs[i, 2:] = segs[i, :2] + rng.normal(0, 1e-9, 2)        # tiny
        elif c == 2: circ[i, 2] = -CLEAR / 2 + 10.0 ** rng.uniform(-14, -3)    # tiny thr
        elif c == 3: circ[i, 2] = 10.0 ** rng.uniform(3, 30)                   # huge radius
        elif c == 4: segs[i] *= 10.0 ** rng.uniform(8, 30)                     # huge coordinates (f32: may overflow to inf)
        elif c == 5: segs[i, rng.integers(0, 4)] = rng.choice([np.nan, np.inf, -np.inf])
        elif c == 6: circ[i, rng.integers(0, 3)] = rng.choice([np.nan, np.inf, -np.inf])
        else: circ[i, :2] *= 10.0 ** rng.uniform(8, 30)
    with np.errstate(over="ignore", invalid="ignore"):
        return segs.astype(ftype), circ


@pytest.mark.parametrize("flavour", ["f64_fused", "f64_unfused", "f32"])
def test_segcheck_filter_margins_vs_oracle(ops, flavour):
    """The division-free decisions and the bin culling must never change a verdict: 400 k pairs whose outcome
    hinges on quantities within 1e-17..1e-2 (relative) of the fast path's thresholds, plus degenerate operands."""
    rng = np.random.default_rng({"f64_fused": 1, "f64_unfused": 2, "f32": 3}[flavour])
    n_maps, spm, omax = 25000, 16, 3
    ftype = np.float32 if flavour == "f32" else np.float64
    segs, circ = _margin_cases(rng, n_maps * spm, ftype)
    # map m holds the circles of its first 3 segments: every segment meets its own circle (rows 0..2) or
    # sharp circles built for a neighbour
    obs = np.ascontiguousarray(circ.reshape(n_maps, spm, 3)[:, :omax])
    cnt = rng.integers(0, omax + 1, n_maps).astype(np.int32)
    cnt[: n_maps // 2] = omax
    seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), spm)
    for bound in (224.0, 1e9):
        if flavour == "f32":
            v, st = ops.segcheck_mpnet_f32(dev(segs), dev(obs), dev(cnt), CLEAR, bound=bound, want_steer=True)
            want, want_st = c_oracle.segcheck_f32(segs, seg_map, obs, cnt, CLEAR, bound=bound, threads=8)
            assert np.array_equal(st.cpu().numpy(), want_st)
        else:
            mode = 0 if flavour == "f64_fused" else 1
            v = ops.segcheck_edage_f64(dev(segs), dev(obs), dev(cnt), CLEAR, bound=bound, dot_mode=mode)
            want = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, CLEAR, bound=bound, dot_mode=mode, threads=8)
        bad = np.nonzero(v.cpu().numpy() != want)[0]
        assert len(bad) == 0, (flavour, bound, bad[:10], segs[bad[:3]], obs[seg_map[bad[:3]]])
    assert 0.05 < want.mean() < 0.95


def test_mpnet_feasible_and_lvc_golden(ops, golden):
    g = golden("segcheck_f32")
    args = (dev(g["path_pts"]), dev(g["path_off"]), dev(g["path_map"]), dev(g["obs"]), dev(g["obs_cnt"]),
            float(g["clearance"]))
    feas, chk = ops.path_feasible_f32(*args)
    assert np.array_equal(feas.cpu().numpy(), g["feasible"])
    _, want_chk = c_oracle.feasible(g["path_pts"], g["path_off"], g["path_map"], g["obs"], g["obs_cnt"],
                                    float(g["clearance"]))
    assert np.array_equal(chk.cpu().numpy(), want_chk)
    out, out_len = ops.lvc_f32(*args)
    out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
    po, lo = g["path_off"], g["lvc_off"]
    for p in range(len(g["path_map"])):
        want = g["lvc_pts"][lo[p]:lo[p + 1]]
        assert out_len[p] == len(want), p
        assert np.array_equal(out[po[p]:po[p] + out_len[p]], want), p


def test_mpnet_config3_vs_oracle(ops):
    """BASELINE config 3 shape: problems x ragged f32 waypoint lists, <= 50 circles each."""
    rng = np.random.default_rng(11)
    n_prob = 400
    obs = np.zeros([n_prob, 50, 3])
    obs[..., 0] = rng.uniform(0, 224, (n_prob, 50))
    obs[..., 1] = rng.uniform(0, 224, (n_prob, 50))
    obs[..., 2] = rng.uniform(0, 9, (n_prob, 50))
    cnt = rng.integers(0, 51, n_prob).astype(np.int32)
    lens = rng.integers(2, 65, n_prob)
    lens[:3] = [2, 3, 160]
    wps = []
    for L in lens:
        a, b = rng.uniform(5, 219, 2), rng.uniform(5, 219, 2)
        t = np.linspace(0, 1, L)[:, None]
        wps.append((a + t * (b - a) + rng.normal(0, 10, (L, 2))).astype(np.float32))
    wp = np.concatenate(wps)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pm = np.arange(n_prob, dtype=np.int32)
    feas, chk = ops.path_feasible_f32(dev(wp), dev(off), dev(pm), dev(obs), dev(cnt), CLEAR)
    want, want_chk = c_oracle.feasible(wp, off, pm, obs, cnt, CLEAR, threads=4)
    assert np.array_equal(feas.cpu().numpy(), want)
    assert np.array_equal(chk.cpu().numpy(), want_chk)
    out, out_len = ops.lvc_f32(dev(wp), dev(off), dev(pm), dev(obs), dev(cnt), CLEAR)
    w_out, w_len = c_oracle.lvc(wp, off, pm, obs, cnt, CLEAR, threads=8)
    assert np.array_equal(out_len.cpu().numpy(), w_len)
    out = out.cpu().numpy()
    for p in range(n_prob):
        assert np.array_equal(out[off[p]:off[p] + w_len[p]], w_out[off[p]:off[p] + w_len[p]]), p
    assert 0 < want.mean() < 1


# ------------------------------------------------------------------ A14
def test_clearance_filter_golden(ops, golden):
    g = golden("mapgen")
    for gi in range(int(g["n_groups"])):
        pre = "g%d_" % gi
        c = float(g[pre + "clearance"])
        acc, out, cnt = ops.clearance_filter_f64(dev(g[pre + "map_pathpoint"]), dev(g[pre + "map_cand"]), 50.0,
                                                 224.0, c)
        out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
        off, pof = g[pre + "prob_obs_off"], g[pre + "map_pathobs_off"]
        for i in range(len(cnt)):
            n_path = pof[i + 1] - pof[i]
            want = g[pre + "prob_obs"][off[i]:off[i + 1] - n_path]      # what the reference wrote to JSON
            assert cnt[i] == len(want)
            assert np.array_equal(out[i, :cnt[i]], want)


def test_clearance_filter_vs_oracle_sharp(ops):
    rng = np.random.default_rng(3)
    M, O = 64, 50
    pp = rng.uniform(20, 200, (M, 1000, 2))
    cand = np.stack([rng.uniform(0, 50, (M, O)), rng.uniform(0, 50, (M, O)), rng.uniform(0, 5, (M, O))], axis=2)
    thr_c = 1.0 / 50 * 224
    for m in range(M):
        for j in range(1, O, 2):                       # radii within +-2 ulp of the threshold
            q = cand[m, j, :2] / 50 * 224
            d = pp[m, 1::2] - q
            r_img = np.sqrt(np.min(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])) - thr_c
            if r_img > 0:
                r = r_img / 224 * 50
                for _ in range(int(rng.integers(0, 3))):
                    r = np.nextafter(r, np.inf if rng.random() < 0.5 else -np.inf)
                cand[m, j, 2] = r
    acc, out, cnt = ops.clearance_filter_f64(dev(pp), dev(cand), 50.0, 224.0, 1.0)
    w_acc, w_out, w_cnt = c_oracle.clearance_filter(pp, cand, 50, 224, 1.0, threads=4)
    assert np.array_equal(acc.cpu().numpy(), w_acc)
    assert np.array_equal(cnt.cpu().numpy(), w_cnt)
    assert np.array_equal(out.cpu().numpy(), w_out)
    assert 0 < w_acc.mean() < 1
    # ragged sizes: odd Np, O not a multiple of 32, O = 0
    for np_, O2 in ((7, 3), (1001, 33), (2, 1), (1000, 0)):
        pp2 = rng.uniform(0, 224, (5, np_, 2))
        cd2 = np.stack([rng.uniform(0, 50, (5, O2)), rng.uniform(0, 50, (5, O2)), rng.uniform(0, 5, (5, O2))], axis=2)
        acc, out, cnt = ops.clearance_filter_f64(dev(pp2), dev(cd2), 50.0, 224.0, 3.0)
        w_acc, w_out, w_cnt = c_oracle.clearance_filter(pp2, cd2, 50, 224, 3.0)
        assert np.array_equal(acc.cpu().numpy(), w_acc) and np.array_equal(cnt.cpu().numpy(), w_cnt)
        assert np.array_equal(out.cpu().numpy(), w_out)


# ------------------------------------------------------------------ A4 / A5
def test_grid_index_golden(ops, golden):
    g = golden("grid")
    pts = dev(g["pts"])
    for res, off, key in ((224, 224, "idx224"), (224, 112.0, "idx112"), (1024, 1024, "idx1024")):
        idx = ops.grid_index_f64(pts, 50.0, float(res), float(off)).cpu().numpy()
        assert np.array_equal(idx, g[key])
    idx = ops.grid_index_f64(dev(g["halves"]), 1.0, 1.0, 0.0).cpu().numpy()
    assert np.array_equal(idx, g["idx_half"])
    # odd element count + big seeded batch against the oracle
    rng = np.random.default_rng(1)
    big = rng.uniform(-60, 60, 2_000_001)
    idx = ops.grid_index_f64(dev(big), 50.0, 224.0, 224.0).cpu().numpy()
    assert np.array_equal(idx, np.rint(big / (50 / 224) + 224.0).astype(np.int32))


def test_corridor_paint_golden(ops, golden):
    g = golden("grid")
    off = g["ray_off"]
    # every golden ray painted into its own image
    n = len(g["ray_x0"])
    sp = ops.corridor_paint(dev(g["ray_x0"].reshape(n, 1, 2)), dev(g["ray_dir"].reshape(n, 1, 2)),
                            dev(g["ray_step_num"]), 50.0, 224.0, 224.0, 448, 448).cpu().numpy()
    for i in range(n):
        want = {tuple(c) for c in g["ray_cells"][off[i]:off[i + 1]]}
        assert {tuple(c) for c in np.argwhere(sp[i])} == want, i
    # whole corridors of real reference paths: identical 448x448 images
    gp = golden("paths")
    x0s, drs, sns, wants = [], [], [], []
    for k in range(int(gp["n_paths"])):
        pre = "p%d_" % k
        c = float(gp[pre + "clearance"])
        bnd = dict(init=gp[pre + "init"], end=gp[pre + "end"], up=gp[pre + "up"], up_dir=gp[pre + "up_dir"],
                   down=gp[pre + "down"], down_dir=-1 * gp[pre + "up_dir"])
        x0, dr = orc.corridor_rays(bnd, gp[pre + "SegPoint_raw"][-1], 50, 224)
        x0s.append(x0), drs.append(dr), sns.append(0.8 * c / (1 / 224 * 50)), wants.append(gp[pre + "space_raw"])
    sp = ops.corridor_paint(dev(np.stack(x0s)), dev(np.stack(drs)), dev(np.asarray(sns)), 50.0, 224.0, 224.0,
                            448, 448).cpu().numpy()
    for k in range(len(wants)):
        assert np.array_equal(sp[k], wants[k]), k


def test_compact_survivors(ops):
    rng = np.random.default_rng(5)
    for n in (0, 1, 31, 1024, 1025, 3_000_001):
        f = (rng.random(n) < 0.37).astype(np.uint8)
        if n > 2000:
            f[5000:9000] = 1                                  # whole CTAs without survivors
            f[20000:26000] = 0                                # ... and full ones
        for keep in (0, 1):
            idx, cnt = ops.compact_u8(dev(f) if n else torch.zeros([0], dtype=torch.uint8, device="cuda"), keep)
            want = np.nonzero(f == keep)[0]
            assert int(cnt.item()) == len(want)
            assert np.array_equal(idx[:len(want)].cpu().numpy(), want)


@pytest.mark.parametrize("flavour", ["f64", "f32"])
def test_verdict_monotonicity_properties(ops, flavour):
    """Size-independent properties of the verdict (an OR over circles of two monotone tests): adding a circle or
    growing radii / clearance can only turn 'free' into 'collision'; removing every circle leaves only the bounds rule."""
    rng = np.random.default_rng(12)
    M, spm, omax = 500, 1024, 40
    ft = np.float64 if flavour == "f64" else np.float32
    segs, obs, cnt = _config2_inputs(rng, M, spm, ft, omax=omax)
    cnt = np.minimum(cnt, omax - 1).astype(np.int32)
    fn = (lambda s, o, c, cl: ops.segcheck_edage_f64(s, o, c, cl)) if flavour == "f64" else \
         (lambda s, o, c, cl: ops.segcheck_mpnet_f32(s, o, c, cl))
    S, Ob, Cn = dev(segs), dev(obs), dev(cnt)
    v = fn(S, Ob, Cn, CLEAR)
    obs2 = obs.copy()
    obs2[np.arange(M), cnt] = np.stack([rng.uniform(0, 224, M), rng.uniform(0, 224, M), rng.uniform(0, 22, M)], axis=1)
    v_more = fn(S, dev(obs2), dev(cnt + 1), CLEAR)
    assert bool((v_more >= v).all()) and bool((v_more > v).any())
    obs3 = obs.copy()
    obs3[..., 2] *= 1.25
    assert bool((fn(S, dev(obs3), Cn, CLEAR) >= v).all())
    assert bool((fn(S, Ob, Cn, 2 * CLEAR) >= v).all())
    v_none = fn(S, Ob, dev(np.zeros(M, dtype=np.int32)), CLEAR).cpu().numpy()
    a = segs.astype(np.float64)
    oob = (a[:, 0] < 0) | (a[:, 1] > 224) | (a[:, 2] < 0) | (a[:, 3] > 224)      # same rule on the raw inputs in both flavours
    assert np.array_equal(v_none.astype(bool), oob)
