"""GPU parity tests, part 2: hull, boundary check, Philox draws, GMM, raster, DDA and the fused map generator."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import philox as ph
from oracle import ppnet_oracle as orc

pytestmark = pytest.mark.gpu
SEED = 0x5050_4E45_54


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


# ------------------------------------------------------------------ A6 hull
def test_hull_golden_and_scipy(ops, golden):
    from scipy.spatial import ConvexHull
    g = golden("paths")
    pts = []
    for k in range(int(g["n_paths"])):
        pts.append(orc.grid_index_vec(g["p%d_PathPoint_raw" % k], 50, 224, 224))
    rng = np.random.default_rng(5)
    for _ in range(40):                                   # random integer walks with duplicates / collinear runs
        steps = rng.integers(-3, 4, (1000, 2))
        pts.append(np.cumsum(steps, axis=0) + 224)
    pts = np.stack(pts).astype(np.int32)
    hull, cnt = ops.hull2d_i32(dev(pts), hmax=64)
    hull, cnt = hull.cpu().numpy(), cnt.cpu().numpy()
    for k in range(len(pts)):
        h = hull[k, :cnt[k]]
        want = orc.hull2d(pts[k])
        assert np.array_equal(h, want), k                 # same vertices, same CCW order, same start
        sv = pts[k][ConvexHull(pts[k].astype(np.float64)).vertices]
        assert {tuple(p) for p in h} == {tuple(p) for p in sv}, k
    for k in range(int(g["n_paths"])):                    # the reference's own hull (cyclic shift allowed)
        want = [tuple(p) for p in g["p%d_hull_raw" % k].astype(np.int64)]
        got = [tuple(p) for p in hull[k, :cnt[k]]]
        j = want.index(got[0])
        assert got == want[j:] + want[:j]
    # degenerate inputs
    deg = np.zeros([3, 8, 2], dtype=np.int32)
    deg[1, :, 0] = np.arange(8)                           # collinear
    deg[2] = [[0, 0], [4, 0], [4, 4], [0, 4], [2, 2], [4, 0], [2, 0], [0, 2]]
    hull, cnt = ops.hull2d_i32(dev(deg), hmax=8)
    assert cnt.cpu().tolist() == [1, 2, 4]
    assert hull[2, :4].cpu().tolist() == [[0, 0], [4, 0], [4, 4], [0, 4]]


# ------------------------------------------------------------------ A10
def test_boundary_check_golden(ops, golden):
    g = golden("mapgen")
    for gi in range(int(g["n_groups"])):
        pre = "g%d_" % gi
        P = int(g[pre + "P"])
        hmax = 40
        hull = np.zeros([P, hmax, 2])
        hcnt = np.zeros(P, dtype=np.int32)
        for j in range(P):
            h = g[pre + "tp%d_hull" % j]
            hull[j, :len(h)] = h
            hcnt[j] = len(h)
        ok, ho = ops.boundary_check(dev(hull), dev(hcnt), dev(g[pre + "bc_path"]), dev(g[pre + "bc_angle_arg"]),
                                    dev(g[pre + "bc_trans_arg"].astype(np.float64)), 224.0, want_hull=True)
        assert np.array_equal(ok.cpu().numpy(), g[pre + "bc_ok"])
        ho = ho.cpu().numpy()
        off = g[pre + "bc_hull_off"]
        for i in range(len(g[pre + "bc_ok"])):
            np.testing.assert_allclose(ho[i, :off[i + 1] - off[i]], g[pre + "bc_hull_out"][off[i]:off[i + 1]],
                                       rtol=1e-9, atol=1e-9)


def test_boundary_check_vs_oracle_bulk(ops, golden):
    """Verdict mismatch rate vs the oracle on 20 000 random placements (device sincos differs from glibc by <= 1 ulp;
    SURVEY 8(a) A10 accepts <= 1e-6 and asks for the rate to be reported)."""
    g = golden("mapgen")
    hull_np = g["g0_tp0_hull"]
    rng = np.random.default_rng(8)
    n = 20000
    ang = -(rng.random(n) * 360 - 180)
    tr = (rng.random((n, 2)) * 224 - 112).astype(np.int64).astype(np.float64)
    hull = np.zeros([1, 40, 2])
    hull[0, :len(hull_np)] = hull_np
    ok = ops.boundary_check(dev(hull), dev(np.asarray([len(hull_np)], dtype=np.int32)), None, dev(ang), dev(tr),
                            224.0).cpu().numpy()
    want = np.asarray([c_oracle.boundary_check(hull_np, ang[i], tr[i, 0], tr[i, 1], 224)[0] for i in range(n)])
    mism = int((ok.astype(bool) != want).sum())
    print("boundary_check verdict mismatches: %d / %d" % (mism, n))
    assert mism <= 1
    assert 0.05 < want.mean() < 0.95


# ------------------------------------------------------------------ Philox / GMM
def test_uniform_matches_philox_oracle(ops):
    u = ops.uniform_f64(SEED, ph.STREAM_UNIFORM, 1000, 33, 7).cpu().numpy()
    assert np.array_equal(u, ph.uniform(SEED, ph.STREAM_UNIFORM, 1000, 33, 7))
    # shard invariance: two halves == the whole
    a = ops.uniform_f64(SEED, ph.STREAM_UNIFORM, 1000, 16, 7).cpu().numpy()
    b = ops.uniform_f64(SEED, ph.STREAM_UNIFORM, 1016, 17, 7).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b]), u)


def test_gmm_params_and_samples(ops):
    from scipy import stats
    mean, std, w = ops.gmm_params(SEED, 10, 2, 70.0, 5.0)
    m_o, s_o, w_o = ph.gmm_params(SEED, 10, 2, 70.0, 5.0)
    assert np.array_equal(mean.cpu().numpy(), m_o) and np.array_equal(std.cpu().numpy(), s_o)
    assert np.array_equal(w.cpu().numpy(), w_o)
    n = 200_000
    x, comp = ops.gmm_sample(SEED, 0, n, mean, std, w, want_comp=True)
    x, comp = x.cpu().numpy(), comp.cpu().numpy()
    # same counters -> same components, same values to float32 transcendental accuracy
    xo, co = ph.gmm_sample(SEED, 0, 4096, m_o, s_o, w_o)
    assert np.array_equal(comp[:4096], co)
    np.testing.assert_allclose(x[:4096], xo, rtol=2e-4, atol=2e-4)
    # shard invariance
    x2 = ops.gmm_sample(SEED, 1000, 500, mean, std, w).cpu().numpy()
    assert np.array_equal(x2, x[1000:1500])
    # statistical parity with the reference distribution (MixtureSameFamily of the same parameters), alpha = 0.01:
    # chi^2 on the component frequencies, KS on each marginal against the analytic mixture CDF
    w64 = w_o.astype(np.float64)
    p = w64 / w64.sum()
    chi = stats.chisquare(np.bincount(comp, minlength=10), p * n)
    assert chi.pvalue > 0.01, chi
    for d in range(2):
        sub = x[::20, d].astype(np.float64)
        ks = stats.kstest(sub, lambda v: orc.gmm_marginal_cdf(v, w_o, m_o, s_o, d))
        assert ks.pvalue > 0.01, (d, ks)
    # and against torch's own MixtureSameFamily sampler with the same parameters (two-sample KS)
    torch.manual_seed(0)
    mix = torch.distributions.Categorical(torch.from_numpy(w_o))
    compd = torch.distributions.Independent(torch.distributions.Normal(torch.from_numpy(m_o), torch.from_numpy(s_o)), 1)
    ref = torch.distributions.MixtureSameFamily(mix, compd).sample([10000]).numpy()
    for d in range(2):
        ks2 = stats.ks_2samp(x[:10000, d], ref[:, d])
        assert ks2.pvalue > 0.01, (d, ks2)


@pytest.mark.parametrize("K,D", [(1, 2), (2, 2), (15, 2), (16, 2), (17, 2), (64, 2), (10, 1), (10, 3), (16, 6)])
def test_gmm_component_pick_and_dims(ops, K, D):
    """Every order / dimension class of the sampler (planar fast path, padded 16-entry search, long scan, extra Philox
    blocks for dims >= 2) picks the oracle's components and lands on its values."""
    mean, std, w = ops.gmm_params(SEED + K, K, D, 70.0, 5.0)
    m_o, s_o, w_o = ph.gmm_params(SEED + K, K, D, 70.0, 5.0)
    assert np.array_equal(mean.cpu().numpy(), m_o) and np.array_equal(w.cpu().numpy(), w_o)
    x, comp = ops.gmm_sample(SEED, 77, 4096, mean, std, w, want_comp=True)
    xo, co = ph.gmm_sample(SEED, 77, 4096, m_o, s_o, w_o)
    assert np.array_equal(comp.cpu().numpy(), co)
    assert len(np.unique(co)) == min(K, len(np.unique(co))) and co.max() <= K - 1
    np.testing.assert_allclose(x.cpu().numpy(), xo, rtol=2e-4, atol=2e-4)


# ------------------------------------------------------------------ raster + DDA
def test_raster_row_ends_on_the_pixel_lattice(ops):
    """The row-span shortcut (SFU sqrt + certainty band, raster.cuh) must hand every tie to the exact rule: centres on
    the half-pixel lattice with Pythagorean / integer / half-integer radii put many pixel centres exactly ON the
    outline (dx^2 + dy^2 == r^2), others 1 ulp and 1e-4 px off it; huge, tiny and off-screen disks ride along."""
    R, M, omax = 224, 16, 64
    rng = np.random.default_rng(11)
    radii = np.asarray([1.0, 2.0, 2.5, 5.0, 6.5, 10.0, 13.0, 12.5, 25.0, 0.5, 1.5, np.sqrt(2.0) / 2, np.sqrt(50.0), 65.0])
    obs = np.zeros([M, omax, 3])
    obs[..., 0] = rng.integers(-20, 2 * R + 40, (M, omax)) * 0.5
    obs[..., 1] = rng.integers(-20, 2 * R + 40, (M, omax)) * 0.5
    obs[..., 2] = radii[rng.integers(0, len(radii), (M, omax))]
    obs[1::4, :, 2] = np.nextafter(obs[1::4, :, 2], 0)                  # 1 ulp inside the tie
    obs[2::4, :, 2] = np.nextafter(obs[2::4, :, 2], 1e9)                # 1 ulp outside
    obs[3::4, :, 2] += rng.uniform(-2e-3, 2e-3, obs[3::4, :, 2].shape)  # inside the certainty band, either side
    obs[0, 0] = [112.0, 112.0, 1e7]
    obs[0, 1] = [1e15, 5.0, 1e15]
    obs[0, 2] = [-1e6, 50.5, 1e6 + 30.0]
    obs[0, 3] = [50.5, 50.5, 1e-300]
    obs[0, 4] = [50.5, 50.5, 1e300]
    obs[4, 0] = [1e9, 1e9, 5.0]
    cnt = np.full(M, omax, np.int32)
    for inflate in (0.0, 0.5, 2.24):
        got = ops.raster_circles_bits(dev(obs[1:]), dev(cnt[1:]), R, inflate).cpu().numpy().view(np.uint32)
        want = c_oracle.raster_circles_bits(obs[1:], cnt[1:], R, inflate, threads=4)
        assert np.array_equal(got, want), inflate
    got = ops.raster_circles_bits(dev(obs[:1, 5:]), dev(cnt[:1] - 5), R, 0.0).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, c_oracle.raster_circles_bits(obs[:1, 5:], cnt[:1] - 5, R, 0.0, threads=1))
    for k in range(5):                                                    # the extreme disks one at a time
        o1 = np.ascontiguousarray(obs[:1, k:k + 1])
        got = ops.raster_circles_bits(dev(o1), dev(np.ones(1, np.int32)), R, 0.0).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, c_oracle.raster_circles_bits(o1, np.ones(1, np.int32), R, 0.0, threads=1)), k


@pytest.mark.parametrize("R,omax", [(224, 50), (1024, 400), (33, 5)])
def test_raster_and_dda_vs_oracle(ops, R, omax):
    rng = np.random.default_rng(R)
    M = 12 if R < 1000 else 4
    obs = np.zeros([M, omax, 3])
    obs[..., 0] = rng.uniform(-0.05 * R, 1.05 * R, (M, omax))
    obs[..., 1] = rng.uniform(-0.05 * R, 1.05 * R, (M, omax))
    obs[..., 2] = rng.uniform(0, 0.1 * R * (0.3 if R > 1000 else 1), (M, omax))
    obs[0, 0] = [R / 2, R / 2, 0.0]                        # r = 0 (+ inflate) and degenerate entries
    obs[0, 1] = [np.nan, 3, 4]
    obs[0, 2] = [10.5, 10.5, 0.5]                          # exactly tangent to pixel centres
    cnt = rng.integers(0, omax + 1, M).astype(np.int32)
    cnt[0] = omax
    for inflate in (0.0, 2.24):
        bits = ops.raster_circles_bits(dev(obs), dev(cnt), R, inflate).cpu().numpy().view(np.uint32)
        want = c_oracle.raster_circles_bits(obs, cnt, R, inflate, threads=4)
        assert np.array_equal(bits, want), inflate
    occ = np.unpackbits(want.view(np.uint8), bitorder="little").mean()
    assert 0.02 < occ < 0.98
    n_per = 4096 if R > 1000 else 700
    s = rng.uniform(-4, R + 4, (M * n_per, 2))
    e = s + rng.normal(0, 0.15 * R, (M * n_per, 2))
    segs = np.concatenate([s, e], axis=1).astype(np.float32)
    segs[::50, 2:] = segs[::50, :2]
    segs[::77] = np.floor(segs[::77]) + 0.5
    segs[5] = [np.nan, 1, 2, 3]
    segs[6] = [3, 3, 1e30, -1e30]
    segs[7] = [R / 2, R / 2, 3e9, 17]
    sm = np.repeat(np.arange(M, dtype=np.int32), n_per)
    v, fh = ops.dda_gridcheck(dev(want.view(np.int32)), R, dev(segs))
    wv, wfh = c_oracle.dda_gridcheck(want, R, segs, sm, threads=4)
    assert np.array_equal(v.cpu().numpy(), wv)
    assert np.array_equal(fh.cpu().numpy(), wfh)
    # ragged CSR
    off = np.sort(rng.integers(0, len(segs) + 1, M - 1))
    off = np.concatenate([[0], off, [len(segs)]]).astype(np.int64)
    sm2 = np.repeat(np.arange(M, dtype=np.int32), np.diff(off))
    v2, fh2 = ops.dda_gridcheck(dev(want.view(np.int32)), R, dev(segs), seg_off=dev(off))
    wv2, wfh2 = c_oracle.dda_gridcheck(want, R, segs, sm2, threads=4)
    assert np.array_equal(v2.cpu().numpy(), wv2) and np.array_equal(fh2.cpu().numpy(), wfh2)


def test_dda_vs_analytic_verdict_disagreement_is_boundary_only(ops):
    """the DDA verdict on maps rasterised from the same circles (inflated by clearance/2) may differ from the
    analytic A12 verdict only for segments that graze a disk boundary within ~1 pixel.  Report the rate."""
    rng = np.random.default_rng(2)
    M, spm, omax = 64, 1024, 50
    obs = np.zeros([M, omax, 3])
    obs[..., 0] = rng.uniform(0, 224, (M, omax))
    obs[..., 1] = rng.uniform(0, 224, (M, omax))
    obs[..., 2] = rng.uniform(0, 22.4, (M, omax))
    cnt = np.full(M, omax, dtype=np.int32)
    s = rng.uniform(8, 216, (M * spm, 2))
    e = np.clip(s + rng.normal(0, 15, s.shape), 1, 223)
    segs = np.concatenate([s, e], axis=1).astype(np.float32)
    clear = 1 / 50 * 224
    bits = ops.raster_circles_bits(dev(obs), dev(cnt), 224, clear / 2)
    v_dda = ops.dda_gridcheck(bits, 224, dev(segs), want_first=False).cpu().numpy()
    v_an = ops.segcheck_mpnet_f32(dev(segs), dev(obs), dev(cnt), clear).cpu().numpy()
    dis = v_dda != v_an
    print("DDA vs analytic disagreement: %.4f (analytic positives %.3f)" % (dis.mean(), v_an.mean()))
    # the analytic test ignores the START vertex (reference quirk) while the grid sees it; exclude those
    so = segs[:, None, 0:2].reshape(M, spm, 1, 2) - obs[:, None, :, 0:2]
    start_in = (np.sqrt((so ** 2).sum(-1)) < obs[:, None, :, 2] + clear / 2 + 1.0).any(-1).reshape(-1)
    assert dis[~start_in].mean() < 0.03


# ------------------------------------------------------------------ fused generator
def _bank_from_golden(g, pre, ops, hmax=40, pomax=32):
    P = int(g[pre + "P"])
    pp = np.stack([g[pre + "tp%d_PathPoint" % j] for j in range(P)])
    sp = np.stack([g[pre + "tp%d_SegPointImage" % j] for j in range(P)])
    hull = np.zeros([P, hmax, 2])
    hcnt = np.zeros(P, dtype=np.int32)
    pobs = np.zeros([P, pomax, 3])
    pcnt = np.zeros(P, dtype=np.int32)
    for j in range(P):
        h, o = g[pre + "tp%d_hull" % j], g[pre + "tp%d_obstacles" % j]
        hull[j, :len(h)], hcnt[j] = h, len(h)
        pobs[j, :len(o)], pcnt[j] = o, len(o)
    return ops.PathBank(dev(pp), dev(sp), dev(hull), dev(hcnt), dev(pobs), dev(pcnt)), P


def test_generator_parity_mode_vs_reference(ops, golden):
    """Feed the reference's own draws (angle, translation, candidate circles) of a real MapGenerate.generate run:
    labels within 1e-5 rel (measured ~1e-13), accepted-obstacle sets bit-exact, placed path obstacles ~1e-9."""
    g = golden("mapgen")
    for gi in range(int(g["n_groups"])):
        pre = "g%d_" % gi
        bank, P = _bank_from_golden(g, pre, ops)
        n, O, c = len(g[pre + "map_index"]), int(g[pre + "O"]), float(g[pre + "clearance"])
        assert list(g[pre + "map_index"]) == list(range(n))
        out = ops.generate_maps(bank, 0, n, reps=P, obstacles_num=O, clearance=c,
                                in_angle=dev(g[pre + "label_angle"]),
                                in_trans=dev(g[pre + "label_translation"].astype(np.int32)),
                                in_cand=dev(g[pre + "map_cand"]))
        torch.cuda.synchronize()
        assert out.valid.cpu().numpy().all() and (out.tries.cpu().numpy() == 1).all()
        np.testing.assert_allclose(out.pathpt.cpu().numpy(), g[pre + "label_pathpoint"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(out.segpt.cpu().numpy(), g[pre + "label_segpoint"], rtol=1e-5, atol=1e-9)
        assert np.abs(out.pathpt.cpu().numpy() - g[pre + "label_pathpoint"]).max() < 1e-9
        obs, cnt, rc = out.obs.cpu().numpy(), out.obs_cnt.cpu().numpy(), out.rand_cnt.cpu().numpy()
        off, pof = g[pre + "prob_obs_off"], g[pre + "map_pathobs_off"]
        for i in range(n):
            n_path = pof[i + 1] - pof[i]
            want = g[pre + "prob_obs"][off[i]:off[i + 1]]
            assert cnt[i] == len(want) and rc[i] == len(want) - n_path
            assert np.array_equal(obs[i, :rc[i]], want[:rc[i]])                      # random obstacles: bit-exact
            np.testing.assert_allclose(obs[i, rc[i]:cnt[i]], want[rc[i]:], rtol=1e-9, atol=1e-7)
        cts = out.counters.cpu().numpy()
        assert cts[0] == n and cts[1] == n and cts[2] == rc.sum() and cts[3] == n
        # the occupancy bitmap equals the oracle raster of the emitted obstacle list
        bits = out.bits.cpu().numpy().view(np.uint32)
        want_bits = c_oracle.raster_circles_bits(obs, cnt, 224, 0.0)
        assert np.array_equal(bits, want_bits)


def _oracle_generate(g, pre, P, O, c, seed, gidx, max_tries=4096):
    """Oracle composition for Philox mode: sequential rejection loop + place + clearance filter."""
    j = (gidx // P) % P
    hull = g[pre + "tp%d_hull" % j]
    for t in range(max_tries):
        ang, t0, t1 = ph.placement_draw(seed, gidx, t, 224)
        ok, _ = orc.boundary_check(hull, -ang, [t1, t0], 224)
        if ok:
            break
    else:
        return None
    pp = orc.place_points(g[pre + "tp%d_PathPoint" % j], ang, [t0, t1], 224)
    sp = orc.place_points(g[pre + "tp%d_SegPointImage" % j], ang, [t0, t1], 224)
    cand = ph.candidates(seed, gidx, O, 50.0, 5.0)
    acc, out = orc.clearance_filter(pp, cand, 50, 224, c)
    po = orc.place_obstacles(g[pre + "tp%d_obstacles" % j], ang, [t0, t1], 224)
    return dict(angle=ang, trans=(t0, t1), tries=t + 1, pathpt=pp, segpt=sp, rand=out, pobs=po)


def test_generator_philox_mode_vs_oracle_and_shard_invariance(ops, golden):
    g = golden("mapgen")
    pre = "g0_"
    bank, P = _bank_from_golden(g, pre, ops)
    O, c, n, map0 = 50, 1.0, 96, 1000
    out = ops.generate_maps(bank, map0, n, reps=P, obstacles_num=O, clearance=c, seed=SEED)
    torch.cuda.synchronize()
    ang, tr, tries = out.angle.cpu().numpy(), out.trans.cpu().numpy(), out.tries.cpu().numpy()
    pp, obs = out.pathpt.cpu().numpy(), out.obs.cpu().numpy()
    cnt, rc = out.obs_cnt.cpu().numpy(), out.rand_cnt.cpu().numpy()
    assert out.valid.cpu().numpy().all()
    for i in range(0, n, 5):
        w = _oracle_generate(g, pre, P, O, c, SEED, map0 + i)
        assert tries[i] == w["tries"] and ang[i] == w["angle"] and tuple(tr[i]) == w["trans"], i
        np.testing.assert_allclose(pp[i], w["pathpt"], rtol=1e-9, atol=1e-9)
        assert rc[i] == len(w["rand"]) and np.array_equal(obs[i, :rc[i]], w["rand"]), i
        np.testing.assert_allclose(obs[i, rc[i]:cnt[i]], w["pobs"], rtol=1e-9, atol=1e-7)
    assert tries.mean() > 1.2                              # the rejection loop really rejects
    # any split of the index range reproduces the same maps bit for bit (counter-based RNG keyed by g)
    a = ops.generate_maps(bank, map0, 40, reps=P, obstacles_num=O, clearance=c, seed=SEED)
    b = ops.generate_maps(bank, map0 + 40, n - 40, reps=P, obstacles_num=O, clearance=c, seed=SEED)
    for name in ("angle", "trans", "pathpt", "segpt", "obs", "obs_cnt", "bits", "tries"):
        whole = getattr(out, name).cpu().numpy()
        parts = np.concatenate([getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy()])
        assert np.array_equal(whole, parts), name
    # every accepted random obstacle satisfies the reference's clearance invariant w.r.t. the emitted path points
    for i in range(0, n, 7):
        for (x, y, r) in obs[i, :rc[i]]:
            d = np.sqrt(((pp[i, 1::2] - [y, x]) ** 2).sum(1)).min()
            assert d > r + c / 50 * 224 - 1e-9
    # retry budget exhausted => flagged invalid, not silently dropped
    bad = ops.generate_maps(bank, map0, 8, reps=P, obstacles_num=O, clearance=c, seed=SEED, max_tries=1)
    v = bad.valid.cpu().numpy()
    assert (v == (tries[:8] == 1)).all()
    assert (bad.obs_cnt.cpu().numpy()[v == 0] == 0).all()
