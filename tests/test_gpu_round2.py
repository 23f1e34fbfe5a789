"""GPU parity tests of the round-2 hot path: the fused A11 + A12 verdict kernel, the DDA on the A11 array, bit-packed
verdicts, warp-scan compaction of the survivors, the device-side segment source and the one-array host pipeline --
all through the C ABI, against the C oracle / the one-flavour kernels on the same seeded inputs.  Bit-exact."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import philox
from oracle import ppnet_oracle as orc

from test_gpu_parity import _config2_inputs, _margin_cases, dev   # same seeded generators as the one-flavour tests

pytestmark = pytest.mark.gpu

CLEAR = 1 / 50 * 224


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "these tests need a B200"
    from ppnet_b200 import ops as _ops
    return _ops


def xy32(segs_rc):
    """The float32 caller's view of an A11 array: (x, y) = (float32(col), float32(row))."""
    return np.ascontiguousarray(segs_rc[:, [1, 0, 3, 2]].astype(np.float32))


def unpack(words, n):
    w = np.asarray(words).view(np.uint32)
    return ((w[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.uint8).reshape(-1)[:n]


ALL = ("u8_64", "u8_32", "bits64", "bits32")


def _check(out, n, want64, want32):
    for k, v in out.items():
        got = unpack(v.cpu().numpy(), n) if k.startswith("bits") else v.cpu().numpy()
        want = want64 if k.endswith("64") else want32
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (k, bad[:10])


@pytest.mark.parametrize("dot_mode", [0, 1])
def test_verdict_fused_vs_oracle_seeded(ops, dot_mode):
    rng = np.random.default_rng(142 + dot_mode)
    n_maps, spm = 300, 1024
    segs, obs, cnt = _config2_inputs(rng, n_maps, spm, np.float64)
    seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), spm)
    want64 = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, CLEAR, dot_mode=dot_mode, threads=8)
    want32 = c_oracle.segcheck_f32(xy32(segs), seg_map, obs, cnt, CLEAR, threads=8, want_steer=False)
    for want in (ALL, ("bits64", "bits32"), ("u8_64",), ("bits32",)):
        _check(ops.verdict_fused(dev(segs), dev(obs), dev(cnt), CLEAR, dot_mode=dot_mode, want=want), len(segs), want64, want32)
    assert 0.05 < want64.mean() < 0.6


def test_verdict_fused_equals_the_one_flavour_kernels_and_goldens(ops, golden):
    """Each flavour of the fused kernel == its own entry point == the reference's golden verdicts."""
    g = golden("segcheck_f64")
    order = np.argsort(g["seg_map"], kind="stable")
    counts = np.bincount(g["seg_map"], minlength=len(g["obs_cnt"]))
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    pts = np.ascontiguousarray(g["segs_xy"][order][:, [1, 0, 3, 2]])              # (row, col)
    o = ops.verdict_fused(dev(pts), dev(g["obs"]), dev(g["obs_cnt"]), float(g["clearance"]), seg_off=dev(off),
                          dot_mode=int(g["dot_mode"]), want=ALL)
    assert np.array_equal(o["u8_64"].cpu().numpy(), g["verdict"][order])         # real reference outputs (A11)
    assert np.array_equal(unpack(o["bits64"].cpu().numpy(), len(pts)), g["verdict"][order])
    v32 = ops.segcheck_mpnet_f32(dev(xy32(pts)), dev(g["obs"]), dev(g["obs_cnt"]), float(g["clearance"]), seg_off=dev(off))
    assert np.array_equal(o["u8_32"].cpu().numpy(), v32.cpu().numpy())
    # the A12 golden (float32 inputs are their own cast: feed them promoted, swapped into the A11 layout)
    g = golden("segcheck_f32")
    order = np.argsort(g["seg_map"], kind="stable")
    counts = np.bincount(g["seg_map"], minlength=len(g["obs_cnt"]))
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rc = np.ascontiguousarray(g["segs_xy"][order][:, [1, 0, 3, 2]].astype(np.float64))
    o = ops.verdict_fused(dev(rc), dev(g["obs"]), dev(g["obs_cnt"]), float(g["clearance"]), seg_off=dev(off), want=("u8_32", "bits32"))
    assert np.array_equal(o["u8_32"].cpu().numpy(), g["verdict"][order])         # real reference outputs (A12)
    assert np.array_equal(unpack(o["bits32"].cpu().numpy(), len(rc)), g["verdict"][order])


def test_verdict_fused_ragged_tiles_and_edge_cases(ops):
    rng = np.random.default_rng(5)
    # ragged CSR with empty rows: output words are shared between maps (atomic path)
    n_maps = 400
    lens = rng.integers(0, 300, n_maps)
    lens[[3, 17, n_maps - 1]] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(off[-1])
    segs, obs, cnt = _config2_inputs(rng, n_maps, 300, np.float64)
    segs = segs[:n]
    seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), lens)
    w64 = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, CLEAR, threads=8)
    w32 = c_oracle.segcheck_f32(xy32(segs), seg_map, obs, cnt, CLEAR, threads=8, want_steer=False)
    _check(ops.verdict_fused(dev(segs), dev(obs), dev(cnt), CLEAR, seg_off=dev(off), want=ALL), n, w64, w32)
    _check(ops.verdict_fused(dev(segs), dev(obs), dev(cnt), CLEAR, seg_off=dev(off), want=("bits64", "bits32")), n, w64, w32)
    # more than 128 circles per map: three circle tiles, later tiles continue from the stored verdicts (bytes or words)
    n_maps, spm, omax = 48, 2048, 400
    segs, obs, _ = _config2_inputs(rng, n_maps, spm, np.float64, omax=omax)
    obs[..., 2] *= 0.25
    cnt = rng.integers(129, omax + 1, n_maps).astype(np.int32)
    seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), spm)
    w64 = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, CLEAR, threads=8)
    w32 = c_oracle.segcheck_f32(xy32(segs), seg_map, obs, cnt, CLEAR, threads=8, want_steer=False)
    assert 0.2 < w64.mean() < 0.98
    _check(ops.verdict_fused(dev(segs), dev(obs), dev(cnt), CLEAR, want=ALL), len(segs), w64, w32)
    _check(ops.verdict_fused(dev(segs), dev(obs), dev(cnt), CLEAR, want=("bits64", "bits32")), len(segs), w64, w32)
    # NaN / inf coordinates, maps without circles, zero segments
    obs = np.zeros([4, 3, 3])
    obs[1, 0] = [50, 14, 2]
    obs[3, :, :] = [[50, 14, 2], [np.nan, 3, 1], [10, 10, np.inf]]
    cnt = np.asarray([0, 1, 0, 3], dtype=np.int32)
    pts = np.asarray([[10, 10, 10, 100], [10, 10, 10, 100], [np.nan, 10, 10, 100], [10, 10, 10, np.inf],
                      [10, 10, 10, 100], [60, 60, 60, 60]], dtype=np.float64)
    off = np.asarray([0, 1, 4, 4, 6], dtype=np.int64)
    seg_map = np.asarray([0, 1, 1, 1, 3, 3], dtype=np.int32)
    with np.errstate(all="ignore"):
        w64 = c_oracle.segcheck_f64(pts, seg_map, obs, cnt, CLEAR)
        w32 = c_oracle.segcheck_f32(xy32(pts), seg_map, obs, cnt, CLEAR, want_steer=False)
    _check(ops.verdict_fused(dev(pts), dev(obs), dev(cnt), CLEAR, seg_off=dev(off), want=ALL), len(pts), w64, w32)
    z = ops.verdict_fused(torch.zeros([0, 4], dtype=torch.float64, device="cuda"), dev(obs), dev(cnt), CLEAR,
                          seg_off=dev(np.zeros(5, dtype=np.int64)), want=("u8_64",))
    assert z["u8_64"].numel() == 0


@pytest.mark.parametrize("dot_mode", [0, 1])
def test_verdict_fused_filter_margins_vs_oracle(ops, dot_mode):
    """The shared culling (one query with the float32 margins for both flavours) and the data-encoded exact-only
    fall-throughs must never change a verdict: the 400 k sharp / degenerate pairs of the one-flavour tests, both
    flavours at once, bound 224 and bound 1e9 (no grid)."""
    rng = np.random.default_rng(11 + dot_mode)
    n_maps, spm, omax = 25000, 16, 3
    segs, circ = _margin_cases(rng, n_maps * spm, np.float64)
    obs = np.ascontiguousarray(circ.reshape(n_maps, spm, 3)[:, :omax])
    cnt = rng.integers(0, omax + 1, n_maps).astype(np.int32)
    cnt[: n_maps // 2] = omax
    seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), spm)
    with np.errstate(all="ignore"):
        s32 = xy32(segs)
    for bound in (224.0, 1e9):
        w64 = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, CLEAR, bound=bound, dot_mode=dot_mode, threads=8)
        w32 = c_oracle.segcheck_f32(s32, seg_map, obs, cnt, CLEAR, bound=bound, threads=8, want_steer=False)
        _check(ops.verdict_fused(dev(segs), dev(obs), dev(cnt), CLEAR, bound=bound, dot_mode=dot_mode, want=ALL),
               len(segs), w64, w32)
    # sharp cases built in float32 (the A12 thresholds sit at float32 ulps): promote them into the A11 layout
    segs32, circ = _margin_cases(np.random.default_rng(3), n_maps * spm, np.float32)
    obs = np.ascontiguousarray(circ.reshape(n_maps, spm, 3)[:, :omax])
    rc = np.ascontiguousarray(segs32[:, [1, 0, 3, 2]].astype(np.float64))
    w32 = c_oracle.segcheck_f32(segs32, seg_map, obs, cnt, CLEAR, threads=8, want_steer=False)
    w64 = c_oracle.segcheck_f64(rc, seg_map, obs, cnt, CLEAR, dot_mode=dot_mode, threads=8)
    _check(ops.verdict_fused(dev(rc), dev(obs), dev(cnt), CLEAR, dot_mode=dot_mode, want=ALL), len(rc), w64, w32)
    assert 0.05 < w32.mean() < 0.95


def test_a12_cmp_mode_numpy1_vs_nep50(ops):
    """neuralplanner.py:54,66 compare a float32 offset with the Python float `size + clearance/2`: float32 under NumPy >= 2
    (NEP 50), float64 under the reference-era NumPy 1.x.  Cases whose offset falls between float32(thr) and thr flip."""
    rng = np.random.default_rng(77)
    n = 200000
    # one circle per map, one segment per map; the vertex distance |e - o| is an exact float32 placed at the threshold
    o = rng.uniform(40, 180, (n, 2)).astype(np.float32)
    r = rng.uniform(3, 20, n)
    thr = r + CLEAR / 2                                              # Python-float threshold (float64)
    t32 = thr.astype(np.float32)
    e = o.copy()
    e[:, 0] = o[:, 0] + t32                                          # |e - o| == fl32(o + t32) - o: within an ulp of t32
    s = e + rng.normal(0, 20, (n, 2)).astype(np.float32)
    segs32 = np.concatenate([s, e], axis=1).astype(np.float32)
    obs = np.concatenate([o.astype(np.float64), r[:, None]], axis=1).reshape(n, 1, 3)
    cnt = np.ones(n, dtype=np.int32)
    seg_map = np.arange(n, dtype=np.int32)
    w_nep = c_oracle.segcheck_f32_cmp(segs32, seg_map, obs, cnt, CLEAR, 0, threads=8)
    w_np1 = c_oracle.segcheck_f32_cmp(segs32, seg_map, obs, cnt, CLEAR, 1, threads=8)
    assert np.array_equal(w_nep, c_oracle.segcheck_f32(segs32, seg_map, obs, cnt, CLEAR, threads=8, want_steer=False))
    flips = int((w_nep != w_np1).sum())
    assert flips > 100, flips                                        # the two NumPy generations really disagree here
    rc = np.ascontiguousarray(segs32[:, [1, 0, 3, 2]].astype(np.float64))
    for mode, want in ((ops.CMP_F32_NEP50, w_nep), (ops.CMP_F64_NUMPY1, w_np1)):
        got = ops.verdict_fused(dev(rc), dev(obs), dev(cnt), CLEAR, cmp_mode=mode, want=("u8_32",))["u8_32"].cpu().numpy()
        assert np.array_equal(got, want), mode
    # the Python restatement agrees with the C one on a sample of the flipping cases
    idx = np.nonzero(w_nep != w_np1)[0][:40]
    for i in idx:
        for mode, want in ((0, w_nep), (1, w_np1)):
            assert orc.segcheck_mpnet_f32(segs32[i, :2], segs32[i, 2:], obs[i].tolist(), CLEAR, cmp_mode=mode) == bool(want[i])


@pytest.mark.parametrize("bound", [224.0, 33.0, 100.5, 1024.0, 100000.0, 0.75])
def test_inner_disk_grid_end_points_around_every_circle_rim(ops, bound):
    """The verdict kernel resolves a segment whose END point lies in a grid cell that is wholly inside a circle's
    threshold disk (verdict.cu, `inner`) without looking at any pair.  Sharp cases for that shortcut: end points at
    distance thr + delta from a centre, delta from -1 px to +0.5 px down to 1e-7 px, end points on the grid's cell
    corners, circles hanging over the map's border, tiny and huge radii -- a cell marked by mistake would block a
    segment the reference leaves free.  Bit-exact against the C oracle in both flavours and both cmp modes."""
    rng = np.random.default_rng(int(bound * 7))
    n_maps, omax, spm = 1200, 24, 512           # >= 8 x 148 maps: one CTA per map, 512 segments each (the grid needs >= 256)
    obs = np.zeros([n_maps, omax, 3])
    obs[..., 0] = rng.uniform(-0.1 * bound, 1.1 * bound, (n_maps, omax))
    obs[..., 1] = rng.uniform(-0.1 * bound, 1.1 * bound, (n_maps, omax))
    obs[..., 2] = rng.uniform(0, bound / 5, (n_maps, omax)) * rng.choice([0.02, 0.3, 1.0], (n_maps, omax))
    cnt = rng.integers(1, omax + 1, n_maps).astype(np.int32)
    clear = CLEAR * bound / 224
    segs = np.empty([n_maps, spm, 4])
    for m in range(n_maps):
        j = rng.integers(0, cnt[m], spm)
        thr = obs[m, j, 2] + clear / 2
        delta = rng.choice([-1.0, -0.3, -0.12, -0.1, -0.08, -1e-3, -1e-7, 1e-7, 1e-3, 0.5], spm) * rng.uniform(0.5, 1.0, spm) * bound / 224
        th = rng.uniform(0, 2 * np.pi, spm)
        ex = obs[m, j, 0] + (thr + delta) * np.cos(th)
        ey = obs[m, j, 1] + (thr + delta) * np.sin(th)
        corner = rng.random(spm) < 0.15                         # end points on (or a hair beside) the 64 x 64 cell corners
        ex[corner] = np.rint(ex[corner] * 64 / bound) * bound / 64 + rng.choice([0.0, 1e-6, -1e-6], corner.sum()) * bound / 224
        ey[corner] = np.rint(ey[corner] * 64 / bound) * bound / 64 + rng.choice([0.0, 1e-6, -1e-6], corner.sum()) * bound / 224
        far = rng.uniform(0, 2 * np.pi, spm)
        sx = ex + rng.uniform(bound / 400, bound / 3, spm) * np.cos(far)
        sy = ey + rng.uniform(bound / 400, bound / 3, spm) * np.sin(far)
        segs[m] = np.stack([sy, sx, ey, ex], axis=1)            # (row, col) pairs
    segs = segs.reshape(-1, 4)
    segs[::997, 0] = np.nan                                     # a NaN start does not stop the vertex test on e
    segs[::1499, 1] = 1e12
    seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), spm)
    with np.errstate(all="ignore"):
        s32 = xy32(segs)
    want64 = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, clear, bound=bound, threads=8)
    want32 = c_oracle.segcheck_f32(s32, seg_map, obs, cnt, clear, bound=bound, threads=8, want_steer=False)
    want32c = c_oracle.segcheck_f32_cmp(s32, seg_map, obs, cnt, clear, 1, bound=bound, threads=8)
    assert 0.2 < want64.mean() < 0.9
    _check(ops.verdict_fused(dev(segs), dev(obs), dev(cnt), clear, bound=bound, want=ALL), len(segs), want64, want32)
    _check(ops.verdict_fused(dev(segs), dev(obs), dev(cnt), clear, bound=bound, cmp_mode=ops.CMP_F64_NUMPY1, want=("bits32",)),
           len(segs), want64, want32c)
    assert np.array_equal(ops.segcheck_edage_f64(dev(segs), dev(obs), dev(cnt), clear, bound=bound).cpu().numpy(), want64)
    assert np.array_equal(ops.segcheck_mpnet_f32(dev(s32), dev(obs), dev(cnt), clear, bound=bound).cpu().numpy(), want32)


def test_dda_on_the_a11_array_equals_the_float32_walk(ops):
    rng = np.random.default_rng(9)
    for R, n_maps, spm, omax in ((224, 200, 1024, 50), (33, 64, 96, 6), (1024, 6, 4096, 300)):
        obs = np.zeros([n_maps, omax, 3])
        obs[..., 0] = rng.uniform(0, R, (n_maps, omax))
        obs[..., 1] = rng.uniform(0, R, (n_maps, omax))
        obs[..., 2] = rng.uniform(0, R / 12, (n_maps, omax))
        cnt = rng.integers(0, omax + 1, n_maps).astype(np.int32)
        s = rng.uniform(-3, R + 3, (n_maps * spm, 2))
        segs = np.concatenate([s, s + rng.normal(0, R / 8, s.shape)], axis=1)
        segs[::211, 0] = np.nan
        segs[5::307, 2] = 1e20; segs[7::401, 3] = -np.inf; segs[11::503, 1] = 3e9; segs[13::601, 2] = np.nan
        bits = ops.raster_circles_bits(dev(obs), dev(cnt), R, 1.5)
        seg_map = np.repeat(np.arange(n_maps, dtype=np.int32), spm)
        with np.errstate(all="ignore"):
            s32 = xy32(segs)
        want_v, want_f = c_oracle.dda_gridcheck(bits.cpu().numpy().view(np.uint32), R, s32, seg_map)
        v32, f32 = ops.dda_gridcheck(bits, R, dev(s32))
        assert np.array_equal(v32.cpu().numpy(), want_v) and np.array_equal(f32.cpu().numpy(), want_f)
        # verdict only: the kernel variant that also resolves on the END cell at park time
        assert np.array_equal(ops.dda_gridcheck(bits, R, dev(s32), want_first=False).cpu().numpy(), want_v)
        o = ops.dda_gridcheck_rc64(bits, R, dev(segs), want=("u8", "bits", "first"))
        assert np.array_equal(o["u8"].cpu().numpy(), want_v)
        assert np.array_equal(o["first"].cpu().numpy(), want_f)
        assert np.array_equal(unpack(o["bits"].cpu().numpy(), len(segs)), want_v)
        assert np.array_equal(unpack(ops.dda_gridcheck_rc64(bits, R, dev(segs), want=("bits",))["bits"].cpu().numpy(), len(segs)), want_v)
    # ragged CSR: shared output words
    lens = rng.integers(0, 200, n_maps)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(off[-1])
    sm = np.repeat(np.arange(n_maps, dtype=np.int32), lens)
    want_v, _ = c_oracle.dda_gridcheck(bits.cpu().numpy().view(np.uint32), R, s32[:n], sm)
    o = ops.dda_gridcheck_rc64(bits, R, dev(segs[:n]), seg_off=dev(off), want=("u8", "bits"))
    assert np.array_equal(o["u8"].cpu().numpy(), want_v)
    assert np.array_equal(unpack(o["bits"].cpu().numpy(), n), want_v)


def test_compaction_of_bit_packed_survivors(ops):
    rng = np.random.default_rng(21)
    for n in (1, 31, 32, 33, 8191, 8192, 100003, 3_000_001):
        nw = (n + 31) // 32
        a, b, c = (rng.integers(0, 2 ** 32, nw, dtype=np.uint64).astype(np.uint32) & rng.integers(0, 2 ** 32, nw, dtype=np.uint64).astype(np.uint32)
                   for _ in range(3))
        if n > 1000:
            a[10:200] = 0xFFFFFFFF                                   # a long run without survivors
            b[300:320] = 0
        free = (unpack(a, n) | unpack(b, n) | unpack(c, n)) == 0
        want = np.nonzero(free)[0]
        idx, cnt, _ = ops.compact_bits(dev(a.view(np.int32)), dev(b.view(np.int32)), dev(c.view(np.int32)), n=n, idx_base=7)
        k = int(cnt.item())
        assert k == len(want) and np.array_equal(idx[:k].cpu().numpy(), want + 7)
        want1 = np.nonzero(unpack(a, n) == 0)[0]
        idx, cnt, _ = ops.compact_bits(dev(a.view(np.int32)), n=n)
        assert int(cnt.item()) == len(want1) and np.array_equal(idx[:len(want1)].cpu().numpy(), want1)
    for n in (0, 1, 1023, 1024, 1025, 5000):
        f = (rng.random(n) < 0.4).astype(np.uint8)
        idx, cnt = ops.compact_u8_i32(dev(f) if n else torch.zeros([0], dtype=torch.uint8, device="cuda"), keep=1, idx_base=100)
        want = np.nonzero(f == 1)[0] + 100
        assert int(cnt.item()) == len(want) and np.array_equal(idx[:len(want)].cpu().numpy(), want)


def test_device_side_segment_source(ops):
    seed, spm, R, sigma = 0xABCDEF12345, 257, 224, 15.0
    whole = ops.propose_segments(1000, 40, spm, R, sigma, seed=seed).cpu().numpy()
    # sharding invariance: any split of the map range gives the same bytes
    parts = [ops.propose_segments(1000 + a, b, spm, R, sigma, seed=seed).cpu().numpy() for a, b in ((0, 7), (7, 1), (8, 32))]
    assert np.array_equal(whole, np.concatenate(parts))
    # the restatement: starts bit-exact (integer -> double arithmetic only), ends to the last ulps of log / sincospi
    for g in (1000, 1017, 1039):
        want = philox.propose_segments(seed, g, spm, R, sigma)
        got = whole[(g - 1000) * spm:(g - 1000 + 1) * spm]
        assert np.array_equal(got[:, :2], want[:, :2])
        assert np.abs(got[:, 2:] - want[:, 2:]).max() < 1e-10
    # distribution: uniform starts, N(0, sigma^2) offsets (KS, alpha = 0.01)
    from scipy import stats
    big = ops.propose_segments(5, 400, 1024, R, sigma, seed=seed + 1).cpu().numpy()
    assert stats.kstest(big[:, 0] / R, "uniform").pvalue > 0.01 and stats.kstest(big[:, 1] / R, "uniform").pvalue > 0.01
    assert stats.kstest((big[:, 2] - big[:, 0]) / sigma, "norm").pvalue > 0.01
    assert stats.kstest((big[:, 3] - big[:, 1]) / sigma, "norm").pvalue > 0.01
    assert abs(np.corrcoef(big[:, 2] - big[:, 0], big[:, 3] - big[:, 1])[0, 1]) < 0.01


def _host_setup(M, SPM, seed_bank=3, n_bank=20):
    from ppnet_b200 import host, ops
    from ppnet_b200.synthetic import synthetic_bank
    bk = synthetic_bank(n_bank, seed=seed_bank)
    keys = ("pathpt", "segpt", "hull", "hull_cnt", "obs", "obs_cnt")
    dbank = ops.PathBank(*[torch.from_numpy(bk[k]).cuda() for k in keys])
    hbank = host.HostBank(*[bk[k] for k in keys], device=0)
    pomax, np_, ns1 = bk["obs"].shape[1], bk["pathpt"].shape[1], bk["segpt"].shape[1]
    return bk, dbank, hbank, pomax, np_, ns1


@pytest.mark.parametrize("SPM", [96, 100])       # 100: rows do not start on word boundaries (atomic word path)
def test_one_array_host_pipeline_equals_the_device_ops(SPM):
    """ppnet_generate_and_check_host in one-array mode: ONE float64 upload, bit-packed + byte verdicts, survivor lists
    == generate_maps + the verdict kernels on device tensors; several slices with a ragged last one."""
    from ppnet_b200 import host, ops
    from ppnet_b200.synthetic import synthetic_segments
    M, R, O, reps, seed, map0 = 2500, 224, 50, 10, 99, 12345
    bk, dbank, hbank, pomax, np_, ns1 = _host_setup(M, SPM)
    ctx = host.HostContext(0)
    s64 = synthetic_segments(M, SPM, seed=5)
    n = M * SPM
    nw = (n + 31) // 32
    hout = dict(angle=np.empty(M), trans=np.empty([M, 2], np.int32), segpt=np.empty([M, ns1, 2]), pathpt=np.empty([M, np_, 2]),
                obs=np.zeros([M, O + pomax, 3]), obs_cnt=np.empty(M, np.int32), rand_cnt=np.empty(M, np.int32),
                bits=np.empty([M, R, 7], np.int32), tries=np.empty(M, np.int32), valid=np.empty(M, np.uint8),
                counters=np.zeros(4, np.uint64))
    chk = dict(segs_rc_f64=s64, clearance_px=CLEAR,
               verdict_f64=np.empty(n, np.uint8), verdict_f32=np.empty(n, np.uint8), verdict_dda=np.empty(n, np.uint8),
               vbits_f64=np.empty(nw, np.uint32), vbits_f32=np.empty(nw, np.uint32), vbits_dda=np.empty(nw, np.uint32),
               free_idx=np.full(n, -1, np.int32), free_count=np.zeros(1, np.int64),
               valid_idx=np.full(M, -1, np.int32), valid_count=np.zeros(1, np.int64))
    host.generate_maps_host(ctx, hbank, map0, M, reps, O, hout, R, 50.0, 5.0, 1.0, seed, raster_inflate=2.24, checks=chk)
    gen = ops.generate_maps(dbank, map0, M, reps, O, R, 50.0, 5.0, 1.0, seed, raster_inflate=2.24)
    S = torch.from_numpy(s64).cuda()
    w64 = ops.segcheck_edage_f64(S, gen.obs, gen.obs_cnt, CLEAR).cpu().numpy()
    S32 = torch.from_numpy(xy32(s64)).cuda()
    w32 = ops.segcheck_mpnet_f32(S32, gen.obs, gen.obs_cnt, CLEAR).cpu().numpy()
    wd = ops.dda_gridcheck(gen.bits, R, S32, want_first=False).cpu().numpy()
    for name in ("angle", "trans", "segpt", "pathpt", "obs_cnt", "rand_cnt", "bits", "tries", "valid"):
        assert np.array_equal(hout[name], getattr(gen, name).cpu().numpy()), name
    assert np.array_equal(chk["verdict_f64"], w64) and np.array_equal(chk["verdict_f32"], w32) and np.array_equal(chk["verdict_dda"], wd)
    assert np.array_equal(unpack(chk["vbits_f64"], n), w64) and np.array_equal(unpack(chk["vbits_f32"], n), w32)
    assert np.array_equal(unpack(chk["vbits_dda"], n), wd)
    free = np.nonzero((w64 | w32 | wd) == 0)[0]
    assert int(chk["free_count"][0]) == len(free) and np.array_equal(chk["free_idx"][:len(free)], free)
    valid = np.nonzero(hout["valid"] == 1)[0]
    assert int(chk["valid_count"][0]) == len(valid) and np.array_equal(chk["valid_idx"][:len(valid)], valid)
    assert 0.2 < len(free) / n < 0.9
    h2d, d2h = ctx.bytes_moved()
    assert h2d == n * 32                                             # ONE upload: 32 B per segment, nothing else
    # bits-only outputs (the bench's e2e configuration) give the same words
    ctx2 = host.HostContext(0)
    chk2 = dict(segs_rc_f64=s64, clearance_px=CLEAR, vbits_f64=np.empty(nw, np.uint32), vbits_f32=np.empty(nw, np.uint32),
                vbits_dda=np.empty(nw, np.uint32), valid_idx=np.full(M, -1, np.int32), valid_count=np.zeros(1, np.int64))
    host.generate_maps_host(ctx2, hbank, map0, M, reps, O, dict(valid=np.empty(M, np.uint8)), R, 50.0, 5.0, 1.0, seed,
                            raster_inflate=2.24, checks=chk2)
    for k in ("vbits_f64", "vbits_f32", "vbits_dda"):
        assert np.array_equal(chk2[k], chk[k]), k
    assert np.array_equal(chk2["valid_idx"][:len(valid)], valid)


def test_generator_mode_host_pipeline_uploads_nothing():
    """Device-side segment source inside the host pipeline: no host->device bytes, verdicts equal the device ops on the
    proposed segments, identical for any split of the map range."""
    from ppnet_b200 import host, ops
    M, SPM, R, O, reps, seed, map0 = 1500, 128, 224, 50, 10, 4242, 777
    bk, dbank, hbank, pomax, np_, ns1 = _host_setup(M, SPM, seed_bank=8)
    n = M * SPM
    nw = n // 32

    def run(m0, m):
        ctx = host.HostContext(0)
        k = m * SPM
        chk = dict(propose_sigma=15.0, segs_per_map=SPM, clearance_px=CLEAR, vbits_f64=np.empty(k // 32, np.uint32),
                   vbits_f32=np.empty(k // 32, np.uint32), vbits_dda=np.empty(k // 32, np.uint32), out_segs_rc=np.empty([k, 4]),
                   free_idx=np.empty(k, np.int32), free_count=np.zeros(1, np.int64))
        out = dict(valid=np.empty(m, np.uint8), obs_cnt=np.empty(m, np.int32))
        host.generate_maps_host(ctx, hbank, m0, m, reps, O, out, R, 50.0, 5.0, 1.0, seed, raster_inflate=2.24, checks=chk)
        return chk, out, ctx.bytes_moved()

    chk, out, (h2d, d2h) = run(map0, M)
    assert h2d == 0
    segs = ops.propose_segments(map0, M, SPM, R, 15.0, seed=seed)
    assert np.array_equal(chk["out_segs_rc"], segs.cpu().numpy())
    gen = ops.generate_maps(dbank, map0, M, reps, O, R, 50.0, 5.0, 1.0, seed, raster_inflate=2.24)
    o = ops.verdict_fused(segs, gen.obs, gen.obs_cnt, CLEAR, want=("bits64", "bits32"))
    d = ops.dda_gridcheck_rc64(gen.bits, R, segs, want=("bits",))
    assert np.array_equal(chk["vbits_f64"], o["bits64"].cpu().numpy().view(np.uint32))
    assert np.array_equal(chk["vbits_f32"], o["bits32"].cpu().numpy().view(np.uint32))
    assert np.array_equal(chk["vbits_dda"], d["bits"].cpu().numpy().view(np.uint32))
    # ... and against the oracle on the downloaded segments
    seg_map = np.repeat(np.arange(M, dtype=np.int32), SPM)
    w64 = c_oracle.segcheck_f64(chk["out_segs_rc"], seg_map, gen.obs.cpu().numpy(), gen.obs_cnt.cpu().numpy(), CLEAR, threads=8)
    assert np.array_equal(unpack(chk["vbits_f64"], n), w64)
    a, _, _ = run(map0, 600)
    b, _, _ = run(map0 + 600, 900)
    assert np.array_equal(np.concatenate([a["vbits_f64"], b["vbits_f64"]]), chk["vbits_f64"])
    assert np.array_equal(np.concatenate([a["vbits_dda"], b["vbits_dda"]]), chk["vbits_dda"])
    fa, fb = int(a["free_count"][0]), int(b["free_count"][0])
    assert fa + fb == int(chk["free_count"][0])
    assert np.array_equal(np.concatenate([a["free_idx"][:fa], b["free_idx"][:fb] + 600 * SPM]), chk["free_idx"][:fa + fb])


def test_torch_ops_equal_the_ctypes_ops(ops):
    """torch.ops.ppnet_b200.* (TORCH_LIBRARY, CUDA key) call the same C ABI as ppnet_b200.ops."""
    from ppnet_b200 import torch_ops
    ns = torch_ops.load()
    rng = np.random.default_rng(31)
    n_maps, spm = 64, 256
    segs, obs, cnt = _config2_inputs(rng, n_maps, spm, np.float64)
    S, Ob, C = dev(segs), dev(obs), dev(cnt)
    S32 = dev(xy32(segs))
    b64, b32 = ns.verdict_fused(S, Ob, C, CLEAR)
    o = ops.verdict_fused(S, Ob, C, CLEAR, want=("bits64", "bits32"))
    assert torch.equal(b64, o["bits64"]) and torch.equal(b32, o["bits32"])
    assert torch.equal(ns.segcheck_edage_f64(S, Ob, C, CLEAR), ops.segcheck_edage_f64(S, Ob, C, CLEAR))
    assert torch.equal(ns.segcheck_mpnet_f32(S32, Ob, C, CLEAR), ops.segcheck_mpnet_f32(S32, Ob, C, CLEAR))
    bits = ns.raster_circles_bits(Ob, C, 224, 2.24)
    assert torch.equal(bits, ops.raster_circles_bits(Ob, C, 224, 2.24))
    wd = ns.dda_gridcheck_rc64(bits, S)
    assert torch.equal(wd, ops.dda_gridcheck_rc64(bits, 224, S, want=("bits",))["bits"])
    assert torch.equal(ns.dda_gridcheck(bits, S32), ops.dda_gridcheck(bits, 224, S32, want_first=False))
    idx, k = ns.compact_bits(b64, b32, wd, len(segs))
    idx2, k2, _ = ops.compact_bits(b64, b32, wd, n=len(segs))
    assert int(k.item()) == int(k2.item()) and torch.equal(idx[:int(k.item())], idx2[:int(k.item())])
    # CSR form: the caller states the longest row
    lens = rng.integers(0, 200, n_maps)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(off[-1])
    a, b = ns.verdict_fused(S[:n].contiguous(), Ob, C, CLEAR, seg_off=dev(off), max_segs_per_map=int(lens.max()))
    o = ops.verdict_fused(S[:n].contiguous(), Ob, C, CLEAR, seg_off=dev(off), want=("bits64", "bits32"))
    assert torch.equal(a, o["bits64"]) and torch.equal(b, o["bits32"])
    assert torch.equal(ns.propose_segments(S, 7, 100, 8, 64), ops.propose_segments(100, 8, 64, seed=7))
    pts = dev(rng.uniform(-5, 55, (1000, 2)))
    assert torch.equal(ns.grid_index_f64(pts, 50.0, 224.0, 112.0), ops.grid_index_f64(pts, 50.0, 224.0, 112.0))
    with pytest.raises(RuntimeError):
        ns.verdict_fused(S, Ob, C.long(), CLEAR)                      # wrong dtype is refused, not reinterpreted


def test_a15_canvas_model_vs_restatement_and_vs_centre_in_disk(ops):
    """A15 second mode: the canvas model of plot_obstacles (576 x 432 canvas, axes affine, crop, bilinear resize) equals
    its numpy + torch restatement pixel for pixel, and its disagreement with the centre-in-disk rule is the +-1 px band
    SURVEY 8(c) predicted -- quantified per radius class."""
    rng = np.random.default_rng(15)
    R, M, O = 224, 24, 20
    obs = np.zeros([M, O, 3])
    obs[..., 0] = rng.uniform(-5, R + 5, (M, O))
    obs[..., 1] = rng.uniform(-5, R + 5, (M, O))
    obs[..., 2] = rng.uniform(0.2, 22.4, (M, O))
    obs[3, 0] = [np.nan, 5, 3]
    obs[3, 1] = [50, 50, 0.0]
    cnt = rng.integers(0, O + 1, M).astype(np.int32)
    cnt[:4] = O
    got = ops.raster_canvas_bits(dev(obs), dev(cnt), (R, R), R, 0.0).cpu().numpy().view(np.uint32)
    for m in range(M):
        want = orc.raster_canvas_bits(obs[m, :cnt[m]].tolist(), (R, R), R)
        assert np.array_equal(got[m], want), m
    # one circle per map, by radius class: pixels where the two A15 modes disagree, relative to the disk's perimeter
    report = {}
    for lo, hi in ((1, 3), (3, 8), (8, 16), (16, 23)):
        n = 200
        one = np.zeros([n, 1, 3])
        one[:, 0, 0] = rng.uniform(30, R - 30, n)
        one[:, 0, 1] = rng.uniform(30, R - 30, n)
        one[:, 0, 2] = rng.uniform(lo, hi, n)
        c1 = np.ones(n, dtype=np.int32)
        a = ops.raster_canvas_bits(dev(one), dev(c1), (R, R), R, 0.0)
        b = ops.raster_circles_bits(dev(one), dev(c1), R, 0.0)
        diff = unpack((a ^ b).cpu().numpy().reshape(-1), n * R * 7 * 32).reshape(n, -1).sum(axis=1)
        area = unpack(b.cpu().numpy().reshape(-1), n * R * 7 * 32).reshape(n, -1).sum(axis=1)
        perim = 2 * np.pi * one[:, 0, 2]
        report[(lo, hi)] = (float(diff.mean()), float((diff / perim).mean()), float(area.mean()))
        assert (diff / perim).mean() < 1.0, report                    # under one pixel of disagreement per perimeter pixel
        assert diff.max() > 0                                         # the modes are genuinely different rasters
    print("A15 canvas vs centre-in-disk, (radius class) -> (pixels differing, per perimeter px, disk area):", report)


def test_content_digest_kernel_vs_restatement_and_additivity(ops):
    """ppnet_digest_u32 == its numpy restatement, and digest([a, c)) == digest([a, b)) + digest([b, c)) mod 2^64 -- on raw
    arrays and on real generator output (the property config 4's cross-rank identity proof rests on)."""
    rng = np.random.default_rng(64)
    n = 777
    f64 = rng.normal(0, 100, (n, 5, 2))
    i32 = rng.integers(-1000, 1000, (n, 3)).astype(np.int32)
    rows = rng.integers(0, 6, n).astype(np.int32)
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.digest(dev(f64), 1000, acc, salt=3)
    ops.digest(dev(i32), 1000, acc, salt=9)
    ops.digest(dev(f64), 1000, acc, rows=dev(rows), row_elems=2, salt=5)
    want = (orc.digest_u32(f64, 1000, salt=3) + orc.digest_u32(i32, 1000, salt=9) +
            orc.digest_u32(f64, 1000, rows=rows, row_words=4, salt=5)) & 0xFFFFFFFFFFFFFFFF
    assert (int(acc.item()) & 0xFFFFFFFFFFFFFFFF) == want
    bank = ops.path_synthesize(0, 20, clearance=1.0, seed=5, pomax=24).to_bank()
    whole = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.digest_maps(ops.generate_maps(bank, 500, 3000, 10, 50, 224, 50.0, 5.0, 1.0, seed=5, raster_inflate=2.24), whole)
    parts = torch.zeros(1, dtype=torch.int64, device="cuda")
    for a, b in ((500, 1203), (1703, 1), (1704, 1796)):
        ops.digest_maps(ops.generate_maps(bank, a, b, 10, 50, 224, 50.0, 5.0, 1.0, seed=5, raster_inflate=2.24), parts)
    assert int(whole.item()) == int(parts.item()) != 0
    other = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.digest_maps(ops.generate_maps(bank, 501, 3000, 10, 50, 224, 50.0, 5.0, 1.0, seed=5, raster_inflate=2.24), other)
    assert int(other.item()) != int(whole.item())                      # a shifted range is a different dataset
