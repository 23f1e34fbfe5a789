"""When the reference tree is present (the build container; never the GPU box), run the REAL PPNet functions again and
check that the committed golden fixtures and the oracle still say what the reference says.  Skipped elsewhere."""
import numpy as np
import pytest

from oracle import ppnet_oracle as orc
from oracle.ref_loader import find_reference

pytestmark = [pytest.mark.skipif(find_reference() is None, reason="PPNet reference tree not available"),
              pytest.mark.filterwarnings("ignore::DeprecationWarning")]

C = 1 / 50 * 224


def _rc(p):
    return np.asarray([p[1], p[0]], dtype=np.float64)


def test_live_reference_f64_verdicts_match_goldens_and_oracle(golden):
    import torch
    from oracle.ref_loader import load_edage
    pm = load_edage()["process_map"]
    g = golden("segcheck_f64")
    rng = np.random.default_rng(0)
    for i in rng.choice(len(g["verdict"]), 120, replace=False):
        m = int(g["seg_map"][i])
        obs = g["obs"][m, :g["obs_cnt"][m]].tolist()
        s, e = g["segs_xy"][i, :2], g["segs_xy"][i, 2:]
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            live = bool(pm.collision_check_circle_edge(torch.tensor(_rc(s)), torch.tensor(_rc(e)), obs, float(g["clearance"])))
        assert live == bool(g["verdict"][i])
        assert live == bool(orc.segcheck_edage_f64(_rc(s), _rc(e), obs, float(g["clearance"]), dot_mode=int(g["dot_mode"])))


def test_live_reference_mpnet_checker_matches_goldens_and_oracle(golden):
    import torch
    from oracle.ref_loader import load_mpnet_checker
    g = golden("segcheck_f32")
    obc = [g["obs"][m, :g["obs_cnt"][m]].tolist() for m in range(len(g["obs_cnt"]))]
    ns = load_mpnet_checker(obc, float(g["clearance"]))
    rng = np.random.default_rng(1)
    for i in rng.choice(len(g["verdict"]), 120, replace=False):
        m = int(g["seg_map"][i])
        s, e = torch.from_numpy(g["segs_xy"][i, :2].copy()), torch.from_numpy(g["segs_xy"][i, 2:].copy())
        live = bool(ns["collision_check_circle_edge"](s, e, m))
        assert live == bool(g["verdict"][i])
        assert int(ns["steerTo"](s, e, m)) == int(g["steer"][i])
        assert live == bool(orc.segcheck_mpnet_f32(g["segs_xy"][i, :2], g["segs_xy"][i, 2:], obc[m], float(g["clearance"])))


def test_live_reference_grid_index_and_pathseg(golden):
    from oracle.ref_loader import load_edage
    mods = load_edage()
    p = mods["Path"].Path(seg_num=3, poly_order=4, clearance=1, is_straight=False)
    p.MapSize, p.Resolution = 50, 224
    g = golden("grid")
    got = p.coord_euclidean2image(g["pts"][:300], 224)
    assert np.array_equal(got, g["idx224"][:300])
    assert np.array_equal(got, orc.grid_index_vec(g["pts"][:300], 50, 224, 224))
    np.random.seed(0)
    seg = mods["PathSeg"].PathSeg(4, 2)
    poly, end = seg.random()
    m = golden("misc")
    np.testing.assert_allclose(poly, m["seg0_poly"], rtol=1e-9, atol=1e-12)
    assert abs(float(end[0]) - float(m["seg0_end"])) < 1e-12


def test_live_reference_read_folder_equals_the_mirror(tmp_path):
    """process_map.read_folder (:75-102) through the real reference and through the mirror on the same folder: the same
    image lists (sorted by numeric stem when the folder is complete, as listed otherwise), labels and corridor arrays."""
    import contextlib, io
    import torch
    from PIL import Image
    from oracle.ref_loader import load_edage
    from ppnet_b200.edage import process_map as mirror
    pm = load_edage()["process_map"]
    folder = tmp_path / "0"
    (folder / "data").mkdir(parents=True)
    n = 12                                                       # two-digit stems: string order != numeric order
    rng = np.random.default_rng(4)
    for i in range(n):
        Image.fromarray(rng.integers(0, 255, (6, 6, 3), dtype=np.uint8)).save(str(folder / ("%d.png" % i)))
    for i in range(2):
        Image.fromarray(rng.integers(0, 255, (6, 6, 3), dtype=np.uint8)).save(str(folder / "data" / ("%d.png" % i)))
    torch.save([[i, float(i), [i, i], np.full([11, 2], float(i)), None] for i in range(n + 3)], str(folder / "data" / "MapLabel"))
    real_load = torch.load                  # the reference predates torch.load's weights_only=True default
    torch.load = lambda *a, **k: real_load(*a, **dict(k, weights_only=False))
    old_ref, old_mir = pm.NUM_PER_FOLDER, mirror.NUM_PER_FOLDER
    try:
        for per in (n, n + 1):                                   # complete folder / incomplete folder
            pm.NUM_PER_FOLDER = mirror.NUM_PER_FOLDER = per
            with contextlib.redirect_stdout(io.StringIO()):
                want = pm.read_folder(str(folder), is_read_path=True)
            got = mirror.read_folder(str(folder), is_read_path=True)
            assert got[0] == want[0]
            assert len(got[1]) == len(want[1]) and all(a[0] == b[0] for a, b in zip(got[1], want[1]))
            assert len(got[2]) == len(want[2]) and all(np.array_equal(a, b) for a, b in zip(got[2], want[2]))
    finally:
        torch.load = real_load
        pm.NUM_PER_FOLDER, mirror.NUM_PER_FOLDER = old_ref, old_mir
