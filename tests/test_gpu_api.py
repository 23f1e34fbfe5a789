"""GPU tests of the drop-in Python surface (ppnet_b200.edage.*, ppnet_b200.mpnet): the reference's own names and
signatures, answered by the CUDA path.  Known-answer table: SURVEY.md 8(a) (values produced by the reference)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import ppnet_oracle as orc

pytestmark = pytest.mark.gpu

C = 1 / 50 * 224

# (s_xy, e_xy, circle or None, A12 verdict, A11 verdict)
KAT = [
    ((10, 10), (100, 10), (50, 14, 2), True, True),          # edge_hit
    ((10, 10), (100, 10), (50, 14.3, 2), False, False),      # edge_miss
    ((10, 10), (100, 10), (103, 10, 1), True, True),         # vertex_hit_e
    ((10, 10), (100, 10), (8, 10, 3), False, False),         # start_inside_only: the start point is never vertex-tested
    ((10, 10), (100, 10), (110, 10, 1), False, False),       # beyond_end_proj
    ((10, 10), (100, 10), (50, 10, 1), True, True),          # on_line_center
    ((-1, 10), (100, 10), None, True, False),                # oob x = -1: A11 tests row < 0 / col > 224
    ((10, 225), (100, 10), None, True, False),               # oob y = 225
    ((230, 10), (100, 10), None, False, True),               # x = 230: only A11 (col > 224)
]


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    assert torch.cuda.is_available(), "these tests need a B200"


def test_known_answer_table_through_the_reference_signatures():
    from ppnet_b200 import mpnet
    from ppnet_b200.edage import process_map
    mpnet.clearance = C
    for s, e, circ, want12, want11 in KAT:
        obs = [list(circ)] if circ else []
        got11 = process_map.collision_check_circle_edge((s[1], s[0]), (e[1], e[0]), obs, C)       # (row, col) inputs
        assert got11 == want11, ("A11", s, e, circ)
        mpnet.obc = [obs]
        got12 = mpnet.collision_check_circle_edge(torch.tensor(s, dtype=torch.float32), torch.tensor(e, dtype=torch.float32), 0)
        assert got12 == want12, ("A12", s, e, circ)
        assert mpnet.steerTo(np.float32(s), np.float32(e), 0) == (0 if want12 else 1)
    # a whole path through the batched form == edge by edge
    path = [[10.0, 10.0], [10.0, 100.0], [60.0, 150.0], [200.0, 150.0]]
    obs = [[50, 14, 2], [150, 100, 9]]
    v = process_map.collision_check_path(path, obs, C)
    assert v.tolist() == [process_map.collision_check_circle_edge(path[i], path[i + 1], obs, C) for i in range(3)]


def test_add_init_end_single_golden(golden):
    from ppnet_b200.edage import process_map
    g = golden("misc")
    for i in range(len(g["init"])):
        img = torch.from_numpy(g["img_in"][i].copy())
        out = process_map.add_init_end_single(img, g["init"][i], g["end"][i])
        assert out is img
        assert np.array_equal(out.numpy(), g["img_out"][i])
    dev_img = torch.from_numpy(g["img_in"][0].copy()).cuda()
    process_map.add_init_end_single(dev_img, g["init"][0], g["end"][0])
    assert np.array_equal(dev_img.cpu().numpy(), g["img_out"][0])


def test_mpnet_module_api_vs_oracle(golden):
    from ppnet_b200 import mpnet
    g = golden("segcheck_f32")
    mpnet.clearance = float(g["clearance"])
    mpnet.obc = [g["obs"][m, :g["obs_cnt"][m]].tolist() for m in range(len(g["obs_cnt"]))]
    po, lo = g["path_off"], g["lvc_off"]
    for p in range(min(12, len(g["path_map"]))):
        idx = int(g["path_map"][p])
        path = [torch.from_numpy(w.copy()) for w in g["path_pts"][po[p]:po[p + 1]]]
        assert mpnet.feasibility_check(path, idx) == int(g["feasible"][p])
        out = mpnet.lvc(path, idx)
        want = g["lvc_pts"][lo[p]:lo[p + 1]]
        assert len(out) == len(want)
        assert np.array_equal(np.stack([o.numpy() for o in out]), want)


def test_gmm_mirror():
    from ppnet_b200 import edage
    from ppnet_b200.edage.GMM import GMM
    edage.seed(3)
    m = GMM(10, 2)
    a = m.Distribution.sample([1000, ])
    b = m.Distribution.sample([1000])
    assert a.shape == (1000, 2) and a.is_cuda and a.dtype == torch.float32
    assert not torch.equal(a, b)                                  # the stream moves on
    edage.seed(3)
    a2 = GMM(10, 2).Distribution.sample([1000, ])
    assert torch.equal(a, a2)                                     # reproducible from the seed
    assert m.Order == 10 and m.Dim == 2
    assert float(a.min()) > -40 and float(a.max()) < 110         # means in [0, 70), std < 5


def test_path_and_pathseg_mirrors():
    from ppnet_b200 import edage
    from ppnet_b200.edage.Path import Path, plot_obstacles
    from ppnet_b200.edage.PathSeg import PathSeg
    edage.seed(11)
    seg = PathSeg(4, 2)
    poly, end = seg.random()
    assert poly.shape == (5,) and poly[-1] == 0 and 0 <= float(end[0]) <= 7
    assert seg.gradient() == (seg.GradSt, seg.GradEnd)
    ref = orc.pathseg_from_poly(poly, float(end[0]), seg.is_straight)
    assert abs(float(seg.length()[0]) - ref["Length"]) <= 1e-9 * max(ref["Length"], 1)
    p = Path(seg_num=10, poly_order=4, clearance=1, is_straight=False)
    p.generate(show_now=False)
    assert p.PathPoint.shape == (1000, 2) and p.SegPoint.shape == (11, 2) and len(p.PathSeg) == 10
    p.draw_boundary(show_now=False)
    assert p.BoundaryPoint.shape == (1100, 2)
    assert p.path_obstacles(resolution=224, map_size=50, map_offset=112)
    assert p.Space.shape == (3, 224, 224) and p.PathPoint.shape == (1000, 2)
    hull = p.ConvexHull.numpy()
    assert abs(hull.mean(axis=0) - 112).max() < 1e-6             # space_normalization centres the hull
    ok, placed = p.boundary_check(0, [0, 0])
    assert ok and np.allclose(placed, hull)
    odd = p.PathPoint[1::2]
    for x, y, r in p.obstacles:
        assert np.sqrt((odd[:, 0] - y) ** 2 + (odd[:, 1] - x) ** 2).min() >= r + 4.48 - 1e-4
    cells = p.coord_euclidean2image(np.array([[0.1, 0.2], [3.0, -4.0]]), 224)
    assert np.array_equal(cells, orc.grid_index_vec(np.array([[0.1, 0.2], [3.0, -4.0]]), 50, 224, 224))
    img = plot_obstacles((224, 224), [[50, 60, 10]], resolution=(224, 224))
    assert img.shape == (3, 224, 224) and float(img[0, 60, 50]) == 0.0 and float(img[0, 5, 5]) == 1.0


def test_mapgenerate_mirror_end_to_end(tmp_path):
    from ppnet_b200 import edage
    from ppnet_b200.edage import MapGenerate as MG
    edage.seed(5)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        mg = MG.MapGenerate(path_num=3, resolution=224, map_size=50, obstacles_num=20, clearance=3)
        mg.generate(map_num=18, folder_path=str(tmp_path), round_index=0)
        P = 3
        assert len(mg.MapLabel) == 18 and len(mg.PathGroup.TargetPaths) == P
        lines = open("unsolved_problems.txt").read().strip().split("\n")
        assert len(lines) == 18
        c_px = 3 / 50 * 224
        for g, ln in enumerate(lines):
            prob = json.loads(ln)
            label, angle, translation, segpoint, pathpoint = mg.MapLabel[g]
            assert prob["Index"] == g and len(label) == 10 and pathpoint.shape == (1000, 2) and segpoint.shape == (11, 2)
            assert np.allclose(prob["Init"], segpoint[0]) and np.allclose(prob["End"], segpoint[10])
            j = (g // P) % P                                       # index rule MapGenerate.py:68
            assert prob["Length"] == mg.PathGroup.TargetPaths[j].Length
            # the label is the rigid placement of target path j (oracle A13)
            want = orc.place_points(mg.PathGroup.TargetPaths[j].PathPoint, float(angle[0]), translation, 224)
            assert np.abs(want - pathpoint).max() < 1e-7
            odd = pathpoint[1::2]
            for x, y, r in prob["Obstacles"]:
                assert np.sqrt((odd[:, 0] - y) ** 2 + (odd[:, 1] - x) ** 2).min() > r + c_px - 1e-4
            # the reference's own post-hoc checker accepts the label path against these obstacles (every 25th edge)
            pts = pathpoint[::25]
            seg_map = np.zeros(len(pts) - 1, dtype=np.int32)
            segs = np.concatenate([pts[:-1], pts[1:]], axis=1)
            obs = np.asarray(prob["Obstacles"], dtype=np.float64).reshape(1, -1, 3)
            if obs.shape[1]:
                v = c_oracle.segcheck_f64(segs, seg_map, obs, np.asarray([obs.shape[1]], dtype=np.int32), c_px)
                assert not v.any()
        imgs = mg.map_images(0, 4)
        assert imgs.shape == (4, 3, 224, 224)
        r0, c0 = (int(np.round(v)) for v in mg.MapLabel[0][3][0])
        assert imgs[0, :, r0, c0].tolist() == [255.0, 0.0, 0.0]   # A16 stamp at the init point
        one = mg.generate_map_randomly(mg.MapLabel[0][4], mg.MapLabel[0][3][0], mg.MapLabel[0][3][10], 1.0, [], 7)
        assert one.shape == (3, 224, 224) and one.is_cuda
    finally:
        os.chdir(cwd)


def test_fused_host_call_equals_the_device_ops():
    """ppnet_generate_and_check_host (HOST buffers, slices on two streams) == generate_maps + the three verdict
    kernels called one by one on device tensors, byte for byte; ragged last slice included."""
    from ppnet_b200 import host, ops
    from ppnet_b200.synthetic import synthetic_bank, synthetic_segments
    M, SPM, R, O, reps, seed, map0 = 2500, 96, 224, 50, 10, 99, 12345
    bk = synthetic_bank(20, seed=3)
    keys = ("pathpt", "segpt", "hull", "hull_cnt", "obs", "obs_cnt")
    dbank = ops.PathBank(*[torch.from_numpy(bk[k]).cuda() for k in keys])
    hbank = host.HostBank(*[bk[k] for k in keys], device=0)
    ctx = host.HostContext(0)
    s64 = synthetic_segments(M, SPM, seed=5)
    s32 = s64.astype(np.float32)
    pomax, np_, ns1 = bk["obs"].shape[1], bk["pathpt"].shape[1], bk["segpt"].shape[1]
    hout = dict(angle=np.empty(M), trans=np.empty([M, 2], np.int32), segpt=np.empty([M, ns1, 2]), pathpt=np.empty([M, np_, 2]),
                obs=np.zeros([M, O + pomax, 3]), obs_cnt=np.empty(M, np.int32), rand_cnt=np.empty(M, np.int32),
                bits=np.empty([M, R, 7], np.int32), tries=np.empty(M, np.int32), valid=np.empty(M, np.uint8),
                counters=np.zeros(4, np.uint64))
    v64, v32, vd = (np.empty(M * SPM, np.uint8) for _ in range(3))
    host.generate_maps_host(ctx, hbank, map0, M, reps, O, hout, R, 50.0, 5.0, 1.0, seed, raster_inflate=2.24,
                            checks=dict(segs_rc_f64=s64, segs_xy_f32=s32, clearance_px=C, verdict_f64=v64, verdict_f32=v32,
                                        verdict_dda=vd))
    gen = ops.generate_maps(dbank, map0, M, reps, O, R, 50.0, 5.0, 1.0, seed, raster_inflate=2.24)
    w64 = ops.segcheck_edage_f64(torch.from_numpy(s64).cuda(), gen.obs, gen.obs_cnt, C)
    w32 = ops.segcheck_mpnet_f32(torch.from_numpy(s32).cuda(), gen.obs, gen.obs_cnt, C)
    wd = ops.dda_gridcheck(gen.bits, R, torch.from_numpy(s32).cuda(), want_first=False)
    for name, dname in (("angle", "angle"), ("trans", "trans"), ("segpt", "segpt"), ("pathpt", "pathpt"), ("obs_cnt", "obs_cnt"),
                        ("rand_cnt", "rand_cnt"), ("bits", "bits"), ("tries", "tries"), ("valid", "valid")):
        assert np.array_equal(hout[name], getattr(gen, dname).cpu().numpy()), name
    cnt = hout["obs_cnt"]
    dobs = gen.obs.cpu().numpy()
    for m in range(0, M, 97):
        assert np.array_equal(hout["obs"][m, :cnt[m]], dobs[m, :cnt[m]])
    assert np.array_equal(v64, w64.cpu().numpy()) and np.array_equal(v32, w32.cpu().numpy())
    assert np.array_equal(vd, wd.cpu().numpy())
    assert hout["counters"][0] == M and hout["counters"][1] == int(hout["valid"].sum())
    h2d, d2h = ctx.bytes_moved()
    assert h2d == M * SPM * 48 and d2h > M * SPM * 3


def test_next_rows_placement_label_masks_and_path_extraction(golden, tmp_path):
    """N1-N3 against the real reference (tests/golden/masks.npz): corridor placement pixel-exact, label masks with the
    same painted pixels, extract_path with the same waypoints / the same failures."""
    from PIL import Image
    from ppnet_b200 import ops
    from ppnet_b200.edage import process_map
    g = golden("masks")
    n = len(g["n1_angle"])
    space = torch.from_numpy(g["n1_space"]).cuda()
    idx = torch.from_numpy(g["n1_path"]).long().cuda()
    placed = ops.mask_rigid(space[idx].contiguous(), torch.from_numpy(-g["n1_angle"]).cuda(),
                            torch.from_numpy(g["n1_translation"].astype(np.float64)).cuda(), 224)
    assert np.array_equal(placed.cpu().numpy(), g["n1_placed"])                       # N1 (MapGenerate.py:102-106)
    m = process_map.gen_path_masks(g["n2_pathpoint"])
    assert np.array_equal(m.cpu().numpy() != 0, g["n2_gen_path"] != 0)               # N2a painted set
    process_map.generate_gen_path(g["n2_pathpoint"], 0, root=str(tmp_path / "mask_path"))
    for i in range(n):
        assert np.array_equal(np.asarray(Image.open(tmp_path / "mask_path" / ("%d.png" % i))), g["n2_gen_path"][i])
    pil_spaces = [Image.fromarray(np.stack([sp] * 3, axis=2)) for sp in g["n1_space"]]
    process_map.generate_seg_space(pil_spaces, g["n2_pathpoint"], list(g["n1_angle"]), g["n1_translation"].tolist(), 0,
                                   root=str(tmp_path / "mask_space"))
    for i in range(n):
        assert np.array_equal(np.asarray(Image.open(tmp_path / "mask_space" / ("%d.png" % i))), g["n2_seg_space"][i])
    # N3: batched kernel and the reference-signature wrapper
    K, ds = int(g["n3_count"]), int(g["n3_ds"])
    masks = torch.from_numpy(np.stack([g["n3_%d_mask" % k] for k in range(K)])).cuda()
    init = torch.from_numpy(np.stack([g["n3_%d_init" % k] for k in range(K)])).cuda()
    end = torch.from_numpy(np.stack([g["n3_%d_end" % k] for k in range(K)])).cuda()
    out, ln, ok = ops.extract_path(masks, init, end, float(ds), max_len=2048)
    for k in range(K):
        assert bool(ok[k].item()) == bool(g["n3_%d_ok" % k])
        want = g["n3_%d_path" % k]
        assert int(ln[k].item()) == len(want)
        assert np.array_equal(out[k, :len(want)].cpu().numpy(), want)
    assert ok.sum().item() >= 2 and (ok == 0).sum().item() >= 2
    # ... and the verdicts extract_path_image would draw from it (process_map.py:491-495): batched == edge by edge
    k0 = next(k for k in range(K) if bool(g["n3_%d_ok" % k]))
    path = g["n3_%d_path" % k0]
    obs = [[60.0, 60.0, 9.0], [150.0, 120.0, 14.0], [float(path[len(path) // 2][1]), float(path[len(path) // 2][0]), 3.0]]
    v = process_map.collision_check_path(path, obs, C)
    want_v = c_oracle.segcheck_f64(np.concatenate([path[:-1], path[1:]], axis=1), np.zeros(len(path) - 1, dtype=np.int32),
                                   np.asarray(obs)[None], np.asarray([3], dtype=np.int32), C)
    assert np.array_equal(v, want_v.astype(bool)) and v.any()


def test_extract_path_image_writes_the_references_solved_problems(golden, tmp_path, capsys):
    """N3's driver (process_map.py:452-506) through the mirror == the file the REAL reference wrote on the same dataset
    (tests/golden/make_golden.py: gen_extract_image): one path solved per line with its Length and Waypoint list, the
    collided and the failed extraction left out -- byte for byte."""
    from PIL import Image
    from ppnet_b200.edage import process_map
    g = golden("extract_image")
    n = len(g["masks"])
    folder = tmp_path / "original_data" / "0"
    (folder / "data").mkdir(parents=True)
    labels = [[None, 0.0, [0, 0], g["segpoint"][i], None] for i in range(n)]
    torch.save(labels, str(folder / "data" / "MapLabel"))
    (tmp_path / "masks").mkdir()
    for i in range(n):
        Image.fromarray(np.zeros([8, 8, 3], dtype=np.uint8)).save(str(folder / ("%d.jpg" % i)))     # the map images themselves are not read
        Image.fromarray(g["masks"][i], mode="L").save(str(tmp_path / "masks" / ("%d.png" % i)))
    Image.fromarray(np.zeros([8, 8, 3], dtype=np.uint8)).save(str(folder / "data" / "0.jpg"))
    (tmp_path / "unsolved.txt").write_bytes(g["unsolved"].tobytes())
    old = process_map.NUM_PER_FOLDER
    process_map.NUM_PER_FOLDER = n
    try:
        failed = process_map.extract_path_image(str(tmp_path / "masks"), str(tmp_path / "original_data"), str(tmp_path / "result"),
                                                str(tmp_path / "unsolved.txt"), float(g["clearance"]))
    finally:
        process_map.NUM_PER_FOLDER = old
    capsys.readouterr()
    got = (tmp_path / "result" / "solved_problems.txt").read_bytes()
    assert got == g["solved"].tobytes()
    assert failed == ["1", "2"]                     # 1: an obstacle sits on its path; 2: the heat-map has a gap


def test_extract_path_image_skips_what_the_reference_skips(golden, tmp_path, capsys):
    """process_map.py:469-483: a mask without a problem is reported and skipped; a folder that does not hold
    NUM_PER_FOLDER images is skipped silently; nothing is written when nothing is solved."""
    from PIL import Image
    from ppnet_b200.edage import process_map
    g = golden("extract_image")
    n = len(g["masks"])
    folder = tmp_path / "original_data" / "0"
    (folder / "data").mkdir(parents=True)
    torch.save([[None, 0.0, [0, 0], g["segpoint"][i], None] for i in range(n)], str(folder / "data" / "MapLabel"))
    (tmp_path / "masks").mkdir()
    for i in range(n - 1):                                       # one image short of NUM_PER_FOLDER
        Image.fromarray(np.zeros([8, 8, 3], dtype=np.uint8)).save(str(folder / ("%d.jpg" % i)))
    Image.fromarray(np.zeros([8, 8, 3], dtype=np.uint8)).save(str(folder / "data" / "0.jpg"))
    Image.fromarray(g["masks"][0], mode="L").save(str(tmp_path / "masks" / "0.png"))
    Image.fromarray(g["masks"][0], mode="L").save(str(tmp_path / "masks" / "77.png"))      # no problem carries index 77
    (tmp_path / "unsolved.txt").write_bytes(g["unsolved"].tobytes())
    old = process_map.NUM_PER_FOLDER
    process_map.NUM_PER_FOLDER = n
    try:
        failed = process_map.extract_path_image(str(tmp_path / "masks"), str(tmp_path / "original_data"), str(tmp_path / "result"),
                                                str(tmp_path / "unsolved.txt"), float(g["clearance"]))
    finally:
        process_map.NUM_PER_FOLDER = old
    out = capsys.readouterr().out
    assert "No matched problem(Index:77)" in out
    assert failed == [] and not (tmp_path / "result" / "solved_problems.txt").exists()


def test_planner_solution_masks_vs_reference(golden):
    """N4: the corridor / path label masks of planner solutions == the PNGs the real generated_by_planners wrote."""
    from ppnet_b200 import ops
    g = golden("planner_masks")
    sp, pm = ops.planner_masks(torch.from_numpy(g["wp"]).cuda(), torch.from_numpy(g["off"]).cuda())
    assert np.array_equal(sp.cpu().numpy() != 0, g["mask_space"] != 0)
    assert np.array_equal(pm.cpu().numpy() != 0, g["mask_path"] != 0)


def test_generated_by_planners_mirror_writes_the_reference_files(golden, tmp_path):
    from PIL import Image
    from ppnet_b200.edage import gerated_by_planners as gp
    g = golden("planner_masks")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with open("solved.txt", "w") as f:
            for i in range(int(g["n"])):
                w = g["wp"][g["off"][i]:g["off"][i + 1]]
                f.write(json.dumps({"Obstacles": [[50.0, 60.0, 8.0]], "Solution": [
                    {"Planner": "BITstar", "Waypoint": w.tolist(), "Time": 1.0},
                    {"Planner": "RRTstar", "Waypoint": w.tolist(), "Time": 60.0}]}) + "\n")
        gp.generated_by_planners("solved.txt")
        for i in range(int(g["n"])):
            assert np.array_equal(np.asarray(Image.open("data_BITstar/mask_space/%d.png" % i)), g["mask_space"][i])
            assert np.array_equal(np.asarray(Image.open("data_BITstar/mask_path/%d.png" % i)), g["mask_path"][i])
            assert os.path.exists("data_BITstar/map/%d.jpg" % i)
        assert os.listdir("data_RRTstar/mask_path") == []                  # Time >= 59 is filtered out
    finally:
        os.chdir(cwd)


def test_host_buffer_entry_points_equal_the_device_ops():
    """Every *_host symbol (pageable numpy in / out, slices on two streams) == the device op it wraps, incl. ragged CSR
    batches that span several slices."""
    from ppnet_b200 import host, ops
    rng = np.random.default_rng(17)
    ctx = host.HostContext(0)
    M, omax = 4000, 50
    obs = np.zeros([M, omax, 3])
    obs[..., 0] = rng.uniform(0, 224, (M, omax)); obs[..., 1] = rng.uniform(0, 224, (M, omax)); obs[..., 2] = rng.uniform(0, 22, (M, omax))
    cnt = rng.integers(0, omax + 1, M).astype(np.int32)
    per = rng.integers(0, 700, M)                               # ~1.4 M segments -> two 2^20-segment slices
    per[7] = 0
    off = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    n = int(off[-1])
    s = rng.uniform(0, 224, (n, 2))
    segs = np.concatenate([s, s + rng.normal(0, 15, (n, 2))], axis=1)
    segs32 = segs.astype(np.float32)
    d = lambda a: torch.from_numpy(a).cuda()
    want64 = ops.segcheck_edage_f64(d(segs), d(obs), d(cnt), C, seg_off=d(off)).cpu().numpy()
    assert np.array_equal(ctx.segcheck_edage_f64(segs, obs, cnt, C, seg_off=off), want64)
    w32, wst = ops.segcheck_mpnet_f32(d(segs32), d(obs), d(cnt), C, seg_off=d(off), want_steer=True)
    steer = np.empty(n, dtype=np.uint8)
    assert np.array_equal(ctx.segcheck_mpnet_f32(segs32, obs, cnt, C, seg_off=off, steer=steer), w32.cpu().numpy())
    assert np.array_equal(steer, wst.cpu().numpy())
    bits = ops.raster_circles_bits(d(obs), d(cnt), 224, C / 2)
    wv, wf = ops.dda_gridcheck(bits, 224, d(segs32), seg_off=d(off))
    fh = np.empty(n, dtype=np.int32)
    assert np.array_equal(ctx.dda_gridcheck(bits.cpu().numpy(), 224, segs32, seg_off=off, first_hit=fh), wv.cpu().numpy())
    assert np.array_equal(fh, wf.cpu().numpy())
    # uniform grouping, A14, GMM
    Mu = 5000                                                    # > one 4096-map slice
    pp = rng.uniform(20, 200, (Mu, 200, 2))
    cand = np.stack([rng.uniform(0, 50, (Mu, 20)), rng.uniform(0, 50, (Mu, 20)), rng.uniform(0, 5, (Mu, 20))], axis=2)
    acc, out, oc = ctx.clearance_filter_f64(pp, cand, 50.0, 224.0, 1.0)
    wa, wo, wc = ops.clearance_filter_f64(d(pp), d(cand), 50.0, 224.0, 1.0)
    assert np.array_equal(acc, wa.cpu().numpy()) and np.array_equal(oc, wc.cpu().numpy())
    wo = wo.cpu().numpy()
    for m in range(0, Mu, 13):                                   # rows >= out_cnt are unspecified
        assert np.array_equal(out[m, :oc[m]], wo[m, :oc[m]])
    mean, std, w = ops.gmm_params(3, 10, 2, 70.0, 5.0)
    got = ctx.gmm_sample(3, 1234, 5_000_000, mean.cpu().numpy(), std.cpu().numpy(), w.cpu().numpy())
    assert np.array_equal(got, ops.gmm_sample(3, 1234, 5_000_000, mean, std, w).cpu().numpy())
    h2d, d2h = ctx.bytes_moved()
    assert h2d > 0 and d2h > 0
    ctx.close()


def test_per_item_kernels_accept_more_than_65535_items():
    """Launches whose grid.y is the item index are chunked inside the C ABI."""
    from ppnet_b200 import ops
    n = 70001
    rng = np.random.default_rng(2)
    pp = rng.uniform(0, 16, (n, 10, 2))
    m = ops.path_mask(torch.from_numpy(pp).cuda(), resolution=16, stride=5).cpu().numpy()
    for i in (0, 65534, 65535, 65536, n - 1):
        assert np.array_equal(m[i], orc.gen_path_mask(pp[i], 16))
    src = torch.from_numpy((rng.random((n, 8, 8)) < 0.5).astype(np.uint8) * 255).cuda()
    ang = torch.from_numpy(rng.uniform(-180, 180, n)).cuda()
    tr = torch.from_numpy(rng.integers(-2, 3, (n, 2)).astype(np.float64)).cuda()
    out = ops.mask_rigid(src, ang, tr, 8).cpu().numpy()
    for i in (0, 65535, 65536, n - 1):
        assert np.array_equal(out[i], orc.mask_rigid(src[i].cpu().numpy(), float(ang[i]), tr[i].cpu().numpy(), 8))
    bits = torch.from_numpy(rng.integers(0, 2 ** 31, (n, 8, 1)).astype(np.int32)).cuda()
    img = ops.bits_to_image(bits, 8)
    assert img.shape == (n, 3, 8, 8) and float(img[n - 1].max()) <= 1.0
    b_last = int(bits[n - 1, 3, 0].item())
    assert [float(v) for v in img[n - 1, 0, 3]] == [0.0 if (b_last >> j) & 1 else 1.0 for j in range(8)]
