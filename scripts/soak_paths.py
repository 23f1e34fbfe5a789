"""Randomised soak of target-path synthesis (A1-A9) and the fused generator (A10/A13/A14): GPU vs the oracle pipeline
on Philox draws, over clearances / resolutions / seeds.  python scripts/soak_paths.py [paths_per_config] [seed]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import c_oracle, philox
from oracle import ppnet_oracle as orc
from ppnet_b200 import ops

npaths = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bad_cells = 0
for ci, (c, R, M, S) in enumerate([(1.0, 224, 50.0, 10), (3.0, 224, 50.0, 10), (1.0, 448, 50.0, 10), (2.0, 224, 50.0, 6), (2.0, 224, 100.0, 16)]):
    seed = seed0 * 1000 + ci
    out = ops.path_synthesize(5000, npaths, seg_num=S, clearance=c, map_size=M, resolution=R, seed=seed, want_space=True, hmax=96, pomax=48)
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in vars(out).items() if isinstance(v, torch.Tensor)}
    for i in range(npaths):
        forced, straight, ys, ue = philox.path_draws(seed, 5000 + i, S)
        w = orc.synthesize_path(ys, ue, straight, c, M, R, philox.path_obst_draws(seed, 5000 + i, 600), path_straight=forced)
        tol = lambda a, b, s: np.abs(np.asarray(a) - np.asarray(b)).max() <= 1e-5 * s
        assert tol(o["pathpoint_raw"][i], w["chain"]["PathPoint"], np.abs(w["chain"]["PathPoint"]).max()), ("pp_raw", ci, i)
        if not np.array_equal(o["cells"][i], w["cells"]):
            bad_cells += int((o["cells"][i] != w["cells"]).sum())       # a point within 1e-13 of a .5 tie may flip; report
            continue
        H = int(o["hull_cnt"][i])
        assert np.array_equal(o["hull_raw"][i, :H], np.asarray(w["hull_raw"])), ("hull", ci, i)
        assert np.array_equal(o["space_raw"][i], orc.corridor_paint(w["ray_x0"], w["ray_dir"], w["step_num"], M, R)), ("space", ci, i)
        assert tol(o["pathpoint"][i], w["norm"]["PathPoint"], R), ("pp", ci, i)
        assert [tuple(x) for x in o["isle"][i, :int(o["isle_cnt"][i])].tolist()] == [tuple(x) for x in w["isles"]], ("isle", ci, i)
        assert int(o["obs_cnt"][i]) == len(w["obstacles"]) and int(o["obst_rand_used"][i]) == w["used"], ("obs cnt", ci, i, o["status"][i])
        if len(w["obstacles"]):
            assert tol(o["obs"][i, :len(w["obstacles"])], w["obstacles"], R), ("obs", ci, i)
    # generator on this bank vs the oracle (A10 / A13 / A14) with Philox draws
    bank = out.to_bank()
    O, nm = 30, 40
    gen = ops.generate_maps(bank, 77, nm, 3, O, R, M, 5.0, c, seed=seed)
    torch.cuda.synchronize()
    assert int(gen.valid.sum().item()) == nm
    pp, rc, tries = gen.pathpt.cpu().numpy(), gen.rand_cnt.cpu().numpy(), gen.tries.cpu().numpy()
    ang, tr = gen.angle.cpu().numpy(), gen.trans.cpu().numpy()
    cand = np.stack([philox.candidates(seed, 77 + g, O, M, 5.0) for g in range(nm)])
    _, w_out, w_cnt = c_oracle.clearance_filter(pp, cand, M, float(R), c, threads=4)
    assert np.array_equal(rc, w_cnt)
    obs = gen.obs.cpu().numpy()
    for g in range(nm):
        assert np.array_equal(obs[g, :rc[g]], w_out[g, :rc[g]])
        j = ((77 + g) // 3) % npaths
        a, t0, t1 = philox.placement_draw(seed, 77 + g, int(tries[g]) - 1, R)
        assert a == ang[g] and (t0, t1) == (int(tr[g][0]), int(tr[g][1]))
        Hj = int(o["hull_cnt"][j])
        ok, _ = orc.boundary_check(o["hull"][j, :Hj], -a, [t1, t0], R)
        assert ok                                               # the accepted try passes the reference's boundary_check
        for t in range(int(tries[g]) - 1):                       # ... and every earlier try fails it
            a2, u0, u1 = philox.placement_draw(seed, 77 + g, t, R)
            assert not orc.boundary_check(o["hull"][j, :Hj], -a2, [u1, u0], R)[0]
        want = orc.place_points(o["pathpoint"][j], a, [t0, t1], R)
        assert np.abs(want - pp[g]).max() <= 1e-9 * R
    print("config %d ok: c=%g R=%d S=%d, %d paths, %d maps" % (ci, c, R, S, npaths, nm), flush=True)
print("soak ok; cells that flipped at a rounding tie:", bad_cells)
