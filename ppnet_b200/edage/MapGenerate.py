"""EDaGe-PP/MapGenerate.py:28-151 -- dataset driver: `path_num` target paths, P*P maps per round, each a random rigid
placement of a target path + random obstacles that keep clearance to it.  One fused launch per generate() call
(ppnet_b200/csrc/generate.cu) instead of the reference's per-map Python loop."""
import json
import os

import numpy as np
import torch

from .. import ops
from . import _state
from .Path import _dev, plot_obstacles          # noqa: F401  (re-exported like the reference's `from Path import ...`)
from .PathGenerate import DIM, ORDER, PATHSEGNUM, PathGroup
from .process_map import add_init_end_single    # noqa: F401

total_record = 10000
cnt = 0


class MapGenerate:
    def __init__(self, path_num=5, resolution=224, map_size=50, obstacles_size=5, obstacles_num=50, clearance=1):
        self.device = _dev()
        self.Resolution = resolution
        self.MapSize = map_size
        self.ObstaclesNum = obstacles_num
        self.ObstacleSize = obstacles_size
        self.Clearance = clearance
        self.MapData = []
        self.PathGroup = PathGroup(path_num=path_num, resolution=resolution, map_size=map_size)
        self.PathGroup.generate(path_seg_num=PATHSEGNUM, poly_order=ORDER, dim=DIM, clearance=clearance)
        self.MapLabel = []
        self.Maps = None              # last ops.MapBatch (device): labels, obstacle sets, bit-packed occupancy
        self._problems, self._problems_list = None, []

    @property
    def Problems(self):
        """The unsolved_problems.txt records of the last generate() as dicts (built on first access)."""
        if self._problems_list is None:
            pr = self._problems
            oc = pr["obs_cnt"].tolist()
            self._problems_list = [{"Index": i, "Init": a, "End": b, "Length": l, "Obstacles": o[:c]} for i, a, b, l, o, c in
                                   zip(pr["index"].tolist(), pr["init"].tolist(), pr["end"].tolist(), pr["length"].tolist(),
                                       pr["obs"].tolist(), oc)]
        return self._problems_list

    def generate(self, map_num=100, folder_path='./', round_index=0, *, write_problems=True, save_images=False,
                 max_tries=1000000):
        """Same index rule as the reference (MapGenerate.py:42-68): rounds = round(map_num / P^2), map
        index = i P^2 + j P + k uses target path j.  Appends [label, angle, translation, segpoint, pathpoint] to
        MapLabel per map; appends the JSON problems to ./unsolved_problems.txt (first `total_record` only)."""
        global cnt
        P = len(self.PathGroup.TargetPaths)
        rounds = int(np.round(map_num / P ** 2))
        n_maps = rounds * P * P
        if n_maps == 0:
            return
        R = int(self.Resolution)
        # a distinct Philox key per (generate call, round_index): the reference's global stream simply moves on
        seed = (_state.current_seed() + 0x9E3779B97F4A7C15 * (1 + _state.next_map_call()) + 0xD1B54A32D192ED03 * round_index) \
            & 0xFFFFFFFFFFFFFFFF
        bank = self.PathGroup.bank
        gen = ops.generate_maps(bank, 0, n_maps, P, self.ObstaclesNum, R, float(self.MapSize), float(self.ObstacleSize),
                                float(self.Clearance), seed=seed, max_tries=min(int(max_tries), 2 ** 31 - 1), want_bits=True)
        self.Maps = gen
        # ONE device->host transfer per array, then everything below is array slicing: no per-map Python arithmetic
        host = {k: getattr(gen, k).cpu().numpy() for k in ("angle", "trans", "segpt", "pathpt", "obs", "obs_cnt", "valid")}
        labels = [[[seg.Poly, seg.EndPoint] for seg in tp.PathSeg] for tp in self.PathGroup.TargetPaths]
        valid = host["valid"].astype(bool)
        if not valid.all():                                # retry budget exhausted (reference: print + break)
            for _ in range(int((~valid).sum())):
                print('Error:Repeated over {} times! path:'.format(max_tries), folder_path)
            idx = np.nonzero(valid)[0]
            host = {k: v[idx] for k, v in host.items()}
        else:
            idx = np.arange(n_maps)
        jj = ((idx // P) % P).tolist()                     # index rule MapGenerate.py:68: map g uses target path (g // P) % P
        ang = host["angle"].reshape(-1, 1)
        # MapLabel entries [label, angle (1-element array), [t0, t1], segpoint, pathpoint]: rows are views of the downloads
        self.MapLabel.extend([labels[j], a, t, sp, pp] for j, a, t, sp, pp in
                             zip(jj, ang, host["trans"].tolist(), host["segpt"], host["pathpt"]))
        # unsolved_problems.txt records: only the first `total_record` maps of the process are logged (MapGenerate.py:144-149).
        # The lines are written by the library's native writer straight from the downloaded arrays (byte for byte what
        # json.dumps produces); `Problems` materialises the same records as dicts only when somebody reads it.
        n_log = max(0, min(len(idx), total_record - cnt))
        lengths = np.asarray([tp.Length for tp in self.PathGroup.TargetPaths], dtype=np.float64)
        self._problems = dict(index=np.ascontiguousarray(idx[:n_log].astype(np.int64) + round_index * 100),
                              init=np.ascontiguousarray(host["segpt"][:n_log, 0]), end=np.ascontiguousarray(host["segpt"][:n_log, -1]),
                              length=np.ascontiguousarray(lengths[np.asarray(jj[:n_log], dtype=np.int64)]) if n_log else np.zeros(0),
                              obs=np.ascontiguousarray(host["obs"][:n_log]), obs_cnt=np.ascontiguousarray(host["obs_cnt"][:n_log]))
        self._problems_list = None
        cnt += n_log
        if write_problems and n_log:
            pr = self._problems
            ops.write_problems_jsonl("./unsolved_problems.txt", pr["index"], pr["init"], pr["end"], pr["length"], pr["obs"],
                                     pr["obs_cnt"], append=True)
        if save_images:
            self.save_images(folder_path)

    def map_images(self, first=0, count=None):
        """Map images f32[n,3,R,R] of the last generate(): obstacles (A15) + placed corridor + init/end stamps (A16)."""
        gen, R, P = self.Maps, int(self.Resolution), len(self.PathGroup.TargetPaths)
        count = gen.n_maps - first if count is None else count
        sl = slice(first, first + count)
        # the corridor of target path j, placed like MapGenerate.py:102-106: rotate by -angle, translate by (t0, t1)
        j = (torch.arange(first, first + count, device=self.device) // P) % P
        space = self.PathGroup.batch.space[j].contiguous()
        tr = gen.trans[sl].to(torch.float64).contiguous()
        mask = ops.mask_rigid(space, (-gen.angle[sl]).contiguous(), tr, R)
        add = (mask.to(torch.float32) / 255.0)[:, None].repeat(1, 3, 1, 1).contiguous()
        img = ops.bits_to_image(gen.bits[sl].contiguous(), R, add=add)
        return ops.add_init_end(img, gen.segpt[sl, 0].contiguous(), gen.segpt[sl, -1].contiguous())

    def save_images(self, folder_path):
        import torchvision
        os.makedirs(os.path.join(folder_path, 'data'), exist_ok=True)
        os.makedirs(os.path.join(folder_path, 'GMM'), exist_ok=True)         # created (and left empty) by the reference too
        for j, tp in enumerate(self.PathGroup.TargetPaths):                  # MapGenerate.py:56
            torchvision.utils.save_image(tp.Space, r'{}/data/{}.jpg'.format(folder_path, j))
        imgs = self.map_images()
        for g in range(imgs.shape[0]):
            torchvision.utils.save_image(imgs[g], r'{}/{}.jpg'.format(folder_path, g))

    def generate_map_randomly(self, path_point, init, end, length, path_obstacles, index):
        """Single-map form (MapGenerate.py:126-151): draw ObstaclesNum circles, keep those that clear the path, log the
        problem, return the obstacle image Tensor[3,R,R]."""
        global cnt
        R, M, O = int(self.Resolution), float(self.MapSize), int(self.ObstaclesNum)
        seed = (_state.current_seed() + 0x9E3779B97F4A7C15 * (1 + _state.next_map_call())) & 0xFFFFFFFFFFFFFFFF
        u = ops.uniform_f64(seed, ops.STREAM_OBST, int(index), 1, 3 * O, device=self.device).reshape(3, O)
        cand = torch.stack([u[0] * M, u[1] * M, u[2] * float(self.ObstacleSize)], dim=1)[None].contiguous()
        pp = torch.from_numpy(np.ascontiguousarray(np.asarray(path_point, dtype=np.float64))[None]).to(self.device)
        _, out, n = ops.clearance_filter_f64(pp, cand, M, float(R), float(self.Clearance))
        obstacles = [[float(a), float(b), float(c)] for a, b, c in out[0, :int(n.item())].cpu().numpy()]
        if cnt < total_record:
            problem = {"Index": index, "Init": [float(v) for v in init], "End": [float(v) for v in end], "Length": length,
                       "Obstacles": obstacles + list(path_obstacles)}
            with open("./unsolved_problems.txt", "a") as f:
                f.write(json.dumps(problem) + "\n")
            cnt += 1
        return plot_obstacles((R, R), obstacles + list(path_obstacles), resolution=(R, R)).to(self.device)
