"""EDaGe-PP/Path.py -- a C1 path of polynomial pieces, its clearance boundary, corridor, hull, isles and
path-hugging obstacles.  Same attributes and method signatures as the reference; every method is a view of one
`ppnet_path_synthesize` launch (ppnet_b200/csrc/path_synth.cu)."""
import numpy as np
import torch

from .. import ops
from . import _state
from .PathSeg import PathSeg

DIM = 2
SEGLENGTH = 3
MAPSIZE = 100


class BoundaryOneSide:
    def __init__(self):
        self.point = []
        self.direction = []


class Boundary:
    def __init__(self):
        self.upboundary = BoundaryOneSide()
        self.downboundary = BoundaryOneSide()
        self.initboundary = []
        self.endboundary = []


def _dev():
    if not torch.cuda.is_available():
        raise ops.PPNetError("ppnet_b200 needs a CUDA device (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


RASTER_MODE = "disk"    # "disk": pixel centre inside the disk (what the generator's bit-packed maps use);
                        # "canvas": the matplotlib canvas model of Path.py:36-49 (axes affine, crop, bilinear resize)


def plot_obstacles(size: tuple, obstacles, resolution: tuple = (224, 224)):
    """Path.py:36-49 restated geometrically (Agg anti-aliasing / JPEG / PIL dither are not reproducible): white (1) free,
    black (0) obstacle.  RASTER_MODE picks the rule: "disk" = pixel centre inside a disk [x, y, r]; "canvas" = the
    reference's canvas geometry (ppnet_raster_canvas_bits).  -> Tensor[3, H, W] on the GPU."""
    r = int(resolution[0])
    obs = np.asarray([[float(o[0]), float(o[1]), float(o[2])] for o in obstacles], dtype=np.float64).reshape(-1, 3)
    omax = max(len(obs), 1)
    o = torch.zeros([1, omax, 3], dtype=torch.float64, device=_dev())
    cnt = torch.tensor([len(obs)], dtype=torch.int32, device=o.device)
    if RASTER_MODE == "canvas":
        if len(obs):
            o[0, :len(obs)] = torch.from_numpy(obs).to(o.device)
        bits = ops.raster_canvas_bits(o, cnt, (float(size[0]), float(size[1])), r)
    elif RASTER_MODE == "disk":
        scale = r / float(size[0])
        if len(obs):
            o[0, :len(obs)] = torch.from_numpy(obs * scale).to(o.device)
        bits = ops.raster_circles_bits(o, cnt, r)
    else:
        raise ops.PPNetError("RASTER_MODE must be 'disk' or 'canvas'")
    return ops.bits_to_image(bits, r)[0]


class Path:
    def __init__(self, seg_num=3, poly_order=3, dim=2, clearance=1, is_straight=True):
        self.device = _dev()
        self.PathSeg = []
        self.SegPoint = [[0, 0]]
        self.PathPoint = []
        self.SegPointImage = []
        self.obstacles = []
        self.SegNum = seg_num
        self.PolyOrder = poly_order
        self.Dim = dim
        self.Boundary = Boundary()
        self.BoundaryPoint = []
        self.Clearance = clearance
        self.EndPoint = np.array([0, 0])
        self.Translation = np.array([0, 0])
        self.Rotation = 0
        self.Space = torch.zeros([1])
        self.PathObs = torch.zeros([1])
        self.Resolution = 0
        self.MapSize = 0
        self.MapOffset = 0
        self.ConvexHull = []
        self.is_straight = is_straight
        self.Length = 0
        self._id = None
        self._polys = None
        self._b = None            # host copies of the last PathBatch
        self._key = None          # (resolution, map_size, width_coef) it was computed for

    # ------------------------------------------------------------------ device work
    def _run(self, resolution=224, map_size=50, batch=None, index=0, width_coef=0.2):
        """(Re)compute everything for this path at the given raster; `batch`/`index` adopt a row of a PathGroup launch."""
        key = (int(resolution), float(map_size), float(width_coef))
        if batch is None:
            if self._b is not None and self._key == key:
                return self._b
            if self._id is None:
                self._id = _state.next_path_ids(1)
            kw = {}
            if self._polys is not None:                  # generate(polys=...): PathSeg.random(poly, endpoint)
                pl = np.asarray(self._polys, dtype=np.float64)
                kw = dict(in_poly=torch.from_numpy(np.ascontiguousarray(pl[:, :self.PolyOrder + 1])[None]).to(self.device),
                          in_uend=torch.from_numpy(np.ascontiguousarray(pl[:, self.PolyOrder + 1])[None]).to(self.device),
                          in_straight=torch.full([1, self.SegNum], 1 if self.is_straight else 0, dtype=torch.uint8,
                                                 device=self.device))
            if self.is_straight:
                kw["force_straight"] = torch.ones([1], dtype=torch.uint8, device=self.device)
            else:
                kw.setdefault("force_straight", torch.zeros([1], dtype=torch.uint8, device=self.device))
            out = ops.path_synthesize_checked(self._id, 1, seg_num=self.SegNum, poly_order=self.PolyOrder, clearance=self.Clearance,
                                              map_size=map_size, resolution=resolution, seed=_state.current_seed(), want_space=True,
                                              device=self.device, width_coef=width_coef, **kw)
            self._b = {k: v[0] for k, v in out.to_host().items()}
        else:
            self._b = {k: v[index] for k, v in batch.items()}
        self._key = key
        return self._b

    # ------------------------------------------------------------------ reference API
    def generate(self, show_now=True, polys=None):
        self._polys = polys
        b = self._run() if self._b is None else self._b
        # the pieces are rows of the path's arrays: one vectorised read per attribute, then plain object construction
        poly, endp, straight = b["poly"], b["endpoint"].reshape(-1, 1), b["is_straight"].tolist()
        gst, gend, slen = b["grad_st"], b["grad_end"], b["seg_length"].reshape(-1, 1)
        rot, trans = b["seg_rot"], b["seg_trans"]
        self.PathSeg = []
        for i in range(self.SegNum):
            seg = PathSeg(self.PolyOrder, self.Dim, is_straight=self.is_straight)
            seg.Poly, seg.EndPoint, seg.is_straight = poly[i], endp[i], bool(straight[i])
            seg.GradSt, seg.GradEnd, seg.Length = gst[i], gend[i], slen[i]
            seg.Rotation = rot[i] if i else False
            seg.Translation = trans[i]
            self.PathSeg.append(seg)
        self.SegPoint = b["segpoint_raw"]
        self.PathPoint = b["pathpoint_raw"]
        self.Length = float(b["length"])
        self.EndPoint = self.SegPoint[-1]

    def draw_boundary(self, show_now=True):
        b = self._b if self._b is not None else self._run()
        self.Boundary = Boundary()
        self.Boundary.upboundary.point = [p for p in b["up"]]
        self.Boundary.upboundary.direction = [p for p in b["up_dir"]]
        self.Boundary.downboundary.point = [p for p in b["down"]]
        self.Boundary.downboundary.direction = [-p for p in b["up_dir"]]
        self.Boundary.initboundary = [p for p in b["cap_init"]]
        self.Boundary.endboundary = [p for p in b["cap_end"]]
        self.BoundaryPoint = b["boundary_raw"]

    def boundary_check(self, angle, translation):
        hull = np.asarray(self.ConvexHull, dtype=np.float64).reshape(1, -1, 2)
        h = torch.from_numpy(np.ascontiguousarray(hull)).to(self.device)
        cnt = torch.tensor([hull.shape[1]], dtype=torch.int32, device=self.device)
        ang = torch.tensor([float(np.reshape(angle, -1)[0])], dtype=torch.float64, device=self.device)
        tr = torch.tensor([[float(translation[0]), float(translation[1])]], dtype=torch.float64, device=self.device)
        ok, out = ops.boundary_check(h, cnt, None, ang, tr, float(self.Resolution), want_hull=True)
        return bool(ok.item()), out[0].cpu().numpy()

    def path_space(self, resolution=224, map_size=50, map_offset=112):
        # the reference only ever calls this with map_offset = resolution / 2 (PathGenerate.py:29; the default 112 = 224 / 2):
        # boundary_check (Path.py:100-111) centres its rotation there, and so do the kernels
        if float(map_offset) != float(resolution) / 2:
            raise ops.PPNetError("path_space: map_offset must be resolution / 2 (got %r for resolution %r)" % (map_offset, resolution))
        self.Resolution, self.MapSize, self.MapOffset = resolution, map_size, map_offset
        b = self._run(resolution, map_size)
        H = int(b["hull_cnt"])
        self.Rotation = float(b["rotation"])
        self.Translation = [b["translation"][0], b["translation"][1]]
        self.ConvexHull = torch.from_numpy(b["hull"][:H])
        self.SegPointImage = b["segpoint_img"]
        self.PathPoint = b["pathpoint"]
        self.BoundaryPoint = b["boundary"]
        dev_space = getattr(self, "_dev_space", None)            # a PathGroup launch keeps the corridor masks (already float, 0..1) on the device
        if dev_space is not None and self._key == (int(resolution), float(map_size), 0.2):
            mask = dev_space
        else:
            mask = torch.from_numpy(b["space"].astype(np.float32) / 255.0).to(self.device)
        self.Space = mask[None].expand(3, -1, -1)                # three identical channels (ToTensor of an RGB copy, Path.py:136-137)
        return True, self.Space

    def path_obstacles(self, resolution=224, map_size=50, map_offset=112):
        rst, _ = self.path_space(resolution, map_size, map_offset)
        if not rst:
            return False
        if self.is_straight:
            self.PathObs = torch.ones([3, resolution, resolution])
        else:
            b = self._b
            self.obstacles = [[float(o[0]), float(o[1]), float(o[2])] for o in b["obs"][:int(b["obs_cnt"])]]
        return True

    @staticmethod
    def coord_rotation(x, radians):
        rotation = np.reshape([[np.cos(radians), -np.sin(radians)], [np.sin(radians), np.cos(radians)]], [2, 2])
        return np.dot(rotation, x)

    def coord_euclidean2image(self, x, mapoffset):
        pts = torch.from_numpy(np.ascontiguousarray(np.reshape(np.asarray(x, dtype=np.float64), [-1, 2]))).to(self.device)
        return ops.grid_index_f64(pts, float(self.MapSize), float(self.Resolution), float(mapoffset)).cpu().numpy().astype(np.int64)

    def convexhull(self):
        cells = torch.from_numpy(self.coord_euclidean2image(self.PathPoint, mapoffset=self.Resolution).astype(np.int32))
        hull, cnt = ops.hull2d_i32(cells[None].contiguous().to(self.device), hmax=128)
        hull_point = hull[0, :int(cnt.item())].float().cpu()
        return hull_point, torch.mean(hull_point, dim=0)

    def free_space_bydirection(self, space, x_init, dir, step_num, mapoffset, value=255):
        w, h = int(space.shape[0]), int(space.shape[1])
        mk = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64).reshape(1, 1, 2)).to(self.device)
        painted = ops.corridor_paint(mk(x_init), mk(dir), torch.tensor([float(step_num)], dtype=torch.float64, device=self.device),
                                     float(self.MapSize), float(self.Resolution), float(mapoffset), w, h, value=value)[0]
        on = painted.to(space.device) != 0
        space[on] = value
        return space

    def search_isle(self, width_coef=0.2):
        """Path.py:502-537.  A width_coef other than the cached launch's recomputes this path (same Philox draws, so the
        geometry is unchanged; only the isle depth threshold int(round(c / step * width_coef)) moves)."""
        if self._b is None or self._key is None:
            raise ops.PPNetError("search_isle: call path_space / path_obstacles first (it needs Resolution, MapSize and the hull)")
        if float(width_coef) != self._key[2]:
            if self._id is None:
                raise ops.PPNetError("search_isle: this Path was adopted from a batch without an id; cannot recompute")
            self._b = None
            self._run(self._key[0], self._key[1], width_coef=width_coef)
        b = self._b
        return [self.PathPoint[lo:hi] for lo, hi in b["isle"][:int(b["isle_cnt"])]]

    def set_obstacles(self, boundarys):
        """Path.py:463-500 on the isles search_isle returned (the reference's only call, Path.py:151-152).  The kernel
        places the obstacles of exactly those isles; other boundary lists are refused rather than silently ignored."""
        b = self._b
        if b is None:
            raise ops.PPNetError("set_obstacles: call path_space / path_obstacles first")
        isles = [self.PathPoint[lo:hi] for lo, hi in b["isle"][:int(b["isle_cnt"])]]
        given = list(boundarys)
        same = len(given) == len(isles) and all(np.shape(g) == np.shape(i) and np.array_equal(np.asarray(g), np.asarray(i))
                                                for g, i in zip(given, isles))
        if not same:
            raise ops.PPNetError("set_obstacles: only the isles returned by search_isle() of this path are supported")
        return [[float(o[0]), float(o[1]), float(o[2])] for o in b["obs"][:int(b["obs_cnt"])]]
