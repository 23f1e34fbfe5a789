"""EDaGe-PP/PathGenerate.py:20-50 -- PathGroup: `path_num` accepted target paths, one batched launch."""
import numpy as np
import torch

from .. import ops
from . import _state
from .Path import Path, _dev

ORDER = 4
PATHSEGNUM = 10
DIM = 2


class PathGroup:
    def __init__(self, path_num=5, resolution=224, map_size=50):
        self.device = _dev()
        self.Paths = []
        self.TargetPaths = []
        self.PathNum = path_num
        self.Resolution = resolution
        self.MapSize = map_size
        self.MapOffset = self.Resolution / 2
        self.SpacesObs = []
        self.SpacesFree = []
        self.Rotation = []
        self.Translation = []
        self.bank = None              # device-resident ops.PathBank (what MapGenerate.generate consumes)
        self.batch = None

    def generate(self, path_seg_num=3, poly_order=4, dim=2, clearance=1):
        # the reference loops until path_num paths are accepted; path_obstacles() always accepts (Path.py:174 tests a
        # tuple), so exactly path_num paths are drawn.  The 1 % forced-straight draw is part of each path's stream.
        first = _state.next_path_ids(self.PathNum)
        # capacities / iteration guards are checked (and grown) here: a truncated hull would let boundary_check accept
        # placements that leave the image
        out = ops.path_synthesize_checked(first, self.PathNum, seg_num=path_seg_num, poly_order=poly_order, clearance=clearance,
                                          map_size=self.MapSize, resolution=self.Resolution, seed=_state.current_seed(),
                                          want_space=True, device=self.device)
        self.batch = out
        self.bank = out.to_bank()
        host = out.to_host()                               # one device->host copy for everything the Path objects mirror
        space_f = out.space.to(torch.float32) / 255.0      # the corridor masks as the reference holds them (ToTensor), on the device
        for i in range(self.PathNum):
            path = Path(seg_num=path_seg_num, poly_order=poly_order, dim=dim, clearance=clearance,
                        is_straight=bool(host["path_straight"][i]))
            path._id = first + i
            path._run(self.Resolution, self.MapSize, batch=host, index=i)
            path._dev_space = space_f[i]
            path.generate(show_now=False)
            path.draw_boundary(show_now=False)
            rst = path.path_obstacles(resolution=self.Resolution, map_size=self.MapSize, map_offset=self.MapOffset)
            if rst:
                self.Paths.append(path)
                self.TargetPaths.append(path)
        return bool(np.size(self.TargetPaths))
