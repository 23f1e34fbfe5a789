"""EDaGe-PP/PathSeg.py:10-58 -- one polynomial curve piece."""
import numpy as np

from .. import ops
from . import _state

SegLenRange = 7
MinLen = 0


class PathSeg:
    def __init__(self, polyorder=4, dim=2, is_straight=False):
        self.PolyOrder = polyorder
        self.Poly = np.zeros([polyorder + 1, 1])
        self.EndPoint = 0
        self.Length = 0
        self.Translation = np.zeros([dim, 1])
        self.Rotation = 0
        self.GradSt = 0
        self.GradEnd = 0
        self._forced = bool(is_straight)
        self.is_straight = bool(is_straight)              # resolved by random() (the 20 % draw is part of the piece's stream)

    def _fill(self, b, p, i):
        """Take piece i of path p out of a PathBatch (host copies)."""
        self.Poly = b["poly"][p, i].copy()
        self.EndPoint = np.array([b["endpoint"][p, i]])
        self.is_straight = bool(b["is_straight"][p, i])
        self.Translation = b["seg_trans_local"][p, i].copy()
        self.GradSt, self.GradEnd = b["grad_st"][p, i], b["grad_end"][p, i]
        self.Length = np.array([b["seg_length"][p, i]])

    def random(self, poly=None, endpoint=None):
        import torch
        kw = {}
        if poly is not None:
            kw = dict(in_poly=torch.as_tensor(np.asarray(poly, dtype=np.float64).reshape(1, 1, -1)).cuda(),
                      in_uend=torch.as_tensor(np.asarray(endpoint, dtype=np.float64).reshape(1, 1)).cuda(),
                      in_straight=torch.tensor([[1 if self._forced else 0]], dtype=torch.uint8).cuda())
        elif self._forced:
            kw = dict(force_straight=torch.ones([1], dtype=torch.uint8).cuda())
        out = ops.path_synthesize(_state.next_path_ids(1), 1, seg_num=1, poly_order=self.PolyOrder,
                                  seed=_state.current_seed(), **kw)
        b = out.to_host()
        self._fill(b, 0, 0)
        return self.Poly, self.EndPoint

    def translation(self):
        return self.Translation

    def gradient(self):
        return self.GradSt, self.GradEnd

    def length(self):
        return self.Length
