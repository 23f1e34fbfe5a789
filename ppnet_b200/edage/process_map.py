"""EDaGe-PP/process_map.py -- the two functions on the hot path: the segment-vs-circles verdict (:383-425) and the
init/end stamps (:119-145).  Scalar signatures as in the reference, plus batched forms."""
import numpy as np
import torch

from .. import ops
from .Path import _dev


def _obs_tensor(obs, dev):
    o = np.asarray([[float(v[0]), float(v[1]), float(v[2])] for v in obs], dtype=np.float64).reshape(-1, 3)
    omax = max(len(o), 1)
    t = torch.zeros([1, omax, 3], dtype=torch.float64, device=dev)
    if len(o):
        t[0, :len(o)] = torch.from_numpy(o).to(dev)
    return t, torch.tensor([len(o)], dtype=torch.int32, device=dev)


def collision_check_circle_edge(s, e, obs, clearance):
    """process_map.py:383-425: s, e in (row, col); obs = [[x, y, r], ...]; -> bool (True = collision)."""
    dev = _dev()
    pts = torch.tensor([[float(s[0]), float(s[1]), float(e[0]), float(e[1])]], dtype=torch.float64, device=dev)
    o, c = _obs_tensor(obs, dev)
    return bool(ops.segcheck_edage_f64(pts, o, c, float(clearance)).item())


def collision_check_path(path, obs, clearance):
    """The checker loop of extract_path_image (process_map.py:491-495) in one launch: path = [[row, col], ...]
    -> bool array, one verdict per consecutive pair."""
    dev = _dev()
    p = np.asarray(path, dtype=np.float64).reshape(-1, 2)
    if len(p) < 2:
        return np.zeros(0, dtype=bool)
    pts = torch.from_numpy(np.ascontiguousarray(np.concatenate([p[:-1], p[1:]], axis=1))).to(dev)
    o, c = _obs_tensor(obs, dev)
    return ops.segcheck_edage_f64(pts, o, c, float(clearance)).cpu().numpy().astype(bool)


def add_init_end_single(image, init, end):
    """process_map.py:119-145: 7x7 red squares at round(init), round(end) on image Tensor[3,R,R] (in place when the
    tensor already lives on the GPU; a CPU tensor is updated through a device copy)."""
    assert len(image.shape) == 3, "Image shape incorrect"
    assert init is not None, "Init is None"
    assert end is not None, "End is None"
    dev = _dev()
    img = image.to(device=dev, dtype=torch.float32)
    work = img.contiguous()[None]
    mk = lambda p: torch.tensor([[float(p[0]), float(p[1])]], dtype=torch.float64, device=dev)
    ops.add_init_end(work, mk(init), mk(end))
    if work.data_ptr() != image.data_ptr():
        image.copy_(work[0].to(device=image.device, dtype=image.dtype))
    return image


# ---------------------------------------------------------------------------------------------------------------
# "next" rows: label-mask rasterisers and post-hoc path extraction (process_map.py:148-191, 293-365)
# ---------------------------------------------------------------------------------------------------------------
NUM_PER_FOLDER = 400


def _label_colormap(n=256):
    """The VOC-style label colormap imgviz.label_colormap() returns (bit-interleaved label index)."""
    cm = np.zeros([n, 3], dtype=np.uint8)
    for i in range(n):
        c, r, g, b = i, 0, 0, 0
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        cm[i] = [r, g, b]
    return cm


def gen_path_masks(path_point, resolution=224):
    """generate_gen_path without the file I/O: path_point [n][Np][2] (row, col) -> uint8 tensor [n,R,R] on the GPU,
    255 at every 5th label point (strictly inside the image)."""
    dev = _dev()
    pp = torch.from_numpy(np.ascontiguousarray(np.asarray(path_point, dtype=np.float64))).to(dev)
    return ops.path_mask(pp, resolution)


def generate_gen_path(path_point, folder_index, root: str = './'):
    """process_map.py:148-163: one 'L' PNG per map, {root}/{folder_index*NUM_PER_FOLDER + i}.png.  The reference's
    float image goes through ToPILImage().convert('L'), which stores painted pixels as 1 -- kept."""
    import os
    from PIL import Image
    os.makedirs(root, exist_ok=True)
    masks = (gen_path_masks(path_point) != 0).to(torch.uint8).cpu().numpy()
    for i, m in enumerate(masks):
        Image.fromarray(m, mode='L').save('{}/{}.png'.format(root, int(folder_index) * NUM_PER_FOLDER + i))


def seg_space_masks(spaces, n_maps, rotation, translation, resolution=224):
    """generate_seg_space without the file I/O: spaces = corridor masks (PIL / array / tensor, one per target path),
    map i uses spaces[int(i / (n_maps / len(spaces)))], rotated by -rotation[i] and translated by translation[i]
    (torchvision semantics), thresholded at 0.5 -> uint8 tensor [n,R,R] in {0, 1}."""
    dev = _dev()
    src = []
    for sp in spaces:
        a = np.asarray(sp)
        if a.ndim == 3:
            a = a[..., 0] if a.shape[-1] in (1, 3, 4) else a[0]
        if a.dtype != np.uint8:
            a = (np.asarray(a, dtype=np.float64) * 255).round().astype(np.uint8)
        src.append(a)
    src = torch.from_numpy(np.ascontiguousarray(np.stack(src))).to(dev)
    idx = torch.tensor([int(i / (n_maps / len(spaces))) for i in range(n_maps)], device=dev)
    ang = torch.tensor([-float(np.reshape(r, -1)[0]) for r in rotation], dtype=torch.float64, device=dev)
    tr = torch.tensor([[float(t[0]), float(t[1])] for t in translation], dtype=torch.float64, device=dev)
    placed = ops.mask_rigid(src[idx].contiguous(), ang, tr, resolution)
    return (placed > 127).to(torch.uint8)            # ToTensor scales to [0, 1]; cv2.threshold(0.5)


def generate_seg_space(spaces, path_point, rotation, translation, folder_index, root: str = './'):
    """process_map.py:166-191: one palette PNG per map (index 255 = corridor, as the reference's float -> 'P' conversion
    stores it)."""
    import os
    from PIL import Image
    os.makedirs(root, exist_ok=True)
    masks = seg_space_masks(spaces, len(path_point), rotation, translation).cpu().numpy() * 255
    cm = _label_colormap().flatten()
    for i, m in enumerate(masks):
        img = Image.fromarray(m.astype(np.uint8), mode='P')
        img.putpalette(cm)
        img.save('{}/{}.png'.format(root, int(folder_index) * NUM_PER_FOLDER + i))


def extract_path(mask, init_state, end_state, down_sample_rate=8, max_len=4096):
    """process_map.py:293-365: mask = PIL heat-map; greedy walk from init_state to end_state on the bilinearly
    down-sampled mask.  -> (True, Tensor[L, 2]) or (False, None).  The reference's 1 s timeout is a step budget here."""
    from PIL import Image
    dev = _dev()
    small = mask.resize((int(mask.size[0] / down_sample_rate), int(mask.size[1] / down_sample_rate)), Image.BILINEAR)
    a = np.asarray(small)
    a = (a.astype(np.float32) / 255.0) if a.dtype == np.uint8 else a.astype(np.float32)       # ToTensor
    if a.ndim == 3:
        a = a[..., 0]
    m = torch.from_numpy(np.ascontiguousarray(a))[None].to(dev)
    mk = lambda p: torch.tensor([[float(p[0]), float(p[1])]], dtype=torch.float64, device=dev)
    out, ln, ok = ops.extract_path(m, mk(init_state), mk(end_state), float(down_sample_rate), max_len=max_len)
    if not bool(ok.item()):
        return False, None
    return True, out[0, :int(ln.item())].cpu()
