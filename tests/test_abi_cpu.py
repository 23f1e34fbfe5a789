"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/ppnet_b200.h declares; the product never routes through the oracle."""
import ctypes
import os
import re

import pytest

from ppnet_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    names = _lib.exported_symbols()
    assert len(names) >= 8
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_error_string():
    L = _lib.lib()
    assert L.ppnet_version() >= 100
    assert isinstance(L.ppnet_last_error(), bytes)
    assert _lib.launch_count() >= 0


def test_invalid_arguments_are_rejected_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    L = _lib.lib()
    rc = L.ppnet_segcheck_edage_f64(None, ctypes.c_int64(4), None, ctypes.c_int64(0), ctypes.c_int64(1), None,
                                    None, ctypes.c_int32(0), ctypes.c_double(4.48), ctypes.c_double(224.0),
                                    ctypes.c_int32(0), None, None)
    assert rc == -1 and L.ppnet_last_error()
    rc = L.ppnet_segcheck_edage_f64(None, ctypes.c_int64(0), None, ctypes.c_int64(0), ctypes.c_int64(0), None,
                                    None, ctypes.c_int32(0), ctypes.c_double(4.48), ctypes.c_double(224.0),
                                    ctypes.c_int32(7), None, None)
    assert rc == -1          # bad dot_mode
    rc = L.ppnet_grid_index_f64(None, ctypes.c_int64(0), ctypes.c_double(50), ctypes.c_double(224),
                                ctypes.c_double(112), None, None)
    assert rc == 0           # empty input is a no-op


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; nothing under ppnet_b200/ may reference it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[./]_build|liboracle", re.M)
    for root, _, files in os.walk(os.path.join(REPO, "ppnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(root, f)) as fh:
                    assert not pat.search(fh.read()), os.path.join(root, f)


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch

    from ppnet_b200 import PPNetError, ops
    with pytest.raises(PPNetError):
        ops.grid_index_f64(torch.zeros(4, 2, dtype=torch.float64), 50, 224, 112)


def test_parameter_struct_layouts_match_the_library():
    import ctypes
    from ppnet_b200 import _lib, ops
    L = ctypes.CDLL(_lib.LIB_PATH)
    L.ppnet_sizeof_params.restype = ctypes.c_int64
    assert L.ppnet_sizeof_params(ctypes.c_int32(0)) == ctypes.sizeof(ops.GenParams)
    assert L.ppnet_sizeof_params(ctypes.c_int32(1)) == ctypes.sizeof(ops.PathParams)
    from ppnet_b200 import host
    assert L.ppnet_sizeof_params(ctypes.c_int32(2)) == ctypes.sizeof(host.PipelineIO)


def test_new_entry_points_validate_before_touching_cuda():
    import ctypes
    from ppnet_b200 import _lib, ops
    L = _lib.lib()
    assert L.ppnet_path_synthesize(None, None) == -1 and b"null params" in L.ppnet_last_error()
    p = ops.PathParams()
    p.n_paths, p.seg_num, p.poly_order, p.hmax, p.pomax, p.max_obst_iter = 1, 65, 4, 64, 32, 8
    p.resolution, p.map_size = 224.0, 50.0
    assert L.ppnet_path_synthesize(ctypes.byref(p), None) == -1 and b"seg_num" in L.ppnet_last_error()
    p.seg_num = 10
    assert L.ppnet_path_synthesize(ctypes.byref(p), None) == -1 and b"output pointer" in L.ppnet_last_error()
    p.n_paths = 0
    assert L.ppnet_path_synthesize(ctypes.byref(p), None) == 0            # empty batch: no-op, nothing dereferenced
    assert L.ppnet_mask_rigid(None, ctypes.c_int32(448), None, None, ctypes.c_int64(3), ctypes.c_int32(224), None, None) == -1
    assert L.ppnet_add_init_end(None, ctypes.c_int32(224), None, None, ctypes.c_int64(0), None) == 0
    assert L.ppnet_bits_to_image(None, ctypes.c_int32(224), ctypes.c_int64(2), None, None, None) == -1
    assert L.ppnet_generate_and_check_host(None, None, None, None) == -1


def test_mirror_modules_import_without_a_gpu_and_refuse_to_compute():
    import pytest
    from ppnet_b200 import PPNetError, edage, mpnet
    from ppnet_b200.edage import GMM, MapGenerate, Path, PathGenerate, PathSeg, process_map   # noqa: F401
    edage.seed(1)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(PPNetError):
            Path.Path(seg_num=3)
        with pytest.raises(PPNetError):
            process_map.collision_check_circle_edge((1, 1), (2, 2), [], 4.48)
        mpnet.obc = [[]]
        with pytest.raises(PPNetError):
            mpnet.steerTo([1.0, 1.0], [2.0, 2.0], 0)


def test_torch_ops_are_registered_and_have_no_cpu_kernel():
    """torch.ops.ppnet_b200.* (SURVEY 8(b)): registered by the in-tree extension, CUDA dispatch key only."""
    import torch
    from ppnet_b200 import torch_ops
    ns = torch_ops.load()
    for name in torch_ops.OPS:
        assert hasattr(ns, name), name
    with pytest.raises(NotImplementedError):
        ns.grid_index_f64(torch.zeros(4, dtype=torch.float64), 50.0, 224.0, 112.0)
    with pytest.raises(NotImplementedError):
        ns.verdict_fused(torch.zeros([32, 4], dtype=torch.float64), torch.zeros([1, 1, 3], dtype=torch.float64),
                         torch.zeros([1], dtype=torch.int32), 4.48)


def test_native_problem_writer_is_byte_equal_to_json_dumps(tmp_path):
    """N1 dataset writer (host code in the C-ABI library, no GPU needed): the lines it appends to unsolved_problems.txt are
    byte for byte what the reference's json.dumps(problem) writes (EDaGe-PP/MapGenerate.py:144-149) -- floats as repr."""
    import json
    import numpy as np
    from ppnet_b200 import ops
    rng = np.random.default_rng(0)
    n, omax = 300, 24
    vals = np.concatenate([rng.uniform(-300, 300, 4000), 10.0 ** rng.uniform(-12, 22, 2000) * rng.choice([-1, 1], 2000),
                           np.round(rng.uniform(-50, 50, 500)),
                           [0.0, -0.0, 1e16, 1e-4, 1e-5, 123456789012345678.0, 0.1, 1 / 3, 5e-324, 1.7976931348623157e308,
                            9999999999999998.0, 1e15, 99999.99999999999, float("inf"), float("-inf"), float("nan")]])
    obs = rng.choice(vals, (n, omax, 3))
    init, end, length = rng.choice(vals, (n, 2)), rng.choice(vals, (n, 2)), rng.choice(vals, n)
    cnt = rng.integers(0, omax + 1, n).astype(np.int32)
    index = np.arange(n, dtype=np.int64) * 7 - 3
    path = tmp_path / "unsolved_problems.txt"
    half = n // 2
    nb = ops.write_problems_jsonl(path, index[:half], init[:half], end[:half], length[:half], obs[:half], cnt[:half], append=False)
    nb += ops.write_problems_jsonl(path, index[half:], init[half:], end[half:], length[half:], obs[half:], cnt[half:], append=True)
    want = "".join(json.dumps({"Index": int(index[i]), "Init": init[i].tolist(), "End": end[i].tolist(), "Length": float(length[i]),
                               "Obstacles": obs[i, :cnt[i]].tolist()}) + "\n" for i in range(n))
    got = path.read_text()
    assert got == want and nb == len(want.encode())
