"""Builds ppnet_b200/lib/libppnet_torch.so: the TORCH_LIBRARY registration (torch_ops.cpp) linked against the C-ABI
library next to it.  Plain g++ with torch's own include / library paths (no JIT cache: the .so must live in-tree)."""
import os
import subprocess
import sys

import torch
from torch.utils import cpp_extension

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.abspath(os.path.join(HERE, "..", "..", "lib"))
SRC = os.path.join(HERE, "torch_ops.cpp")
OUT = os.path.join(LIB, "libppnet_torch.so")


def build(force=False):
    core = os.path.join(LIB, "libppnet_b200.so")
    if not os.path.exists(core):
        raise RuntimeError("build the C-ABI library first (make -C ppnet_b200/csrc)")
    hdr = os.path.abspath(os.path.join(HERE, "..", "..", "..", "include", "ppnet_b200.h"))
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        return OUT
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = cpp_extension.include_paths() + [os.path.join(cuda_home, "include")]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
            SRC, "-o", OUT] + ["-I" + p for p in inc] +
           ["-L" + tlib, "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda", "-L" + LIB, "-lppnet_b200",
            "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib, "-Wl,--no-as-needed"])
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
