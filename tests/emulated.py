"""Test infrastructure: the drop-in classes on the HOST build of the device sources (tests/emul/libsfl_emul.so).

The build container has no GPU; ``EmulEngine`` / ``EmulSwitchEnv`` let the ``-m "not gpu"`` tests drive the kernel logic
and the Python host layer against the golden vectors.  They live here, not in the product package: the product's
``backend.Engine`` only ever loads ``libswitchfl_b200.so`` and raises without a CUDA device."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from switchfl_b200 import api, backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "network-distributed-q-learning_b200", "csrc")
EMUL = os.path.join(ROOT, "tests", "emul", "libsfl_emul.so")


def build_emul() -> str:
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cu", ".cuh", ".h"))] + [os.path.join(ROOT, "include", "switchfl_b200.h")]
    if not os.path.exists(EMUL) or any(os.path.getmtime(s) > os.path.getmtime(EMUL) for s in srcs):
        os.makedirs(os.path.dirname(EMUL), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DSFL_HOST_EMUL", "-x", "c++",
                               "-I", os.path.join(ROOT, "include"), "-I", SRC, "-o", EMUL, os.path.join(SRC, "sfl_api.cu")])
    return EMUL


def emul_library() -> C.CDLL:
    return backend.load_library(build_emul())


def emul_distance_map(grid, target_cells):
    return backend.device_distance_map(grid, target_cells, lib=emul_library())


class EmulEngine(backend.Engine):
    """``backend.Engine`` with the host build underneath: buffers are CPU tensors, there is no stream."""

    def _open(self, device):
        return emul_library(), self.torch.device("cpu"), 0

    @staticmethod
    def bfs_device(device):
        return None                                       # host BFS (railmap.distance_to)

    def _stream(self):
        return None

    def _side_stream(self):
        return None

    def _upload(self, name, host):
        raw = np.ascontiguousarray(host).view(np.uint8).reshape(-1)
        self.h2d_bytes += raw.size
        self.buf[name][:raw.size].copy_(self.torch.from_numpy(raw))

    def _download(self, name, nbytes=None):
        t = self.buf[name] if nbytes is None else self.buf[name][:nbytes]
        return t.numpy().copy()


class EmulSwitchEnv(api.ASyncSwitchEnv):
    engine_cls = EmulEngine
