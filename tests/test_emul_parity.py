"""CPU: the kernel LOGIC (host build of the device sources, tests/emul) against the reference goldens."""
import os
import subprocess

import pytest

from switchfl_b200 import backend
from tests._parity import check_replay
from tests._util import golden_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "network-distributed-q-learning_b200", "csrc")
EMUL = os.path.join(ROOT, "tests", "emul", "libsfl_emul.so")


def build_emul():
    srcs = [os.path.join(SRC, f) for f in ("sfl_api.cu", "sfl_core.cuh")] + [os.path.join(ROOT, "include", "switchfl_b200.h")]
    if not os.path.exists(EMUL) or any(os.path.getmtime(s) > os.path.getmtime(EMUL) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DSFL_HOST_EMUL", "-x", "c++",
                               "-I", os.path.join(ROOT, "include"), "-I", SRC, "-o", EMUL, os.path.join(SRC, "sfl_api.cu")])
    return EMUL


@pytest.fixture(scope="session")
def emul_lib():
    return build_emul()


@pytest.mark.parametrize("name", golden_names())
def test_emul_replay_matches_reference(name, emul_lib):
    check_replay(name, lambda rm, **kw: backend.Engine(rm, _emul_lib=emul_lib, **kw), n_envs=2)


def test_emul_chunked_launches_equal_one_launch(emul_lib):
    check_replay("slips24_t6", lambda rm, **kw: backend.Engine(rm, _emul_lib=emul_lib, **kw), n_envs=1, chunk=7)
