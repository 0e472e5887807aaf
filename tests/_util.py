"""Shared helpers for the tests: golden loading and the reference-table namespace for the oracle."""
from __future__ import annotations

import glob
import os
from types import SimpleNamespace

import numpy as np

from switchfl_b200 import mapgen

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-len(".fixture.npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.fixture.npz")))


def load_golden(name):
    fx = mapgen.load_fixture(os.path.join(GOLDEN_DIR, name + ".fixture.npz"))
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    return fx, g


def ref_tables(fx, g) -> SimpleNamespace:
    """The reference's own port-graph tables (row A0), as recorded in the golden file."""
    H, W = fx["grid"].shape
    sw = g["ref_switch"]
    ports = g["ref_ports"]
    S = len(sw)
    sw_port0 = np.concatenate([[0], np.cumsum(sw[:, 0])]).astype(np.int32)
    acts = g["ref_actions"]
    n_act = np.bincount(acts[:, 0], minlength=S)
    sw_act0 = np.concatenate([[0], np.cumsum(n_act)]).astype(np.int32)
    cells = []
    for s in range(S):
        p = int(sw_port0[s])
        # switch cell from the neighbour's prev-cell is not stored; recover it from the first decision-free field
    off = g["ref_rail_nodes_off"]
    rn = [[tuple(int(x) for x in rc) for rc in g["ref_rail_nodes"][off[i]:off[i + 1]]] for i in range(len(off) - 1)]
    from switchfl_b200 import railmap
    cells = railmap.build_switch_tables(fx["grid"]).switch_cells     # names are asserted equal at golden time
    cell_switch = np.full(H * W, -1, np.int32)
    for i, (r, c) in enumerate(cells):
        cell_switch[r * W + c] = i
    return SimpleNamespace(H=H, W=W, grid=fx["grid"], switch_cells=cells, sw_P=sw[:, 0], sw_A=sw[:, 1], sw_port0=sw_port0,
                           sw_act0=sw_act0, port_switch=ports[:, 0], port_side=ports[:, 1], port_dir=ports[:, 2],
                           port_nbr=g["ref_port_nbr"], port_dist=g["ref_port_dist"], port_prev_cell=g["ref_port_prev"],
                           port_n_intra=g["ref_port_nintra"], port_intra0=g["ref_port_intra0"], act_in=acts[:, 1],
                           act_out=acts[:, 2], act_move=acts[:, 3], cell_switch=cell_switch, rail_nodes=rn)


def hparams(g) -> dict:
    return {k[3:]: float(g[k]) for k in g if k.startswith("hp_")}


def q_dict(keys, vals) -> dict:
    out = {}
    for k, v in zip(keys, vals):
        kk = tuple(int(x) for x in k if x != -9)
        out[kk] = [float(x) for x in v if not np.isnan(x)]
    return out
