"""Grid description the B200 backend derives from PetIBM's YAML configuration.

Only what the pressure-Poisson operator needs: the pressure-cell
widths per axis (parser::parseMesh, src/parser/parser.cpp:239-356 -- the arithmetic itself is the
C ABI's b200ls_axis_from_subdomains), the periodicity flags (misc::checkPeriodicBC,
src/misc/misc.cpp:19-83), the time-step size (navierstokes.cpp:112) and the DMDA slab ownership
rule."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

_DIRS = {"x": 0, "y": 1, "z": 2}
_LOCS = {"xMinus": 0, "xPlus": 1, "yMinus": 2, "yPlus": 3, "zMinus": 4, "zPlus": 5,
         "left": 0, "right": 1, "bottom": 2, "top": 3, "back": 4, "front": 5}
_FIELDS = {"u": 0, "v": 1, "w": 2}


def axis_from_subdomains(start: float, subdomains) -> np.ndarray:
    """Cell widths of one axis from the YAML sub-domain list (parser.cpp:297-356, misc.h:148-163)."""
    L = _lib.lib()
    ends = np.ascontiguousarray([float(s["end"]) for s in subdomains], dtype=np.float64)
    cells = np.ascontiguousarray([int(s["cells"]) for s in subdomains], dtype=np.int32)
    ratios = np.ascontiguousarray([float(s["stretchRatio"]) for s in subdomains], dtype=np.float64)
    cap = int(cells.sum())
    out = np.empty(max(cap, 1), dtype=np.float64)
    n = C.c_int(0)
    _lib.check(L.b200ls_axis_from_subdomains(
        float(start), len(subdomains), ends.ctypes.data_as(_lib._dp), cells.ctypes.data_as(_lib._ip),
        ratios.ctypes.data_as(_lib._dp), out.ctypes.data_as(_lib._dp), cap, C.byref(n)))
    return out[: n.value].copy()


@dataclass
class Grid:
    """Pressure grid of a CartesianMesh: widths per axis, periodic flags, time step."""

    widths: list                      # [dx, dy(, dz)] numpy arrays
    periodic: tuple = (False, False, False)
    dt: float = 1.0
    bounds: list = field(default_factory=list)   # [(start, end)] per axis

    @property
    def dim(self) -> int:
        return len(self.widths)

    @property
    def n(self):
        return tuple(int(w.size) for w in self.widths)

    @property
    def size(self) -> int:
        return int(np.prod(self.n))

    @staticmethod
    def uniform(n, lo=0.0, hi=1.0, periodic=(False, False, False), dt=0.01) -> "Grid":
        widths = [axis_from_subdomains(lo, [{"end": hi, "cells": int(m), "stretchRatio": 1.0}]) for m in n]
        per = tuple(bool(p) for p in periodic)[: len(n)] + (False,) * (3 - len(n))
        return Grid(widths, per, float(dt), [(lo, hi)] * len(n))

    @staticmethod
    def from_config(node) -> "Grid":
        """node: the merged YAML settings (parser::getSettings) as nested dicts/lists."""
        mesh = node["mesh"]
        dim = len(mesh)
        widths = [None] * dim
        bounds = [None] * dim
        for ax in mesh:
            d = _DIRS[str(ax["direction"])]
            start = float(ax["start"])
            widths[d] = axis_from_subdomains(start, ax["subDomains"])
            bounds[d] = (start, float(ax["subDomains"][-1]["end"]))
        if any(w is None for w in widths):
            raise ValueError("mesh node does not describe every direction once")
        periodic = parse_periodic(node, dim)
        dt = float(node["parameters"]["dt"])
        return Grid(widths, periodic, dt, bounds)


def parse_periodic(node, dim: int):
    """misc::checkPeriodicBC: a direction is periodic iff both faces are PERIODIC for every field."""
    types = [[None] * 6 for _ in range(3)]
    for sub in node["flow"]["boundaryConditions"]:
        loc = _LOCS[str(sub["location"])]
        for key, val in sub.items():
            if key == "location":
                continue
            types[_FIELDS[key]][loc] = str(val[0]).upper()
    per = [False, False, False]
    for d in range(dim):
        flags = []
        for f in range(dim):
            p1 = types[f][2 * d] == "PERIODIC"
            p2 = types[f][2 * d + 1] == "PERIODIC"
            if p1 != p2:
                raise ValueError(f"periodic BC on one side only (direction {d}, field {f})")
            flags.append(p1)
        if any(flags) and not all(flags):
            raise ValueError(f"direction {d}: periodic for some velocity fields only")
        per[d] = all(flags)
    return tuple(per)


def slab_range(nslow: int, rank: int, nranks: int):
    """PETSc DMDA ownership along one axis: the first (M mod m) ranks own one more plane."""
    base, rem = divmod(int(nslow), int(nranks))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
