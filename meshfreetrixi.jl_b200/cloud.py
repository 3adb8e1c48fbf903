"""Point clouds: Medusa text I/O (the reference's on-disk format) and the synthetic clouds of BASELINE.json.

Setup-time host code; nothing here is on the timed path.
Reference: src/auxiliary/medusa/read_medusa_file.jl:2-64 (format), SURVEY.md section 8d (synthetic configs).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class Cloud:
    points: np.ndarray                 # (N,2) float64
    boundary_idxs: list                # per group: 0-based int64 indices
    boundary_normals: list             # per group: (n,2)
    interior_idx: np.ndarray | None = None
    extent: tuple | None = None


def read_medusa_file(casename: str) -> Cloud:
    """`<case>_positions.txt` "x , y"; `_types.txt` (>0 interior, -g boundary group g); `_boundary.txt` /
    `_interior.txt` 0-based indices; `_normals.txt` "nx , ny" (row j <-> `_boundary.txt` row j)."""
    positions = np.loadtxt(casename + "_positions.txt", delimiter=",", dtype=np.float64, ndmin=2)
    types = np.loadtxt(casename + "_types.txt", dtype=np.int64, ndmin=1)
    boundary_idx = np.loadtxt(casename + "_boundary.txt", dtype=np.int64, ndmin=1)
    interior_idx = np.loadtxt(casename + "_interior.txt", dtype=np.int64, ndmin=1)
    normals = np.loadtxt(casename + "_normals.txt", delimiter=",", dtype=np.float64, ndmin=2)
    ngroups = int(-types.min()) if types.min() < 0 else 0
    btype = -types[boundary_idx] - 1
    idxs, nrms = [], []
    for g in range(ngroups):
        sel = btype == g
        idxs.append(boundary_idx[sel].astype(np.int64))
        nrms.append(normals[sel].reshape(-1, 2))
    return Cloud(positions, idxs, nrms, interior_idx)


def write_medusa_file(casename: str, cloud: Cloud) -> None:
    n = cloud.points.shape[0]
    types = np.ones(n, dtype=np.int64)
    for g, idx in enumerate(cloud.boundary_idxs):
        types[idx] = -(g + 1)
    bidx = np.concatenate(cloud.boundary_idxs) if cloud.boundary_idxs else np.zeros(0, dtype=np.int64)
    bnrm = np.concatenate(cloud.boundary_normals) if cloud.boundary_normals else np.zeros((0, 2))
    order = np.argsort(bidx, kind="stable")
    interior = np.nonzero(types > 0)[0]
    os.makedirs(os.path.dirname(os.path.abspath(casename)), exist_ok=True)
    with open(casename + "_positions.txt", "w") as f:
        for x, y in cloud.points:
            f.write(f"{float(x)!r} , {float(y)!r}\n")
    np.savetxt(casename + "_types.txt", types, fmt="%d")
    np.savetxt(casename + "_boundary.txt", bidx[order], fmt="%d")
    np.savetxt(casename + "_interior.txt", interior, fmt="%d")
    with open(casename + "_normals.txt", "w") as f:
        for nx, ny in bnrm[order]:
            f.write(f"{float(nx)!r} , {float(ny)!r}\n")


def jittered_lattice(nx: int, ny: int, lx: float, ly: float, seed: int = 0, jitter: float = 0.5,
                     x0: float = 0.0, y0: float = 0.0, ring: bool = True) -> Cloud:
    """Synthetic scattered cloud of SURVEY.md 8d: interior x=(i+0.5+jitter*(U-0.5))*lx/nx (PCG64 `default_rng(seed)`),
    plus (ring=True) a boundary ring of exact edge points with outward unit normals, groups
    0=left(inlet) 1=right(outlet) 2=bottom 3=top."""
    rng = np.random.default_rng(seed)
    hx, hy = lx / nx, ly / ny
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    U = rng.random((nx, ny, 2))
    X = np.empty((nx * ny, 2))
    X[:, 0] = x0 + ((ii + 0.5 + jitter * (U[..., 0] - 0.5)) * hx).reshape(-1)
    X[:, 1] = y0 + ((jj + 0.5 + jitter * (U[..., 1] - 0.5)) * hy).reshape(-1)
    if not ring:
        return Cloud(X, [], [], np.arange(nx * ny), (x0, y0, lx, ly))
    ys = y0 + (np.arange(ny) + 0.5) * hy
    xs = x0 + (np.arange(nx) + 0.5) * hx
    left = np.stack([np.full(ny, x0), ys], 1)
    right = np.stack([np.full(ny, x0 + lx), ys], 1)
    bottom = np.stack([xs, np.full(nx, y0)], 1)
    top = np.stack([xs, np.full(nx, y0 + ly)], 1)
    pts = np.concatenate([X, left, right, bottom, top])
    n0 = X.shape[0]
    offs = np.cumsum([n0, ny, ny, nx, nx])
    idxs = [np.arange(offs[g], offs[g + 1], dtype=np.int64) for g in range(4)]
    nrm = [np.tile([-1.0, 0.0], (ny, 1)), np.tile([1.0, 0.0], (ny, 1)), np.tile([0.0, -1.0], (nx, 1)),
           np.tile([0.0, 1.0], (nx, 1))]
    return Cloud(pts, idxs, nrm, np.arange(n0), (x0, y0, lx, ly))


def reorder(cloud: Cloud, perm: np.ndarray) -> Cloud:
    """Renumber the points: new point d is old point perm[d] (e.g. a Hilbert order from mft_sfc_order).  The numbering
    of a scattered cloud is arbitrary; a space-filling-curve numbering makes "ascending neighbour index" -- the
    order in which the reference's CSC SpMV sums a row -- spatially coherent, which is what the gather hardware likes."""
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm), dtype=perm.dtype)
    idxs = [inv[b] for b in cloud.boundary_idxs]
    interior = inv[cloud.interior_idx] if cloud.interior_idx is not None else None
    return Cloud(np.ascontiguousarray(cloud.points[perm]), idxs, [n.copy() for n in cloud.boundary_normals], interior,
                 cloud.extent)


# ---- initial conditions (return conservative variables, shape (4,N)) -----------------------------------
def prim2cons(rho, v1, v2, p, gamma):
    return np.stack([rho, rho * v1, rho * v2, p / (gamma - 1.0) + 0.5 * rho * (v1 * v1 + v2 * v2)])


def isentropic_vortex(points: np.ndarray, gamma: float = 1.4, center=(5.0, 5.0), beta: float = 5.0,
                      base=(1.0, 1.0, 1.0, 25.0)):
    """Shu isentropic vortex (the reference has no examples directory)."""
    rho0, v10, v20, p0 = base
    dx, dy = points[:, 0] - center[0], points[:, 1] - center[1]
    r2 = dx * dx + dy * dy
    du = beta / (2.0 * np.pi) * np.exp(0.5 * (1.0 - r2))
    dT = -(gamma - 1.0) / (2.0 * gamma * p0 / rho0) * du * du
    rho = rho0 * (1.0 + dT) ** (1.0 / (gamma - 1.0))
    v1 = v10 + du * (-dy)
    v2 = v20 + du * dx
    p = p0 * (1.0 + dT) ** (gamma / (gamma - 1.0))
    return prim2cons(rho, v1, v2, p, gamma)


def sod(points: np.ndarray, gamma: float = 1.4, x_mid: float = 0.5):
    left = points[:, 0] < x_mid
    rho = np.where(left, 1.0, 0.125)
    p = np.where(left, 1.0, 0.1)
    z = np.zeros_like(rho)
    return prim2cons(rho, z, z, p, gamma)
