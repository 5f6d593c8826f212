"""Synthetic coarse-grid matrices (host side, numpy) in the reference's input format.

The reference's driver (serial_amg.c:76-96) feeds ``amg_setup`` an assembled matrix as COO
triplets read from amgdmp_{i,j,p}.dat.  These generators produce the same kind of input for
the configurations BASELINE.json names: 7-point and 27-point Poisson, an anisotropic
variable-coefficient diffusion operator and a Nek5000-style Q1 vertex-mesh Laplacian.

All generators return ``(Ai, Aj, Av)`` with 0-based int32 indices, float64 values, entries
sorted by (row, col) and no duplicates.
"""
from __future__ import annotations

import numpy as np

__all__ = ["poisson7", "poisson27", "aniso7", "sem_hex", "read_amgdmp", "write_amgdmp", "by_name"]


def _grid_index(nx, ny, nz):
    return np.arange(nx * ny * nz, dtype=np.int64).reshape(nz, ny, nx)


def _finish(rows, cols, vals, n):
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    vals = np.concatenate(vals)
    key = rows * n + cols
    order = np.argsort(key, kind="stable")
    return rows[order].astype(np.int32), cols[order].astype(np.int32), vals[order].astype(np.float64)


def _stencil(nx, ny, nz, offsets, weight_fn, diag_fn):
    """Generic constant-topology stencil assembly (Dirichlet truncation at the boundary)."""
    idx = _grid_index(nx, ny, nz)
    n = nx * ny * nz
    rows, cols, vals = [], [], []
    for (dz, dy, dx) in offsets:
        z0, z1 = max(0, -dz), nz - max(0, dz)
        y0, y1 = max(0, -dy), ny - max(0, dy)
        x0, x1 = max(0, -dx), nx - max(0, dx)
        if z0 >= z1 or y0 >= y1 or x0 >= x1:
            continue
        src = idx[z0:z1, y0:y1, x0:x1].ravel()
        dst = idx[z0 + dz:z1 + dz, y0 + dy:y1 + dy, x0 + dx:x1 + dx].ravel()
        rows.append(src)
        cols.append(dst)
        vals.append(weight_fn(src, dst, (dz, dy, dx)))
    rows.append(idx.ravel())
    cols.append(idx.ravel())
    vals.append(diag_fn(idx.ravel()))
    return _finish(rows, cols, vals, n)


def poisson7(nx, ny=None, nz=None):
    """7-point finite-difference Laplacian, Dirichlet boundary: 6 on the diagonal, -1 off."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    offs = [(0, 0, 1), (0, 0, -1), (0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0)]
    return _stencil(nx, ny, nz, offs,
                    lambda s, d, o: np.full(s.shape, -1.0),
                    lambda i: np.full(i.shape, 6.0))


def poisson27(nx, ny=None, nz=None):
    """27-point Laplacian: 26 on the diagonal, -1 to each of the 26 neighbours."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    offs = [(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
            if (dz, dy, dx) != (0, 0, 0)]
    return _stencil(nx, ny, nz, offs,
                    lambda s, d, o: np.full(s.shape, -1.0),
                    lambda i: np.full(i.shape, 26.0))


def aniso7(nx, ny=None, nz=None, seed=0, eps=0.1, spread=1.0):
    """-div(K grad u) with a cell-wise diagonal K: finite volumes, harmonic face averages.

    K = diag(kx, ky, kz); kx is log-uniform in [10^-spread, 10^spread] per cell, ky = eps*kx on one half of
    the domain (strong x/z coupling) and ky = kx/eps on the other, kz = 1.  Symmetric M-matrix;
    the Dirichlet boundary adds the boundary-face conductance to the diagonal.
    """
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    rng = np.random.default_rng(seed)
    kx = 10.0 ** rng.uniform(-spread, spread, size=(nz, ny, nx))
    half = (np.arange(nx)[None, None, :] < nx // 2)
    ky = np.where(half, eps * kx, kx / eps)
    kz = np.ones_like(kx)
    K = {2: kx, 1: ky, 0: kz}
    idx = _grid_index(nx, ny, nz)
    n = nx * ny * nz
    diag = np.zeros((nz, ny, nx))
    rows, cols, vals = [], [], []
    for axis in (0, 1, 2):
        k = K[axis]
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis] = slice(0, -1)
        hi[axis] = slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        t = 2.0 * k[lo] * k[hi] / (k[lo] + k[hi])     # face transmissibility
        rows += [idx[lo].ravel(), idx[hi].ravel()]
        cols += [idx[hi].ravel(), idx[lo].ravel()]
        vals += [-t.ravel(), -t.ravel()]
        diag[lo] += t
        diag[hi] += t
        first = [slice(None)] * 3
        last = [slice(None)] * 3
        first[axis] = slice(0, 1)
        last[axis] = slice(-1, None)
        diag[tuple(first)] += 2.0 * k[tuple(first)]   # Dirichlet faces
        diag[tuple(last)] += 2.0 * k[tuple(last)]
    rows.append(idx.ravel())
    cols.append(idx.ravel())
    vals.append(diag.ravel())
    return _finish(rows, cols, vals, n)


def _q1_hex_stiffness(xyz):
    """8x8 stiffness matrices of trilinear hexes; xyz has shape (ne, 8, 3), 2x2x2 Gauss."""
    g = 1.0 / np.sqrt(3.0)
    corners = np.array([[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [1, 1, -1],
                        [-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]], dtype=np.float64)
    ne = xyz.shape[0]
    Ke = np.zeros((ne, 8, 8))
    for q in corners * g:
        dN = np.empty((8, 3))
        for a in range(8):
            ca = corners[a]
            dN[a, 0] = 0.125 * ca[0] * (1 + ca[1] * q[1]) * (1 + ca[2] * q[2])
            dN[a, 1] = 0.125 * ca[1] * (1 + ca[0] * q[0]) * (1 + ca[2] * q[2])
            dN[a, 2] = 0.125 * ca[2] * (1 + ca[0] * q[0]) * (1 + ca[1] * q[1])
        J = np.einsum("ak,eaj->ekj", dN, xyz)          # d x_j / d xi_k
        detJ = np.linalg.det(J)
        Jinv = np.linalg.inv(J)
        G = np.einsum("ejk,ak->eaj", Jinv, dN)         # physical gradients
        Ke += np.einsum("eaj,ebj->eab", G, G) * detJ[:, None, None]
    return Ke


def sem_hex(nex, ney=None, nez=None, seed=0, jitter=0.15, neumann=True):
    """Nek5000-style coarse problem: Q1 Laplacian on the vertices of a (jittered) hex mesh.

    ``nex*ney*nez`` elements, ``(nex+1)(ney+1)(nez+1)`` vertices.  With ``neumann=True`` the
    operator is singular (constants), which is the pressure coarse problem crs_setup is given
    with ``null_space=1`` (crs.h:16); otherwise the boundary vertices are removed.
    """
    ney = nex if ney is None else ney
    nez = nex if nez is None else nez
    nvx, nvy, nvz = nex + 1, ney + 1, nez + 1
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(nvz, dtype=np.float64), np.arange(nvy, dtype=np.float64),
                          np.arange(nvx, dtype=np.float64), indexing="ij")
    pts = np.stack([x, y, z], axis=-1)
    interior = np.zeros((nvz, nvy, nvx), dtype=bool)
    interior[1:-1, 1:-1, 1:-1] = True
    pts[interior] += rng.uniform(-jitter, jitter, size=(int(interior.sum()), 3))
    vid = _grid_index(nvx, nvy, nvz)
    ez, ey, ex = np.meshgrid(np.arange(nez), np.arange(ney), np.arange(nex), indexing="ij")
    ez, ey, ex = ez.ravel(), ey.ravel(), ex.ravel()
    conn = np.stack([vid[ez + dz, ey + dy, ex + dx] for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)],
                    axis=1)                                                   # (ne, 8)
    xyz = pts.reshape(-1, 3)[conn]
    Ke = _q1_hex_stiffness(xyz)
    rows = np.repeat(conn, 8, axis=1).ravel()
    cols = np.tile(conn, (1, 8)).ravel()
    vals = Ke.reshape(-1)
    n = nvx * nvy * nvz
    key = rows * n + cols
    order = np.argsort(key, kind="stable")
    key, vals = key[order], vals[order]
    ukey, start = np.unique(key, return_index=True)
    v = np.add.reduceat(vals, start)
    r, c = ukey // n, ukey % n
    if not neumann:
        keep = interior.ravel()
        remap = np.cumsum(keep) - 1
        m = keep[r] & keep[c]
        r, c, v = remap[r[m]], remap[c[m]], v[m]
    nzm = v != 0.0
    return r[nzm].astype(np.int32), c[nzm].astype(np.int32), v[nzm].astype(np.float64)


def read_amgdmp(dirname):
    """Read amgdmp_{i,j,p}.dat (serial_amg.c:76-96): doubles, first one is the 3.14159
    endianness marker, ids are 1-based.  Returns 0-based (Ai, Aj, Av)."""
    import os
    out = []
    for k in "ijp":
        d = np.fromfile(os.path.join(dirname, "amgdmp_%s.dat" % k), dtype="<f8")
        if d.size and abs(d[0] - 3.14159) > 1e-6:
            d = d.byteswap()
            if abs(d[0] - 3.14159) > 1e-6:
                raise ValueError("amgdmp_%s.dat: endianness marker not found" % k)
        out.append(d[1:])
    Ai = out[0].astype(np.int64).astype(np.int32) - 1
    Aj = out[1].astype(np.int64).astype(np.int32) - 1
    return Ai, Aj, out[2].astype(np.float64)


def write_amgdmp(dirname, Ai, Aj, Av):
    """Write the amgdmp_{i,j,p}.dat triple (amg.c:998 dump_matrix) from 0-based COO."""
    import os
    for k, arr in zip("ijp", (np.asarray(Ai) + 1.0, np.asarray(Aj) + 1.0, np.asarray(Av))):
        d = np.concatenate([[3.14159], np.asarray(arr, dtype=np.float64)])
        d.astype("<f8").tofile(os.path.join(dirname, "amgdmp_%s.dat" % k))


def by_name(name, n, seed=0):
    """Workload lookup used by bench.py and the tests."""
    if name == "poisson7":
        return poisson7(n)
    if name == "poisson27":
        return poisson27(n)
    if name == "aniso7":
        return aniso7(n, seed=seed)
    if name == "sem_hex":
        return sem_hex(n, seed=seed)
    raise ValueError("unknown workload %r" % name)
