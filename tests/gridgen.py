"""Synthetic Cartesian grids in the reference's table format (test helper, numpy, brute force).

Restates, on integer cell coordinates, what the reference's grid pipeline produces for a uniform level-L grid
(SURVEY.md section 8a rows G1, G2, G4, G5):

  G1  cell order = ascending key of the reference's space-filling curve (include/common/math/hilbert.h:16-48):
      per level the quadrant bits x + 2y + 4z go through the LUT {0,3,1,2,5,4,6,7} and become one base-2^D digit.
  G2  axis neighbours: the same-level cell one step along +-x, +-y, +-z, -1 if it is not part of the domain;
      grid-level periodic links (cartesiangrid.h:608-706) are written into the axis table BEFORE diagonals exist.
  G5  diagonal neighbours by composition of axis steps (cartesiangrid.h:451-493): 2D slots 4..7 =
      N(N(c,+x),+y), N(N(c,+x),-y), N(N(c,-x),-y), N(N(c,-x),+y).  3D is not implemented in the reference
      (TERMM at :489-491); here the same rule is extended: slot i of LBMethod<D3Q27>::m_dirs
      (src/lbm/constants.h:368-399) = axis steps in x, then y, then z order, -1 as soon as one step is missing.
  G4  surfaces: for direction d, cells (ascending) whose d-neighbour is missing, normal = unit vector of d.

Used by the tests to build 2D/3D boxes of arbitrary size with periodic or wall sides, the shapes BASELINE.json's
synthetic benchmarks name.  Not used by the product.
"""
import numpy as np

LUT = np.array([0, 3, 1, 2, 5, 4, 6, 7], dtype=np.int64)

D2_DIRS = np.array([[-1, 0], [1, 0], [0, -1], [0, 1], [1, 1], [1, -1], [-1, -1], [-1, 1]], dtype=np.int64)
D3_DIRS = np.array([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1], [-1, -1, 0], [-1, 1, 0],
                    [1, -1, 0], [1, 1, 0], [-1, 0, -1], [-1, 0, 1], [1, 0, -1], [1, 0, 1], [0, -1, -1], [0, -1, 1],
                    [0, 1, -1], [0, 1, 1], [-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1], [1, -1, -1],
                    [1, -1, 1], [1, 1, -1], [1, 1, 1]], dtype=np.int64)
DIR_NAMES = ["-x", "+x", "-y", "+y", "-z", "+z"]


def sfc_key(coords, level):
    """coords [n, D] integer cell coordinates at `level` -> key of the reference's curve."""
    coords = np.asarray(coords, dtype=np.int64)
    ndim = coords.shape[1]
    key = np.zeros(coords.shape[0], dtype=np.int64)
    for l in range(level):
        bit = level - 1 - l
        q = np.zeros(coords.shape[0], dtype=np.int64)
        for d in range(ndim):
            q |= ((coords[:, d] >> bit) & 1) << d
        key = (key << ndim) | LUT[q]
    return key


def sfc_key_from_unit(x, level):
    """hilbert::index on unit-cube coordinates, written like the reference (float recursion)."""
    pos = np.array(x, dtype=np.float64)
    ndim = len(pos)
    index = 0
    for l in range(level):
        quadrant = 0
        for d in range(ndim):
            if pos[d] >= 0.5:
                quadrant |= 1 << d
        lut16 = [0, 3, 1, 2, 5, 4, 6, 7, 10, 9, 11, 8, 15, 14, 12, 13]
        index += (2 ** (ndim * (level - 1 - l))) * lut16[quadrant]
        for d in range(ndim):
            pos[d] = 2 * pos[d] - ((quadrant >> d) & 1)
    return index


def box_grid(shape, periodic=None):
    """Uniform box of `shape` cells. Returns dict with nghbr [n, 8|26] (int64), coords, level, surfaces."""
    shape = tuple(int(s) for s in shape)
    ndim = len(shape)
    periodic = tuple(periodic) if periodic is not None else (False,) * ndim
    level = int(np.ceil(np.log2(max(shape))))
    level = max(level, 1)
    grids = np.meshgrid(*[np.arange(s, dtype=np.int64) for s in shape], indexing="ij")
    coords = np.stack([g.ravel() for g in grids], axis=1)
    order = np.argsort(sfc_key(coords, level), kind="stable")
    coords = coords[order]
    n = coords.shape[0]
    strides = np.array([int(np.prod(shape[d + 1:])) for d in range(ndim)], dtype=np.int64)
    lin2cell = np.full(int(np.prod(shape)), -1, dtype=np.int64)
    lin2cell[coords @ strides] = np.arange(n)

    def step(cells, d, sgn):
        """axis neighbour of `cells` (-1 allowed) along dimension d"""
        out = np.full(cells.shape, -1, dtype=np.int64)
        ok = cells >= 0
        c = coords[cells[ok]].copy()
        c[:, d] += sgn
        if periodic[d]:
            c[:, d] %= shape[d]
            inside = np.ones(c.shape[0], dtype=bool)
        else:
            inside = (c[:, d] >= 0) & (c[:, d] < shape[d])
        res = np.full(c.shape[0], -1, dtype=np.int64)
        res[inside] = lin2cell[c[inside] @ strides]
        out[ok] = res
        return out

    dirs = D2_DIRS if ndim == 2 else D3_DIRS
    nn = dirs.shape[0]
    nghbr = np.full((n, nn), -1, dtype=np.int64)
    cells = np.arange(n, dtype=np.int64)
    for i in range(nn):
        cur = cells
        for d in range(ndim):  # x, then y, then z
            if dirs[i, d] != 0:
                cur = step(cur, d, int(dirs[i, d]))
        nghbr[:, i] = cur
    surfaces = {}
    for d in range(2 * ndim):
        miss = np.nonzero(nghbr[:, d] < 0)[0].astype(np.int64)
        normal = np.zeros(ndim)
        normal[d // 2] = -1.0 if d % 2 == 0 else 1.0
        surfaces[DIR_NAMES[d]] = (miss, np.tile(normal, (len(miss), 1)))
    h = 1.0 / max(shape)
    return dict(ndim=ndim, nghbr=nghbr, coords=coords, level=level, surfaces=surfaces, shape=shape,
                center=(coords + 0.5) * h, cell_length=h,
                bbmin=np.zeros(ndim), bbmax=np.array(shape, float) * h)
