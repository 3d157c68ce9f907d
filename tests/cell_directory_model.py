"""Test infrastructure (not product code): numpy models of the map's hashed cell directory, pinning the
exact-k-NN termination rule on the CPU.  CellDirectoryModel is the general ring search the rule was derived with;
HaloRuleModel (below) restates what the kernel does -- ring 1 only, float32 cell arithmetic, safety margin.

Points live in cubic cells of side `cell`; a query looks at its own cell and the rings of cells around it.  After ring
r the region examined is the cube of (2r+1)^3 cells centred on the query's cell; every point OUTSIDE that cube is at
least   g(r) = min over axes of the distance from the query to the cube's faces   away.  The k best found so far are
final as soon as the k-th squared distance is < g(r)^2 -- strict, because the reference's heap keeps the first of two
equidistant candidates (ikd_Tree.cpp:1088) and a tie outside the cube could otherwise be the rightful owner of the slot.
"""
import numpy as np


class CellDirectoryModel:
    def __init__(self, pts4, cell=1.0):
        self.pts = np.asarray(pts4, dtype=np.float32).reshape(-1, 4)
        self.cell = np.float32(cell)
        ijk = np.floor(self.pts[:, :3] / self.cell).astype(np.int64)
        self.lo = ijk.min(axis=0) if len(ijk) else np.zeros(3, np.int64)
        self.hi = ijk.max(axis=0) if len(ijk) else np.zeros(3, np.int64)
        self.cells = {}
        for i, c in enumerate(map(tuple, ijk)):
            self.cells.setdefault(c, []).append(i)

    def knn(self, q, k=5):
        """Exact k nearest neighbours of q (3 floats): (indices, float32 squared distances), ascending."""
        q = np.asarray(q, dtype=np.float32)[:3]
        c0 = np.floor(q / self.cell).astype(np.int64)
        best = []                                       # (d2 float32, index) ascending, ties keep the earlier find
        probes = 0
        r = 0
        r_max = int(max(np.abs(self.hi - c0).max(), np.abs(self.lo - c0).max())) + 1
        while True:
            for dx in range(-r, r + 1):
                for dy in range(-r, r + 1):
                    for dz in range(-r, r + 1):
                        if max(abs(dx), abs(dy), abs(dz)) != r:
                            continue                    # only the shell of ring r is new
                        probes += 1
                        for i in self.cells.get((c0[0] + dx, c0[1] + dy, c0[2] + dz), ()):
                            p = self.pts[i, :3]
                            d = q - p
                            d2 = np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
                            if len(best) < k or d2 < best[-1][0]:
                                j = len(best)
                                while j > 0 and d2 < best[j - 1][0]:
                                    j -= 1
                                best.insert(j, (d2, i))
                                del best[k:]
            lo_face = (c0 - r).astype(np.float32) * self.cell
            hi_face = (c0 + r + 1).astype(np.float32) * self.cell
            g = float(min((q - lo_face).min(), (hi_face - q).min()))
            if len(best) == k and float(best[-1][0]) < g * g:
                break
            if r >= r_max:
                break                                   # the whole map has been seen
            r += 1
        idx = np.array([b[1] for b in best], dtype=np.int64)
        d2 = np.array([b[0] for b in best], dtype=np.float32)
        return idx, d2, r, probes


class HaloRuleModel:
    """The rule the product's thread search uses (fast_lio_b200/csrc/map.cuh, cell_knn), restated in float32 numpy with the
    kernel's own arithmetic: points are filed by floor(fl(x * fl(1/cell))), a query sees the points filed in the 3x3x3 block of
    cells around its own cell (its cell's halo list), and its k best are declared EXACT when the k-th squared distance is
    strictly below g^2, g = (cell - marg) + the shrunk distance from the query to the nearest face of its own cell.  The float32
    cell arithmetic can file a point one cell off when it sits within rounding of a face; marg = 4e-6 * (max|q| + 2 cell) must
    cover that.  `knn` returns (indices, d2, proven); an unproven query is answered by the BVH walk in the product."""

    def __init__(self, pts4, cell=1.0, marg_scale=1.0):
        self.pts = np.asarray(pts4, dtype=np.float32).reshape(-1, 4)
        self.cell = np.float32(cell)
        self.inv = np.float32(1.0) / self.cell
        self.marg_scale = np.float32(marg_scale)        # 1: the product's margin; 0: the naive rule (tests show it is wrong)
        ijk = np.floor(self.pts[:, :3] * self.inv).astype(np.int64)
        self.cells = {}
        for i, c in enumerate(map(tuple, ijk)):
            self.cells.setdefault(c, []).append(i)

    def knn(self, q, k=5):
        f = np.float32
        q = np.asarray(q, dtype=np.float32)[:3]
        c0 = np.floor(q * self.inv).astype(np.int64)
        cand = []
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    cand += self.cells.get((c0[0] + dx, c0[1] + dy, c0[2] + dz), [])
        if len(cand) < k:
            return np.zeros(0, np.int64), np.zeros(0, np.float32), False
        cand = np.array(sorted(cand), dtype=np.int64)
        d = q[None, :] - self.pts[cand, :3]
        d2 = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]).astype(np.float32)
        order = np.argsort(d2, kind="stable")[:k]
        c = self.cell
        marg = self.marg_scale * f(4e-6) * (f(np.abs(q).max()) + f(2) * c)
        lo = np.maximum(q - c0.astype(np.float32) * c - marg, f(0))
        hi = np.maximum((c0 + 1).astype(np.float32) * c - q - marg, f(0))
        g = f(np.minimum(lo, hi).min() + (c - marg))
        proven = bool(d2[order[-1]] < g * g)
        return cand[order], d2[order], proven
