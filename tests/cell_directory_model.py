"""Round-2 groundwork (not product code): a numpy model of the hashed cell directory DESIGN.md §6c proposes for the
map, to pin down the exact-k-NN termination rule before it is written as a kernel.

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
