"""Pure-numpy restatement of the MAP-MUTATION semantics of the reference's ikd-Tree, used as a
second, independent oracle next to oracle/_ref (test infrastructure only).

  Add_Points(downsample_on=true)   include/ikd-Tree/ikd_Tree.cpp:478-548
  Add_Points(downsample_on=false)  include/ikd-Tree/ikd_Tree.cpp:549-568
  Delete_Point_Boxes               include/ikd-Tree/ikd_Tree.cpp:632-658 -> Delete_by_range :767-851
"""
import numpy as np

F = np.float32


class VoxelMapModel:
    def __init__(self, pts4, ds=0.5):
        self.ds = F(ds)
        self.pts = [np.array(p, dtype=np.float32) for p in np.asarray(pts4, dtype=np.float32).reshape(-1, 4)]
        self.valid = [True] * len(self.pts)
        self.cells = {}
        for i, p in enumerate(self.pts):
            self.cells.setdefault(self._cell(p), []).append(i)

    def _cell(self, p):
        return tuple(float(np.floor(F(p[a]) / self.ds)) for a in range(3))

    def _box(self, p):
        bmin = np.array([F(np.floor(F(p[a]) / self.ds)) * self.ds for a in range(3)], dtype=np.float32)
        bmax = (bmin + self.ds).astype(np.float32)
        mid = np.array([F(float(bmin[a]) + float(F(bmax[a] - bmin[a])) / 2.0) for a in range(3)], dtype=np.float32)
        return bmin, bmax, mid

    @staticmethod
    def _dist(a, b):
        d = (a[:3] - b[:3]).astype(np.float32)
        return F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))

    def _in_box(self, p, bmin, bmax):
        return all(bmin[a] <= p[a] and bmax[a] > p[a] for a in range(3))

    def _insert(self, p):
        self.pts.append(np.array(p, dtype=np.float32))
        self.valid.append(True)
        self.cells.setdefault(self._cell(p), []).append(len(self.pts) - 1)

    def add_points(self, batch4, downsample_on):
        counter = 0
        for p in np.asarray(batch4, dtype=np.float32).reshape(-1, 4):
            if not downsample_on:
                self._insert(p)
                continue
            bmin, bmax, mid = self._box(p)
            storage = [i for i in self.cells.get(self._cell(p), []) if self.valid[i] and self._in_box(self.pts[i], bmin, bmax)]
            min_dist = self._dist(p, mid)
            result, result_is_new = p, True
            for i in storage:
                d = self._dist(self.pts[i], mid)
                if d < min_dist:
                    min_dist, result, result_is_new = d, self.pts[i].copy(), False
            same = result_is_new or bool(np.all(np.abs(p[:3].astype(np.float64) - result[:3].astype(np.float64)) < 1e-6))
            if len(storage) > 1 or same:
                for i in storage:
                    self.valid[i] = False
                self._insert(result)
                counter += 1
        return counter

    def delete_boxes(self, boxes6):
        cnt = 0
        for b in np.asarray(boxes6, dtype=np.float32).reshape(-1, 6):
            for i, p in enumerate(self.pts):
                if self.valid[i] and self._in_box(p, b[:3], b[3:]):
                    self.valid[i] = False
                    cnt += 1
        return cnt

    def flatten(self):
        out = [p for p, v in zip(self.pts, self.valid) if v]
        return np.array(out, dtype=np.float32).reshape(-1, 4)


def sort_rows(a):
    a = np.asarray(a, dtype=np.float32).reshape(-1, 4)
    if len(a) == 0:
        return a
    order = np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))
    return a[order]


def map_incremental(scan4, x26, nearest, nearest_cnt, fsm=0.5, ekf_inited=True):
    """Restatement of map_incremental's decision logic (reference src/laserMapping.cpp:427-474).
    Returns (PointToAdd, PointNoNeedDownsample) as float32 arrays of world points, in scan order."""
    from oracle import bind
    L = bind.lib()
    to_add, no_need = [], []
    w = np.zeros(3, dtype=np.float32)
    x26 = np.ascontiguousarray(x26, dtype=np.float64)
    fsm = float(fsm)
    for i in range(len(scan4)):
        L.oracle_transform_point(x26, np.ascontiguousarray(scan4[i, :3]), w)                  # pointBodyToWorld :177-186
        pw = np.array([w[0], w[1], w[2], scan4[i, 3]], dtype=np.float32)
        cnt = int(nearest_cnt[i])
        if cnt > 0 and ekf_inited:                                                                # :438
            mid = np.array([F(np.floor(float(w[a]) / fsm) * fsm + 0.5 * fsm) for a in range(3)], dtype=np.float32)   # :444-446
            dist = VoxelMapModel._dist(pw, mid)                                                   # :447
            n0 = nearest[i, 0]
            if all(float(abs(F(n0[a] - mid[a]))) > 0.5 * fsm for a in range(3)):                  # :448
                no_need.append(pw)
                continue
            need_add = True
            if cnt >= 5:
                for j in range(5):                                                                # :452-459
                    if VoxelMapModel._dist(nearest[i, j], mid) < dist:
                        need_add = False
                        break
            if need_add:
                to_add.append(pw)
        else:
            to_add.append(pw)
    return (np.array(to_add, dtype=np.float32).reshape(-1, 4), np.array(no_need, dtype=np.float32).reshape(-1, 4))


def voxelgrid_model(pts4, leaf):
    """pcl::VoxelGrid::applyFilter (PCL 1.8-1.12, pcl/filters/impl/voxel_grid.hpp) in numpy float32, written
    independently of oracle/frontend_oracle.cpp: bounding box -> min_b/div_b -> cell index -> stable sort ->
    sequential float32 centroid per cell, cells in ascending index."""
    f32 = np.float32
    pts = np.asarray(pts4, dtype=f32).reshape(-1, 4)
    if len(pts) == 0:
        return pts.copy()
    inv = f32(1.0) / f32(leaf)
    mn, mx = pts[:, :3].min(axis=0), pts[:, :3].max(axis=0)
    d = [int(f32(f32(mx[c] - mn[c]) * inv)) + 1 for c in range(3)]
    if d[0] * d[1] * d[2] > 2 ** 31 - 1:
        return pts.copy()
    min_b = np.floor(mn * inv).astype(np.int32)
    div = np.floor(mx * inv).astype(np.int32) - min_b + 1
    ijk = (np.floor(pts[:, :3] * inv) - min_b.astype(f32)).astype(np.int32)
    idx = (ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]).astype(np.uint32)
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
    ends = np.r_[starts[1:], len(sidx)]
    out = np.zeros((len(starts), 4), dtype=f32)
    for k, (a, b) in enumerate(zip(starts, ends)):
        s = np.zeros(4, dtype=f32)
        for j in order[a:b]:
            s = s + pts[j]
        out[k] = s / f32(b - a)
    return out
