"""ctypes binding of libfastlio_b200.so (the C ABI in include/fastlio_b200.h).

The reference is C++ (no Python anywhere), so this module is only the harness the tests,
bench.py and __graft_entry__ use to reach the product; names mirror the reference's two
classes: `KdTree` ~ KD_TREE<PointType> (include/ikd-Tree/ikd_Tree.h:48-341) and `Esekf` ~
esekfom::esekf<state_ikfom,12,input_ikfom> (include/IKFoM_toolkit/esekfom/esekfom.hpp:105).

There is no CPU fallback: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class PassLog(C.Structure):
    _fields_ = [("searched", C.c_int), ("valid", C.c_int), ("effct", C.c_int), ("converged", C.c_int),
                ("res_sum", C.c_double), ("HtH", C.c_double * 144), ("Hth", C.c_double * 12),
                ("x_after", C.c_double * 26)]


class FastLioError(RuntimeError):
    pass


_lib = None

# every symbol include/fastlio_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "fl_last_error", "fl_device_count", "fl_version", "fl_host_register", "fl_host_unregister", "fl_filter_debug_prof",
    "fl_map_create", "fl_map_destroy", "fl_map_set_downsample", "fl_map_build", "fl_map_size", "fl_map_validnum",
    "fl_map_knn", "fl_map_add_points", "fl_map_delete_boxes", "fl_map_flatten", "fl_map_tree_range",
    "fl_map_rebuild", "fl_map_stats", "fl_map_add_boxes", "fl_map_acquire_removed", "fl_map_set_cell_directory", "fl_map_dir_stats",
    "fl_filter_create", "fl_filter_destroy", "fl_filter_set_params", "fl_filter_set_solver", "fl_filter_set_search", "fl_filter_set_fused", "fl_filter_update",
    "fl_filter_map_incremental", "fl_filter_get_nearest", "fl_filter_get_selected", "fl_filter_get_pass_logs", "fl_filter_upload_scan",
    "fl_filter_upload_state", "fl_filter_run", "fl_filter_download_state", "fl_filter_sync",
    "fl_filter_time_resident", "fl_filter_time_search_pass", "fl_filter_time_e2e", "fl_filter_gpu_launches",
    "fl_scan_create", "fl_scan_destroy", "fl_scan_upload", "fl_scan_undistort", "fl_scan_voxel_downsample", "fl_scan_download",
    "fl_filter_update_scan", "fl_localmap_create", "fl_localmap_destroy", "fl_localmap_segment", "fl_localmap_get",
    "fl_comm_unique_id", "fl_filter_comm_init", "fl_filter_set_shard", "fl_filter_p2p_handle", "fl_filter_p2p_connect",
]


def lib_path() -> str:
    return _build.LIB


def load():
    """Load libfastlio_b200.so (never builds implicitly on a GPU box: the .so ships with the tree)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB):
        raise FastLioError(f"{_build.LIB} is missing -- run `python -m fast_lio_b200.build` (nvcc, sm_100a); "
                           "there is no CPU fallback")
    L = C.CDLL(_build.LIB)
    L.fl_last_error.restype = C.c_char_p
    L.fl_host_register.argtypes = [C.c_void_p, C.c_ulonglong]
    L.fl_host_unregister.argtypes = [C.c_void_p]
    L.fl_map_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_float]
    L.fl_map_destroy.argtypes = [C.c_void_p]
    L.fl_map_set_downsample.argtypes = [C.c_void_p, C.c_float]
    L.fl_map_build.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.fl_map_size.argtypes = [C.c_void_p]
    L.fl_map_validnum.argtypes = [C.c_void_p]
    L.fl_map_knn.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, _f32p, _i32p]
    L.fl_map_add_points.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int]
    L.fl_map_delete_boxes.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.fl_map_flatten.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.fl_map_add_boxes.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.fl_map_acquire_removed.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.fl_map_tree_range.argtypes = [C.c_void_p, _f32p]
    L.fl_map_rebuild.argtypes = [C.c_void_p]
    L.fl_map_stats.argtypes = [C.c_void_p, _i32p]
    L.fl_map_set_cell_directory.argtypes = [C.c_void_p, C.c_int, C.c_float]
    L.fl_map_dir_stats.argtypes = [C.c_void_p, _i32p]
    L.fl_filter_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
    L.fl_filter_destroy.argtypes = [C.c_void_p]
    L.fl_filter_set_params.argtypes = [C.c_void_p, C.c_int, _f64p, C.c_int]
    L.fl_filter_set_solver.argtypes = [C.c_void_p, C.c_int]
    L.fl_filter_set_search.argtypes = [C.c_void_p, C.c_int]
    L.fl_filter_set_fused.argtypes = [C.c_void_p, C.c_int]
    L.fl_filter_update.argtypes = [C.c_void_p, _f32p, C.c_int, _f64p, _f64p, C.c_double, C.POINTER(C.c_double)]
    L.fl_filter_map_incremental.argtypes = [C.c_void_p, C.c_double, C.c_int, _i32p]
    L.fl_filter_get_nearest.argtypes = [C.c_void_p, _f32p, _i32p, C.c_int]
    L.fl_filter_get_selected.argtypes = [C.c_void_p, _u8p, C.c_int]
    L.fl_filter_get_pass_logs.argtypes = [C.c_void_p, C.POINTER(PassLog), C.c_int]
    L.fl_filter_upload_scan.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.fl_filter_upload_state.argtypes = [C.c_void_p, _f64p, _f64p, C.c_double]
    L.fl_filter_run.argtypes = [C.c_void_p]
    L.fl_filter_download_state.argtypes = [C.c_void_p, _f64p, _f64p, C.POINTER(C.c_int)]
    L.fl_filter_sync.argtypes = [C.c_void_p]
    L.fl_filter_time_resident.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.fl_filter_time_search_pass.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.fl_filter_gpu_launches.argtypes = [C.c_void_p]
    L.fl_filter_time_e2e.argtypes = [C.c_void_p, _f32p, C.c_int, _f64p, _f64p, C.c_double, C.c_int, C.POINTER(C.c_double), _f64p, _f64p]
    L.fl_scan_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p]
    L.fl_scan_destroy.argtypes = [C.c_void_p]
    L.fl_scan_upload.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
    L.fl_scan_undistort.argtypes = [C.c_void_p, _f64p, C.c_int, _f64p]
    L.fl_scan_voxel_downsample.argtypes = [C.c_void_p, C.c_float]
    L.fl_scan_download.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int]
    L.fl_filter_update_scan.argtypes = [C.c_void_p, C.c_void_p, _f64p, _f64p, C.c_double, C.POINTER(C.c_double)]
    L.fl_localmap_create.argtypes = [C.POINTER(C.c_void_p), C.c_double, C.c_float]
    L.fl_localmap_destroy.argtypes = [C.c_void_p]
    L.fl_localmap_segment.argtypes = [C.c_void_p, C.c_void_p, _f64p, _f32p, C.POINTER(C.c_int)]
    L.fl_localmap_get.argtypes = [C.c_void_p, _f32p]
    L.fl_comm_unique_id.argtypes = [C.c_char_p]
    L.fl_filter_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    L.fl_filter_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.fl_filter_p2p_handle.argtypes = [C.c_void_p, C.c_char_p]
    L.fl_filter_p2p_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    _lib = L
    return L


def _check(rc: int) -> int:
    if rc < 0:
        raise FastLioError(f"fastlio_b200 error {rc}: {load().fl_last_error().decode(errors='replace')}")
    return rc


def device_count() -> int:
    return load().fl_device_count()


class KdTree:
    """Device point map with the KD_TREE<PointType> call surface used by laserMapping.cpp."""

    def __init__(self, device: int = 0, downsample: float = 0.5, cell_directory: bool = True, cell_size: float = 0.0):
        self._L = load()
        h = C.c_void_p()
        _check(self._L.fl_map_create(C.byref(h), device, downsample))
        self.h = h
        if not cell_directory or cell_size > 0.0:
            _check(self._L.fl_map_set_cell_directory(self.h, int(cell_directory), cell_size))

    def close(self):
        if getattr(self, "h", None):
            self._L.fl_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # KD_TREE::set_downsample_param
    def set_downsample_param(self, v: float):
        _check(self._L.fl_map_set_downsample(self.h, v))

    # KD_TREE::Build
    def Build(self, pts4):
        pts4 = np.ascontiguousarray(pts4, dtype=np.float32).reshape(-1, 4)
        _check(self._L.fl_map_build(self.h, pts4, len(pts4)))

    def size(self) -> int:
        return _check(self._L.fl_map_size(self.h))

    def validnum(self) -> int:
        return _check(self._L.fl_map_validnum(self.h))

    # KD_TREE::Nearest_Search, batched
    def Nearest_Search(self, q4, k: int = 5):
        q4 = np.ascontiguousarray(q4, dtype=np.float32).reshape(-1, 4)
        nq = len(q4)
        pts = np.zeros((nq, k, 4), dtype=np.float32)
        d2 = np.zeros((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        _check(self._L.fl_map_knn(self.h, q4, nq, k, pts, d2, cnt))
        return pts, d2, cnt

    # KD_TREE::Add_Points
    def Add_Points(self, pts4, downsample_on: bool) -> int:
        pts4 = np.ascontiguousarray(pts4, dtype=np.float32).reshape(-1, 4)
        return _check(self._L.fl_map_add_points(self.h, pts4, len(pts4), int(downsample_on)))

    # KD_TREE::Delete_Point_Boxes
    def Delete_Point_Boxes(self, boxes6) -> int:
        boxes6 = np.ascontiguousarray(boxes6, dtype=np.float32).reshape(-1, 6)
        return _check(self._L.fl_map_delete_boxes(self.h, boxes6, len(boxes6)))

    # KD_TREE::Add_Point_Boxes
    def Add_Point_Boxes(self, boxes6) -> int:
        boxes6 = np.ascontiguousarray(boxes6, dtype=np.float32).reshape(-1, 6)
        return _check(self._L.fl_map_add_boxes(self.h, boxes6, len(boxes6)))

    # KD_TREE::acquire_removed_points
    def acquire_removed_points(self, cap: int = 1 << 20) -> np.ndarray:
        out = np.zeros((max(cap, 1), 4), dtype=np.float32)
        n = _check(self._L.fl_map_acquire_removed(self.h, out, cap))
        return out[:min(n, cap)].copy()

    # KD_TREE::flatten(Root_Node, ..., NOT_RECORD)
    def flatten(self) -> np.ndarray:
        n = self.validnum()
        out = np.zeros((max(n, 1), 4), dtype=np.float32)
        got = _check(self._L.fl_map_flatten(self.h, out, len(out)))
        return out[:got].copy()

    def tree_range(self) -> np.ndarray:
        box = np.zeros(6, dtype=np.float32)
        _check(self._L.fl_map_tree_range(self.h, box))
        return box

    def rebuild(self):
        _check(self._L.fl_map_rebuild(self.h))

    def stats(self) -> dict:
        s = np.zeros(4, dtype=np.int32)
        _check(self._L.fl_map_stats(self.h, s))
        return dict(main_leaves=int(s[0]), overflow_leaves=int(s[1]), levels=int(s[2]), rebuilds=int(s[3]))

    def dir_stats(self) -> dict:
        s = np.zeros(6, dtype=np.int32)
        _check(self._L.fl_map_dir_stats(self.h, s))
        return dict(cells=int(s[0]), ext_buckets=int(s[1]), crowded_cells=int(s[2]), capacity=int(s[3]), relists=int(s[4]), walked=int(s[5]), enabled=bool(s[5] >= 0))


class Esekf:
    """esekf::update_iterated_dyn_share_modified with the fused device measurement model."""

    def __init__(self, tree: KdTree, max_points: int = 100000, max_iter: int = 4, limit: float = 0.001,
                 extrinsic_est_en: bool = False, solver: int = 1, search: int = -1, fused: int = -1):
        self._L = load()
        self.tree = tree
        h = C.c_void_p()
        _check(self._L.fl_filter_create(C.byref(h), tree.h, max_points))
        self.h = h
        self.max_iter = max_iter
        lim = np.full(23, limit, dtype=np.float64)
        _check(self._L.fl_filter_set_params(self.h, max_iter, lim, int(extrinsic_est_en)))
        _check(self._L.fl_filter_set_solver(self.h, solver))
        if search >= 0:
            _check(self._L.fl_filter_set_search(self.h, search))
        if fused >= 0:
            _check(self._L.fl_filter_set_fused(self.h, fused))
        self._fused = (fused != 0) and solver == 1 and os.environ.get('FASTLIO_B200_LEGACY', '0') != '1'

    def close(self):
        if getattr(self, "h", None):
            self._L.fl_filter_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_iterated_dyn_share_modified(self, scan4, x26, P, R: float = 0.001):
        """Returns (x, P, solve_time_s).  Host buffers in, host buffers out."""
        scan4 = np.ascontiguousarray(scan4, dtype=np.float32).reshape(-1, 4)
        x = np.array(x26, dtype=np.float64).copy()
        Pm = np.ascontiguousarray(np.array(P, dtype=np.float64).copy())
        st = C.c_double(0.0)
        _check(self._L.fl_filter_update(self.h, scan4, len(scan4), x, Pm, R, C.byref(st)))
        return x, Pm, st.value

    def map_incremental(self, filter_size_map_min: float = 0.5, flg_EKF_inited: bool = True):
        """laserMapping.cpp:427-474 on the device; returns (|PointToAdd|, |PointNoNeedDownsample|, Add_Points return)."""
        out = np.zeros(3, dtype=np.int32)
        _check(self._L.fl_filter_map_incremental(self.h, filter_size_map_min, int(flg_EKF_inited), out))
        return int(out[0]), int(out[1]), int(out[2])

    def nearest(self, nq: int):
        pts = np.zeros((nq, 5, 4), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        _check(self._L.fl_filter_get_nearest(self.h, pts, cnt, nq))
        return pts, cnt

    def selected(self, nq: int):
        out = np.zeros(nq, dtype=np.uint8)
        _check(self._L.fl_filter_get_selected(self.h, out, nq))
        return out

    def pass_logs(self):
        logs = (PassLog * 16)()
        n = _check(self._L.fl_filter_get_pass_logs(self.h, logs, 16))
        out = []
        for i in range(n):
            l = logs[i]
            out.append(dict(searched=l.searched, valid=l.valid, effct=l.effct, converged=l.converged,
                            res_sum=l.res_sum, HtH=np.array(l.HtH).reshape(12, 12).copy(),
                            Hth=np.array(l.Hth).copy(), x_after=np.array(l.x_after).copy()))
        return out

    # device-resident pieces
    def upload_scan(self, scan4):
        scan4 = np.ascontiguousarray(scan4, dtype=np.float32).reshape(-1, 4)
        _check(self._L.fl_filter_upload_scan(self.h, scan4, len(scan4)))

    def upload_state(self, x26, P, R: float = 0.001):
        _check(self._L.fl_filter_upload_state(self.h, np.ascontiguousarray(x26, dtype=np.float64),
                                              np.ascontiguousarray(P, dtype=np.float64), R))

    def run(self):
        _check(self._L.fl_filter_run(self.h))

    def download_state(self):
        x = np.zeros(26, dtype=np.float64)
        P = np.zeros((23, 23), dtype=np.float64)
        n = C.c_int(0)
        _check(self._L.fl_filter_download_state(self.h, x, P, C.byref(n)))
        return x, P, n.value

    def time_resident(self, reps: int, flush_l2: bool = True) -> float:
        ms = C.c_float(0.0)
        _check(self._L.fl_filter_time_resident(self.h, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def time_search_pass(self, reps: int, flush_l2: bool = True) -> float:
        ms = C.c_float(0.0)
        _check(self._L.fl_filter_time_search_pass(self.h, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def time_e2e(self, scan4, x26, P, R: float, reps: int):
        """Seconds for `reps` native back-to-back fl_filter_update calls with host buffers; returns (seconds, x, P)."""
        scan4 = np.ascontiguousarray(scan4, dtype=np.float32).reshape(-1, 4)
        x_out = np.zeros(26, dtype=np.float64)
        P_out = np.zeros((23, 23), dtype=np.float64)
        sec = C.c_double(0.0)
        _check(self._L.fl_filter_time_e2e(self.h, scan4, len(scan4), np.ascontiguousarray(x26, dtype=np.float64),
                                          np.ascontiguousarray(P, dtype=np.float64), R, reps, C.byref(sec), x_out, P_out))
        return sec.value, x_out, P_out

    def gpu_launches(self) -> int:
        return _check(self._L.fl_filter_gpu_launches(self.h))

    def fused(self) -> bool:
        return self._fused

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        _check(self._L.fl_filter_comm_init(self.h, nranks, rank, unique_id))

    def p2p_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self._L.fl_filter_p2p_handle(self.h, buf))
        return buf.raw

    def p2p_connect(self, nranks: int, rank: int, handles: bytes):
        _check(self._L.fl_filter_p2p_connect(self.h, nranks, rank, handles))

    def set_shard(self, q_begin: int, q_end: int):
        _check(self._L.fl_filter_set_shard(self.h, q_begin, q_end))


class Scan:
    """feats_undistort / feats_down_body kept in HBM: UndistortPcl's backward pass (IMU_Processing.hpp:232-346) and the
    pcl::VoxelGrid down-sampling (laserMapping.cpp:904-905) in front of the update."""

    def __init__(self, tree: KdTree):
        self._L = load()
        self.tree = tree
        h = C.c_void_p()
        _check(self._L.fl_scan_create(C.byref(h), tree.h))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._L.fl_scan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, xyzi, offset_ms):
        xyzi = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
        offset_ms = np.ascontiguousarray(offset_ms, dtype=np.float32).reshape(-1)
        if len(offset_ms) != len(xyzi):
            raise ValueError("one offset time per point")
        _check(self._L.fl_scan_upload(self.h, xyzi, offset_ms, len(xyzi)))
        self.n = len(xyzi)

    def undistort(self, imu_pose22, x26_end):
        poses = np.ascontiguousarray(imu_pose22, dtype=np.float64).reshape(-1, 22)
        _check(self._L.fl_scan_undistort(self.h, poses, len(poses), np.ascontiguousarray(x26_end, dtype=np.float64)))

    def voxel_downsample(self, leaf: float) -> int:
        return _check(self._L.fl_scan_voxel_downsample(self.h, leaf))

    def download(self, which: int = 1) -> np.ndarray:
        n = _check(self._L.fl_scan_download(self.h, which, np.zeros((1, 4), dtype=np.float32), 0))
        out = np.zeros((max(n, 1), 4), dtype=np.float32)
        _check(self._L.fl_scan_download(self.h, which, out, n))
        return out[:n].copy()

    def update(self, filt: "Esekf", x26, P, R: float = 0.001):
        """fl_filter_update on the down-sampled cloud without a host hop; returns (x, P, solve_time_s)."""
        x = np.array(x26, dtype=np.float64).copy()
        Pm = np.ascontiguousarray(np.array(P, dtype=np.float64).copy())
        st = C.c_double(0.0)
        _check(self._L.fl_filter_update_scan(filt.h, self.h, x, Pm, R, C.byref(st)))
        return x, Pm, st.value


class LocalMap:
    """lasermap_fov_segment() (laserMapping.cpp:229-277): the sliding cube that issues the delete boxes."""

    def __init__(self, cube_len: float, det_range: float):
        self._L = load()
        h = C.c_void_p()
        _check(self._L.fl_localmap_create(C.byref(h), cube_len, det_range))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._L.fl_localmap_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def segment(self, pos_lid, tree: "KdTree | None" = None):
        """Returns (cub_needrm as an (nb, 6) array, kdtree_delete_counter)."""
        boxes = np.zeros((3, 6), dtype=np.float32)
        nd = C.c_int(0)
        nb = _check(self._L.fl_localmap_segment(self.h, tree.h if tree is not None else None,
                                                np.ascontiguousarray(pos_lid, dtype=np.float64), boxes, C.byref(nd)))
        return boxes[:nb].copy(), nd.value

    def box(self) -> np.ndarray:
        b = np.zeros(6, dtype=np.float32)
        _check(self._L.fl_localmap_get(self.h, b))
        return b


def host_register(arr: np.ndarray):
    """Page-lock a numpy array that will be passed repeatedly (the reference's scan buffer)."""
    _check(load().fl_host_register(arr.ctypes.data, arr.nbytes))


def host_unregister(arr: np.ndarray):
    _check(load().fl_host_unregister(arr.ctypes.data))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load().fl_comm_unique_id(buf))
    return buf.raw


def shard_range(n: int, nranks: int, rank: int):
    """Contiguous shard of n scan points for `rank`."""
    base, rem = divmod(n, nranks)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
