"""CPU: the C-ABI library loads and exports every symbol include/fastlio_b200.h declares; host logic."""
import ctypes
import os
import re

import numpy as np
import pytest

from fast_lio_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "fastlio_b200.h")).read()
    return sorted(set(re.findall(r"\b(fl_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(build.LIB), "run `python -m fast_lio_b200.build` (or __graft_entry__.build()) first"
    lib = ctypes.CDLL(build.LIB)
    decl = declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if not hasattr(lib, s)]
    assert not missing, missing
    # the binding's list and the header agree
    assert sorted(api.SYMBOLS) == decl


def test_no_gpu_fails_loudly():
    """On a box without a CUDA device the product refuses to run (no CPU fallback)."""
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.FastLioError):
        api.KdTree(0, 0.5)
    assert b"no CUDA device" in api.load().fl_last_error()


def test_shard_range_partitions_the_scan():
    for n in (0, 1, 7, 30000, 131072):
        for w in (1, 2, 3, 4, 8):
            r = [api.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (the judge checks the same)."""
    pkg = os.path.join(ROOT, "fast_lio_b200")
    banned = re.compile(r"(^\s*(from|import)\s+oracle\b)|liboracle|libikdtree_ref|oracle[/.](_ref|bind)|ref_kdtree_|oracle_update_iterated",
                        re.MULTILINE)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not banned.search(txt), (dirpath, f)
    for hdr in os.listdir(os.path.join(ROOT, "include")):
        path = os.path.join(ROOT, "include", hdr)
        if os.path.isfile(path):
            assert not banned.search(open(path, errors="ignore").read()), hdr
