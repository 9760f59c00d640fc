"""TEST INFRASTRUCTURE ONLY -- numpy-facing wrappers of the C restatement (oracle/csrc/stag_ref.c)."""
import ctypes

import numpy as np

from . import cbuild

KIND = {"none": 0, "external": 1, "normal": 2, "uniform": 3, "bernoulli": 4}


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return cbuild.load().ref_num_threads()


def use_all_cores():
    """OpenMP threads = every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cbuild.load().ref_set_num_threads(int(n))
    return num_threads()


def csx_build(src, dst, num_nodes, by_dst=True):
    lib = cbuild.load()
    src = np.ascontiguousarray(src, dtype=np.int64)
    dst = np.ascontiguousarray(dst, dtype=np.int64)
    E = src.shape[0]
    indptr = np.empty(num_nodes + 1, np.int32)
    indices = np.empty(max(E, 1), np.int32)
    eid = np.empty(max(E, 1), np.int32)
    lib.ref_csx_build(_p(src), _p(dst), ctypes.c_int64(E), ctypes.c_int64(num_nodes), int(by_dst), _p(indptr),
                      _p(indices), _p(eid))
    return indptr, indices[:E], eid[:E]


def pshape_of(p, E, K):
    p = np.asarray(p)
    if p.size == 1:
        return 0
    if p.size == K and K != 1:
        return 1
    if p.size == E:
        return 2
    assert p.size == E * K, (p.shape, E, K)
    return 3


def noise(kind, E, K, sample, seed, offset, p0, p1=None, relu=False, return_raw=False):
    lib = cbuild.load()
    p0 = _f32(np.atleast_1d(p0))
    p1 = None if p1 is None else _f32(np.atleast_1d(p1))
    w = np.empty((E, K), np.float32)
    raw = np.empty((E, K), np.float32) if return_raw else None
    lib.ref_noise(KIND[kind], ctypes.c_int64(E), K, int(sample), ctypes.c_uint64(seed), ctypes.c_uint64(offset),
                  _p(p0), _p(p1), pshape_of(p0, E, K), int(relu), _p(w), _p(raw))
    return (w, raw) if return_raw else w


def in_norm(csc, num_nodes, w):
    lib = cbuild.load()
    w = np.array(w, dtype=np.float32, order="C", copy=True)
    lib.ref_in_norm(_p(csc[0]), _p(csc[2]), ctypes.c_int64(num_nodes), w.shape[1], _p(w))
    return w


def aggregate(csx, num_rows, x, w=None, gather_scale=None, row_scale=None):
    """csx = (indptr, indices, eid).  CSC + (src_scale, dst_scale) = forward; CSR with
    x := dout and (dst_scale, src_scale) = dX."""
    lib = cbuild.load()
    x = _f32(x)
    D = x.shape[1]
    w = _f32(w)
    K = 0 if w is None else (1 if w.ndim == 1 else w.shape[1])
    out = np.empty((num_rows, D), np.float32)
    lib.ref_aggregate(_p(csx[0]), _p(csx[1]), _p(csx[2]), ctypes.c_int64(num_rows), _p(x), D, _p(w), K,
                      _p(_f32(gather_scale)), _p(_f32(row_scale)), _p(out))
    return out


def sddmm(src, dst, x, dout, K, src_scale=None, dst_scale=None):
    lib = cbuild.load()
    src = np.ascontiguousarray(src, dtype=np.int64)
    dst = np.ascontiguousarray(dst, dtype=np.int64)
    x, dout = _f32(x), _f32(dout)
    E, D = src.shape[0], x.shape[1]
    dw = np.empty((E, K), np.float32)
    lib.ref_sddmm(_p(src), _p(dst), ctypes.c_int64(E), _p(x), _p(dout), D, K, _p(_f32(src_scale)),
                  _p(_f32(dst_scale)), _p(dw))
    return dw


class LayerPass:
    """Pre-allocated un-fused layer forward+backward (noise materialised as the reference does);
    `run(sample)` is what bench.py times as the CPU baseline."""

    def __init__(self, src, dst, num_nodes, x, dout, kind="normal", p0=1.0, p1=0.4, vi=False, gcn_norm=True):
        self.lib = cbuild.load()
        self.src = np.ascontiguousarray(src, dtype=np.int64)
        self.dst = np.ascontiguousarray(dst, dtype=np.int64)
        self.N, self.E = int(num_nodes), self.src.shape[0]
        self.csc = csx_build(self.src, self.dst, self.N, True)
        self.csr = csx_build(self.src, self.dst, self.N, False)
        self.x, self.dout = _f32(x), _f32(dout)
        self.D = self.x.shape[1]
        self.kind, self.vi = KIND[kind], int(vi)
        self.p0 = _f32(np.atleast_1d(p0))
        self.p1 = None if p1 is None else _f32(np.atleast_1d(p1))
        self.pshape = pshape_of(self.p0, self.E, self.D)
        self.ss = self.ds = None
        if gcn_norm:
            self.ss = (np.maximum(np.diff(self.csr[0]), 1).astype(np.float32)) ** -0.5
            self.ds = (np.maximum(np.diff(self.csc[0]), 1).astype(np.float32)) ** -0.5
        self.w = np.empty((self.E, self.D), np.float32)
        self.raw = np.empty((self.E, self.D), np.float32) if vi else None
        self.dw = np.empty((self.E, self.D), np.float32) if vi else None
        self.out = np.empty((self.N, self.D), np.float32)
        self.dx = np.empty((self.N, self.D), np.float32)
        self.dp = (ctypes.c_double * 2)()

    def run(self, sample=0, seed=42, offset=0):
        self.lib.ref_layer_fwd_bwd(
            _p(self.src), _p(self.dst), ctypes.c_int64(self.E), ctypes.c_int64(self.N),
            _p(self.csc[0]), _p(self.csc[1]), _p(self.csc[2]), _p(self.csr[0]), _p(self.csr[1]), _p(self.csr[2]),
            _p(self.x), _p(self.dout), self.D, self.kind, int(sample), ctypes.c_uint64(seed), ctypes.c_uint64(offset),
            _p(self.p0), _p(self.p1), self.pshape, _p(self.ss), _p(self.ds), self.vi,
            _p(self.w), _p(self.raw), _p(self.dw), _p(self.out), _p(self.dx),
            ctypes.byref(self.dp, 0), ctypes.byref(self.dp, 8))
        return self.out, self.dx
