"""TEST INFRASTRUCTURE ONLY -- index oracle (bit-exact target of ``stag_csx_build``).

Restates the COO -> CSC / CSR conversion the reference inherits from DGL when
``graph.update_all`` (stag/zoo/gcn.py:95, stag/layers.py:12-15) first needs the
in-edge (CSC) or out-edge (CSR) adjacency: edges are grouped by destination
(source) and, inside a group, keep increasing edge-id order (a stable sort).  DGL
is third-party and absent from /root/reference, so this is the published behaviour
of ``g.adj_tensors('csc')`` / ``('csr')``; SURVEY.md 8(c) "Index oracle".

Parity pin: tests/test_oracle_cpu.py checks this against ``torch.sort(stable=True)``
and against the order in which the dgl shim's ``index_add`` reduction visits edges.
"""
import numpy as np


def csx_build(src, dst, num_nodes, by_dst=True):
    """Return (indptr[N+1] int32, indices[E] int32, eid[E] int32).

    by_dst=True  -> CSC: rows are destinations, ``indices`` holds sources.
    by_dst=False -> CSR: rows are sources, ``indices`` holds destinations.
    ``eid[j]`` is the original COO position of the j-th stored edge.
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    key, other = (dst, src) if by_dst else (src, dst)
    perm = np.argsort(key, kind="stable")
    counts = np.bincount(key, minlength=num_nodes) if key.size else np.zeros(num_nodes, np.int64)
    indptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr.astype(np.int32), other[perm].astype(np.int32), perm.astype(np.int32)


def degrees(src, dst, num_nodes):
    """(in_degrees, out_degrees) as int64 -- stag/zoo/gcn.py:68,101; stag/layers.py:21."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    return (np.bincount(dst, minlength=num_nodes).astype(np.int64),
            np.bincount(src, minlength=num_nodes).astype(np.int64))


def hub_segments(indptr, hub_threshold, seg_len):
    """Host restatement of the hub-row split the CUDA library performs at build time
    (stag_b200/csrc/csx_build.cu): rows with more than ``hub_threshold`` stored edges
    are cut into consecutive segments of at most ``seg_len`` edges.  Returns
    (hub_rows, seg_ptr) with hub rows in increasing row order."""
    indptr = np.asarray(indptr, dtype=np.int64)
    deg = indptr[1:] - indptr[:-1]
    hub_rows = np.nonzero(deg > hub_threshold)[0]
    nseg = (deg[hub_rows] + seg_len - 1) // seg_len
    seg_ptr = np.zeros(len(hub_rows) + 1, dtype=np.int64)
    np.cumsum(nseg, out=seg_ptr[1:])
    return hub_rows.astype(np.int32), seg_ptr.astype(np.int32)
