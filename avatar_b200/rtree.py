"""RTree model files (file formats): the reference's binary '.srtr' format and its legacy
text format (RTree::loadFile / exportFile, RTree.cpp:2967-3094), leafBestMatch (updateBestMatchTable, RTree.cpp:3452-3463)
and the '.partmap' side file (RTree::readPartMap, RTree.cpp:3465-3510).  Host-side parsing only; the arrays go to
Fitter.set_rtree (avb_fitter_set_rtree)."""
import struct

import numpy as np


def _best_match(leaf_data):
    """first strict maximum per leaf (RTree.cpp:3452-3463)"""
    best = np.zeros(len(leaf_data), np.uint8)
    for i, row in enumerate(leaf_data):
        b, arg = np.finfo(np.float32).min, 0
        for j, v in enumerate(row):
            if v > b:
                b, arg = v, j
        best[i] = arg
    return best


def load_rtree(path):
    """-> dict(u, v, thresh, lnode, rnode, leafid, leaf_best, leaf_data, num_parts) as RTree::loadFile fills RTree::nodes"""
    with open(path, "rb") as fh:
        raw = fh.read()
    if raw[:1] == b"R":                                     # binary format (RTree.cpp:2971-3015)
        n_nodes, n_leafs, num_parts = struct.unpack_from("<IIi", raw, 1)
        pos = 13
        u = np.zeros((n_nodes, 2), np.float32); v = np.zeros((n_nodes, 2), np.float32)
        thresh = np.zeros(n_nodes, np.float32)
        lnode = np.full(n_nodes, -1, np.int32); rnode = np.full(n_nodes, -1, np.int32); leafid = np.full(n_nodes, -1, np.int32)
        leaf_data = np.zeros((n_leafs, num_parts), np.float32)
        last = 0
        for i in range(n_nodes):
            is_leaf = raw[pos]; pos += 1
            if is_leaf:
                cnt = raw[pos]; pos += 1
                if cnt > num_parts:
                    raise ValueError(f"leaf has {cnt} parts, expected {num_parts} at most")
                for _ in range(cnt):
                    k = raw[pos]; pos += 1
                    if k > num_parts:
                        raise ValueError(f"leaf index {k} is out of bounds, at most {num_parts}")
                    leaf_data[last, k] = struct.unpack_from("<f", raw, pos)[0]; pos += 4
                leafid[i] = last
                last += 1
            else:
                lnode[i], rnode[i] = struct.unpack_from("<ii", raw, pos); pos += 8
                thresh[i] = struct.unpack_from("<f", raw, pos)[0]; pos += 4
                u[i] = struct.unpack_from("<ff", raw, pos); pos += 8
                v[i] = struct.unpack_from("<ff", raw, pos); pos += 8
        if raw[pos:pos + 1] != b"T":
            raise ValueError("incorrect RTree format, T end marker missing")
    else:                                                   # legacy text format (RTree.cpp:3017-3047)
        tok = raw.decode().split()
        it = iter(tok)
        n_nodes, n_leafs, num_parts = int(next(it)), int(next(it)), int(next(it))
        u = np.zeros((n_nodes, 2), np.float32); v = np.zeros((n_nodes, 2), np.float32)
        thresh = np.zeros(n_nodes, np.float32)
        lnode = np.full(n_nodes, -1, np.int32); rnode = np.full(n_nodes, -1, np.int32); leafid = np.zeros(n_nodes, np.int32)
        for i in range(n_nodes):
            leafid[i] = int(next(it))
            if leafid[i] < 0:
                lnode[i], rnode[i] = int(next(it)), int(next(it))
                thresh[i] = np.float32(next(it))
                u[i] = (np.float32(next(it)), np.float32(next(it)))
                v[i] = (np.float32(next(it)), np.float32(next(it)))
        leaf_data = np.array([[np.float32(next(it)) for _ in range(num_parts)] for _ in range(n_leafs)], np.float32).reshape(n_leafs, num_parts)
    return dict(u=u, v=v, thresh=thresh, lnode=lnode, rnode=rnode, leafid=leafid, leaf_best=_best_match(leaf_data),
                leaf_data=leaf_data, num_parts=int(num_parts))


def save_rtree(path, tree, leaf_data, legacy_text=False):
    """RTree::exportFile (RTree.cpp:3063-3094), or the legacy text layout the loader still accepts"""
    leaf_data = np.asarray(leaf_data, np.float32)
    n_nodes, (n_leafs, num_parts) = len(tree["thresh"]), leaf_data.shape
    if legacy_text:
        with open(path, "w") as fh:
            fh.write(f"{n_nodes} {n_leafs} {num_parts}\n")
            for i in range(n_nodes):
                fh.write(f" {int(tree['leafid'][i])}")
                if tree["leafid"][i] < 0:
                    fh.write("  %d %d %.9g %.9g %.9g %.9g %.9g" % (tree["lnode"][i], tree["rnode"][i], tree["thresh"][i],
                                                               tree["u"][i][0], tree["u"][i][1], tree["v"][i][0], tree["v"][i][1]))
                fh.write("\n")
            for row in leaf_data:
                fh.write(" ".join("%.9g" % x for x in row) + "\n")
        return
    out = bytearray(b"R") + struct.pack("<IIi", n_nodes, n_leafs, num_parts)
    for i in range(n_nodes):
        if tree["leafid"][i] < 0:
            out += bytes([0]) + struct.pack("<iif", int(tree["lnode"][i]), int(tree["rnode"][i]), float(tree["thresh"][i]))
            out += struct.pack("<ffff", *[float(x) for x in (*tree["u"][i], *tree["v"][i])])
        else:
            row = leaf_data[tree["leafid"][i]]
            nz = [j for j in range(num_parts) if row[j] != 0.0]
            out += bytes([255, len(nz)])
            for j in nz:
                out += bytes([j]) + struct.pack("<f", float(row[j]))
    out += b"T"
    with open(path, "wb") as fh:
        fh.write(bytes(out))


def read_partmap(path):
    """RTree::readPartMap (RTree.cpp:3465-3510) -> (part_map list over the old parts, number of new parts, type) with
    type 1 = 'disjoint', 0 = 'contiguous'; None if the file does not start with the expected markers"""
    with open(path) as fh:
        tok = fh.read().split()
    it = iter(tok)
    try:
        if next(it) != "partmap":
            return None
        kind = next(it)
        if kind not in ("disjoint", "contiguous"):
            return None
        if next(it) != "src":
            return None
        n_old = int(next(it))
        old = {next(it): i for i in range(n_old)}
        if next(it) != "dest":
            return None
        n_new = int(next(it))
        new = {next(it): i for i in range(n_new)}
        result = [0] * n_old
        for _ in range(n_old):
            try:
                a, b = next(it), next(it)
            except StopIteration:
                break
            result[old[a]] = new[b]
        return result, n_new, 1 if kind == "disjoint" else 0
    except StopIteration:
        return None
