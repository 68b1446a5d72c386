"""An INDEPENDENT restatement of the octomap file formats (octomap 1.9: OcTree::writeBinaryNode / OcTreeBaseImpl::
writeNodesRecurs, keyToCoord, calcMinMax, the leaf iterator's depth-first child order), in plain Python, written from the
published library source -- NOT from amcl3d_b200/host/compat/octomap/OcTree.h, which it is used to cross-check
(tests/test_octomap_independent.py).  octomap itself is a third-party dependency of the reference (package.xml) that is
absent from /root/reference and from this image, and the reference's release maps are not available offline.

Tree model: TREE_DEPTH 16, key origin 32768.  A leaf is (kx, ky, kz, depth, occupied): kx.. are the leaf's `depth`-bit
path prefixes per axis (for depth 16: voxel index + 32768)."""
import struct

import numpy as np

TREE_DEPTH = 16
TREE_MAX_VAL = 32768
OCC_LOGODDS, FREE_LOGODDS = np.float32(3.5), np.float32(-2.0)     # clamping thresholds of a default OcTree


class Node:
    __slots__ = ("children", "leaf")

    def __init__(self):
        self.children = [None] * 8
        self.leaf = None            # None (inner / empty) or True (occupied) / False (free)

    def has_children(self):
        return any(c is not None for c in self.children)


def build_tree(leaves):
    """leaves: iterable of (px, py, pz, depth, occupied) with depth-bit path prefixes.  Later leaves must not lie below
    earlier ones (the caller supplies a consistent, pruned tree)."""
    root = Node()
    for px, py, pz, depth, occ in leaves:
        node = root
        for level in range(depth):
            shift = depth - 1 - level
            idx = ((px >> shift) & 1) | (((py >> shift) & 1) << 1) | (((pz >> shift) & 1) << 2)
            if node.leaf is not None:
                raise ValueError("leaf below a leaf")
            if node.children[idx] is None:
                node.children[idx] = Node()
            node = node.children[idx]
        if node.has_children():
            raise ValueError("leaf above existing nodes")
        node.leaf = bool(occ)
    return root


def count_nodes(node):
    return 1 + sum(count_nodes(c) for c in node.children if c is not None)


def _inner_value(node):
    """updateInnerOccupancy: an inner node holds the maximum of its children's log-odds."""
    if node.leaf is not None:
        return OCC_LOGODDS if node.leaf else FREE_LOGODDS
    return max(_inner_value(c) for c in node.children if c is not None)


def write_bt(path, root, res):
    """OcTree::writeBinary: text header, then per inner node two bytes -- two bits per child: 00 unknown, 01 occupied leaf
    (bit 2i = 0, bit 2i+1 = 1), 10 free leaf (bit 2i = 1), 11 inner node -- followed by the inner children in order."""
    out = bytearray()

    def rec(node):
        bits = 0
        for i, c in enumerate(node.children):
            if c is None:
                continue
            if c.has_children():
                bits |= 3 << (2 * i)
            elif c.leaf:
                bits |= 2 << (2 * i)       # bit 2i = 0, bit 2i + 1 = 1
            else:
                bits |= 1 << (2 * i)       # bit 2i = 1, bit 2i + 1 = 0
        out.append(bits & 0xFF)            # children 0..3
        out.append((bits >> 8) & 0xFF)     # children 4..7
        for c in node.children:
            if c is not None and c.has_children():
                rec(c)

    rec(root)
    with open(path, "wb") as f:
        f.write(b"# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n")
        f.write(("id OcTree\nsize %d\nres %s\ndata\n" % (count_nodes(root), repr(float(res)))).encode())
        f.write(bytes(out))


def write_ot(path, root, res):
    """AbstractOcTree::write: text header, then per node its float log-odds and one byte with a bit per existing child,
    followed by the children in order."""
    out = bytearray()

    def rec(node):
        out.extend(struct.pack("<f", float(_inner_value(node))))
        mask = 0
        for i, c in enumerate(node.children):
            if c is not None:
                mask |= 1 << i
        out.append(mask)
        for c in node.children:
            if c is not None:
                rec(c)

    rec(root)
    with open(path, "wb") as f:
        f.write(b"# Octomap OcTree file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n")
        f.write(("id OcTree\nsize %d\nres %s\ndata\n" % (count_nodes(root), repr(float(res)))).encode())
        f.write(bytes(out))


def key_to_coord(key, depth, res):
    """OcTreeBaseImpl::keyToCoord(key, depth), in double like the library."""
    if depth == 0:
        return 0.0
    if depth == TREE_DEPTH:
        return (float(int(key) - TREE_MAX_VAL) + 0.5) * res
    return (np.floor((float(key) - float(TREE_MAX_VAL)) / float(1 << (TREE_DEPTH - depth))) + 0.5) * (res * float(1 << (TREE_DEPTH - depth)))


def iterate_leaves(root, res):
    """The leaf iterator: depth first, children 0..7; keys by computeChildKey.  Yields (x, y, z, size, occupied) doubles."""
    stack = [(root, (TREE_MAX_VAL,) * 3, 0)]
    while stack:
        node, key, depth = stack.pop()
        if node.leaf is not None or not node.has_children():
            if node.leaf is not None:
                yield tuple(key_to_coord(k, depth, res) for k in key) + (res * float(1 << (TREE_DEPTH - depth)), node.leaf)
            continue
        off = TREE_MAX_VAL >> (depth + 1)
        for i in range(7, -1, -1):
            c = node.children[i]
            if c is None:
                continue
            ck = tuple(k + off if (i >> a) & 1 else k - off - (0 if off else 1) for a, k in enumerate(key))
            stack.append((c, ck, depth + 1))


def expected_point_cloud(root, res):
    """What PointCloudTools.cpp:51-82 must return for this tree: occupied leaf centres narrowed to float, in iterator
    order, and the metric bounds of calcMinMax over ALL leaves ((centre - size/2), then (+ size), in double)."""
    pts = []
    lo, hi = [np.inf] * 3, [-np.inf] * 3
    for x, y, z, size, occ in iterate_leaves(root, res):
        half = size / 2.0
        for a, c in enumerate((x, y, z)):
            v = c - half
            lo[a] = min(lo[a], v)
            v += size
            hi[a] = max(hi[a], v)
        if occ:
            pts.append((np.float32(x), np.float32(y), np.float32(z)))
    return np.array(pts, np.float32).reshape(-1, 3), np.array(lo + hi + [res], np.float64)


def read_file(path):
    """Parses a .bt / .ot file into (root, res): the inverse of write_bt / write_ot (OcTree::readBinaryNode /
    OcTreeBaseImpl::readNodesRecurs).  A node is occupied when its log-odds reach the 0.5-probability threshold (>= 0)."""
    data = open(path, "rb").read()
    binary = data.startswith(b"# Octomap OcTree binary file")
    if not binary and not data.startswith(b"# Octomap OcTree file"):
        raise ValueError("not an octomap file")
    pos, res = 0, None
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if line.startswith("res"):
            res = float(line.split()[1])
        if line == "data":
            break
    root = Node()
    if binary:
        def rec(node):
            nonlocal pos
            bits = data[pos] | (data[pos + 1] << 8)
            pos += 2
            inner = []
            for i in range(8):
                b0, b1 = (bits >> (2 * i)) & 1, (bits >> (2 * i + 1)) & 1
                if not b0 and not b1:
                    continue
                node.children[i] = Node()
                if b0 and b1:
                    inner.append(node.children[i])
                else:
                    node.children[i].leaf = bool(b1 and not b0)
            for c in inner:
                rec(c)
        rec(root)
    else:
        def rec(node):
            nonlocal pos
            (value,) = struct.unpack_from("<f", data, pos)
            mask = data[pos + 4]
            pos += 5
            if mask == 0:
                node.leaf = bool(value >= 0.0)
            for i in range(8):
                if (mask >> i) & 1:
                    node.children[i] = Node()
                    rec(node.children[i])
        rec(root)
    return root, res
