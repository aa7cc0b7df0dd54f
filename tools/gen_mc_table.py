"""Generate the marching-cubes case table used by csrc/mesh.cu and oracle/mesh_mc.py.

The reference extracts its meshes with skimage.measure.marching_cubes (utils.py:231-288), which is not installable
offline, and the classic 256 x 16 triangle table is not reproduced from memory here.  Instead the table is DERIVED: for
every sign configuration of the 8 cube corners the isosurface crosses exactly the edges whose end points differ in sign;
on each of the 6 faces the crossing points are joined by segments (2 crossings: one segment; 4 crossings -- the ambiguous
face with diagonally opposite inside corners -- two segments that cut off the INSIDE corners, a rule that depends only on
the face's own corner signs, so neighbouring cubes always agree and the surface has no holes); the segments close into
loops of 3..7 edge points, and every loop is fan-triangulated with the inside (value < level) on the back side.

Output: eprecon_b200/csrc/mc_table.cuh (int8 kMcTris[256][16]: edge ids, -1 terminated; uint8 kMcCount[256]) and
oracle/mc_table.py (the same table as Python lists for the numpy oracle).

Conventions.  Corner c = (cx, cy, cz) bits: c = cx | cy << 1 | cz << 2.  Edge e in 0..11: e = axis * 4 + j, the edge along
`axis` whose other two coordinates (in cyclic order: axis+1, axis+2) are (j & 1, j >> 1); it starts at the corner with that
axis coordinate 0.  Cube index bit c is set when corner c is INSIDE (value < level).
"""
import itertools
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def corner_xyz(c):
    return (c & 1, (c >> 1) & 1, (c >> 2) & 1)


def edge_corners(e):
    axis, j = divmod(e, 4)
    a1, a2 = (axis + 1) % 3, (axis + 2) % 3
    p = [0, 0, 0]
    p[a1], p[a2] = j & 1, j >> 1
    q = list(p)
    q[axis] = 1
    idx = lambda v: v[0] | v[1] << 1 | v[2] << 2  # noqa: E731
    return idx(p), idx(q)


EDGE_BY_CORNERS = {frozenset(edge_corners(e)): e for e in range(12)}


def faces():
    """6 faces as cyclic lists of 4 corners (any orientation; orientation is fixed per loop later)."""
    out = []
    for axis in range(3):
        a1, a2 = (axis + 1) % 3, (axis + 2) % 3
        for side in (0, 1):
            cyc = []
            for (u, v) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                p = [0, 0, 0]
                p[axis], p[a1], p[a2] = side, u, v
                cyc.append(p[0] | p[1] << 1 | p[2] << 2)
            out.append(cyc)
    return out


FACES = faces()


def edge_mid(e):
    a, b = edge_corners(e)
    pa, pb = corner_xyz(a), corner_xyz(b)
    return tuple((x + y) / 2.0 for x, y in zip(pa, pb))


def case_triangles(index):
    inside = [(index >> c) & 1 for c in range(8)]
    adj = {}   # crossing edge -> the (two) crossing edges it is joined to

    def link(e0, e1):
        adj.setdefault(e0, []).append(e1)
        adj.setdefault(e1, []).append(e0)
    for cyc in FACES:
        s = [inside[c] for c in cyc]
        cross = [EDGE_BY_CORNERS[frozenset((cyc[i], cyc[(i + 1) % 4]))] for i in range(4) if s[i] != s[(i + 1) % 4]]
        if len(cross) == 2:
            link(cross[0], cross[1])
        elif len(cross) == 4:
            # ambiguous face: cut off each INSIDE corner (join the two face edges that meet at it)
            for i in range(4):
                if s[i]:
                    e_prev = EDGE_BY_CORNERS[frozenset((cyc[(i - 1) % 4], cyc[i]))]
                    e_next = EDGE_BY_CORNERS[frozenset((cyc[i], cyc[(i + 1) % 4]))]
                    link(e_prev, e_next)
    assert all(len(v) == 2 for v in adj.values()), (index, adj)
    tris, seen = [], set()
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            a, b = adj[cur]
            nxt = a if (prev is None or a != prev) else b
            if nxt == start:
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        assert 3 <= len(loop) <= 7, (index, loop)
        # orientation: the loop's polygon normal (Newell) must point away from the inside corners it separates
        pts = [edge_mid(e) for e in loop]
        n = [0.0, 0.0, 0.0]
        for i in range(len(pts)):
            a, b = pts[i], pts[(i + 1) % len(pts)]
            n[0] += (a[1] - b[1]) * (a[2] + b[2])
            n[1] += (a[2] - b[2]) * (a[0] + b[0])
            n[2] += (a[0] - b[0]) * (a[1] + b[1])
        cen = [sum(p[k] for p in pts) / len(pts) for k in range(3)]
        # an inside corner adjacent to the loop: one end of the loop's first edge
        ca, cb = edge_corners(loop[0])
        cin = ca if inside[ca] else cb
        d = [corner_xyz(cin)[k] - cen[k] for k in range(3)]
        if sum(n[k] * d[k] for k in range(3)) > 0:     # normal points towards the inside corner -> flip
            loop = loop[::-1]
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


def build():
    table = [case_triangles(i) for i in range(256)]
    assert max(len(t) for t in table) <= 5
    return table


def main():
    table = build()
    rows = []
    for t in table:
        flat = list(itertools.chain.from_iterable(t))
        rows.append(flat + [-1] * (16 - len(flat)))
    counts = [len(t) for t in table]
    cuh = ["// GENERATED by tools/gen_mc_table.py -- do not edit.  Marching-cubes case table (derivation: see the generator).",
           "#pragma once", "#include <stdint.h>",
           "__device__ __constant__ int8_t kMcTris[256][16] = {"]
    cuh += ["  {" + ", ".join(f"{v:2d}" for v in r) + "}," for r in rows]
    cuh += ["};", "__device__ __constant__ uint8_t kMcCount[256] = {" + ", ".join(str(c) for c in counts) + "};", ""]
    with open(os.path.join(ROOT, "eprecon_b200", "csrc", "mc_table.cuh"), "w") as f:
        f.write("\n".join(cuh))
    with open(os.path.join(ROOT, "oracle", "mc_table.py"), "w") as f:
        f.write('"""GENERATED by tools/gen_mc_table.py -- marching-cubes case table for oracle/mesh_mc.py (test infrastructure)."""\n')
        f.write("MC_TRIS = " + repr(rows) + "\n")
        f.write("MC_COUNT = " + repr(counts) + "\n")
    print("cases with triangles:", sum(1 for c in counts if c), "max triangles", max(counts), "total", sum(counts))


if __name__ == "__main__":
    main()
