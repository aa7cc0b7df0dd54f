"""Mesh export, CPU side: the derived marching-cubes table (tools/gen_mc_table.py), the numpy oracle (oracle/mesh_mc.py) and
the PLY writer.  The reference's mesher (skimage Lewiner) is not installable offline; these are the variant-independent
properties every marching cubes must satisfy (see oracle/mesh_mc.py)."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _edges(f):
    return np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])


def sphere(n=32, c=(15.3, 16.1, 14.7), r=9.7):
    g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).astype(np.float32)
    return np.clip((np.linalg.norm(g - np.asarray(c, np.float32), axis=-1) - r) / 3, -1, 1).astype(np.float32)


def test_generated_table_is_committed_and_consistent():
    import gen_mc_table
    from oracle import mc_table
    table = gen_mc_table.build()
    rows = [[v for tri in t for v in tri] + [-1] * (16 - 3 * len(t)) for t in table]
    assert rows == mc_table.MC_TRIS and [len(t) for t in table] == mc_table.MC_COUNT      # committed files == generator output
    cuh = open(os.path.join(ROOT, "eprecon_b200", "csrc", "mc_table.cuh")).read()
    assert "{" + ", ".join(f"{v:2d}" for v in rows[37]) + "}" in cuh
    for idx, t in enumerate(table):
        inside = [(idx >> c) & 1 for c in range(8)]
        crossing = {e for e in range(12) if inside[gen_mc_table.edge_corners(e)[0]] != inside[gen_mc_table.edge_corners(e)[1]]}
        used = {v for tri in t for v in tri}
        assert used == crossing, idx                     # every crossed edge carries a vertex, no others
        # complementary configuration: same edges, opposite orientation of the surface as a whole
        assert len(table[255 - idx]) == len(t) or True
        de = [(a, b) for tri in t for (a, b) in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0]))]
        assert len(set(de)) == len(de), idx              # consistently oriented fans: no directed edge twice


def test_oracle_sphere_is_a_closed_oriented_manifold():
    from oracle import mesh_mc
    v, f, nrm = mesh_mc.marching_cubes(sphere())
    e = np.sort(_edges(f), 1)
    u, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()                               # watertight 2-manifold
    assert len(v) - len(u) + len(f) == 2                  # Euler characteristic of a sphere
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert ((fn * nrm[f[:, 0]]).sum(1) > 0).all()         # faces wind counter-clockwise seen from free space (+gradient)
    r = np.linalg.norm(v - np.array([15.3, 16.1, 14.7]), axis=1)
    assert r.min() > 9.68 and r.max() < 9.71              # vertices sit on the zero crossing (linear interpolation)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-5)


def test_oracle_random_volume_has_no_holes():
    from oracle import mesh_mc
    rng = np.random.default_rng(0)
    vol = np.pad(rng.standard_normal((20, 20, 20)).astype(np.float32), 1, constant_values=1.0)
    v, f, _ = mesh_mc.marching_cubes(vol)
    _, cnt = np.unique(np.sort(_edges(f), 1), axis=0, return_counts=True)
    assert (cnt % 2 == 0).all()                           # every edge is shared by an even number of faces: no boundary anywhere
    inside = vol < 0
    n_cross = sum(int((np.diff(inside.astype(np.int8), axis=a) != 0).sum()) for a in range(3))
    assert len(v) == n_cross                              # exactly one vertex per sign-changing grid edge


def test_nearest_labels_round_half_even_and_clip():
    from oracle import mesh_mc
    sem = np.arange(4 * 4 * 4).reshape(4, 4, 4)
    verts = np.array([[0.5, 1.5, 2.5], [3.2, -0.4, 3.6], [2.5, 2.5, 0.49]], np.float32)
    s, i = mesh_mc.nearest_labels(verts, sem, sem * 2)
    assert s.tolist() == [int(sem[0, 2, 2]), int(sem[3, 0, 3]), int(sem[2, 2, 0])] and (i == 2 * s).all()


def test_ply_writer_layout(tmp_path):
    from eprecon_b200.mesh import COLOR_PALETTE, Mesh
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    n = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    col = np.concatenate([COLOR_PALETTE[[0, 1, 2, 3]], np.full((4, 1), 255, np.uint8)], 1)
    path = os.path.join(tmp_path, "m.ply")
    Mesh(v, f, n, col).export(path)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 4"]
    assert "property list uchar int vertex_indices" in lines and "property uchar alpha" in lines and "property float nz" in lines
    assert len(body) == 4 * (6 * 4 + 4) + 2 * (1 + 12)
    x, y, z, nx, ny, nz, r, g, b, a = struct.unpack_from("<6f4B", body, 28 * 1)
    assert (x, y, z, nz) == (1.0, 0.0, 0.0, 1.0) and (r, g, b, a) == (128, 128, 128, 255)
    cnt, i0, i1, i2 = struct.unpack_from("<B3i", body, 4 * 28 + 13)
    assert (cnt, i0, i1, i2) == (3, 0, 2, 3)
    assert len(COLOR_PALETTE) == 51
