"""TEST INFRASTRUCTURE (oracle): numpy marching cubes + nearest-voxel labels, the CPU restatement the CUDA mesh export
(eprecon_b200/csrc/mesh.cu) is checked against.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import it.

PARITY UNPINNED against the reference's mesher: the reference calls skimage.measure.marching_cubes(tsdf_vol, level=0)
(utils.py:231-236: Lewiner variant), which is neither vendored nor installable offline.  What is restated exactly is
everything AROUND the mesher -- the dense scene volume with default 1 (models/gru_fusion.py:217-257), `verts * voxel_size +
origin`, the nearest-voxel semantic / instance lookup `np.round -> np.clip -> index` (utils.py:240-244) and the colour
mapping (utils.py:250-268) -- and the mesher itself is a classic marching cubes over a DERIVED, face-consistent case table
(tools/gen_mc_table.py).  Any marching-cubes variant, skimage's included, puts its vertices on the sign-changing grid edges
at the linearly interpolated zero crossing, which is what `vertices()` computes; the triangulation inside ambiguous cells
(and Lewiner's extra interior vertices there) may differ.  tests/test_mesh_cpu.py checks the properties that do not depend
on the variant: watertightness, Euler characteristic, orientation, vertex positions.
"""
import numpy as np

from .mc_table import MC_COUNT, MC_TRIS


def _grad(vol, axis):
    """np.gradient along one axis: central differences inside, one-sided at the borders (float32)."""
    if vol.shape[axis] == 1:
        return np.zeros_like(vol)
    return np.gradient(vol.astype(np.float32), axis=axis).astype(np.float32)


def marching_cubes(vol, level=0.0):
    """vol float32 [dx,dy,dz] -> verts f32 [nv,3] (index coordinates), faces int32 [nf,3], normals f32 [nv,3].
    Vertex order: ascending (voxel raster index, axis) of the crossed grid edge; face order: ascending cell raster index,
    table order inside a cell."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    dx, dy, dz = vol.shape
    level = np.float32(level)
    inside = vol < level
    n = dx * dy * dz
    flags = np.zeros((n, 3), dtype=bool)
    lin = np.arange(n).reshape(dx, dy, dz)
    for axis in range(3):
        sl0 = [slice(None)] * 3
        sl1 = [slice(None)] * 3
        sl0[axis], sl1[axis] = slice(0, -1), slice(1, None)
        cross = inside[tuple(sl0)] != inside[tuple(sl1)]
        flags[lin[tuple(sl0)][cross], axis] = True
    edge_ids = np.flatnonzero(flags.reshape(-1))            # 3 * voxel + axis, ascending
    pos_of_edge = np.full(3 * n, -1, dtype=np.int64)
    pos_of_edge[edge_ids] = np.arange(len(edge_ids))
    vox, axis = edge_ids // 3, edge_ids % 3
    x, y, z = vox // (dy * dz), (vox // dz) % dy, vox % dz
    p0 = np.stack([x, y, z], 1)
    p1 = p0.copy()
    p1[np.arange(len(axis)), axis] += 1
    v0 = vol[p0[:, 0], p0[:, 1], p0[:, 2]]
    v1 = vol[p1[:, 0], p1[:, 1], p1[:, 2]]
    t = ((level - v0) / (v1 - v0)).astype(np.float32)
    verts = p0.astype(np.float32)
    verts[np.arange(len(axis)), axis] = verts[np.arange(len(axis)), axis] + t
    g = np.stack([_grad(vol, a) for a in range(3)], -1)     # [dx,dy,dz,3]
    g0 = g[p0[:, 0], p0[:, 1], p0[:, 2]]
    g1 = g[p1[:, 0], p1[:, 1], p1[:, 2]]
    gv = (g0 + t[:, None] * (g1 - g0)).astype(np.float32)
    ln = np.sqrt((gv[:, 0] * gv[:, 0] + gv[:, 1] * gv[:, 1]) + gv[:, 2] * gv[:, 2]).astype(np.float32)
    inv = np.where(ln > 0, np.float32(1.0) / np.maximum(ln, np.float32(1e-38)), np.float32(0.0)).astype(np.float32)
    normals = (gv * inv[:, None]).astype(np.float32)
    # faces
    if min(dx, dy, dz) < 2:
        return verts, np.zeros((0, 3), np.int32), normals
    idx = np.zeros((dx - 1, dy - 1, dz - 1), dtype=np.int32)
    for c in range(8):
        cx, cy, cz = c & 1, (c >> 1) & 1, c >> 2
        idx |= inside[cx:cx + dx - 1, cy:cy + dy - 1, cz:cz + dz - 1].astype(np.int32) << c
    count = np.asarray(MC_COUNT, dtype=np.int32)[idx]
    cells = np.argwhere(count > 0)                            # ascending raster order of (x,y,z)
    table = np.asarray(MC_TRIS, dtype=np.int32)
    faces = []
    for (cx, cy, cz) in cells:
        case = idx[cx, cy, cz]
        row = table[case]
        for k in range(3 * MC_COUNT[case]):
            e = int(row[k])
            ax, j = e >> 2, e & 3
            o = [0, 0, 0]
            o[(ax + 1) % 3], o[(ax + 2) % 3] = j & 1, j >> 1
            owner = ((cx + o[0]) * dy + (cy + o[1])) * dz + (cz + o[2])
            faces.append(pos_of_edge[3 * owner + ax])
    faces = np.asarray(faces, dtype=np.int32).reshape(-1, 3)
    return verts, faces, normals


def nearest_labels(verts, semantic_vol, instance_vol):
    """utils.py:240-244: round half to even, clip to the volume, index."""
    r = np.round(verts).astype(int)
    r = np.clip(r, [0, 0, 0], np.array(semantic_vol.shape) - 1)
    return semantic_vol[r[:, 0], r[:, 1], r[:, 2]], instance_vol[r[:, 0], r[:, 1], r[:, 2]]
