"""Scene TSDF -> triangle meshes (SURVEY.md section 8 f row 3): drop-ins for the mesh tail of the reference's SaveScene
(utils.py:231-288 `tsdf2mesh` / `tsdf_panoptic2mesh`, :362-388 `save_scene_eval`) on the dense scene volumes that
GRUFusion.save_mesh returns in `outputs` (models/gru_fusion.py:217-257: 'scene_tsdf', 'scene_semantic', 'scene_instance',
'origin').

The reference runs skimage.measure.marching_cubes on the CPU and wraps the result in trimesh objects; here the surface
extraction, the vertex normals and the nearest-voxel label lookup run on the GPU (csrc/mesh.cu, three kernels) and a small
`Mesh` container writes the same binary PLY layout trimesh exports (float x y z [nx ny nz], uchar red green blue alpha,
list uchar int vertex_indices).  No CPU fallback: CUDA tensors only.

Mesher note (also DESIGN.md): skimage's Lewiner tables are not available offline; the case table is derived
(tools/gen_mc_table.py).  Vertices coincide with any marching-cubes variant's (one per sign-changing grid edge at the linear
zero crossing); triangulation inside ambiguous cells may differ from skimage's.
"""
import os

import numpy as np
import torch

from . import _lib, ops

# the reference's palette (utils.py:250-262); semantic label s -> COLOR_PALETTE[s], instance id i -> COLOR_PALETTE[i % 50]
COLOR_PALETTE = np.array([
    [255, 192, 203], [128, 128, 128], [144, 238, 144], [0, 0, 255], [255, 255, 0], [0, 255, 255],
    [0, 128, 255], [128, 0, 255], [255, 0, 128], [255, 0, 0], [255, 255, 255],
    [255, 192, 203], [75, 0, 130], [255, 165, 0], [0, 100, 0], [255, 20, 147],
    [100, 149, 237], [255, 105, 180], [205, 92, 92], [186, 85, 211], [124, 252, 0],
    [70, 130, 180], [255, 215, 0], [0, 255, 255], [255, 69, 0], [138, 43, 226],
    [255, 105, 180], [70, 130, 180], [255, 192, 203], [219, 112, 147], [128, 128, 0],
    [255, 105, 180], [255, 20, 147], [255, 99, 71], [255, 69, 0], [255, 215, 0],
    [255, 182, 193], [0, 255, 0], [0, 255, 127], [34, 139, 34], [255, 240, 245],
    [255, 0, 255], [128, 0, 0], [0, 128, 0], [0, 0, 128], [128, 128, 0],
    [0, 128, 128], [128, 0, 128], [255, 128, 0], [128, 255, 0], [0, 255, 128],
], dtype=np.uint8)


class Mesh:
    """Minimal stand-in for the trimesh.Trimesh objects the reference returns: vertices / faces / vertex_normals /
    vertex_colors (host numpy arrays) + export(path) to binary PLY."""

    def __init__(self, vertices, faces, vertex_normals=None, vertex_colors=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32)
        self.faces = np.ascontiguousarray(faces, dtype=np.int32)
        self.vertex_normals = None if vertex_normals is None else np.ascontiguousarray(vertex_normals, dtype=np.float32)
        self.vertex_colors = None if vertex_colors is None else np.ascontiguousarray(vertex_colors, dtype=np.uint8)

    def copy(self):
        return Mesh(self.vertices.copy(), self.faces.copy(), None if self.vertex_normals is None else self.vertex_normals.copy(),
                    None if self.vertex_colors is None else self.vertex_colors.copy())

    def export(self, path):
        nv, nf = len(self.vertices), len(self.faces)
        props, cols = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")], [self.vertices]
        header = ["ply", "format binary_little_endian 1.0", f"element vertex {nv}",
                  "property float x", "property float y", "property float z"]
        if self.vertex_normals is not None:
            header += ["property float nx", "property float ny", "property float nz"]
            props += [("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4")]
        if self.vertex_colors is not None:
            header += ["property uchar red", "property uchar green", "property uchar blue", "property uchar alpha"]
            props += [("red", "u1"), ("green", "u1"), ("blue", "u1"), ("alpha", "u1")]
        header += [f"element face {nf}", "property list uchar int vertex_indices", "end_header"]
        vrec = np.zeros(nv, dtype=props)
        vrec["x"], vrec["y"], vrec["z"] = self.vertices[:, 0], self.vertices[:, 1], self.vertices[:, 2]
        if self.vertex_normals is not None:
            vrec["nx"], vrec["ny"], vrec["nz"] = self.vertex_normals[:, 0], self.vertex_normals[:, 1], self.vertex_normals[:, 2]
        if self.vertex_colors is not None:
            c = self.vertex_colors
            vrec["red"], vrec["green"], vrec["blue"] = c[:, 0], c[:, 1], c[:, 2]
            vrec["alpha"] = c[:, 3] if c.shape[1] > 3 else 255
        frec = np.zeros(nf, dtype=[("n", "u1"), ("v", "<i4", (3,))])
        frec["n"] = 3
        frec["v"] = self.faces
        with open(path, "wb") as f:
            f.write(("\n".join(header) + "\n").encode("ascii"))
            f.write(vrec.tobytes())
            f.write(frec.tobytes())


def marching_cubes(tsdf_vol, level=0.0, semantic_vol=None, instance_vol=None):
    """tsdf_vol CUDA f32 [dx,dy,dz] -> dict(verts f32 [nv,3] in index coordinates, faces int32 [nf,3], normals f32 [nv,3]
    [, semantics int32 [nv], instances int32 [nv]]).  Vertex order: ascending (voxel raster index, axis) of the crossed
    grid edge; face order: ascending cell raster index.  Empty surface -> zero-row tensors."""
    if not tsdf_vol.is_cuda:
        raise _lib.EpreconError("marching_cubes needs a CUDA tensor (eprecon_b200 has no CPU path)")
    L = _lib.lib()
    vol = tsdf_vol.float().contiguous()
    dx, dy, dz = (int(v) for v in vol.shape)
    dev = vol.device
    n = dx * dy * dz
    st = ops.stream_ptr()
    edge_flags = torch.empty(3 * n, dtype=torch.uint8, device=dev)
    cell_ntri = torch.empty(n, dtype=torch.uint8, device=dev)
    _lib.check(L.ep_mc_classify(vol.data_ptr(), dx, dy, dz, float(level), edge_flags.data_ptr(), cell_ntri.data_ptr(), st),
               "ep_mc_classify")
    edge_index, nv, edge_pos = ops.compact_flags(edge_flags, want_pos=True)
    out = {"verts": torch.zeros((nv, 3), dtype=torch.float32, device=dev), "faces": torch.zeros((0, 3), dtype=torch.int32, device=dev),
           "normals": torch.zeros((nv, 3), dtype=torch.float32, device=dev)}
    want_labels = semantic_vol is not None and instance_vol is not None
    if want_labels:
        sem = semantic_vol.to(torch.int32).contiguous()
        ins = instance_vol.to(torch.int32).contiguous()
        out["semantics"] = torch.zeros(nv, dtype=torch.int32, device=dev)
        out["instances"] = torch.zeros(nv, dtype=torch.int32, device=dev)
    if nv == 0:
        return out
    _lib.check(L.ep_mc_vertices(vol.data_ptr(), dx, dy, dz, float(level), edge_index.data_ptr(), nv, out["verts"].data_ptr(),
                                out["normals"].data_ptr(), sem.data_ptr() if want_labels else 0, ins.data_ptr() if want_labels else 0,
                                out["semantics"].data_ptr() if want_labels else 0, out["instances"].data_ptr() if want_labels else 0,
                                st), "ep_mc_vertices")
    cell_index, nc = ops.compact_flags(cell_ntri)            # non-zero counts are truthy flags
    if nc == 0:
        return out
    counts = cell_ntri[cell_index.long()].to(torch.int32)
    tri_offset = (torch.cumsum(counts, 0, dtype=torch.int32) - counts).contiguous()
    n_tri = int(counts.sum().item())
    faces = torch.empty((n_tri, 3), dtype=torch.int32, device=dev)
    _lib.check(L.ep_mc_faces(vol.data_ptr(), dx, dy, dz, float(level), cell_index.data_ptr(), tri_offset.data_ptr(), nc,
                             edge_pos.data_ptr(), faces.data_ptr(), st), "ep_mc_faces")
    out["faces"] = faces
    return out


def tsdf2mesh(voxel_size, origin, tsdf_vol):
    """SaveScene.tsdf2mesh (utils.py:231-236)."""
    mc = marching_cubes(tsdf_vol)
    org = torch.as_tensor(np.asarray(origin, dtype=np.float32), device=mc["verts"].device)
    verts = mc["verts"] * float(voxel_size) + org
    return Mesh(verts.cpu().numpy(), mc["faces"].cpu().numpy(), mc["normals"].cpu().numpy())


def tsdf_panoptic2mesh(voxel_size, origin, tsdf_vol, semantic_vol, instance_vol):
    """SaveScene.tsdf_panoptic2mesh (utils.py:238-288): geometry + per-vertex semantic / instance colours."""
    mc = marching_cubes(tsdf_vol, semantic_vol=semantic_vol, instance_vol=instance_vol)
    org = torch.as_tensor(np.asarray(origin, dtype=np.float32), device=mc["verts"].device)
    verts = (mc["verts"] * float(voxel_size) + org).cpu().numpy()
    faces, normals = mc["faces"].cpu().numpy(), mc["normals"].cpu().numpy()
    sem = mc["semantics"].cpu().numpy().astype(np.int64)
    ins = mc["instances"].cpu().numpy().astype(np.int64)
    mesh = Mesh(verts, faces, normals)
    alpha = np.full((len(verts), 1), 255, dtype=np.uint8)
    mesh_sem = Mesh(verts, faces, normals, np.concatenate([COLOR_PALETTE[sem], alpha], 1))
    mesh_ins = Mesh(verts, faces, normals, np.concatenate([COLOR_PALETTE[ins % len(COLOR_PALETTE)], alpha], 1))
    return mesh, mesh_sem, mesh_ins


def save_scene_eval(outputs, save_path, voxel_size, batch_idx=0):
    """SaveScene.save_scene_eval (utils.py:362-388): <scene>.npz (origin, voxel_size, tsdf, semantic, instance) +
    <scene>.ply + mesh_semantic_<scene>.ply + mesh_instance_<scene>.ply.  Returns the three meshes, or None when the scene
    volume holds no observed voxel (the reference logs a warning and writes nothing)."""
    tsdf = outputs["scene_tsdf"][batch_idx]
    inst = outputs["scene_instance"][batch_idx]
    sem = outputs["scene_semantic"][batch_idx]
    origin = outputs["origin"][batch_idx].detach().cpu().numpy()
    scene = outputs["scene_name"][batch_idx].replace("/", "-")
    if bool((tsdf == 1).all()):
        return None
    meshes = tsdf_panoptic2mesh(voxel_size, origin, tsdf, sem, inst)
    os.makedirs(save_path, exist_ok=True)
    np.savez_compressed(os.path.join(save_path, f"{scene}.npz"), origin=origin, voxel_size=voxel_size,
                        tsdf=tsdf.detach().cpu().numpy(), semantic=sem.detach().cpu().numpy(), instance=inst.detach().cpu().numpy())
    meshes[0].export(os.path.join(save_path, f"{scene}.ply"))
    meshes[1].export(os.path.join(save_path, f"mesh_semantic_{scene}.ply"))
    meshes[2].export(os.path.join(save_path, f"mesh_instance_{scene}.ply"))
    return meshes
