"""GPU marching cubes + label lookup (csrc/mesh.cu) vs the numpy oracle (oracle/mesh_mc.py): vertices, faces and labels
bit-exact (same arithmetic order, same table, same ordering rules), normals to 1e-6; then the reference-shaped entry points
on scene volumes produced by GRUFusion(direct_substitute=True).save_mesh."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from test_mesh_cpu import sphere  # noqa: E402


def _check(vol, sem=None, ins=None):
    from oracle import mesh_mc
    from eprecon_b200 import mesh
    want_v, want_f, want_n = mesh_mc.marching_cubes(vol)
    got = mesh.marching_cubes(torch.from_numpy(vol).cuda(), semantic_vol=None if sem is None else torch.from_numpy(sem).cuda(),
                              instance_vol=None if ins is None else torch.from_numpy(ins).cuda())
    assert np.array_equal(got["verts"].cpu().numpy(), want_v)
    assert np.array_equal(got["faces"].cpu().numpy(), want_f)
    assert np.abs(got["normals"].cpu().numpy() - want_n).max() <= 2e-6
    if sem is not None:
        ws, wi = mesh_mc.nearest_labels(want_v, sem, ins)
        assert np.array_equal(got["semantics"].cpu().numpy(), ws) and np.array_equal(got["instances"].cpu().numpy(), wi)
    return got


def test_sphere_bit_exact(cuda_lib):
    got = _check(sphere())
    assert got["faces"].shape[0] == 3524 and got["verts"].shape[0] == 1764


def test_random_volume_with_labels_bit_exact(cuda_lib):
    rng = np.random.default_rng(1)
    vol = rng.standard_normal((37, 23, 41)).astype(np.float32)
    vol[rng.random(vol.shape) < 0.3] = 1.0                     # unobserved voxels carry the default 1, as in the scene volume
    sem = rng.integers(0, 21, vol.shape).astype(np.int32)
    ins = rng.integers(0, 300, vol.shape).astype(np.int32)
    _check(vol, sem, ins)


def test_degenerate_volumes(cuda_lib):
    from eprecon_b200 import mesh
    empty = mesh.marching_cubes(torch.ones((8, 9, 10), device="cuda"))
    assert empty["verts"].shape == (0, 3) and empty["faces"].shape == (0, 3)
    _check(np.where(np.arange(5 * 1 * 6).reshape(5, 1, 6) % 2 == 0, -0.5, 0.5).astype(np.float32))   # a flat volume: vertices, no cells


def test_scene_export_from_direct_substitute_fusion(cuda_lib, tmp_path):
    """tsdf_panoptic2mesh / save_scene_eval on the scene volumes of three fused fragments (the same fixture inputs that pin
    the panoptic fusion to the reference)."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_panoptic_fusion import N_VOX, fusion_inputs
    from oracle import mesh_mc
    from eprecon_b200 import mesh, synth
    from eprecon_b200.gru_fusion import GRUFusion
    cfg = synth.make_cfg(n_vox=N_VOX)
    fuse = GRUFusion(cfg, direct_substitute=True, trianing=False)
    outputs = {}
    for frag in range(3):
        inputs, coords, tsdf, info = fusion_inputs(frag)
        cin = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inputs.items() if k not in ("occ_list", "tsdf_list")}
        pinfo = {"panoptic_seg": [info["panoptic_seg"][0].cuda(), info["panoptic_seg"][1]]}
        outputs = fuse(coords.cuda(), tsdf.cuda(), cin, 2, outputs, save_mesh=True, panoptic_infos=[pinfo])
    tsdf_vol = outputs["scene_tsdf"][-1]
    sem_vol, ins_vol = outputs["scene_semantic"][-1], outputs["scene_instance"][-1]
    origin = outputs["origin"][-1].cpu().numpy()
    m, ms, mi = mesh.tsdf_panoptic2mesh(cfg.VOXEL_SIZE, origin, tsdf_vol, sem_vol, ins_vol)
    wv, wf, wn = mesh_mc.marching_cubes(tsdf_vol.cpu().numpy())
    ws, wi = mesh_mc.nearest_labels(wv, sem_vol.cpu().numpy(), ins_vol.cpu().numpy())
    assert len(wf) > 1000
    assert np.array_equal(m.faces, wf)
    assert np.array_equal(m.vertices, (torch.from_numpy(wv) * cfg.VOXEL_SIZE + torch.from_numpy(origin.astype(np.float32))).numpy())
    assert np.array_equal(ms.vertex_colors[:, :3], mesh.COLOR_PALETTE[ws.astype(np.int64)])
    assert np.array_equal(mi.vertex_colors[:, :3], mesh.COLOR_PALETTE[wi.astype(np.int64) % 51])
    plain = mesh.tsdf2mesh(cfg.VOXEL_SIZE, origin, tsdf_vol)
    assert np.array_equal(plain.vertices, m.vertices) and plain.vertex_colors is None
    res = mesh.save_scene_eval(outputs, str(tmp_path), cfg.VOXEL_SIZE, batch_idx=len(outputs["scene_name"]) - 1)
    assert res is not None
    scene = outputs["scene_name"][-1]
    for name in (f"{scene}.npz", f"{scene}.ply", f"mesh_semantic_{scene}.ply", f"mesh_instance_{scene}.ply"):
        assert os.path.getsize(os.path.join(tmp_path, name)) > 0
    z = np.load(os.path.join(tmp_path, f"{scene}.npz"))
    assert np.array_equal(z["tsdf"], tsdf_vol.cpu().numpy()) and float(z["voxel_size"]) == cfg.VOXEL_SIZE
