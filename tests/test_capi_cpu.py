"""C-ABI contract checks that need no GPU: the library loads and exports every symbol the header declares; the host-side
mirrors keep the reference's names and state-dict layout; the product path fails loudly without CUDA."""
import os
import re

import pytest
import torch


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from eprecon_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 40
    L = _lib.lib()
    for name in protos:
        assert hasattr(L, name), name
    assert L.ep_version() >= 100
    hdr = open(_lib.HEADER_PATH).read()
    assert len(re.findall(r"\bep_\w+\s*\(", hdr)) == len(protos)      # the parser saw every prototype


def test_dominant_conv_kernel_keeps_three_ctas_per_sm():
    """The default sparse-conv kernel (spconv_hl_cp_kernel<CA=false, TICKET=false>) is tuned for 3 resident CTAs of 256 threads
    per SM, i.e. at most 85 registers per thread and no spills; a silent register growth once cost a third of the occupancy."""
    import shutil
    import subprocess
    import __graft_entry__ as g
    g.build()
    from eprecon_b200 import _lib
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    hits = re.findall(r"Function (\S*spconv_hl_cp_kernelILb0ELb0E\S*):\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert len(hits) == 1, out[:400]
    assert int(hits[0][1]) <= 85 and int(hits[0][2]) == 0, hits


def test_header_has_no_torch_types():
    from eprecon_b200 import _lib
    hdr = re.sub(r"/\*.*?\*/", " ", open(_lib.HEADER_PATH).read(), flags=re.S)   # prototypes only, comments stripped
    assert "torch" not in hdr.lower() and "at::" not in hdr and "Tensor" not in hdr


def test_state_dict_layout_mirrors_reference_names():
    from eprecon_b200 import synth
    from eprecon_b200.neucon_network import NeuConNet
    sd = NeuConNet(synth.make_cfg()).state_dict()
    assert sd["sp_convs.0.stem.0.kernel"].shape == (27, 80, 32)                    # torchsparse [K,Cin,Cout]
    assert sd["sp_convs.1.stage1.0.net.0.kernel"].shape == (8, 16, 16)
    assert sd["sp_convs.2.up1.0.net.0.kernel"].shape == (8, 32, 24)
    assert sd["initialization.subm1.sparsesubmconv3d.weight"].shape == (32, 3, 3, 3, 32)   # spconv [Cout,k,k,k,Cin]
    assert sd["initialization.similary_1.conv7.conv.weight"].shape == (32, 1, 1, 1, 128)
    assert sd["gru_fusion.fusion_nets_voxel.0.convz.net.kernel"].shape == (27, 192, 96)
    assert sd["gru_fusion.fusion_nets_img.2.convq.point_transforms.0.weight"].shape == (24, 48)
    assert sd["tsdf_preds.0.linear1.weight"].shape == (384, 96)
    assert "sp_convs.0.stem.1.running_mean" in sd


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from eprecon_b200 import _lib, synth
    from eprecon_b200.occupancy_initialization import Back_Project
    inputs, fa, fb = synth.make_fragment(seed=1, image_hw=(120, 160), n_vox=(32, 32, 32))
    coords = torch.zeros((8, 4), dtype=torch.int32)
    feats = torch.stack([f[2] for f in fb])
    kr = inputs["proj_matrices"][:, :, 2].permute(1, 0, 2, 3).contiguous()
    with pytest.raises(_lib.EpreconError):
        Back_Project(80)(coords, inputs["vol_origin_partial"], 0.04, feats, kr, 2)


def test_oracle_is_not_imported_by_the_product():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eprecon_b200")
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_replicas_share_parameters_but_not_state():
    from eprecon_b200 import synth
    from eprecon_b200.neucon_network import NeuConNet
    from eprecon_b200.streams import replicate
    net = NeuConNet(synth.make_cfg())
    synth.fill_parameters_(net, 1)
    rep = replicate(net)
    a, b = dict(net.named_parameters()), dict(rep.named_parameters())
    assert a.keys() == b.keys()
    assert all(a[k].data_ptr() == b[k].data_ptr() for k in a)          # same storage, no weight copy
    assert rep.gru_fusion is not net.gru_fusion and rep.gru_fusion.global_volume is not net.gru_fusion.global_volume
    assert rep.training == net.training


def test_executor_descriptors_parse_natively():
    """The flat int64 parameter descriptors executor.py builds are parsed by the SAME C++ reader the native executor uses
    (ep_exec_desc_check, host only): layout, end markers and the channel bookkeeping must agree for every module."""
    from eprecon_b200 import _lib, executor, synth
    from eprecon_b200.neucon_network import NeuConNet
    net = NeuConNet(synth.make_cfg())
    synth.fill_parameters_(net, 1)
    L = _lib.lib()
    for sp in net.sp_convs:
        arr, keep = executor._build_spvcnn(sp)
        assert L.ep_exec_desc_check(0, arr.ctypes.data) == arr.size
    for gv, gi in zip(net.gru_fusion.fusion_nets_voxel, net.gru_fusion.fusion_nets_img):
        arr, keep = executor._build_gru((gv, gi))
        assert L.ep_exec_desc_check(1, arr.ctypes.data) == arr.size
    for i in range(3):
        arr, keep = executor._build_lin4x((net.tsdf_preds[i], net.occ_preds[i]))
        assert L.ep_exec_desc_check(2, arr.ctypes.data) == arr.size
        arr, keep = executor._build_lin4x((net.panoptic_preds[i],))
        assert L.ep_exec_desc_check(2, arr.ctypes.data) == arr.size
    arr, keep = executor._build_init(net.initialization)
    assert L.ep_exec_desc_check(3, arr.ctypes.data) == arr.size
    bad = arr.copy()
    bad[-1] = 0                                                        # broken end marker -> rejected
    assert L.ep_exec_desc_check(3, bad.ctypes.data) < 0


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (build container only)")
def test_state_dict_keys_and_shapes_equal_the_reference_module():
    """Every parameter / buffer of the drop-in NeuConNet exists under the same name and shape in the UNMODIFIED reference
    NeuConNet (incl. the panoptic decoder), so reference checkpoints load by name."""
    from oracle import ref_import
    from eprecon_b200 import synth
    from eprecon_b200.neucon_network import NeuConNet
    ns = ref_import.load()
    cfg = synth.make_cfg()
    ref = {k: tuple(v.shape) for k, v in ns.neucon.NeuConNet(cfg).state_dict().items()}
    ours = {k: tuple(v.shape) for k, v in NeuConNet(cfg).state_dict().items()}
    missing = [k for k in ours if k not in ref]
    assert not missing, missing[:10]
    wrong = [(k, ours[k], ref[k]) for k in ours if ours[k] != ref[k]]
    assert not wrong, wrong[:10]
    # what the reference has and the drop-in does not: only the training criterion (loss weights) and BN bookkeeping
    extra = [k for k in ref if k not in ours and not k.startswith("criterion.") and not k.endswith("num_batches_tracked")]
    assert not extra, extra[:10]
