"""Host-side logic of bench.py and the profile tools (no GPU): input packing, CPU-sample sizing, the reference arm's JSON
contract, and the launch-list summariser on the committed ncu list."""
import io
import json
import os
import subprocess
import sys
from contextlib import redirect_stdout

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_packed_host_round_trip(monkeypatch):
    if not torch.cuda.is_available():
        monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)   # pinning needs a CUDA driver
    g = torch.Generator().manual_seed(0)
    obj = {"inputs": {"a": torch.randn(3, 4, generator=g), "m": torch.rand(5, generator=g) > 0.5, "l": [torch.randn(2, generator=g), torch.randn(7, generator=g)]},
           "fa": [[torch.randn(1, 2, 3, generator=g)], [torch.randn(1, 2, 3, generator=g)]], "tag": "x"}
    host = bench.PackedHost(obj)
    assert host.nbytes == sum(t.numel() * t.element_size() for t in (obj["inputs"]["a"], obj["inputs"]["m"], *obj["inputs"]["l"],
                                                                    obj["fa"][0][0], obj["fa"][1][0]))
    back = host.build({dt: b.clone() for dt, b in host.bufs.items()})     # stands in for the device copies
    assert torch.equal(back["inputs"]["a"], obj["inputs"]["a"]) and torch.equal(back["inputs"]["m"], obj["inputs"]["m"])
    assert torch.equal(back["inputs"]["l"][1], obj["inputs"]["l"][1]) and torch.equal(back["fa"][1][0], obj["fa"][1][0])
    assert back["tag"] == "x"


def test_cpu_sample_level_fits_the_budget():
    # 1 s level-0 probe -> a full fragment is estimated at 17 s
    assert bench.pick_sample_level(1.0, 1, 60.0) == 2
    assert bench.pick_sample_level(1.0, 11, 300.0) == 2
    assert bench.pick_sample_level(2.0, 11, 300.0) == 1
    assert bench.pick_sample_level(20.0, 11, 300.0) == 0
    assert bench.SAMPLE_FRACTION[2] == 1.0 and 0 < bench.SAMPLE_FRACTION[0] < bench.SAMPLE_FRACTION[1] < 1


def test_reference_arm_json_contract(monkeypatch):
    calls = []

    def fake_sample(steps, warmup, max_level=2, workload="fragment"):
        assert workload == "fragment"
        calls.append((steps, warmup, max_level))
        return {0: 0.5, 1: 1.4, 2: 9.0}[max_level]
    monkeypatch.setattr(bench, "cpu_sample", fake_sample)
    monkeypatch.setenv("RANK", "0")
    args = type("A", (), {"steps": 10, "warmup": 3, "gpus": 2, "impl": "reference"})()
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "fragments/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 2 and line["steps"] == 10
    assert calls[0] == (1, 0, 0) and calls[1] == (10, 1, 2)                # probe, then K full fragments
    assert abs(line["value"] - 1.0 / 9.0) < 1e-12 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["fragment_fraction_per_step"] == 1.0 and "FULL" in cb["sample"]
    # ranks other than 0 print nothing
    monkeypatch.setenv("RANK", "1")
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args)
    assert buf.getvalue() == ""


def test_launch_list_summariser_on_the_committed_ncu_list():
    csv_path = os.path.join(ROOT, "profiles", "r01_launches_v8_one_fragment.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), csv_path], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert "1441 launches" in lines[0]
    first = lines[2].split()
    assert first[3].startswith("spconv_tc_kernel") and int(first[2]) == 89   # dominant kernel, 89 launches per fragment
    shares = sum(float(l.split()[1].rstrip("%")) for l in lines[2:])
    assert abs(shares - 100.0) < 1.0
