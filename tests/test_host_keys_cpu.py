"""The key functions the CUDA kernels share with host code (`__host__ __device__` in eprecon_b200/csrc/common.cuh), compiled
for the CPU with nvcc and checked without a GPU:
  * the compact Z-order sort keys of the executor (24 / 32-bit tiers) order rows exactly like the 64-bit keys and round-trip;
  * the 60-bit coordinate hash equals the oracle's restatement of torchsparse v2.0.0's hash (`oracle/shims`), the function
    whose ascending order the reference's ConvGRU quirk makes parity-relevant."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def keys_bin(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host") / "keys_host")
    cmd = [NVCC, "-std=c++17", "-O2", "-I", os.path.join(ROOT, "eprecon_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
           "-o", out, os.path.join(ROOT, "tests", "host", "keys_host.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return out


def test_compact_sort_keys_keep_the_row_order(keys_bin):
    r = subprocess.run([keys_bin, "order"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr


def test_coordinate_hash_matches_the_oracle(keys_bin):
    from oracle.shims.torchsparse.nn.functional import sphash
    g = torch.Generator().manual_seed(2)
    c = torch.randint(-300, 300, (4000, 4), generator=g, dtype=torch.int32)
    c[:, 3] = torch.randint(0, 4, (4000,), generator=g, dtype=torch.int32)
    c[:8] = torch.tensor([[0, 0, 0, 0], [-1, -1, -1, 0], [2147483647, 0, 0, 0], [-2147483648, 5, 5, 1], [1, 2, 3, 0], [3, 2, 1, 0],
                          [95, 95, 95, 0], [-128, 127, 0, 3]], dtype=torch.int32)
    txt = "\n".join(" ".join(str(int(v)) for v in row) for row in c.tolist()) + "\n"
    r = subprocess.run([keys_bin, "hash", str(c.shape[0])], input=txt, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.array([int(x) for x in r.stdout.split()], dtype=np.uint64)
    want = sphash(c).numpy().astype(np.uint64)
    assert np.array_equal(got, want)
