"""Host-side check of the Poseidon table derivations used by the CUDA permutation (no GPU needed):
every dense/sparse split of the partial rounds reproduces the textbook permutation (and its KAT), the
constants folded into the FP64 MDS layers are the next round's, the split-circulant constants are integers."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not found")
def test_sparse_and_hybrid_tables_match_textbook_permutation(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "poseidon_tables_check")
    src = os.path.join(ROOT, "tests", "native", "poseidon_tables_check.cu")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr
