"""Host check of the register FFT codelets (mfcc_b200/csrc/mfcc_rfft.cuh): tests/codelets/check_rfft.cpp is built with
g++ and compares every codelet, and the kernels' two-pass index algebra, with a direct double-precision DFT.  Both
builds of the header are checked: the product-form twiddles and the FMA-fused butterflies (MFCC_RFFT_FUSED)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("fused", [0, 1])
def test_codelets_match_direct_dft(tmp_path, fused):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = tmp_path / f"check_rfft_{fused}"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-DMFCC_RFFT_FUSED={fused}", "-o", str(exe),
                    os.path.join(HERE, "codelets", "check_rfft.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "failures 0" in r.stdout, r.stdout[-2000:]
