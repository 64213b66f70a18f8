"""CPU tests of the host side: the C-ABI library loads and exports every symbol
include/mfcc_b200.h declares (no compute without a GPU), pure-host entry points
agree with the oracle, and the sharding arithmetic is sound."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from mfcc_b200 import api, config_a, config_b, config_c, make_params, sharding, PAD_ZERO_TAIL
from mfcc_b200.params import MfccParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mfcc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfcc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = api.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mfcc_b200.h but not exported"
    assert set(names) == set(api.ABI), "api.ABI and the header disagree"


def test_params_struct_layout_matches_header():
    assert C.sizeof(MfccParams) == 14 * 4
    p = MfccParams()
    assert api.load().mfcc_params_init(C.byref(p), 16000) == 0
    a = config_a()
    assert p.as_dict() == pytest.approx(a.as_dict())
    assert api.load().mfcc_params_init(C.byref(p), 8000) == 0
    assert (p.frame_len, p.hop_len, p.nfft) == (200, 80, 256)
    assert api.load().mfcc_params_init(C.byref(p), 48000) == 0
    assert (p.frame_len, p.hop_len, p.nfft) == (1200, 480, 2048)


def test_num_frames_and_validation_agree_with_oracle():
    lib = api.load()
    rng = np.random.default_rng(0)
    for p in (config_a(), config_b(), config_c(), config_a().copy(pad_mode=PAD_ZERO_TAIL)):
        for n in list(range(0, 3 * p.frame_len, 37)) + rng.integers(0, 10**7, 50).tolist():
            assert api.num_frames(p, int(n)) == oracle.num_frames(p, int(n))
        assert lib.mfcc_out_dim(C.byref(p)) == p.out_dim
    bad = [dict(nfft=500), dict(frame_len=600), dict(n_cep=27), dict(n_mel=0), dict(hop_len=0),
           dict(window=7), dict(log_floor=0.0), dict(f_hi=9000.0), dict(f_lo=8000.0), dict(lifter=-1),
           dict(preemph=1.5), dict(pad_mode=3), dict(output=2)]
    for kw in bad:
        q = make_params(**kw)
        assert lib.mfcc_params_validate(C.byref(q)) == -1, kw
        assert api.num_frames(q, 16000) == -1
    assert api.num_frames(config_a(), -5) == -1
    assert lib.mfcc_strerror(-3).decode().startswith("CUDA")


def test_no_gpu_means_ecuda_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.MfccError) as e:
        api.Plan(config_a())
    assert e.value.code == -3


def test_partition_balances_frames_and_covers_everything():
    p = config_b()
    from mfcc_b200.synth import ragged_batch
    _, off = ragged_batch(1000, 4000, 24000, seed=3)
    nf = sharding.frame_counts(p, off)
    assert np.array_equal(nf, [oracle.num_frames(p, int(n)) for n in np.diff(off)])
    for W in (1, 2, 3, 4, 8):
        parts = sharding.partition(p, off, W)
        assert parts[0][0] == 0 and parts[-1][1] == 1000
        assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))
        loads = [int(nf[a:b].sum()) for a, b in parts]
        assert sum(loads) == int(nf.sum())
        assert max(loads) - min(loads) <= nf.max() + 1
    # fewer utterances than ranks: nothing lost, some ranks empty
    parts = sharding.partition(p, off[:4], 8)
    assert sum(b - a for a, b in parts) == 3
    s0, s1, loc = sharding.local_slice(off, 10, 20)
    assert loc[0] == 0 and loc[-1] == s1 - s0 and np.array_equal(np.diff(loc), np.diff(off[10:21]))


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/mfcc_b200.h compiles as strict C99 and a C caller links against the shared library
    (INTEGRATION.md §3).  Without a GPU the demo stops after the host-only calls (MFCC_ECUDA, no fallback)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "demo")
    libdir = os.path.join(root, "mfcc_b200")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cabi", "demo.c"), "-o", exe, "-L", libdir, "-lmfcc_b200",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "mfcc_b200" in r.stdout
