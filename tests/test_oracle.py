"""CPU tests of the oracle: against numpy/scipy float64, the committed golden
vectors, closed-form known answers (SURVEY.md §8c list) and the framing table.

The reference holds no tests for this path (test/run.sh:1-10 and
test/test.rb:1-44 assert nothing about MFCC) — PARITY UNPINNED.
"""
import numpy as np
import pytest

import np_ref
import oracle
from mfcc_b200 import (config_a, config_b, config_c, make_params, OUT_LOGMEL, PAD_ZERO_TAIL,
                       WINDOW_RECT, WINDOW_HANN)
from mfcc_b200.synth import clip_config1, noise_utterance, hostile_clip, HOSTILE_KINDS
from util import assert_parity, golden, hostile_golden

CFG = {"A": config_a, "B": config_b, "C": config_c}


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_f64_matches_numpy(name):
    p = CFG[name]()
    x = noise_utterance(p.frame_len + 37 * p.hop_len + 5, seed=11)
    got = oracle.mfcc(p, x, np.float64)
    ref = np_ref.mfcc(p, x)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-9


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_f32_within_tolerance_of_f64(name):
    p = CFG[name]()
    x = noise_utterance(p.frame_len + 50 * p.hop_len, seed=12)
    assert_parity(oracle.mfcc(p, x, np.float32), oracle.mfcc(p, x, np.float64), what=name)


def test_golden_vectors():
    g = golden()
    a, b, c = config_a(), config_b(), config_c()
    cases = [
        (a, g["A_pcm"], g["A_cep"]),
        (a.copy(output=OUT_LOGMEL), g["A_pcm"], g["A_logmel"]),
        (a.copy(lifter=22), g["A_pcm"], g["A_lifter22"]),
        (a.copy(pad_mode=PAD_ZERO_TAIL), g["A_pcm"][:15000], g["A_padtail"]),
        (b, g["B_pcm"], g["B_cep"]),
        (c, g["C_pcm"], g["C_cep"]),
    ]
    for p, x, ref in cases:
        assert np.abs(oracle.mfcc(p, x, np.float64) - ref).max() < 1e-9
        assert_parity(oracle.mfcc(p, x, np.float32), ref)
    assert g["A_cep"].shape == (98, 13)


def test_golden_inputs_are_the_documented_signals():
    g = golden()
    assert np.array_equal(g["A_pcm"], clip_config1(1.0, 16000, 0))
    assert np.array_equal(g["B_pcm"], noise_utterance(6000, 3))


def test_tables_match_numpy():
    for p in (config_a(), config_b(), config_c(), config_a().copy(window=WINDOW_HANN),
              make_params(f_lo=64.0, f_hi=7600.0, n_mel=40, n_cep=20)):
        assert np.array_equal(oracle.mel_bins(p), np_ref.mel_bins(p))
        assert np.abs(oracle.window(p) - np_ref.window(p)).max() < 1e-15
        assert np.abs(oracle.mel_weights(p) - np_ref.mel_weights(p)).max() < 1e-15
    D = oracle.dct(config_a())
    M = 26
    ref = np.array([[np.sqrt((1 if k == 0 else 2) / M) * np.cos(np.pi * k * (m + 0.5) / M)
                     for m in range(M)] for k in range(13)])
    assert np.abs(D - ref).max() < 1e-15
    # orthonormal rows
    Dfull = oracle.dct(config_a().copy(n_cep=26))
    assert np.abs(Dfull @ Dfull.T - np.eye(26)).max() < 1e-12


# ---- framing: the bit-exact contract (SURVEY.md §8d "Frame-count convention") ----
def test_frame_count_table():
    p = config_a()
    L, H = p.frame_len, p.hop_len
    table = {0: 0, 1: 0, L - 1: 0, L: 1, L + 1: 1, L + H - 1: 1, L + H: 2, L + H + 1: 2,
             16000: 98, 160000: 998}
    for n, nf in table.items():
        assert oracle.num_frames(p, n) == nf == np_ref.num_frames(p, n), n
    q = p.copy(pad_mode=PAD_ZERO_TAIL)
    table = {0: 0, 1: 1, L - 1: 1, L: 1, L + 1: 2, L + H: 2, L + H + 1: 3, 160000: 999}
    for n, nf in table.items():
        assert oracle.num_frames(q, n) == nf == np_ref.num_frames(q, n), n
    assert oracle.num_frames(config_c(), 28_800_000) == 59_998
    assert oracle.num_frames(config_b(), 16000) == 198           # BASELINE.md §5 row 3 (fixed 2.0 s)


def test_framing_indices_exact():
    """Frame t is samples [t*hop, t*hop + L): an int16 ramp through a rectangular
    window with no pre-emphasis comes back out of the 'framed' tap unchanged."""
    p = make_params(preemph=0.0, window=WINDOW_RECT)
    x = (np.arange(3000) % 30000).astype(np.int16)
    for t in (0, 1, 7, oracle.num_frames(p, x.size) - 1):
        fr, _, _, _ = oracle.stages(p, x, t)
        assert np.array_equal(fr[: p.frame_len], x[t * p.hop_len: t * p.hop_len + p.frame_len].astype(np.float64))
        assert not fr[p.frame_len:].any()


def test_invalid_parameters_rejected():
    bad = [dict(nfft=500), dict(frame_len=600), dict(n_cep=27), dict(n_mel=0), dict(hop_len=0),
           dict(window=7), dict(log_floor=0.0), dict(f_hi=9000.0), dict(f_lo=8000.0), dict(lifter=-1),
           dict(preemph=1.5), dict(pad_mode=3), dict(output=2), dict(nfft=8192, frame_len=400)]
    for kw in bad:
        assert oracle.num_frames(make_params(**kw), 16000) < 0, kw
    assert oracle.num_frames(config_a(), -1) < 0


# ---- known answers (SURVEY.md §8c "Self-made known-answer tests") ----
def test_kat_all_zero_pcm():
    p = config_a()
    c = oracle.mfcc(p, np.zeros(2000, np.int16), np.float64)
    floor = float(np.float32(1e-10))
    assert np.allclose(c[:, 0], np.sqrt(26) * np.log(floor), rtol=0, atol=1e-9)
    assert np.abs(c[:, 1:]).max() < 1e-9


def test_kat_impulse_flat_spectrum():
    p = make_params(preemph=0.0)
    x = np.zeros(400, np.int16)
    x[0] = 1000
    _, pw, _, _ = oracle.stages(p, x, 0)
    w0 = 0.54 - 0.46
    assert np.allclose(pw, (1000 * w0) ** 2 / 512, rtol=1e-12)


def test_kat_bin_centred_sinusoid():
    p = make_params(preemph=0.0, window=WINDOW_RECT, frame_len=512, hop_len=160)
    k0 = 40
    n = np.arange(512)
    x = np.rint(8000 * np.cos(2 * np.pi * k0 * n / 512)).astype(np.int16)
    _, pw, mel, _ = oracle.stages(p, x, 0)
    assert pw.argmax() == k0 and pw[k0] > 1e6 * np.delete(pw, k0).max()
    b = oracle.mel_bins(p)
    lit = {m for m in range(26) if b[m] < k0 < b[m + 2]}
    assert 1 <= len(lit) <= 2
    dark = [m for m in range(26) if m not in lit]
    assert mel[dark].max() < 1e-6 * mel[list(lit)].max()


def test_kat_dc_with_unit_preemphasis():
    p = make_params(preemph=1.0, window=WINDOW_RECT)
    x = np.full(1000, 1234, np.int16)
    fr0, _, _, _ = oracle.stages(p, x, 0)
    assert fr0[0] == 1234 and not fr0[1:].any()
    fr1, _, _, _ = oracle.stages(p, x, 1)
    assert not fr1.any()


def test_kat_parseval():
    p = config_a()
    x = noise_utterance(2000, 5)
    fr, pw, _, _ = oracle.stages(p, x, 3)
    full = pw.copy()
    full[1:-1] *= 2
    assert np.isclose(full.sum(), (fr ** 2).sum(), rtol=1e-12)


def test_kat_dct_of_constant():
    D = oracle.dct(config_a())
    c = D @ np.full(26, 3.5)
    assert np.isclose(c[0], 3.5 * np.sqrt(26)) and np.abs(c[1:]).max() < 1e-12


def test_preemphasis_is_whole_utterance():
    """y[n] = x[n] - a x[n-1] over the utterance: frame t > 0 sees x[t*hop - 1]."""
    p = make_params(window=WINDOW_RECT)
    x = noise_utterance(1000, 9)
    fr, _, _, _ = oracle.stages(p, x, 2)
    s = 2 * p.hop_len
    a = float(np.float32(0.97))
    assert np.allclose(fr[:400], x[s:s + 400].astype(float) - a * x[s - 1:s + 399].astype(float))
    fr0, _, _, _ = oracle.stages(p, x, 0)
    assert fr0[0] == float(x[0])


def test_batch_equals_per_utterance_and_threads_agree():
    p = config_b()
    from mfcc_b200.synth import ragged_batch
    pcm, off = ragged_batch(9, 150, 900, seed=3)   # includes utterances shorter than a frame
    out1, fo1 = oracle.mfcc_batch(p, pcm, off, nthreads=1)
    out4, fo4 = oracle.mfcc_batch(p, pcm, off, nthreads=4)
    assert np.array_equal(fo1, fo4) and np.array_equal(out1, out4)
    for u in range(9):
        ref = oracle.mfcc(p, pcm[off[u]:off[u + 1]])
        assert np.array_equal(out1[fo1[u]:fo1[u + 1]], ref)


# ---- §8(f) widening: CMVN, deltas, G.711 ----
def test_cmvn_and_delta_against_numpy():
    rng = np.random.default_rng(0)
    fo = np.array([0, 5, 5, 40, 41], np.int64)
    f = rng.normal(3, 2, (41, 13)).astype(np.float32)
    got = oracle.cmvn(f, fo, True)
    for u in range(4):
        seg = f[fo[u]:fo[u + 1]].astype(np.float64)
        if len(seg) == 0:
            continue
        ref = (seg - seg.mean(0)) / np.sqrt(np.maximum(seg.var(0), 1e-20))
        assert np.allclose(got[fo[u]:fo[u + 1]], ref, atol=1e-5)
    d = oracle.delta(f, fo, 2)
    for u in range(4):
        seg = f[fo[u]:fo[u + 1]].astype(np.float64)
        T = len(seg)
        for t in range(T):
            ref = sum(n * (seg[min(t + n, T - 1)] - seg[max(t - n, 0)]) for n in (1, 2)) / 10.0
            assert np.allclose(d[fo[u] + t], ref, atol=1e-5)


def test_g711_tables_match_audioop():
    import os
    from util import GOLDEN
    t = np.load(os.path.join(GOLDEN, "g711_tables.npz"))
    codes = np.arange(256, dtype=np.uint8)
    assert np.array_equal(oracle.decode_g711(codes, False), t["ulaw"])
    assert np.array_equal(oracle.decode_g711(codes, True), t["alaw"])
    assert oracle.decode_g711(np.array([0xFF, 0x00, 0x80], np.uint8), False).tolist() == [0, -32124, 32124]


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_hostile_golden_vectors(name):
    """The committed hostile fixtures (tests/golden/make_golden.py: digital silence inside speech-level noise, DC
    offset, sigma = 3 noise, full-scale 1 kHz tone, single-sample clicks; numpy float64 results stored as f32):
    inputs regenerate bit for bit, the double oracle agrees to f32 storage precision, the float oracle to the
    stated tolerance (elements that sit at the f32 noise floor are counted, not waved through)."""
    import zlib
    g = hostile_golden()
    p0 = CFG[name]()
    n = p0.frame_len + 40 * p0.hop_len + p0.hop_len // 3
    for kind in HOSTILE_KINDS:
        x = hostile_clip(kind, n, p0.sample_rate)
        assert zlib.crc32(x.tobytes()) == int(g[f"{name}_{kind}_crc"][0])
        for oname, output in (("cep", 0), ("logmel", OUT_LOGMEL)):
            for pname, pad in (("none", 0), ("tail", PAD_ZERO_TAIL)):
                ref = g[f"{name}_{kind}_{oname}_{pname}"]
                p = p0.copy(output=output, pad_mode=pad)
                f64 = oracle.mfcc(p, x, np.float64)
                assert f64.shape == ref.shape == (41 + (pad == PAD_ZERO_TAIL), p.out_dim)
                assert np.abs(f64 - ref).max() <= 4e-6 * max(1.0, np.abs(ref).max())
                assert_parity(oracle.mfcc(p, x, np.float32), f64, what=f"{name}/{kind}/{oname}/{pname}")


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_cpu_baseline_build_matches_the_plain_oracle(name):
    """oracle/mfcc_cpu_fast.c (bench.py's cpu_baseline / --impl reference: real-input FFT through a half-size complex FFT,
    -O3, AVX2 clones) computes the same features as the plain oracle: frame offsets identical, values within the
    stated tolerance, ragged batch with every framing edge, all output options."""
    p0 = CFG[name]()
    L, H = p0.frame_len, p0.hop_len
    lens = [0, 1, L - 1, L, L + 1, L + H, L + 31 * H + 5, 3 * L, L + 64 * H]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = noise_utterance(int(off[-1]), seed=33)
    for kw in ({}, dict(output=OUT_LOGMEL), dict(pad_mode=PAD_ZERO_TAIL, lifter=22), dict(energy=1), dict(energy=2, output=OUT_LOGMEL),
               dict(window=WINDOW_HANN, preemph=0.0)):
        p = p0.copy(**kw)
        ref, fo = oracle.mfcc_batch(p, pcm, off, nthreads=2)
        got, fo2 = oracle.mfcc_batch(p, pcm, off, nthreads=3, fast=True)
        assert np.array_equal(fo, fo2) and got.shape == ref.shape
        truth = np.concatenate([oracle.mfcc(p, pcm[off[u]:off[u + 1]], dtype=np.float64) for u in range(len(lens))])
        from util import lifter_gains
        assert_parity(got, ref, what=f"{name} {kw}", truth=truth, col_scale=lifter_gains(p), max_escapes=2)
