"""CPU restatement (numpy, float64) of the identities the fused kernels rely on, checked against the oracle's DENSE
tables and per-stage outputs.  None of this runs product code: it pins the algebra the CUDA path is built on
(DESIGN.md §5.1), so that a failure of the GPU parity tests can be told apart from a wrong derivation.

  * filterbank: segment j = bins [b_j, b_j+1) rises into filter j and falls out of filter j - 1 with LINEAR weights, so
    with run_i = P_0 + .. + P_i and acc = sum_i run_i = sum_i (w - i) P_i:   fall_j = acc / w,  rise_j = S - fall_j,
    and band m = rise_m + fall_(m+1)                                          (mfcc_fused_sp.cu S3 + tail)
  * DCT: d[k][M-1-m] = (-1)^k d[k][m], so c_k = sum_(q < M/2) d[k][q] (l_q + (-1)^k l_(M-1-q))   (tail warps)
  * two-pass real FFT: N = RB * RA, n = a + RA b, k = k1 + RB k2; rows 1 .. RB/2 - 1 give bins k and, mirrored, N - k
"""
import numpy as np
import pytest

import oracle
from mfcc_b200 import config_a, config_b, config_c, make_params
from mfcc_b200.synth import noise_utterance

CFG = {"A": config_a, "B": config_b, "C": config_c}


def segment_sums(P, bins):
    """rise[j], fall[j] of every segment j = 0 .. M from the power spectrum P, with the kernel's running sums."""
    M = len(bins) - 2
    rise, fall = np.zeros(M + 1), np.zeros(M + 1)
    for j in range(M + 1):
        w = int(bins[j + 1] - bins[j])
        run = acc = 0.0
        for i in range(w):
            run += P[bins[j] + i]
            acc += run
        fall[j] = acc / w if w else 0.0
        rise[j] = run - fall[j]
    return rise, fall


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_running_sum_segments_equal_the_dense_filterbank(name):
    p = CFG[name]()
    x = noise_utterance(p.frame_len + 3 * p.hop_len, seed=5)
    _, pw, mel, _ = oracle.stages(p, x, 2, np.float64)
    bins = oracle.mel_bins(p)
    W = oracle.mel_weights(p)
    rise, fall = segment_sums(pw, bins)
    bands = rise[:-1] + fall[1:]
    np.testing.assert_allclose(bands, W @ pw, rtol=1e-12, atol=1e-12 * pw.max())
    np.testing.assert_allclose(bands, mel, rtol=1e-12, atol=1e-12 * pw.max())
    # the fall of segment 0 and the rise of segment M belong to no band (their weights are not in the dense matrix)
    assert W[:, : bins[0]].sum() == 0.0 and W[:, bins[-1]:].sum() == 0.0


@pytest.mark.parametrize("mel,cep,lifter", [(26, 13, 0), (20, 13, 0), (26, 13, 22), (80, 40, 0), (40, 16, 0)])
def test_dct_mirror_fold(mel, cep, lifter):
    p = make_params(n_mel=mel, n_cep=cep, lifter=lifter)
    D = oracle.dct(p)
    k = np.arange(cep)[:, None]
    np.testing.assert_allclose(D[:, ::-1], (-1.0) ** k * D, rtol=0, atol=1e-12)   # (the lifter scales rows by up to 1 + Q / 2)
    rng = np.random.default_rng(3)
    l = rng.normal(size=mel) * 10.0
    half = mel // 2
    v_even, v_odd = l[:half] + l[::-1][:half], l[:half] - l[::-1][:half]
    c = np.array([D[kk, :half] @ (v_odd if kk & 1 else v_even) for kk in range(cep)])
    np.testing.assert_allclose(c, D @ l, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("rb,ra,L", [(32, 16, 400), (16, 16, 200), (64, 32, 1200)])
def test_two_pass_real_fft_index_algebra(rb, ra, L):
    N, H = rb * ra, rb // 2
    rng = np.random.default_rng(7)
    y = np.zeros(N)
    y[:L] = rng.normal(size=L)
    ref = np.fft.rfft(y)
    a = np.arange(ra)
    cols = y.reshape(rb, ra)                       # cols[b][a] = y[a + RA b]
    Y = np.fft.fft(cols, axis=0)                   # pass 1: DFT-RB over b, rows k1
    X = np.zeros(N // 2 + 1, complex)
    for k1 in range(1, H):                         # complex rows: twiddle, DFT-RA over a
        z = np.fft.fft(Y[k1] * np.exp(-2j * np.pi * a * k1 / N))
        for k2 in range(ra):
            k = k1 + rb * k2
            if k2 < ra // 2:
                X[k] = z[k2]
            else:
                X[N - k] = np.conj(z[k2])
    assert np.abs(Y[0].imag).max() < 1e-12 and np.abs(Y[H].imag).max() < 1e-12   # rows 0 and H are real before the twiddle
    X[0: N // 2 + 1: rb] = np.fft.fft(Y[0].real)[: ra // 2 + 1]                       # row 0: real DFT-RA -> bins RB k2
    zh = np.fft.fft(Y[H].real * np.exp(-2j * np.pi * a / (2 * ra)))                # row H: bins H + RB k2, k2 < RA/2
    X[H: N // 2: rb] = zh[: ra // 2]
    np.testing.assert_allclose(X, ref, rtol=0, atol=1e-9)
