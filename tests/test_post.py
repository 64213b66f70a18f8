"""Fused post-processing (mfcc_post_batch: per-utterance CMVN + delta + delta-delta, stacked) against the oracle.

The oracle side is a composition of oracle_cmvn_f32 and oracle_delta_f32 (oracle.post); the CPU tests pin that
composition against an independent numpy restatement with explicit clamped indexing, the GPU tests compare the CUDA
kernels with it through the C ABI.  Floating point: the statistics are formed in double on both sides, the regression in
double by the oracle and in f32 by the kernel — stated tolerance max |got - ref| <= 1e-5 for features of MFCC magnitude
(|x| < 128, where one f32 ulp is 7.6e-6); every call is also required to be bit-reproducible.
PARITY UNPINNED: the oracle is this repo's own (the reference has no MFCC code)."""
import numpy as np
import pytest

import oracle

POST_TOL = 1e-5


def np_post(feat, fo, cmvn_mode, window, order):
    """Independent restatement: loops over utterances, float64 throughout, f32 rounding where the spec has it."""
    out = np.zeros((feat.shape[0], feat.shape[1] * (1 + order)), np.float32)
    den = 2.0 * sum(n * n for n in range(1, window + 1))
    for u in range(len(fo) - 1):
        seg = feat[fo[u]:fo[u + 1]].astype(np.float64)
        T = len(seg)
        if T == 0:
            continue
        if cmvn_mode:
            seg = seg - seg.mean(0)
            if cmvn_mode == 2:
                seg = seg / np.sqrt(np.maximum((seg ** 2).mean(0), 1e-20))
        parts = [seg.astype(np.float32)]
        for _ in range(order):
            src = parts[-1].astype(np.float64)
            idx = np.arange(T)
            d = sum(n * (src[np.minimum(idx + n, T - 1)] - src[np.maximum(idx - n, 0)]) for n in range(1, window + 1)) / den
            parts.append(d.astype(np.float32))
        out[fo[u]:fo[u + 1]] = np.concatenate(parts, axis=1)
    return out


def fake_features(frames_per_utt, dim, seed=0):
    rng = np.random.default_rng(seed)
    fo = np.concatenate([[0], np.cumsum(frames_per_utt)]).astype(np.int64)
    f = rng.normal(0.0, 4.0, (int(fo[-1]), dim)).astype(np.float32)
    f[:, 0] += 60.0          # c0-like column: large mean
    return f, fo


@pytest.mark.parametrize("cmvn_mode", [0, 1, 2])
@pytest.mark.parametrize("order,window", [(0, 2), (1, 1), (1, 2), (2, 2), (2, 3), (2, 8)])
def test_oracle_post_against_numpy(cmvn_mode, order, window):
    f, fo = fake_features([0, 1, 2, 3, 4, 5, 9, 40, 0, 17], 13, seed=order + 3 * cmvn_mode)
    got = oracle.post(f, fo, cmvn_mode, window, order)
    ref = np_post(f, fo, cmvn_mode, window, order)
    assert got.shape == ref.shape
    assert np.abs(got.astype(np.float64) - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


# ---------------------------------------------------------------- GPU ----
torch = pytest.importorskip("torch")
from mfcc_b200 import api, config_a, config_b, make_params, OUT_LOGMEL, ENERGY_APPEND   # noqa: E402
from mfcc_b200.synth import ragged_batch   # noqa: E402


def frames_to_samples(p, frames):
    """Utterance lengths (samples) that give exactly these frame counts under PAD_NONE."""
    return [0 if n == 0 else p.frame_len + (n - 1) * p.hop_len for n in frames]


def device_batch(p, frames, seed=5, sigma=3000.0):
    rng = np.random.default_rng(seed)
    lens = np.array(frames_to_samples(p, frames), np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = np.clip(np.rint(rng.standard_normal(int(off[-1])) * sigma), -32768, 32767).astype(np.int16)
    plan = api.Plan(p)
    b = plan.batch(off)
    assert list(np.diff(b.frame_offsets)) == list(frames)
    feat = plan.compute_batch(b, torch.from_numpy(pcm).cuda())
    torch.cuda.synchronize()
    return plan, b, feat


# frame counts around every edge of the kernel: no frames, fewer rows than the regression window, the chunk size (256 rows
# of 13 cepstra) and its neighbours, several chunks with a short last one
EDGE_FRAMES = [0, 1, 2, 3, 4, 5, 8, 9, 255, 256, 257, 258, 260, 511, 512, 513, 1031, 0, 7]


@pytest.mark.gpu
@pytest.mark.parametrize("cmvn_mode", [0, 1, 2])
@pytest.mark.parametrize("order,window", [(0, 2), (1, 1), (1, 2), (2, 2), (2, 3), (2, 8)])
def test_post_matches_oracle(cmvn_mode, order, window):
    plan, b, feat = device_batch(config_a(), EDGE_FRAMES)
    f_host = feat.cpu().numpy()
    got = plan.post(b, feat, cmvn_mode, window, order)
    again = plan.post(b, feat, cmvn_mode, window, order)
    torch.cuda.synchronize()
    assert torch.equal(got, again)                        # deterministic: no floating-point atomics, fixed combine order
    assert torch.equal(feat.cpu(), torch.from_numpy(f_host))   # the input is untouched
    ref = oracle.post(f_host, b.frame_offsets, cmvn_mode, window, order)
    g = got.cpu().numpy()
    assert g.shape == ref.shape and np.isfinite(g).all()
    err = np.abs(g.astype(np.float64) - ref)
    assert err.max() <= POST_TOL, (err.max(), np.unravel_index(err.argmax(), err.shape))
    if cmvn_mode == 0:
        assert np.array_equal(g[:, :plan.out_dim], f_host)  # the static part is a copy


@pytest.mark.gpu
@pytest.mark.parametrize("what", ["logmel80", "append14", "wide129", "telephony"])
def test_post_other_row_widths(what):
    """Row widths that change the chunk size (80 columns -> 48 rows per chunk), are not a divisor of anything (14), need two
    column passes in the stacking sweep (129 x 3 = 387 output columns > 256 threads), and the 8 kHz geometry."""
    p = {"logmel80": make_params(n_mel=80, output=OUT_LOGMEL),
         "append14": make_params(energy=ENERGY_APPEND),
         "wide129": make_params(n_mel=128, output=OUT_LOGMEL, energy=ENERGY_APPEND),
         "telephony": config_b()}[what]
    plan, b, feat = device_batch(p, [1, 47, 48, 49, 0, 96, 97, 31, 32, 33, 300, 5])
    f_host = feat.cpu().numpy()
    for cmvn_mode, order, window in ((2, 2, 2), (1, 2, 5), (0, 1, 2), (2, 0, 2)):
        g = plan.post(b, feat, cmvn_mode, window, order).cpu().numpy()
        ref = oracle.post(f_host, b.frame_offsets, cmvn_mode, window, order)
        err = np.abs(g.astype(np.float64) - ref)
        assert g.shape == ref.shape and err.max() <= POST_TOL, (what, cmvn_mode, order, window, err.max())


@pytest.mark.gpu
def test_post_tiny_rows_and_unaligned_matrix():
    """One and two columns per row (chunks of a few bytes: no bulk copy fits, the ring still has to turn), and a feature
    matrix that starts 4 bytes off a 16-byte boundary (no bulk copies at all): same values."""
    for n_cep in (1, 2, 3):
        plan, b, feat = device_batch(make_params(n_cep=n_cep), [1, 2, 3, 0, 5, 300, 1, 1, 1, 1, 1, 2, 700])
        g = plan.post(b, feat, 2, 2, 2).cpu().numpy()
        ref = oracle.post(feat.cpu().numpy(), b.frame_offsets, 2, 2, 2)
        assert np.abs(g.astype(np.float64) - ref).max() <= POST_TOL, n_cep
    plan, b, feat = device_batch(config_a(), EDGE_FRAMES)
    want = plan.post(b, feat, 2, 2, 2)
    shifted = torch.empty(feat.numel() + 1, dtype=torch.float32, device="cuda")[1:].view_as(feat)
    shifted.copy_(feat)
    assert shifted.data_ptr() % 16 == 4
    assert torch.equal(plan.post(b, shifted, 2, 2, 2), want)


@pytest.mark.gpu
def test_post_equals_the_separate_entries_and_hostile_statistics():
    plan, b, feat = device_batch(config_a(), [300, 1, 40, 700])
    stacked = plan.post(b, feat, 2, 2, 2).cpu().numpy()
    x = plan.cmvn(b, feat.clone(), True)
    d1 = plan.delta(b, x, 2)
    d2 = plan.delta(b, d1, 2)
    sep = torch.cat([x, d1, d2], dim=1).cpu().numpy()
    assert np.abs(stacked - sep).max() <= POST_TOL
    # constant rows (zero variance: 1 / sigma is clamped at 1e10, x - mu is exactly 0), a huge common offset (the
    # variance must not be lost to cancellation), and a one-row utterance
    T = int(b.total_frames)
    hostile = torch.zeros((T, plan.out_dim), dtype=torch.float32, device="cuda")
    fo = b.frame_offsets
    hostile[fo[0]:fo[1]] = 7.25
    hostile[fo[2]:fo[3]] = 1.0e4 + torch.randn((int(fo[3] - fo[2]), plan.out_dim), device="cuda") * 0.01
    hostile[fo[3]:fo[4]] = torch.randn((int(fo[4] - fo[3]), plan.out_dim), device="cuda") * 1e-3 - 3.0
    h = hostile.cpu().numpy()
    for mode in (1, 2):
        g = plan.post(b, hostile, mode, 2, 2).cpu().numpy()
        ref = oracle.post(h, fo, mode, 2, 2)
        assert np.isfinite(g).all()
        assert np.abs(g.astype(np.float64) - ref).max() <= POST_TOL * max(1.0, np.abs(ref).max()), mode
        assert np.array_equal(g[fo[0]:fo[1]], np.zeros_like(g[fo[0]:fo[1]]))


@pytest.mark.gpu
def test_host_pipeline_with_post_equals_the_device_path():
    """mfcc_compute_host_post: PCM in host memory in, stacked rows out, the post-processing kernels run per CHUNK of
    utterances inside the copy / compute pipeline.  Same bits as compute_batch + post on the whole batch (the statistics are
    summed in an order that depends on the utterance alone), and within tolerance of the oracle."""
    p = config_a()
    plan = api.Plan(p)
    # 70 MB of PCM: several pipeline chunks (32 MiB each), ragged lengths incl. empty and sub-frame utterances
    pcm, off = ragged_batch(300, 0, 240000, seed=12)
    b = plan.batch(off)
    feat = plan.compute_batch(b, torch.from_numpy(pcm).cuda())
    for cmvn_mode, window, order in ((2, 2, 2), (1, 3, 1), (0, 2, 2), (2, 2, 0)):
        want = plan.post(b, feat, cmvn_mode, window, order).cpu().numpy()
        got, fo = plan.compute_host(pcm, off, post=(cmvn_mode, window, order))
        assert np.array_equal(fo, b.frame_offsets)
        assert got.shape == want.shape and np.array_equal(got, want), (cmvn_mode, window, order)
    ref = oracle.post(feat.cpu().numpy(), b.frame_offsets, 2, 2, 2)
    got, _ = plan.compute_host(pcm, off, post=(2, 2, 2))
    assert np.abs(got.astype(np.float64) - ref).max() <= POST_TOL
    plain, _ = plan.compute_host(pcm, off)                       # the plain entry still returns the plain rows
    assert np.array_equal(plain, feat.cpu().numpy())
    with pytest.raises(ValueError):
        plan.compute_host(pcm, off, post=(1, 2, 3))


@pytest.mark.gpu
def test_post_argument_errors():
    plan, b, feat = device_batch(config_a(), [10, 20])
    lib = api.load()
    out = torch.empty((30, 39), dtype=torch.float32, device="cuda")
    call = lambda f, c, w, o, dst: lib.mfcc_post_batch(plan._h, b._h, f, c, w, o, dst, None)   # noqa: E731
    assert call(feat.data_ptr(), 1, 2, 2, out.data_ptr()) == 0
    assert call(feat.data_ptr(), 3, 2, 2, out.data_ptr()) == api.MFCC_EINVAL
    assert call(feat.data_ptr(), 1, 2, 3, out.data_ptr()) == api.MFCC_EINVAL
    assert call(feat.data_ptr(), 1, 0, 1, out.data_ptr()) == api.MFCC_EINVAL
    assert call(feat.data_ptr(), 1, 9, 1, out.data_ptr()) == api.MFCC_EINVAL
    assert call(feat.data_ptr(), 1, 0, 0, out.data_ptr()) == 0            # the window is ignored without a regression
    assert call(None, 1, 2, 2, out.data_ptr()) == api.MFCC_EINVAL
    assert call(feat.data_ptr(), 1, 2, 2, None) == api.MFCC_EINVAL
    assert call(feat.data_ptr(), 1, 2, 0, feat.data_ptr()) == api.MFCC_EINVAL   # overlapping input and output
    other = api.Plan(config_b())
    assert lib.mfcc_post_batch(other._h, b._h, feat.data_ptr(), 1, 2, 2, out.data_ptr(), None) == api.MFCC_EINVAL
    torch.cuda.synchronize()
    with pytest.raises(ValueError):
        plan.post(b, feat, 1, 2, 5)
    with pytest.raises(ValueError):
        plan.post(b, feat.double(), 1, 2, 2)
