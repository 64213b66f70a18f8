"""GPU parity tests: the CUDA path, called through the C ABI (ctypes ->
libmfcc_b200.so), against the CPU oracle on the same seeded inputs, against the
committed golden fixtures, and — at BASELINE.json's full sizes — through
size-independent properties.

Tolerance (BASELINE.json north_star): max abs <= 1e-3 and max rel <= 1e-4 on the
cepstra (rel against max(|ref|, 1)); framing (frame counts, row offsets) bit-exact.
PARITY UNPINNED: the oracle is this repo's own (the reference has no MFCC code).
"""
import os

import numpy as np
import pytest

import oracle
from mfcc_b200 import (api, config_a, config_b, config_c, make_params, KERNEL_GENERIC, KERNEL_FUSED,
                       KERNEL_AUTO, OUT_LOGMEL, PAD_ZERO_TAIL, WINDOW_HANN, WINDOW_RECT, ENERGY_APPEND,
                       ENERGY_REPLACE_C0)
from mfcc_b200.synth import (clip_config1, fast_fixed_batch, noise_utterance, ragged_batch, hostile_clip,
                             HOSTILE_KINDS)
from util import assert_parity, golden, hostile_golden, lifter_gains, parity_errors

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CFG = {"A": config_a, "B": config_b, "C": config_c}
# "auto" is what a caller gets: the fused tile kernel of the geometry when there is one, else the generic kernel
KERNELS = {"generic": KERNEL_GENERIC, "auto": KERNEL_AUTO}
ALL_KERNELS = ["generic", "auto"]


def make_plan(p, kernel):
    return api.Plan(p, kernel=KERNELS[kernel])


def run_device(plan, pcm, offsets):
    b = plan.batch(offsets)
    d_pcm = torch.from_numpy(np.ascontiguousarray(pcm)).cuda()
    out = plan.compute_batch(b, d_pcm)
    torch.cuda.synchronize()
    return out.cpu().numpy(), b.frame_offsets.copy()


@pytest.mark.parametrize("kernel", ALL_KERNELS)
@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_ragged_batch_matches_oracle(name, kernel):
    p = CFG[name]()
    plan = make_plan(p, kernel)
    L, H = p.frame_len, p.hop_len
    # lengths hit every framing edge: empty, < L, == L, one hop more, tile boundaries (32, 33 frames), long
    lens = [0, 1, L - 1, L, L + 1, L + H - 1, L + H, L + 31 * H, L + 32 * H, L + 32 * H + 1, L + 100 * H + 7,
            3 * L, L + 63 * H, L + 64 * H]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = noise_utterance(int(off[-1]), seed=21)
    got, fo = run_device(plan, pcm, off)
    ref, fo_ref = oracle.mfcc_batch(p, pcm, off)
    assert np.array_equal(fo, fo_ref)              # framing: bit-exact
    assert got.shape == ref.shape
    assert_parity(got, ref, what=f"{name}/{kernel}")


@pytest.mark.parametrize("name", ["A", "B"])
def test_aligned_ragged_batch_takes_bulk_copy_path(name, kernel="auto"):
    """Utterances whose starts are multiples of 8 samples (16-byte aligned int16): the specialised
    kernel stages these tiles by bulk async copy; mixed with short, tile-boundary and tail tiles.
    Also an utterance that starts right after another one (the 8 lead samples belong to the
    neighbour and must not leak into y[0] = x[0])."""
    p = CFG[name]()
    plan = make_plan(p, kernel)
    L, H = p.frame_len, p.hop_len
    rng = np.random.default_rng(31)
    lens = [8 * int(v) for v in rng.integers(1, 4000, 40)]
    lens += [L + 31 * H, L + 32 * H, L + 33 * H, 0, 8, L + 95 * H + 8, 10 * H * 32 + L]
    lens = [v - v % 8 for v in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = noise_utterance(int(off[-1]), seed=32)
    got, fo = run_device(plan, pcm, off)
    ref, fo_ref = oracle.mfcc_batch(p, pcm, off)
    assert np.array_equal(fo, fo_ref)
    assert_parity(got, ref, what=f"aligned {name}/{kernel}")
    # same data, whole array shifted by one sample: nothing is aligned any more, results identical
    pcm1 = np.concatenate([[0], pcm]).astype(np.int16)
    b = plan.batch(off)
    d = torch.from_numpy(pcm1).cuda()[1:]
    out1 = plan.compute_batch(b, d)
    torch.cuda.synchronize()
    assert np.array_equal(out1.cpu().numpy(), got)


@pytest.mark.parametrize("kernel", ALL_KERNELS)
def test_golden_fixtures(kernel):
    g = golden()
    a, b, c = config_a(), config_b(), config_c()
    cases = [
        ("A_cep", a, g["A_pcm"]), ("A_logmel", a.copy(output=OUT_LOGMEL), g["A_pcm"]),
        ("A_lifter22", a.copy(lifter=22), g["A_pcm"]),
        ("A_padtail", a.copy(pad_mode=PAD_ZERO_TAIL), g["A_pcm"][:15000]),
        ("B_cep", b, g["B_pcm"]), ("C_cep", c, g["C_pcm"]),
    ]
    for key, p, x in cases:
        plan = make_plan(p, kernel)
        got, _ = run_device(plan, x, np.array([0, x.size]))
        assert_parity(got, g[key], what=f"{key}/{kernel}")


@pytest.mark.parametrize("kernel", ALL_KERNELS)
def test_option_matrix_matches_oracle(kernel):
    base = config_a()
    variants = [
        base.copy(window=WINDOW_HANN), base.copy(window=WINDOW_RECT, preemph=0.0),
        base.copy(n_mel=40, n_cep=20), base.copy(n_mel=23, n_cep=13, f_lo=64.0, f_hi=7600.0),
        base.copy(frame_len=512, hop_len=128), base.copy(frame_len=320, hop_len=100, preemph=0.95),
        base.copy(pad_mode=PAD_ZERO_TAIL, lifter=22), base.copy(output=OUT_LOGMEL, n_mel=80, n_cep=80),
        make_params(sample_rate=8000, frame_len=256, hop_len=64, nfft=256, n_mel=24, n_cep=12),
        # cepstral counts around the 8-warp split of the DCT (warp w forms c[w] and c[w + 8])
        base.copy(n_cep=5), base.copy(n_cep=8), base.copy(n_cep=9), base.copy(n_cep=16), base.copy(n_mel=25, n_cep=16),
        config_b().copy(n_cep=1), config_b().copy(n_mel=21, n_cep=13, log_floor=1e-30),
        config_b().copy(pad_mode=PAD_ZERO_TAIL), config_b().copy(pad_mode=PAD_ZERO_TAIL, output=OUT_LOGMEL),
    ]
    off = np.array([0, 5000, 5100, 12345, 12345, 20000], np.int64)
    pcm = noise_utterance(int(off[-1]), seed=22)
    for p in variants:
        plan = make_plan(p, kernel)
        got, fo = run_device(plan, pcm, off)
        ref, fo_ref = oracle.mfcc_batch(p, pcm, off)
        assert np.array_equal(fo, fo_ref)
        truth = np.concatenate([oracle.mfcc(p, pcm[off[u]:off[u + 1]], dtype=np.float64)
                                for u in range(len(off) - 1)]).astype(np.float32)
        # elements allowed to pass on the f64 criterion instead: the lowest band(s) of the 80-filter bank (one or two
        # bins right above DC, 65 dB under the spectrum after pre-emphasis), nothing else
        res = assert_parity(got, ref, what=f"{p.as_dict()}/{kernel}", truth=truth, col_scale=lifter_gains(p),
                            max_escapes=(4 if p.n_mel >= 80 and p.output == OUT_LOGMEL else 0))   # measured: 1 of 9,440
        assert all(c < 2 for c in res.where), res.where


def test_wide_kernel_option_matrix():
    """2048-point geometry (configs[3]): band and cepstrum counts around the tail's pairing (band m with
    M - 1 - m, cepstra slot, slot + 16, slot + 32), log-mel output, lifter, zero-tail padding."""
    base = config_c()
    variants = [base, base.copy(n_mel=41, n_cep=20), base.copy(n_mel=40, n_cep=40), base.copy(n_cep=7),
                base.copy(n_cep=16), base.copy(n_cep=17), base.copy(n_cep=33), base.copy(n_mel=2, n_cep=2),
                base.copy(output=OUT_LOGMEL, n_mel=64), base.copy(lifter=22, pad_mode=PAD_ZERO_TAIL),
                base.copy(n_mel=128, n_cep=40, f_lo=20.0, f_hi=20000.0)]
    off = np.array([0, 30000, 30007, 61234, 61234, 100001], np.int64)
    pcm = noise_utterance(int(off[-1]), seed=23)
    for p in variants:
        plan = api.Plan(p)
        assert plan.kernel_name.startswith("fused_wide"), plan.kernel_name
        got, fo = run_device(plan, pcm, off)
        ref, fo_ref = oracle.mfcc_batch(p, pcm, off)
        assert np.array_equal(fo, fo_ref)
        truth = np.concatenate([oracle.mfcc(p, pcm[off[u]:off[u + 1]], dtype=np.float64)
                                for u in range(len(off) - 1)]).astype(np.float32)
        # the stated tolerance is on plain cepstra; a lifter multiplies cepstrum k and its rounding noise by
        # g_k = 1 + (Q/2) sin(pi k / Q): column k gets g_k times the tolerance (util.lifter_gains), not the largest
        # gain for all.  Elements passing on the f64 criterion are bounded: the 80- and 128-filter banks have one or
        # two filters on the bins right above DC (65 dB under the spectrum after pre-emphasis), which every cepstrum
        # sees through the DCT.
        noisy = p.n_mel >= 64
        assert_parity(got, ref, what=f"{p.as_dict()}", truth=truth, col_scale=lifter_gains(p),
                      max_escapes=(8 if noisy else 0))   # measured: none


def test_known_answers_on_device():
    p = config_a()
    plan = api.Plan(p)
    # all-zero PCM: every band at the floor -> c0 = sqrt(M) ln(floor), c_k = 0
    z = plan.compute(np.zeros(4000, np.int16))
    assert np.allclose(z[:, 0], np.sqrt(26) * np.log(np.float32(1e-10)), atol=1e-4)
    assert np.abs(z[:, 1:]).max() < 1e-4
    # full-scale square wave (clipping extremes) stays finite and matches
    x = np.where(np.arange(8000) % 50 < 25, 32767, -32768).astype(np.int16)
    x = (x.astype(np.int32) + np.random.default_rng(1).integers(-200, 200, x.size)).clip(-32768, 32767).astype(np.int16)
    assert_parity(plan.compute(x), oracle.mfcc(p, x), what="square")


def test_float_pcm_entry_matches_int16():
    p = config_a()
    plan = api.Plan(p)
    pcm, off = ragged_batch(6, 3000, 9000, seed=5)
    b = plan.batch(off)
    d16 = torch.from_numpy(pcm).cuda()
    o16 = plan.compute_batch(b, d16)
    o32 = plan.compute_batch(b, d16.float())
    torch.cuda.synchronize()
    assert torch.equal(o16, o32)


def test_host_entry_points_match_device_path():
    p = config_b()
    plan = api.Plan(p)
    pcm, off = ragged_batch(300, 100, 20000, seed=6)
    dev, fo = run_device(plan, pcm, off)
    host, fo2 = plan.compute_host(pcm, off)
    assert np.array_equal(fo, fo2) and np.array_equal(dev, host)
    one = plan.compute(pcm[off[3]:off[4]])
    assert np.array_equal(one, dev[fo[3]:fo[4]])
    # empty batch and all-too-short batch are fine
    e, foe = plan.compute_host(np.zeros(0, np.int16), np.array([0]))
    assert e.shape == (0, 13) and foe.tolist() == [0]
    e, foe = plan.compute_host(np.zeros(10, np.int16), np.array([0, 5, 10]))
    assert e.shape == (0, 13) and foe.tolist() == [0, 0, 0]


def test_bad_calls_fail_loudly():
    plan = api.Plan(config_a())
    with pytest.raises(api.MfccError):
        plan.batch(np.array([0, 100, 50]))          # decreasing offsets
    with pytest.raises(api.MfccError):
        api.Plan(make_params(nfft=500))
    with pytest.raises(api.MfccError) as e:
        api.Plan(config_a().copy(hop_len=161), kernel=KERNEL_FUSED)   # odd hop: fused unsupported
    assert e.value.code == -4
    b = plan.batch(np.array([0, 16000]))
    with pytest.raises(ValueError):
        plan.compute_batch(b, torch.zeros(100, dtype=torch.int16, device="cuda"))
    assert api.Plan(config_a().copy(hop_len=161)).kernel_name == "generic_radix2"
    assert api.Plan(config_a()).kernel_name.startswith("fused_sp_")
    assert api.Plan(config_b()).kernel_name.startswith("fused_sp_")
    assert api.Plan(config_a().copy(output=OUT_LOGMEL, lifter=22)).kernel_name.startswith("fused_sp_")
    assert api.Plan(config_a().copy(n_mel=40)).kernel_name.startswith("fused_sp_")       # filterbank is data
    assert api.Plan(config_a().copy(n_mel=40, n_cep=20)).kernel_name.startswith("fused_sp_")   # > 16 cepstra: second DCT round
    assert api.Plan(config_a().copy(hop_len=128)).kernel_name == "generic_radix2"
    assert api.Plan(config_c()).kernel_name.startswith("fused_wide_")
    with pytest.raises(api.MfccError):
        api.Plan(config_a(), kernel=3)               # the selector is AUTO / GENERIC / FUSED, nothing else
    # host-buffer entry: offsets past the PCM, wrong dtype / short output
    with pytest.raises(ValueError):
        plan.compute_host(np.zeros(100, np.int16), np.array([0, 16000]))
    with pytest.raises(ValueError):
        plan.compute_host(np.zeros(16000, np.int16), np.array([0, 16000]), out=np.zeros((98, 13), np.float64))
    with pytest.raises(ValueError):
        plan.compute_host(np.zeros(16000, np.int16), np.array([0, 16000]), out=np.zeros((10, 13), np.float32))
    with pytest.raises(ValueError):
        plan.cmvn(b, torch.zeros((98, 13), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        plan.delta(b, torch.zeros((98, 13)))
    # a batch built by a plan with another framing is refused: its tile table would be wrong for this plan
    other = api.Plan(config_a().copy(pad_mode=PAD_ZERO_TAIL))
    with pytest.raises(api.MfccError):
        other.compute_batch(b, torch.zeros(16000, dtype=torch.int16, device="cuda"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """One host thread, the same kernel instantiation on device 0 and then on device 1: the opt-in to > 48 KB of
    dynamic shared memory is per device (it used to be cached per kernel pointer only)."""
    p = config_a()
    pcm, off = ragged_batch(50, 300, 9000, seed=9)
    outs = []
    for dev in (0, 1, 0):
        plan = api.Plan(p, device=dev)
        with torch.cuda.device(dev):
            b = plan.batch(off)
            o = plan.compute_batch(b, torch.from_numpy(pcm).to(f"cuda:{dev}"))
            torch.cuda.synchronize(dev)
            outs.append(o.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_every_start_alignment_takes_the_same_values(name):
    """Utterances may start at any sample of the concatenated array: the bulk-copy staging absorbs the
    shift (first_sample mod 8 = 0..7).  Same clip at 8 alignments -> bit-identical rows, and oracle parity."""
    p = CFG[name]()
    plan = api.Plan(p)
    L, H = p.frame_len, p.hop_len
    clip = noise_utterance(L + 75 * H + 3, seed=77)          # 76 frames: two full tiles and a partial one (A, B)
    ref = oracle.mfcc(p, clip)
    rows = []
    for shift in range(8):
        head = noise_utterance(8 * 50 + shift, seed=78)      # pushes the clip to alignment `shift`
        tail = noise_utterance(1000, seed=79)                # keeps the staged span inside the array
        off = np.cumsum([0, len(head), len(clip), len(tail)]).astype(np.int64)
        batch = np.concatenate([head, clip, tail])
        got, fo = run_device(plan, batch, off)
        mine = got[fo[1]:fo[2]]
        assert mine.shape == ref.shape
        assert_parity(mine, ref, what=f"{name} shift {shift}")
        rows.append(mine)
        got32, _ = run_device(plan, batch.astype(np.float32), off)      # f32 PCM entry: 16-byte vector staging
        assert np.array_equal(got32, got), f"f32 input differs from int16 input at alignment {shift}"
    for shift in range(1, 8):
        assert np.array_equal(rows[shift], rows[0]), f"alignment {shift} changed the values"


@pytest.mark.parametrize("kernel", ALL_KERNELS)
@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_hostile_fixtures(name, kernel):
    """Committed hostile inputs (tests/golden/hostile_golden.npz, numpy float64 results): digital silence inside
    speech-level noise, DC offset, sigma = 3 LSB noise, a full-scale 1 kHz tone, single-sample clicks; cepstra and
    log-mel, with and without the zero-padded tail.  Stated tolerance against the stored truth, nothing waved through."""
    g = hostile_golden()
    p0 = CFG[name]()
    n = p0.frame_len + 40 * p0.hop_len + p0.hop_len // 3
    for kind in HOSTILE_KINDS:
        x = hostile_clip(kind, n, p0.sample_rate)
        # as an utterance in the middle of a batch too (bulk-copy staging; neighbours are loud noise)
        head, tail = noise_utterance(1003, seed=91), noise_utterance(2000, seed=92)
        off = np.cumsum([0, head.size, x.size, tail.size]).astype(np.int64)
        batch = np.concatenate([head, x, tail])
        for oname, output in (("cep", 0), ("logmel", OUT_LOGMEL)):
            for pname, pad in (("none", 0), ("tail", PAD_ZERO_TAIL)):
                p = p0.copy(output=output, pad_mode=pad)
                plan = make_plan(p, kernel)
                ref = g[f"{name}_{kind}_{oname}_{pname}"]
                got, fo = run_device(plan, batch, off)
                assert_parity(got[fo[1]:fo[2]], ref, what=f"hostile {name}/{kind}/{oname}/{pname}/{kernel}")
                alone, _ = run_device(plan, x, np.array([0, x.size], np.int64))
                assert np.array_equal(alone, got[fo[1]:fo[2]])


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_bin_centred_tone_known_answer(name):
    """SURVEY.md 8c known answer 3: a sinusoid exactly on bin k0 of a rectangular window without pre-emphasis.
    (1) frame = NFFT (generic kernel): the energy sits in the <= 2 filters covering k0, every other band holds only
    the int16 quantisation noise, > 60 dB down.  (2) the fused kernel's own frame length (L < NFFT, so the tone leaks
    at the -13 dB level of the zero-padded rectangular window): tones on filter peaks and edges are where a
    filterbank that forms one half of a triangle by subtraction (mfcc_fused_sp.cu S3: rise = S / N - fall) loses the
    most — the bound is eps * w * P_peak / P_neighbour ~ 3e-5 relative for L / NFFT = 400 / 512 — so
    every band must still match the dense-table oracle within the stated tolerance."""
    base = CFG[name]()
    N, M = base.nfft, base.n_mel
    for L in (N, base.frame_len):
        p = base.copy(frame_len=L, window=WINDOW_RECT, preemph=0.0, output=OUT_LOGMEL)
        plan = api.Plan(p)
        assert plan.kernel_name.startswith("fused_") == (L != N)
        bins = oracle.mel_bins(p)
        ks = sorted({int(bins[j]) for j in (1, 2, M // 2, M // 2 + 1, M - 1, M)} | {int(bins[M // 2]) + 1})
        for k0 in ks:
            if k0 <= 0 or k0 >= N // 2:
                continue
            nsamp = L + 40 * p.hop_len
            x = np.rint(30000.0 * np.cos(2 * np.pi * k0 * np.arange(nsamp) / N + 0.3)).astype(np.int16)
            got = plan.compute(x)
            ref = oracle.mfcc(p, x)
            truth = oracle.mfcc(p, x, np.float64)
            what = f"tone {name} L={L} k0={k0} kernel={plan.kernel_name}"
            if L == N:
                # the quiet bands sit at the f32 FFT noise floor (the f32 oracle is 0.01 .. 0.2 off the f64 one there):
                # parity on the loud bands, a level check on the quiet ones
                loud = [m for m in range(M) if bins[m] < k0 < bins[m + 2]]
                quiet = [m for m in range(M) if m not in loud]
                assert_parity(got[:, loud], ref[:, loud], what=what)
                assert got[:, loud].min() > got[:, quiet].max() + np.log(1e6), (k0, loud)
            else:
                assert_parity(got, ref, what=what, truth=truth)     # no element may need the f64 criterion


# ---- BASELINE.json full sizes, through size-independent properties ----
def test_config2_full_size_properties():
    """1,024 x 10 s @ 16 kHz: frame rows exact, fused == generic
    within tolerance everywhere, oracle parity on a sample of utterances, and
    shift-by-one-hop equivariance (tile decomposition independence), bit-exact."""
    p = config_a()
    pcm, off = fast_fixed_batch(1024, 160000, seed=1000)
    plan = api.Plan(p, kernel=KERNEL_AUTO)
    b = plan.batch(off)
    assert b.total_frames == 1_021_952 and np.array_equal(np.diff(b.frame_offsets), np.full(1024, 998))
    d_pcm = torch.from_numpy(pcm).cuda()
    out = plan.compute_batch(b, d_pcm)
    gen = api.Plan(p, kernel=KERNEL_GENERIC)
    out_g = gen.compute_batch(gen.batch(off), d_pcm)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    a, r = parity_errors(out.cpu().numpy(), out_g.cpu().numpy())
    assert a <= 1e-3 and r <= 1e-4, (a, r)
    o = out.cpu().numpy()
    # every one of the 1,021,952 rows against the oracle (all host cores: a few seconds)
    ref, fo_ref = oracle.mfcc_batch(p, pcm, off, nthreads=os.cpu_count() or 8)
    assert np.array_equal(b.frame_offsets, fo_ref)
    assert_parity(o, ref, what="configs[1], all rows")
    # drop the first hop of utterance 7: frame t of the shifted clip == frame t+1 of the original (t >= 1;
    # t = 0 differs through the pre-emphasis boundary y[0] = x[0])
    x = pcm[off[7]:off[8]]
    s = plan.compute(x[p.hop_len:])
    full = o[b.frame_offsets[7]:b.frame_offsets[8]]
    assert np.array_equal(s[1:], full[2:])


def test_config3_ragged_telephony_batch():
    """8 kHz, 16,384 short utterances of 0.5-3 s in ONE launch (full size: 2.8 M frames),
    every row against the oracle."""
    p = config_b()
    pcm, off = ragged_batch(16384, 4000, 24000, seed=3)
    plan = api.Plan(p)
    n0 = api.launch_count()
    got, fo = run_device(plan, pcm, off)
    assert api.launch_count() - n0 == 1
    ref, fo_ref = oracle.mfcc_batch(p, pcm, off, nthreads=os.cpu_count() or 8)
    assert np.array_equal(fo, fo_ref)
    assert_parity(got, ref, what="config3")


def test_config4_long_stream():
    """48 kHz long-form stream, 20 s here: many tiles of one utterance."""
    p = config_c()
    x = noise_utterance(48000 * 20, seed=4000)
    plan = api.Plan(p)
    got = plan.compute(x)
    ref = oracle.mfcc(p, x)
    assert got.shape == ref.shape == (1998, 40)
    assert_parity(got, ref, what="config4")


def test_config4_full_length_stream_properties():
    """One full configs[3] stream: 10 min @ 48 kHz = 28.8 M samples, 59,998 frames, 7,500 work units of one utterance.
    Frame count exact; windows of the result equal the oracle on the matching slice (the slice starts one hop early
    and its first frame is dropped: it alone sees y[0] = x[0]); and the stream cut in two overlapping halves gives
    the same rows bit for bit (tile decomposition independence)."""
    p = config_c()
    L, H = p.frame_len, p.hop_len
    n = 28_800_000
    x = fast_fixed_batch(1, n, seed=4000)[0]
    plan = api.Plan(p)
    off = np.array([0, n], np.int64)
    got, fo = run_device(plan, x, off)
    assert got.shape == (59_998, 40) and fo[-1] == 1 + (n - L) // H
    assert np.isfinite(got).all()
    for t0 in (1, 29_990, 59_900):
        cnt = 64
        s0 = (t0 - 1) * H
        ref = oracle.mfcc(p, x[s0:s0 + (cnt) * H + L])      # frames t0-1 .. t0+cnt-1 of the stream
        assert_parity(got[t0:t0 + cnt], ref[1:1 + cnt], what=f"stream window at frame {t0}")
    # second half as its own utterance, starting one hop before frame 30,000
    t1 = 30_000
    g2, _ = run_device(plan, x[(t1 - 1) * H:], np.array([0, n - (t1 - 1) * H], np.int64))
    assert np.array_equal(g2[1:], got[t1:])


# ---- §8(f) widening ----
@pytest.mark.parametrize("kernel", ALL_KERNELS)
def test_energy_term_and_front_end_shapes(kernel):
    """append-energy / c0 replacement on every geometry and output, and the filter / cepstrum
    counts that have a tail-warp variant of the 512- / 256-point kernel (26/13, 40/13, 23/13 — an odd bank —, log-mel
    26 / 40 / 80; 8 kHz: 20/13, 23/13, log-mel 20 / 40) next to ones that take the generic tail."""
    a, b, c = config_a(), config_b(), config_c()
    variants = [
        a.copy(energy=ENERGY_REPLACE_C0), a.copy(energy=ENERGY_APPEND), a.copy(energy=ENERGY_APPEND, output=OUT_LOGMEL),
        a.copy(energy=ENERGY_REPLACE_C0, lifter=22, pad_mode=PAD_ZERO_TAIL), a.copy(energy=ENERGY_APPEND, f_lo=300.0, f_hi=3400.0),
        a.copy(n_mel=40), a.copy(n_mel=23, f_lo=20.0, f_hi=7800.0), a.copy(n_mel=23, energy=ENERGY_APPEND),
        a.copy(output=OUT_LOGMEL), a.copy(output=OUT_LOGMEL, n_mel=40), a.copy(output=OUT_LOGMEL, n_mel=80, n_cep=80, f_lo=40.0),
        a.copy(output=OUT_LOGMEL, n_mel=80, n_cep=80, f_lo=40.0, energy=ENERGY_APPEND),
        a.copy(n_mel=30, n_cep=12, energy=ENERGY_REPLACE_C0), a.copy(n_mel=31, output=OUT_LOGMEL, energy=ENERGY_APPEND),
        b.copy(energy=ENERGY_REPLACE_C0), b.copy(energy=ENERGY_APPEND), b.copy(n_mel=23), b.copy(output=OUT_LOGMEL),
        b.copy(output=OUT_LOGMEL, n_mel=40, n_cep=40, energy=ENERGY_APPEND), b.copy(n_mel=23, f_lo=100.0, f_hi=3800.0, energy=ENERGY_REPLACE_C0),
        c.copy(energy=ENERGY_REPLACE_C0), c.copy(energy=ENERGY_APPEND), c.copy(energy=ENERGY_APPEND, output=OUT_LOGMEL),
        c.copy(energy=ENERGY_APPEND, n_mel=41, n_cep=13, f_lo=60.0, f_hi=20000.0),
    ]
    for p in variants:
        plan = make_plan(p, kernel)
        L, H = p.frame_len, p.hop_len
        lens = [L + 70 * H + 3, 5, L, L + 31 * H, 0, L + 33 * H + 1]
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        pcm = noise_utterance(int(off[-1]), seed=24)
        got, fo = run_device(plan, pcm, off)
        ref, fo_ref = oracle.mfcc_batch(p, pcm, off)
        assert np.array_equal(fo, fo_ref) and got.shape == ref.shape == (fo[-1], p.out_dim)
        truth = np.concatenate([oracle.mfcc(p, pcm[off[u]:off[u + 1]], dtype=np.float64) for u in range(len(off) - 1)])
        assert_parity(got, ref, what=f"{p.as_dict()}/{kernel}/{plan.kernel_name}", truth=truth, col_scale=lifter_gains(p),
                      max_escapes=(4 if p.n_mel >= 80 and p.f_lo < 20 else 0))


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_g711_codes_are_expanded_inside_the_kernel(name):
    """mfcc_compute_batch_g711 / mfcc_compute_host_g711: mu-law and A-law bytes in, the same rows as decoding to
    int16 first (mfcc_decode_g711, itself bit-exact against the ITU tables) and calling mfcc_compute_batch — bit for
    bit, for every start alignment (the 1-byte codes are bulk-copied in 16-sample units)."""
    p = CFG[name]()
    if name == "C":
        p = p.copy(f_lo=60.0)     # two kernels are compared here: keep the bank off the noise-floor bins right above DC
    plan = api.Plan(p)
    L, H = p.frame_len, p.hop_len
    rng = np.random.default_rng(15)
    lens = [int(v) for v in rng.integers(1, 40 * H, 60)] + [L + 64 * H, 0, 3, L, L + 31 * H + 1]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    codes = rng.integers(0, 256, int(off[-1])).astype(np.uint8)
    b = plan.batch(off)
    for alaw in (False, True):
        lin = oracle.decode_g711(codes, alaw)
        want = plan.compute_batch(b, torch.from_numpy(lin).cuda())
        got = plan.compute_batch(b, torch.from_numpy(codes).cuda(), alaw=alaw)
        torch.cuda.synchronize()
        if plan.kernel_name.startswith("fused_sp"):
            assert torch.equal(got, want)            # same kernel, the codes expanded in its staging
        else:                                        # 2048-point plans take the generic kernel for G.711 codes
            assert_parity(got.cpu().numpy(), want.cpu().numpy(), what=f"g711 {name} vs int16 entry")
        ref, _ = oracle.mfcc_batch(p, lin, off)
        assert_parity(got.cpu().numpy(), ref, what=f"g711 {name} alaw={alaw}")
        host, fo = plan.compute_host(codes, off, alaw=alaw)
        assert np.array_equal(host, got.cpu().numpy()) and np.array_equal(fo, b.frame_offsets)
        # a device view that starts at an odd byte: no tile may take the bulk-copy path, same values
        shifted = torch.from_numpy(np.concatenate([[0], codes]).astype(np.uint8)).cuda()[1:]
        assert torch.equal(plan.compute_batch(b, shifted, alaw=alaw), got)



def test_cmvn_delta_g711_on_device():
    p = config_a()
    plan = api.Plan(p)
    pcm, off = ragged_batch(40, 300, 30000, seed=8)
    b = plan.batch(off)
    feat = plan.compute_batch(b, torch.from_numpy(pcm).cuda())
    torch.cuda.synchronize()
    f_host = feat.cpu().numpy()
    d = plan.delta(b, feat, 2)
    d2 = plan.delta(b, d, 2)
    ref_d = oracle.delta(f_host, b.frame_offsets, 2)
    assert np.abs(d.cpu().numpy() - ref_d).max() < 1e-5
    assert np.abs(d2.cpu().numpy() - oracle.delta(ref_d, b.frame_offsets, 2)).max() < 1e-5
    for nv in (False, True):
        g = plan.cmvn(b, feat.clone(), nv)
        assert np.abs(g.cpu().numpy() - oracle.cmvn(f_host, b.frame_offsets, nv)).max() < 2e-5
    codes = np.random.default_rng(2).integers(0, 256, 100003).astype(np.uint8)
    for alaw in (False, True):
        for view in (codes, codes[1:]):   # aligned and misaligned starts
            got = api.decode_g711(torch.from_numpy(view.copy()).cuda()[...], alaw)
            assert np.array_equal(got.cpu().numpy(), oracle.decode_g711(view, alaw))   # bit-exact


def test_streaming_feed_is_bit_identical_to_offline():
    """mfcc_stream_*: any chunking of a clip gives the rows of the one-shot call, bit for bit;
    zero-tail frames appear at flush; the object is reusable after flush."""
    rng = np.random.default_rng(41)
    for p in (config_a(), config_b(), config_a().copy(pad_mode=PAD_ZERO_TAIL), config_c(),
              config_a().copy(hop_len=450)):      # hop > frame: gaps between frames
        plan = api.Plan(p)
        x = noise_utterance(5 * p.sample_rate // 2 + 123, seed=42)
        full = plan.compute(x)
        st = api.Stream(plan)
        for trial in range(2):
            rows, pos = [], 0
            while pos < x.size:
                n = int(rng.choice([1, 7, p.hop_len - 1, p.hop_len, p.frame_len, 1000, 4096, 40000]))
                rows.append(st.feed(x[pos:pos + n]))
                pos += n
            rows.append(st.flush())
            got = np.concatenate(rows)
            assert got.shape == full.shape, (got.shape, full.shape, p.pad_mode)
            assert np.array_equal(got, full)
        # a stream shorter than one frame: nothing under PAD_NONE, one padded frame under ZERO_TAIL
        assert st.feed(x[:10]).shape[0] == 0
        tail = st.flush()
        assert tail.shape[0] == (1 if p.pad_mode == PAD_ZERO_TAIL else 0)
        if tail.shape[0]:
            assert np.array_equal(tail, plan.compute(x[:10]))


def test_feed_many_streams_in_one_launch():
    """mfcc_stream_feed_many: S live streams, ragged chunk sizes, ONE launch per call; every stream's rows equal the
    offline rows of its own audio, bit for bit, and equal feeding the streams one at a time."""
    rng = np.random.default_rng(51)
    for p in (config_b(), config_a().copy(pad_mode=PAD_ZERO_TAIL), config_c()):
        plan = api.Plan(p)
        S = 37
        clips = [noise_utterance(int(rng.integers(p.frame_len // 2, 4 * p.sample_rate // 10)), seed=500 + i) for i in range(S)]
        want = [plan.compute(c) for c in clips]
        grp = api.StreamGroup(plan, S, max_frames_per_feed=64)
        pos = [0] * S
        got = [[] for _ in range(S)]
        calls = 0
        while any(pos[i] < clips[i].size for i in range(S)):
            chunks = []
            for i in range(S):
                n = int(rng.choice([0, 1, p.hop_len - 1, p.hop_len, 2 * p.hop_len, p.frame_len + 3, 20 * p.hop_len]))
                chunks.append(clips[i][pos[i]:pos[i] + n])
                pos[i] += n
            l0 = api.launch_count()
            rows, counts = grp.feed(chunks)
            assert api.launch_count() - l0 <= 1
            calls += 1
            for i in range(S):
                got[i].append(rows[i, :counts[i]].copy())
        for i in range(S):
            tail = grp.streams[i].flush()
            g = np.concatenate(got[i] + [tail])
            assert g.shape == want[i].shape, (i, g.shape, want[i].shape)
            assert np.array_equal(g, want[i])
        # uniform chunks as one [S, n] array (the serving shape: 20 ms per stream per call)
        n = 2 * p.hop_len
        audio = np.stack([noise_utterance(10 * n + p.frame_len, seed=900 + i) for i in range(S)])
        ref = [plan.compute(audio[i]) for i in range(S)]
        acc = [[] for _ in range(S)]
        for k in range(audio.shape[1] // n):
            rows, counts = grp.feed(np.ascontiguousarray(audio[:, k * n:(k + 1) * n]))
            for i in range(S):
                acc[i].append(rows[i, :counts[i]].copy())
        for i in range(S):
            g = np.concatenate(acc[i])
            assert np.array_equal(g, ref[i][:g.shape[0]]) and g.shape[0] >= ref[i].shape[0] - 3
        # all-or-nothing: a capacity that is too small for one stream feeds none of them
        grp2 = api.StreamGroup(plan, 2, max_frames_per_feed=1)
        with pytest.raises(api.MfccError):
            grp2.feed([clips[0][:p.hop_len], noise_utterance(p.frame_len + 5 * p.hop_len, seed=1)])
        rows, counts = grp2.feed([clips[0][:p.frame_len], np.zeros(0, np.int16)])
        assert counts.tolist() == [1, 0] and np.array_equal(rows[0, :1], plan.compute(clips[0][:p.frame_len]))


def test_torch_custom_op_and_dlpack_entry():
    """torch.ops.mfcc_b200.compute_batch: runs on the CURRENT stream (a side stream here), passes torch's op checks
    (schema, fake tensor shape), traces through torch.compile as an opaque call, and takes DLPack capsules."""
    import mfcc_b200.torch_ops as ops
    p = config_a()
    pcm, off = ragged_batch(64, 300, 30000, seed=71)
    h = ops.Handle(p, off)
    d = torch.from_numpy(pcm).cuda()
    want = h.plan.compute_batch(h.batch, d)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        got = torch.ops.mfcc_b200.compute_batch(d, h.id)
    side.synchronize()
    assert torch.equal(got, want) and got.shape == h.shape
    torch.library.opcheck(torch.ops.mfcc_b200.compute_batch.default, (d, h.id), test_utils=("test_schema", "test_faketensor"))

    @torch.compile(fullgraph=True)
    def front_end(x):
        return torch.ops.mfcc_b200.compute_batch(x, h.id) * 2.0
    assert torch.equal(front_end(d), want * 2.0)
    # the post-processing op on the same handle: stacked rows, same bits as the plan's own call, traceable
    full = torch.ops.mfcc_b200.post(want, h.id, 2, 2, 2)
    assert full.shape == (h.shape[0], 3 * h.shape[1]) and torch.equal(full, h.plan.post(h.batch, want, 2, 2, 2))
    torch.library.opcheck(torch.ops.mfcc_b200.post.default, (want, h.id, 2, 2, 2), test_utils=("test_schema", "test_faketensor"))

    @torch.compile(fullgraph=True)
    def front_end39(x):
        return torch.ops.mfcc_b200.post(torch.ops.mfcc_b200.compute_batch(x, h.id), h.id, 2, 2, 2)
    assert torch.equal(front_end39(d), full)
    # DLPack in (a capsule from another owner of the memory), DLPack out
    cap = torch.utils.dlpack.to_dlpack(d.float())
    class Foreign:          # an object that only speaks the protocol
        def __init__(self, t): self.t = t
        def __dlpack__(self, stream=None): return self.t.__dlpack__(stream=stream)
        def __dlpack_device__(self): return self.t.__dlpack_device__()
    out = ops.mfcc_from_dlpack(Foreign(d.float()), h)
    assert torch.equal(out, want)
    back = torch.from_dlpack(out)
    assert back.data_ptr() == out.data_ptr()
    del cap
    h.close()


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_pieces_of_a_long_recording_equal_the_whole(name):
    """One long recording cut into pieces at frame boundaries (sharding.split_stream / mfcc_piece_span), every piece but the
    first carrying one history sample (mfcc_batch_create_lead): the rows of the pieces, one after the other, are the rows
    of the whole recording bit for bit — what lets one stream be spread over several GPUs."""
    from mfcc_b200.sharding import split_stream
    for pad in (0, PAD_ZERO_TAIL):
        p = CFG[name]()
        p.pad_mode = pad
        plan = api.Plan(p)
        n = p.sample_rate * 20 + 137                       # 20 s and a ragged end
        x = noise_utterance(n, seed=31)
        whole, _ = run_device(plan, x, np.array([0, n]))
        for pieces in (2, 5):
            spans = split_stream(p, n, pieces)
            parts = [x[b:e] for _, _, b, e, _ in spans]
            off = np.concatenate([[0], np.cumsum([len(q) for q in parts])]).astype(np.int64)
            b = plan.batch(off, lead=[s[4] for s in spans])
            assert list(np.diff(b.frame_offsets)) == [f1 - f0 for f0, f1, _, _, _ in spans]
            got = plan.compute_batch(b, torch.from_numpy(np.concatenate(parts)).cuda())
            torch.cuda.synchronize()
            assert np.array_equal(got.cpu().numpy(), whole), (name, pad, pieces)
        # without the history sample the first row of every later piece differs (y[0] = x[0] instead of x[0] - a x[-1])
        spans = split_stream(p, n, 2)
        parts = [x[b + l:e] for _, _, b, e, l in spans]
        off = np.concatenate([[0], np.cumsum([len(q) for q in parts])]).astype(np.int64)
        cut, _ = run_device(plan, np.concatenate(parts), off)
        f1 = spans[0][1]
        assert np.array_equal(cut[:f1], whole[:f1]) and not np.array_equal(cut[f1], whole[f1])
        assert np.array_equal(cut[f1 + 1:], whole[f1 + 1:])


def test_c_caller_runs_the_device_path(tmp_path):
    """The plain C99 demo (tests/cabi/demo.c) through mfcc_compute on the GPU."""
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
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "mfcc_compute: ok, 98 frames" in r.stdout and "fused_sp" in r.stdout
    assert "mfcc_compute_host_post: ok, 98 frames x 39" in r.stdout


POISON_SCRIPT = r"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import oracle
from util import assert_parity
from mfcc_b200 import api, config_a, config_b, config_c, OUT_LOGMEL, PAD_ZERO_TAIL
from mfcc_b200.synth import ragged_batch, noise_utterance
assert api.LIB_PATH.endswith("libmfcc_b200_poison.so"), api.LIB_PATH
worst = {}
# (48 kHz cases: f_lo = 60 Hz keeps the bank off the one or two bins right above DC, which sit at the f32 noise floor)
cases = [("A", config_a()), ("B", config_b()), ("C", config_c().copy(f_lo=60.0)), ("A_logmel", config_a().copy(output=OUT_LOGMEL)),
         ("B_24_12", config_b().copy(n_mel=24, n_cep=12)), ("A_tail", config_a().copy(pad_mode=PAD_ZERO_TAIL)),
         ("A_energy", config_a().copy(energy=2)), ("B_30_energy", config_b().copy(n_mel=30, energy=1, f_lo=100.0)),
         ("C_logmel", config_c().copy(output=OUT_LOGMEL, f_lo=60.0, energy=2))]
for name, p in cases:
    plan = api.Plan(p)
    assert plan.kernel_name.startswith("fused_"), plan.kernel_name
    L, H = p.frame_len, p.hop_len
    pcm, off = ragged_batch(700 if p.nfft < 2048 else 120, L // 2, L + 150 * H, seed=61)
    b = plan.batch(off)
    for dt in (torch.int16, torch.float32, torch.uint8):
        if dt == torch.uint8:     # G.711 entry: the PCM's low bytes as mu-law codes, against their own expansion
            if not plan.kernel_name.startswith("fused_sp"):
                continue
            codes = pcm.view(np.uint8)[::2].copy()
            out = plan.compute_batch(b, torch.from_numpy(codes).cuda())
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            ref, _ = oracle.mfcc_batch(p, oracle.decode_g711(codes, False), off, nthreads=8)
            assert np.isfinite(got).all()
            assert_parity(got, ref, what=name + "/g711")
            continue
        out = plan.compute_batch(b, torch.from_numpy(pcm).cuda().to(dt))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert np.isfinite(got).all(), f"{name}: NaN from a poisoned buffer reached the output"
        ref, _ = oracle.mfcc_batch(p, pcm, off, nthreads=8)
        truth = None
        if p.nfft == 2048:   # 80 bands at 48 kHz: the band on the bins right above DC sits at the f32 noise floor
            truth = np.concatenate([oracle.mfcc(p, pcm[off[u]:off[u + 1]], dtype=np.float64) for u in range(len(off) - 1)])
        res = assert_parity(got, ref, what=name, truth=truth, max_escapes=got.size // 10000)
        worst[name] = [res[0], res[1], res.escapes]
print(json.dumps(worst))
"""


def test_poison_build_finds_no_stale_reads():
    """The sanitizer substitute (compute-sanitizer is refused on the pool's boxes): libmfcc_b200_poison.so is the same
    library built with -DMFCC_POISON=1, which NaN-fills every aliased shared-memory buffer (staged / P, workspace /
    tail scratch, raw PCM) at the point where the kernels' barrier reasoning says it is dead.  A read of a dead or
    not-yet-written value then reaches the output as NaN.  Ragged batches through both fused kernels, int16 and f32
    entries, tail-warp and generic-tail variants: finite and within tolerance of the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "mfcc_b200", "libmfcc_b200_poison.so")
    assert os.path.exists(lib), "poison build missing: make -C mfcc_b200/csrc poison (done by __graft_entry__.build())"
    env = dict(os.environ, MFCC_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", POISON_SCRIPT], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    print("poison build:", r.stdout.strip())
