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
    assert C.sizeof(MfccParams) == 15 * 4
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
           dict(preemph=1.5), dict(pad_mode=3), dict(output=2),
           dict(energy=3), dict(energy=-1), dict(energy=1, output=1)]
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


# ---- WAV header parsing ----
def _riff(fmt_body: bytes, data: bytes, extra_chunks=(), data_size=None) -> bytes:
    import struct
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt_body)) + fmt_body
    for cid, payload in extra_chunks:
        body += cid + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"")
    body += b"data" + struct.pack("<I", len(data) if data_size is None else data_size) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def _fmt(tag, channels, rate, bits, extensible_sub=None):
    import struct
    block = channels * bits // 8
    base = struct.pack("<HHIIHH", 0xFFFE if extensible_sub is not None else tag, channels, rate, rate * block, block, bits)
    if extensible_sub is None:
        return base
    guid_tail = bytes.fromhex("000000001000800000aa00389b71")
    return base + struct.pack("<HHI", 22, bits, 0) + struct.pack("<H", extensible_sub) + guid_tail


def test_wav_parse_pcm16_written_by_the_wave_module(tmp_path):
    """An external writer (CPython's wave module) produces the file; the parser must find its samples."""
    import wave
    x = (np.random.default_rng(3).normal(0, 3000, (4000, 2))).astype("<i2")
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(x.tobytes())
    raw = open(path, "rb").read()
    info, s = api.wav_samples(raw)
    assert (info.format, info.channels, info.sample_rate, info.bits_per_sample) == (api.WAV_PCM16, 2, 16000, 16)
    assert info.n_frames == 4000 and info.data_bytes == 16000 and np.array_equal(s, x)


def test_wav_parse_formats_chunks_and_damage():
    codes = bytes(range(256)) * 3 + b"\x7f"                      # odd length: the parser must not need the pad byte
    # G.711 mu-law and A-law (tags 7 and 6), with LIST and odd-sized chunks before the data
    for tag, fmt in ((7, api.WAV_MULAW), (6, api.WAV_ALAW)):
        raw = _riff(_fmt(tag, 1, 8000, 8) + b"\0\0", codes, extra_chunks=[(b"LIST", b"INFOabc"), (b"fact", b"\1\2\3\4")])
        info, s = api.wav_samples(raw)
        assert (info.format, info.channels, info.sample_rate, info.n_frames) == (fmt, 1, 8000, len(codes))
        assert s.tobytes() == codes
        assert np.array_equal(oracle.decode_g711(s[:, 0], fmt == api.WAV_ALAW)[:4],
                              oracle.decode_g711(np.frombuffer(codes, np.uint8), fmt == api.WAV_ALAW)[:4])
    # IEEE float, plain tag 3 and WAVE_FORMAT_EXTENSIBLE with the float sub-format
    f = np.linspace(-1, 1, 300, dtype="<f4")
    for fmt_body in (_fmt(3, 1, 48000, 32), _fmt(0, 1, 48000, 32, extensible_sub=3)):
        info, s = api.wav_samples(_riff(fmt_body, f.tobytes()))
        assert (info.format, info.sample_rate, info.n_frames) == (api.WAV_F32, 48000, 300) and np.array_equal(s[:, 0], f)
    # extensible PCM16
    x = np.arange(-50, 50, dtype="<i2")
    info, s = api.wav_samples(_riff(_fmt(0, 1, 16000, 16, extensible_sub=1), x.tobytes()))
    assert info.format == api.WAV_PCM16 and np.array_equal(s[:, 0], x)
    # streamed (size 0xFFFFFFFF / 0) and truncated data chunks run to the end of the buffer; partial frames are dropped
    for claimed in (0xFFFFFFFF, 0, 10**6):
        info = api.wav_parse(_riff(_fmt(1, 2, 16000, 16), x.tobytes() + b"\x01", data_size=claimed))
        assert info.n_frames == 50 and info.data_bytes == 200
    # unsupported sample formats say so; damage is EINVAL
    for fmt_body in (_fmt(1, 1, 16000, 8), _fmt(1, 1, 16000, 24), _fmt(2, 1, 16000, 4), _fmt(3, 1, 16000, 64)):
        with pytest.raises(api.MfccError) as e:
            api.wav_parse(_riff(fmt_body, b"\0" * 64))
        assert e.value.code == -4
    good = _riff(_fmt(1, 1, 16000, 16), x.tobytes())
    for bad in (b"", b"RIFF", good[:20], b"RIFX" + good[4:], good[:8] + b"AVI " + good[12:], good.replace(b"data", b"dat_"),
                _riff(_fmt(1, 0, 16000, 16), x.tobytes()), b"RIFF\0\0\0\0WAVEdata\4\0\0\0abcd"):
        with pytest.raises(api.MfccError) as e:
            api.wav_parse(bad)
        assert e.value.code == -1, bad[:24]


def test_piece_spans_of_a_long_recording():
    """mfcc_piece_span / sharding.split_stream: pieces cut at frame boundaries, one history sample in front of every piece but
    the first; the frame counts of the pieces (their samples minus the history sample) add up to the recording's, under
    both padding modes, for lengths around every edge."""
    from mfcc_b200 import PAD_ZERO_TAIL, make_params
    from mfcc_b200.sharding import split_stream, frame_counts
    rng = np.random.default_rng(9)
    for pad in (0, PAD_ZERO_TAIL):
        for p in (make_params(pad_mode=pad), make_params(sample_rate=8000, frame_len=200, hop_len=80, nfft=256, n_mel=20, pad_mode=pad)):
            L, H = p.frame_len, p.hop_len
            for n in [0, 1, L - 1, L, L + 1, L + H - 1, L + H, L + 5 * H + 3, 123457] + rng.integers(0, 300000, 6).tolist():
                total = int(frame_counts(p, [0, n])[0])
                for pieces in (1, 2, 3, 7):
                    spans = split_stream(p, n, pieces)
                    assert spans[0][0] == 0 and spans[-1][1] == total
                    for (f0, f1, b, e, lead), nxt in zip(spans, spans[1:] + [None]):
                        assert (b, e, lead) == api.piece_span(p, n, f0, f1)
                        if nxt is not None:
                            assert nxt[0] == f1
                        if f1 == f0:
                            continue
                        assert 0 <= b < e <= n and lead == (1 if f0 > 0 else 0)
                        assert int(frame_counts(p, [0, e - b - lead])[0]) == f1 - f0     # the piece on its own has its frames
                        assert b + lead == f0 * H                                       # and they start where the whole's do
    with pytest.raises(api.MfccError):
        api.piece_span(make_params(), 160000, 5, 2000)
