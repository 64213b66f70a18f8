"""Generate the fixtures under tests/golden/ from the numpy/scipy float64 restatement.

    python tests/golden/make_golden.py

The reference (simotin13/mfcc) has no MFCC implementation to import or run
(PARITY UNPINNED), so these vectors come from
tests/np_ref.py — an implementation independent of both the C oracle and the
CUDA path.  Inputs are the BASELINE.md §5 synthetic signals (mfcc_b200/synth.py).
G.711 tables come from CPython's audioop (an external implementation).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import np_ref  # noqa: E402
from mfcc_b200 import config_a, config_b, config_c, OUT_LOGMEL, PAD_ZERO_TAIL  # noqa: E402
from mfcc_b200.synth import clip_config1, noise_utterance, hostile_clip, HOSTILE_KINDS  # noqa: E402


def main():
    out = {}
    # config 1: the 1.0 s 16 kHz clip, 98 frames x 13
    a = config_a()
    x = clip_config1(1.0, 16000, seed=0)
    out["A_pcm"] = x
    out["A_cep"] = np_ref.mfcc(a, x)
    out["A_logmel"] = np_ref.mfcc(a.copy(output=OUT_LOGMEL), x)
    out["A_lifter22"] = np_ref.mfcc(a.copy(lifter=22), x)
    out["A_padtail"] = np_ref.mfcc(a.copy(pad_mode=PAD_ZERO_TAIL), x[:15000])
    # config 3 geometry: 8 kHz, 0.75 s of noise
    b = config_b()
    xb = noise_utterance(6000, seed=3, sigma=3000.0)
    out["B_pcm"] = xb
    out["B_cep"] = np_ref.mfcc(b, xb)
    # config 4 geometry: 48 kHz, 0.25 s of noise
    c = config_c()
    xc = noise_utterance(12000, seed=4000, sigma=3000.0)
    out["C_pcm"] = xc
    out["C_cep"] = np_ref.mfcc(c, xc)
    np.savez_compressed(os.path.join(HERE, "mfcc_golden.npz"), **out)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import audioop
    codes = bytes(range(256))
    ulaw = np.frombuffer(audioop.ulaw2lin(codes, 2), np.int16)
    alaw = np.frombuffer(audioop.alaw2lin(codes, 2), np.int16)
    np.savez_compressed(os.path.join(HERE, "g711_tables.npz"), ulaw=ulaw, alaw=alaw)
    for k, v in out.items():
        print(k, v.shape, v.dtype)

    # Hostile inputs: every geometry x {cepstra, log-mel} x {no padding, zero tail}.  The PCM is
    # regenerated from mfcc_b200.synth.hostile_clip by the tests (checked against the stored CRC), only the float64
    # numpy results are stored, as float32.
    import zlib
    hostile = {}
    for name, p in (("A", a), ("B", b), ("C", c)):
        n = p.frame_len + 40 * p.hop_len + p.hop_len // 3          # 41 frames (42 with the zero tail)
        for kind in HOSTILE_KINDS:
            x = hostile_clip(kind, n, p.sample_rate)
            hostile[f"{name}_{kind}_crc"] = np.array([zlib.crc32(x.tobytes())], np.uint32)
            for oname, output in (("cep", 0), ("logmel", OUT_LOGMEL)):
                for pname, pad in (("none", 0), ("tail", PAD_ZERO_TAIL)):
                    q = p.copy(output=output, pad_mode=pad)
                    hostile[f"{name}_{kind}_{oname}_{pname}"] = np_ref.mfcc(q, x).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "hostile_golden.npz"), **hostile)
    print("hostile:", len(hostile), "arrays")


if __name__ == "__main__":
    main()
