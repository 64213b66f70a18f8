import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import oracle
from mfcc_b200 import api, config_b
from mfcc_b200.synth import ragged_batch
print(api.LIB_PATH)
for kw in (dict(n_mel=30, energy=1), dict(n_mel=30), dict(n_mel=24, n_cep=12)):
    p = config_b().copy(**kw)
    plan = api.Plan(p)
    L, H = p.frame_len, p.hop_len
    pcm, off = ragged_batch(700, L // 2, L + 150 * H, seed=61)
    b = plan.batch(off)
    got = plan.compute_batch(b, torch.from_numpy(pcm).cuda()).cpu().numpy()
    ref, fo = oracle.mfcc_batch(p, pcm, off, nthreads=8)
    d = np.abs(got - ref); rel = d / np.maximum(np.abs(ref), 1)
    bad = np.argwhere(rel > 1e-4)
    print(kw, plan.kernel_name, "bad", len(bad), "max rel", rel.max())
    if len(bad):
        cols = sorted(set(bad[:, 1].tolist())); print(" cols", cols)
        r, c = bad[0]
        u = np.searchsorted(fo, r, side="right") - 1
        x = pcm[off[u]:off[u + 1]]
        t = oracle.mfcc(p, x, np.float64)
        print(" row", r, "utt", u, "frame", r - fo[u], "of", fo[u + 1] - fo[u], "got", got[r, c], "ref", ref[r, c], "truth", t[r - fo[u], c])
        # log-mel of that frame
        q = p.copy(output=1, energy=0)
        lm_t = oracle.mfcc(q, x, np.float64)[r - fo[u]]; lm_f = oracle.mfcc(q, x)[r - fo[u]]
        pl2 = api.Plan(q); lm_g = pl2.compute(x)[r - fo[u]]
        print(" logmel truth", np.round(lm_t[:6], 4), "\n f32 oracle err", np.abs(lm_f - lm_t).max(), np.argmax(np.abs(lm_f - lm_t)), "\n gpu err", np.abs(lm_g - lm_t).max(), np.argmax(np.abs(lm_g - lm_t)), pl2.kernel_name)
        print(" frame abs max", np.abs(x[(r-fo[u])*H:(r-fo[u])*H+L]).max())
