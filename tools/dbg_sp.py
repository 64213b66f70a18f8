import numpy as np, torch, sys
sys.path.insert(0, '.')
import oracle
from mfcc_b200 import api, config_a, config_b, KERNEL_FUSED, KERNEL_FUSED_CT
from mfcc_b200.synth import noise_utterance
for name, cfg in (("B", config_b), ("A", config_a)):
    p = cfg()
    L, H = p.frame_len, p.hop_len
    lens = [0, 1, L - 1, L, L + 1, L + H - 1, L + H, L + 31 * H, L + 32 * H, L + 32 * H + 1, L + 100 * H + 7, 3 * L, L + 63 * H, L + 64 * H]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = noise_utterance(int(off[-1]), seed=21)
    ref, fo = oracle.mfcc_batch(p, pcm, off)
    for kern in (KERNEL_FUSED, KERNEL_FUSED_CT):
        plan = api.Plan(p, kernel=kern)
        b = plan.batch(off)
        out = plan.compute_batch(b, torch.from_numpy(pcm).cuda()).cpu().numpy()
        err = np.abs(out - ref)
        bad = np.argwhere(err > 1e-3)
        print(name, plan.kernel_name, "max err", err.max(), "n bad", len(bad), "frames", sorted(set(bad[:, 0].tolist()))[:40], "ceps", sorted(set(bad[:, 1].tolist())))
        print("  frame offsets", fo.tolist())
        if len(bad):
            f = bad[0, 0]
            print("  got", out[f], "\n  ref", ref[f])
