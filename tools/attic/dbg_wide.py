import numpy as np, torch, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle
from mfcc_b200 import api, config_c, PAD_ZERO_TAIL, KERNEL_GENERIC
from mfcc_b200.synth import noise_utterance
p = config_c().copy(lifter=22, pad_mode=PAD_ZERO_TAIL)
off = np.array([0, 30000, 30007, 61234, 61234, 100001], np.int64)
pcm = noise_utterance(int(off[-1]), seed=23)
ref, fo = oracle.mfcc_batch(p, pcm, off)
truth = np.concatenate([oracle.mfcc(p, pcm[off[u]:off[u + 1]], dtype=np.float64) for u in range(len(off) - 1)])
for kern in (0, KERNEL_GENERIC):
    plan = api.Plan(p, kernel=kern)
    b = plan.batch(off)
    got = plan.compute_batch(b, torch.from_numpy(pcm).cuda()).cpu().numpy()
    err = np.abs(got - truth) / np.maximum(np.abs(truth), 1)
    idx = np.argsort(err.ravel())[::-1][:6]
    print(plan.kernel_name, "frame offsets", fo)
    for i in idx:
        r, k = divmod(int(i), got.shape[1])
        print(f"  row {r} k {k} got {got[r,k]:.6f} ref32 {ref[r,k]:.6f} truth {truth[r,k]:.6f} rel {err[r,k]:.2e}")
    e32 = np.abs(ref - truth) / np.maximum(np.abs(truth), 1)
    print("  oracle f32 vs truth max rel", e32.max())
