import numpy as np, torch, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle
from mfcc_b200 import api, config_a, KERNEL_FUSED, KERNEL_FUSED_CT, KERNEL_GENERIC, OUT_LOGMEL
from mfcc_b200.synth import noise_utterance
from util import parity_errors
p = config_a().copy(output=OUT_LOGMEL, n_mel=80, n_cep=80)
off = np.array([0, 5000, 5100, 12345, 12345, 20000], np.int64)
pcm = noise_utterance(int(off[-1]), seed=22)
ref, fo = oracle.mfcc_batch(p, pcm, off)
ref64 = np.concatenate([oracle.mfcc(p, pcm[off[u]:off[u+1]], dtype=np.float64) for u in range(len(off)-1)])
for kern in (KERNEL_FUSED, KERNEL_FUSED_CT, KERNEL_GENERIC):
    plan = api.Plan(p, kernel=kern)
    b = plan.batch(off)
    out = plan.compute_batch(b, torch.from_numpy(pcm).cuda()).cpu().numpy()
    a, r = parity_errors(out, ref)
    err = np.abs(out - ref)
    i = np.unravel_index(err.argmax(), err.shape)
    print(plan.kernel_name, "abs", a, "rel", r, "at", i, "got", out[i], "ref", ref[i], "bins", plan.mel_bins()[max(0,i[1]-1):i[1]+4])
    if ref64 is not None:
        a, r = parity_errors(out, ref64.astype(np.float32)); print("   vs f64 oracle: abs", a, "rel", r)
if ref64 is not None:
    print("f32 oracle vs f64 oracle", parity_errors(ref, ref64.astype(np.float32)))
