"""Report fp32-path error metrics of K1 next to the reference's own fp32 path (goldens)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from conftest import golden
from oracle import features as OF
from deep_audio_mixer_b200 import features as F

W = golden('window.npz')['hann_2048']
g = golden('features.npz')

def metrics(x, truth):
    amp_t = 10.0 ** (truth / 20); amp_x = 10.0 ** (x.astype(np.float64) / 20)
    peak = amp_t.max(axis=-2, keepdims=True)
    rel = np.abs(amp_x - amp_t) / peak
    d = np.abs(x - truth)
    near = amp_t >= peak * 1e-2
    return dict(amp_rel_peak_max=rel.max(), amp_rel_peak_rms=np.sqrt((rel**2).mean()),
                db_max=d.max(), db_rms=np.sqrt((d**2).mean()), db_max_within40dB=d[near].max(), db_p999=np.percentile(d, 99.9))

for kind in ('white', 'music', 'tonal'):
    a = g['in_%s_float32' % kind]
    truth = OF.compute_features(a.astype(np.float64), window=W)
    got = F.stft_db(torch.from_numpy(a).cuda()[None], window=torch.from_numpy(W).cuda())[0].cpu().numpy()
    np32 = OF.compute_features(a, window=W)
    for name, x in (('K1 fp32', got), ('reference MKL fp32', g['out_%s_float32' % kind]), ('numpy pocketfft fp32', np32)):
        m = metrics(x, truth)
        print('%-6s %-22s ' % (kind, name) + ' '.join('%s=%.3g' % kv for kv in m.items()))
