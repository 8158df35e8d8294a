"""K1 float32 compute against the reference's OWN float32 outputs (tests/golden/features.npz: out_*_float32 and
out_batch were produced by the unmodified reference, tools/gen_golden.py), bin for bin, plus the float64
compute path against the reference's float64 outputs.  Used by tests/test_stft_gpu.py (asserted bars) and by
bench.py (the ``parity`` block of the JSON line).  Reads only committed fixtures; needs a CUDA device.

    python tools/parity_report.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# Bars asserted on the float32-compute path against the reference's float32 path (torch.stft on MKL).  Two
# float32 FFTs with different operation orders differ by O(eps * ||frame||) in ABSOLUTE terms, so the dB
# difference is unbounded on bins far below the frame peak; the bars are therefore stated (a) in amplitude
# relative to the frame peak on every bin and (b) in dB on every bin within 40 dB of the frame peak.
BARS_F32 = {'amp_rel_peak_max': 1e-6, 'db_max_within40dB': 1e-4}
# float64 compute: north_star's bar on every bin
BAR_F64_DB = 1e-4


def _metrics(got, ref):
    got = got.astype(np.float64)
    ref = ref.astype(np.float64)
    amp_g, amp_r = 10.0 ** (got / 20), 10.0 ** (ref / 20)
    peak = amp_r.max(axis=-2, keepdims=True)                  # per frame
    rel = np.abs(amp_g - amp_r) / peak
    d = np.abs(got - ref)
    near = amp_r >= peak * 1e-2
    return {'amp_rel_peak_max': float(rel.max()), 'db_max_within40dB': float(d[near].max()),
            'db_max_all_bins': float(d.max()), 'db_rms': float(np.sqrt((d ** 2).mean())),
            'db_median': float(np.median(d)), 'frac_bins_within_1e-4dB': float((d <= 1e-4).mean())}


def k1_vs_reference_goldens(device=None):
    import torch
    from deep_audio_mixer_b200 import features as F
    dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
    g = np.load(os.path.join(GOLDEN, 'features.npz'))
    w = torch.from_numpy(np.load(os.path.join(GOLDEN, 'window.npz'))['hann_2048']).to(dev)

    def run(a, **kw):
        return F.stft_db(torch.from_numpy(np.ascontiguousarray(a)).to(dev), window=w, **kw).cpu().numpy()

    rep = {'float32_compute_vs_reference_float32': {}, 'float64_compute_vs_reference_float64': {},
           'reference_float32_vs_reference_float64': {}}
    for kind in ('white', 'music', 'tonal'):
        a32, ref32 = g['in_%s_float32' % kind], g['out_%s_float32' % kind]
        rep['float32_compute_vs_reference_float32'][kind] = _metrics(run(a32), ref32)
        a64, ref64 = g['in_%s_float64' % kind], g['out_%s_float64' % kind]
        rep['float64_compute_vs_reference_float64'][kind] = {'db_max_all_bins': float(np.abs(run(a64) - ref64).max())}
        # the yardstick: how far the reference's own two dtypes are from each other on the same signal
        rep['reference_float32_vs_reference_float64'][kind] = _metrics(ref32, ref64)
    rep['float32_compute_vs_reference_float32']['batch'] = _metrics(run(g['in_batch']), g['out_batch'])
    rep['bars'] = {'float32': BARS_F32, 'float64_db_max_all_bins': BAR_F64_DB,
                   'note': 'float32 bars are relative to the frame peak / within 40 dB of it: no float32 FFT meets '
                           '1e-4 dB on every bin (see reference_float32_vs_reference_float64)'}
    return rep


if __name__ == '__main__':
    print(json.dumps(k1_vs_reference_goldens(), indent=1))
