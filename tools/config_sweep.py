"""Measure the hot path on the five BASELINE.json configurations (device-resident, CUDA events) and
write profiles/r01_config_sweep.json.  Not the contract benchmark (that is bench.py, config 2);
this is the per-config table DESIGN.md section 5 quotes.

    python tools/config_sweep.py [out.json]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deep_audio_mixer_b200 import features as F, mixing, models, synthetic, inference_utils as IU  # noqa: E402
from deep_audio_mixer_b200.dataset import Featuriser  # noqa: E402
from deep_audio_mixer_b200.loudness import Meter  # noqa: E402

PEAK = 6549.8
SR = 44100


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def k1_case(rows, n, hop, dtype=torch.float32, compute=None, channels=1, label='', interleaved=False):
    shape = ((rows, n, channels) if interleaved else (rows, channels, n)) if channels > 1 else (rows, n)
    if dtype == torch.int16:
        x = (torch.randn(shape, device='cuda') * 3000).clamp(-32768, 32767).to(torch.int16)
    else:
        x = (torch.randn(shape, device='cuda') * 0.1).to(dtype)
    compute = compute or (torch.float64 if dtype == torch.float64 else torch.float32)
    T = 1 + n // hop
    out = torch.empty(rows, 1025, T, device='cuda', dtype=compute)
    kw = dict(hop=hop, compute_dtype=compute, out_dtype=compute, out=out)
    if channels > 1:
        kw['channel_axis'] = -1 if interleaved else -2
    ms = timed(lambda: F.stft_db(x, **kw))
    byts = x.numel() * x.element_size() + out.numel() * out.element_size()
    stem_s = rows * n / SR
    return {'case': label, 'rows': rows, 'n': n, 'hop': hop, 'T': T, 'audio_dtype': str(dtype).replace('torch.', ''),
            'channels': channels, 'compute': str(compute).replace('torch.', ''), 'ms': ms,
            'M_stem_s_per_s': stem_s / ms / 1e3, 'algorithmic_GB': byts / 1e9, 'GBps': byts / ms / 1e6,
            'frac_of_hbm_peak': byts / ms / 1e6 / PEAK}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r01_config_sweep.json')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {'gpu': torch.cuda.get_device_name(0), 'hbm_peak_GBps': PEAK, 'timing': 'CUDA events, device-resident inputs'}

    # ---- config 1: scalar-1s inference on one 4-stem 30 s song through mix_song_smooth (host in, host out)
    song = synthetic.make_song(synthetic.config_seed(1, 0), 4, SR * 30, kind='music', dtype=np.float32)
    tracks = {k: song[i] for i, k in enumerate(IU.STEMS)}
    torch.manual_seed(0)
    net = models.MixingModelScalar1s().cuda().eval()
    IU.mix_song_smooth(Featuriser(), net, tracks, chunk_length=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        IU.mix_song_smooth(Featuriser(), net, tracks, chunk_length=2)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    dev_tracks = {k: torch.from_numpy(v).cuda() for k, v in tracks.items()}
    ms_dev = timed(lambda: IU.mix_song_smooth(Featuriser(), net, dev_tracks, chunk_length=2, return_device=True), reps=5)
    res['config1'] = {'what': 'mix_song_smooth, scalar-1s CNN included, 4 stems x 30 s, chunk 2 s, hop 1024',
                      'host_in_host_out_s': dt, 'stem_s_per_s_host': 120 / dt,
                      'device_resident_ms': ms_dev, 'stem_s_per_s_device': 120 / (ms_dev * 1e-3),
                      'reference_cpu_s_survey': 1.29}

    # ---- config 3: bulk featurisation, one GPU's share of 122 songs x 8 stems x 240 s (5 s chunks, hop 1024)
    rows = 122 * 8 * 240 // 5 // 8
    res['config3'] = [k1_case(rows, 220500, 1024, torch.float32, label='fp32 mono, fp32 FFT'),
                      k1_case(rows, 220500, 1024, torch.int16, label='PCM16 mono, fp32 FFT'),
                      k1_case(rows // 2, 220500, 1024, torch.float64, label='fp64 mono, fp64 FFT (training-path dtype)'),
                      k1_case(rows, 220500, 1024, torch.float32, channels=2, label='fp32 stereo planar, fp32 FFT'),
                      k1_case(rows, 220500, 1024, torch.int16, channels=2, label='PCM16 stereo planar, fp32 FFT'),
                      k1_case(rows, 220500, 1024, torch.int16, channels=2, interleaved=True, label='PCM16 stereo INTERLEAVED (WAV frame order), fp32 FFT'),
                      k1_case(rows, 220500, 1024, torch.float32, channels=2, interleaved=True, label='fp32 stereo interleaved, fp32 FFT'),
                      k1_case(rows // 2, 220500, 1024, torch.float64, channels=2, interleaved=True, label='fp64 stereo interleaved (soundfile.read layout), fp64 FFT')]

    # ---- config 2 / 4 shapes of K1 alone
    res['k1_shapes'] = [k1_case(1280, 88200, 512, torch.float32, channels=2, label='config 2: stereo, 2 s @ hop 512 (T=173)'),
                        k1_case(2560, 176400, 1024, torch.float32, label='4 s @ hop 1024 (T=173)'),
                        k1_case(48 * 5, 220500, 1024, torch.float64, channels=2, label='config 4 batch: 48 x 5 tracks stereo fp64 (T=216)'),
                        k1_case(4096, 88200, 1024, torch.float32, label='2 s @ hop 1024 (T=87)')]

    # ---- config 4: model tail K3 forward + backward on a ResNet batch
    x = torch.rand(48, 4, 1025, 216, device='cuda') * 100 - 80
    m = torch.randn(48, 4, device='cuda', requires_grad=True)
    go = torch.randn(48, 1025, 216, device='cuda')

    def tail():
        y = mixing.feature_mix(x, m)
        y.backward(go)
        m.grad = None

    def tail_ref():
        y = torch.zeros_like(x[:, 0])
        for i in range(4):
            y = y + m[:, i:i + 1].unsqueeze(2) * x[:, i]
        y.backward(go)
        m.grad = None

    ms_k3, ms_ref = timed(tail), timed(tail_ref)
    byts = x.numel() * 4 * 2 + go.numel() * 4 * 2
    res['config4_tail'] = {'what': 'masked = sum m_i x_i forward + backward, B=48, T=216', 'k3_ms': ms_k3,
                           'torch_ops_ms': ms_ref, 'speedup': ms_ref / ms_k3, 'GBps': byts / ms_k3 / 1e6,
                           'frac_of_hbm_peak': byts / ms_k3 / 1e6 / PEAK}

    # ---- config 5: full-song mixdown sweep with known gains (single pass) + loudness + normalise
    meter = Meter(SR)
    sweep = []
    for S in (2, 4, 8, 16, 32):
        for L in (30, 60, 240, 600):
            N = L * SR
            if S * 2 * N * 4 > 8e9:
                continue
            stems = torch.randn(S, 2, N, device='cuda') * 0.1
            K = max(1, L // 2 - 1)
            gains = np.random.default_rng(S * 1000 + L).uniform(0.5, 1.5, size=(S, K))

            def run_mix():
                r = mixing.gain_mix(stems, gains, compute_dtype=torch.float64, want_mix=True, want_stats=True)
                lufs = meter.integrated_loudness_device(r['mix'][None])
                return mixing.scale_rows(r['mix'].reshape(1, -1), torch.pow(10.0, (-20.0 - lufs) / 20.0))

            ms = timed(run_mix, reps=5, warm=2)
            byts = S * 2 * N * 4 + 2 * N * 8 + 2 * N * 8 + 2 * 2 * N * 8        # stems in, mix out, loudness read, scale r+w
            sweep.append({'stems': S, 'seconds': L, 'ms': ms, 'M_stem_s_per_s': S * L / ms / 1e3,
                          'algorithmic_GB': byts / 1e9, 'frac_of_hbm_peak': byts / ms / 1e6 / PEAK})
            del stems
    res['config5_mix_sweep'] = sweep

    with open(out_path, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1)[:6000])


if __name__ == '__main__':
    main()
