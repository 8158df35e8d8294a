"""BASELINE.json config 5: listening-test-style full-song mixdown sweep (2-32 stems x 30 s-10 min), known per-chunk
gains, single pass: dB-STFT features of every stem chunk (2 s chunks, hop 1024) + gain masks + stem sum + BS.1770
loudness + -20 LUFS normalisation, measured as a fraction of the HBM roofline of the fused single pass
(SongBatchMixdown.algorithmic_bytes: stems read once, features and mix written once).

Three ways to run the same songs: one song per call (eager), B songs per call (eager), B songs per call replayed
from a CUDA graph.  Device-resident inputs, CUDA events.      python tools/config5_sweep.py [out.json]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deep_audio_mixer_b200.pipeline import SongBatchMixdown  # noqa: E402

SR = 44100
PEAK = 6549.8


def timed(fn, reps=5, warm=2):
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


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_config5_sweep.json')
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peak = float(json.load(open(pk))['hbm_gbs']) if os.path.exists(pk) else PEAK
    mixer = SongBatchMixdown(chunk_samples=2 * SR, hop=1024)
    rows = []
    for S in (2, 4, 8, 16, 32):
        for L in (30, 60, 240, 600):
            N = L * SR
            B = int(min(32, max(1, round(2400 / (S * L)))))
            while B > 1 and B * S * 2 * N * 4 * 2.2 > 24e9:
                B //= 2
            if S * 2 * N * 4 * 2.2 > 24e9:
                continue
            stems = torch.randn(B, S, 2, N, device='cuda') * 0.1
            K = max(1, L // 2 - 1)
            gains = torch.from_numpy(np.random.default_rng(S * 1000 + L).uniform(0.5, 1.5, size=(B, S, K))).cuda()
            nc = N // mixer.chunk_samples
            feats = mixer.feature_buffer(B * S * nc, 'cuda')
            byts = mixer.algorithmic_bytes(B, S, 2, N, mixer.chunk_samples, mixer.hop)

            def per_song():
                for b in range(B):
                    mixer.run(stems[b:b + 1], gains[b:b + 1], feats_out=feats[b * S * nc:(b + 1) * S * nc])

            def batched():
                return mixer.run(stems, gains, feats_out=feats)

            ms_one, ms_batch = timed(per_song), timed(batched)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                batched()
                side.synchronize()
                with torch.cuda.graph(g, stream=side):
                    batched()
            torch.cuda.current_stream().wait_stream(side)
            ms_graph = timed(g.replay)
            stem_s = B * S * L
            rows.append({'stems': S, 'seconds': L, 'songs_per_batch': B, 'algorithmic_GB': byts / 1e9,
                         'ms_one_song_per_call': ms_one, 'ms_batched': ms_batch, 'ms_batched_graph': ms_graph,
                         'frac_one_song_per_call': byts / ms_one / 1e6 / peak, 'frac_batched': byts / ms_batch / 1e6 / peak,
                         'frac_batched_graph': byts / ms_graph / 1e6 / peak,
                         'M_stem_s_per_s_batched_graph': stem_s / ms_graph / 1e3})
            print(rows[-1])
            del stems, feats, g
            torch.cuda.empty_cache()
    res = {'gpu': torch.cuda.get_device_name(0), 'hbm_peak_GBps': peak,
           'what': 'config 5: known-gain single-pass mixdown (features + mix + loudness + normalise), 2 s chunks @ hop 1024',
           'rows': rows}
    with open(out_path, 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
