"""Quick K1 timing on the GPU box (not the contract bench): python tools/quick_bench.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_audio_mixer_b200 import features as F

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]

def main():
    dev = torch.device('cuda')
    print(torch.cuda.get_device_name(0))
    for rows, n, hop in ((1280, 88200, 512), (1280 * 2, 176400, 1024), (4096, 220500, 1024)):
        x = torch.randn(rows, n, device=dev) * 0.1
        stem_s = rows * n / 44100.0
        for cd in (torch.float32, torch.float64):
            out = torch.empty(rows, 1025, 1 + n // hop, device=dev, dtype=torch.float32)
            med, best = timeit(lambda: F.stft_db(x, hop=hop, compute_dtype=cd, out_dtype=torch.float32, out=out))
            byts = x.numel() * 4 + out.numel() * 4
            print('rows=%d n=%d hop=%d compute=%s: median %.3f ms best %.3f ms -> %.2f M stem-s/s, %.0f GB/s (%.1f%% of 6549.8)'
                  % (rows, n, hop, cd, med, best, stem_s / med / 1e3, byts / med / 1e6, byts / med / 1e6 / 65.498))
        del x, out
    # stereo planar fp32 source, mono features (config-2 shape)
    x = torch.randn(1280, 2, 88200, device=dev) * 0.1
    out = torch.empty(1280, 1025, 173, device=dev)
    med, best = timeit(lambda: F.stft_db(x, hop=512, channel_axis=-2, out=out))
    print('stereo planar rows=1280 hop=512: median %.3f ms -> %.2f M stem-s/s' % (med, 1280 * 2 / med / 1e3))

if __name__ == '__main__':
    main()
