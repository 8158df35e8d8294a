import sys, torch
sys.path.insert(0, '/root/repo')
from deep_audio_mixer_b200 import features as F
n, hop = 220500, 1024
T = 1 + n // hop
for rows in (2048, 8192, 16384, 32768, 49152):
    x = torch.empty(rows, n, device='cuda').normal_(0, 0.1)
    out = torch.empty(rows, 1025, T, device='cuda')
    for _ in range(2):
        F.stft_db(x, hop=hop, out=out.reshape(-1))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        F.stft_db(x, hop=hop, out=out.reshape(-1))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    byts = rows * (n * 4 + 1025 * T * 4)
    print('rows %6d  %.1f GB  %.3f ms  frac %.3f' % (rows, byts / 1e9, ms, byts / ms / 1e6 / 6549.8), flush=True)
    # same rows in 4 sub-launches
    q = rows // 4
    a.record()
    for _ in range(3):
        for i in range(4):
            F.stft_db(x[i * q:(i + 1) * q], hop=hop, out=out[i * q:(i + 1) * q].reshape(-1))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print('   in 4 sub-launches: %.3f ms  frac %.3f' % (ms, byts / ms / 1e6 / 6549.8), flush=True)
    del x, out
