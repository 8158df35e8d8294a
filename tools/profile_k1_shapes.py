"""ncu driver: one profiled K1 launch per shape, bracketed by cudaProfilerStart/Stop
(use ncu --profile-from-start off).  python tools/profile_k1_shapes.py rows,ch,n,hop [rows,ch,n,hop ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_audio_mixer_b200 import features as F

shapes = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]] or [(640, 1, 88200, 512), (640, 1, 89600, 512)]
for rows, ch, n, hop in shapes:
    x = torch.randn((rows, ch, n) if ch > 1 else (rows, n), device='cuda') * 0.1
    out = torch.empty(rows, 1025, 1 + n // hop, device='cuda')
    kw = dict(hop=hop, out=out)
    if ch > 1:
        kw['channel_axis'] = -2
    for _ in range(2):
        F.stft_db(x, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    F.stft_db(x, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('ok', rows, ch, n, hop, float(out[0, 0, 0]))
