"""ncu driver: one profiled K1 launch per shape, bracketed by cudaProfilerStart/Stop
(use ncu --profile-from-start off).  python tools/profile_k1_shapes.py rows,ch,n,hop[,dtype[,pad]] ...
dtype f32 | i16; pad 1 = rows padded to a multiple of 8 frames (features.aligned_pitch)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_audio_mixer_b200 import features as F, _lib

args = sys.argv[1:] or ['1280,1,88200,512,f32,1']
for a in args:
    parts = a.split(',')
    rows, ch, n, hop = (int(v) for v in parts[:4])
    dt = parts[4] if len(parts) > 4 else 'f32'
    pad = len(parts) > 5 and parts[5] == '1'
    x = torch.randn((rows, ch, n) if ch > 1 else (rows, n), device='cuda') * 0.1
    if dt == 'i16':
        x = (x * 32768.0).round().clamp(-32768, 32767).to(torch.int16)
    T = 1 + n // hop
    pitch = F.aligned_pitch(T) if pad else T
    out = torch.empty(rows, 1025, pitch, device='cuda')
    kw = dict(hop=hop, out=out, frame_pitch=pitch)
    if ch > 1:
        kw['channel_axis'] = -2
    for _ in range(2):
        F.stft_db(x, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    F.stft_db(x, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('ok', a, _lib.lib().dam_last_stft_kernel().decode(), float(out[0, 0, 0]))
