"""ncu driver for the loudness kernels (K4) on corpus-shaped input: R mono float32 rows of n samples.
python tools/profile_k4.py [R n dtype]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_audio_mixer_b200.loudness import Meter

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 240 * 44100
dt = torch.float64 if len(sys.argv) > 3 and sys.argv[3] == 'f64' else torch.float32
x = (torch.randn(R, 1, n, device='cuda') * 0.1).to(dt)
m = Meter(44100)
for _ in range(3):
    out = m.integrated_loudness_device(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    out = m.integrated_loudness_device(x)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print('R=%d n=%d %s: %.3f ms, %.1f G samples/s, lufs0 %.4f' % (R, n, dt, ms, R * n / ms / 1e6, float(out[0])))
