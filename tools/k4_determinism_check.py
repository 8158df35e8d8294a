import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from deep_audio_mixer_b200.loudness import Meter
from oracle import loudness as OL
m = Meter(44100)
x = torch.randn(16, 2, 441000, device='cuda', dtype=torch.float64) * 0.1
ref = m.integrated_loudness_device(x).clone()
bad = 0
for i in range(200):
    if not torch.equal(m.integrated_loudness_device(x), ref): bad += 1
print('nondeterministic repeats:', bad, 'of 200')
xh = x[0].cpu().numpy().T
print(abs(float(ref[0]) - OL.integrated_loudness(xh, 44100)))
# tiny segments (shorter than a chunk) through the generic segment entry: compare with numpy
