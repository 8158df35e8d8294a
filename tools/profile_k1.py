"""Tiny driver for ncu captures of K1: python tools/profile_k1.py [f32|f64] [rows] [n] [hop]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_audio_mixer_b200 import features as F

mode = sys.argv[1] if len(sys.argv) > 1 else 'f32'
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2560
n = int(sys.argv[3]) if len(sys.argv) > 3 else 176400
hop = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
cd = torch.float32 if mode == 'f32' else torch.float64
x = torch.randn(rows, n, device='cuda') * 0.1
out = torch.empty(rows, 1025, 1 + n // hop, device='cuda')
for _ in range(4):
    F.stft_db(x, hop=hop, compute_dtype=cd, out_dtype=torch.float32, out=out)
torch.cuda.synchronize()
print('ok', float(out[0, 0, 0]))
