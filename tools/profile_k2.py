import sys, torch
sys.path.insert(0, '/root/repo')
from deep_audio_mixer_b200 import pipeline as P
pipe = P.SegmentBatchPipeline()
B, n = 64, 441000
buses = torch.rand((B, 4, 2, n), device='cuda') - 0.5
gains = torch.rand((B, 4, 5), device='cuda', dtype=torch.float64) + 0.5
for _ in range(2):
    pipe.stage_mix(buses, gains, for_meter=False)
    pipe.stage_mix(buses, gains, for_meter=True)
torch.cuda.synchronize()
