"""One step of the bench pipeline for ncu captures: python tools/profile_pipeline.py [n_seg]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from deep_audio_mixer_b200 import pipeline as P, synthetic

n_seg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda', 0)
n = bench.SECONDS * bench.SR
nc = n // bench.CHUNK_SAMPLES
pipe = P.SegmentBatchPipeline(bench.N_STEMS, bench.GROUP, bench.CHUNK_SAMPLES, bench.HOP, device=dev)
raw = synthetic.make_batch_torch(1, (n_seg, bench.N_STEMS, bench.N_CH, n), dev)
m = torch.from_numpy(bench.synth_m(n_seg * nc)).to(dev)
for _ in range(3):
    res = pipe.run(raw, m)
torch.cuda.synchronize()
print('ok', float(res['lufs'][0]))
