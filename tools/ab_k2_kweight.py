"""A/B on the bench batch: mixdown + loudness with the meter's first pass inside the mixdown kernel
(dam_gain_mix_kweight + dam_kweight_segments_resume) against the two separate calls.  CUDA graphs, CUDA events.

    python tools/ab_k2_kweight.py [segments] > gpurun_out/ab_k2_kweight.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deep_audio_mixer_b200 import pipeline as P          # noqa: E402


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = 441000
    pipe = P.SegmentBatchPipeline()
    buses = torch.rand((B, 4, 2, n), device='cuda') - 0.5
    gains = torch.rand((B, 4, 5), device='cuda', dtype=torch.float64) + 0.5
    res = {}
    t_mix, r0 = timed(lambda: pipe.stage_mix(buses, gains, for_meter=False))
    t_loud, l0 = timed(lambda: pipe.stage_loudness(r0['mix']))
    t_both, l0 = timed(lambda: pipe.stage_loudness(pipe.stage_mix(buses, gains, for_meter=False)['mix']))
    res['separate'] = {'gain_mix_ms': t_mix, 'loudness_ms': t_loud, 'both_ms': t_both}
    t_mix, r1 = timed(lambda: pipe.stage_mix(buses, gains, for_meter=True))
    t_loud, l1 = timed(lambda: pipe.stage_loudness(r1['mix'], r1['kw_workspace']))

    def fused():
        r = pipe.stage_mix(buses, gains, for_meter=True)
        return pipe.stage_loudness(r['mix'], r['kw_workspace'])
    t_both, l1 = timed(fused)
    res['fused'] = {'gain_mix_ms': t_mix, 'loudness_ms': t_loud, 'both_ms': t_both}
    res['identical_lufs'] = bool(torch.equal(l0, l1))
    # the whole bench step, both ways, interleaved
    raw = torch.rand((B, 8, 2, n), device='cuda') - 0.5
    m = torch.rand((B * 5, 4), device='cuda') - 0.5
    feats = pipe.feature_buffer(B * 5)
    del buses
    graphs = {}
    for name, flag in (('separate', False), ('fused', True)):
        pipe.fuse_meter = flag
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            pipe.run(raw, m, feats_out=feats)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = pipe.run(raw, m, feats_out=feats)
        graphs[name] = (g, out)
    step = {'separate': [], 'fused': []}
    for rnd in range(6):
        for name in ('separate', 'fused'):
            g = graphs[name][0]
            for _ in range(3):
                g.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            step[name].append(round(a.elapsed_time(b) / 20, 4))
    res['step_ms'] = step
    res['step_identical_mix'] = bool(torch.equal(graphs['separate'][1]['mix'], graphs['fused'][1]['mix']))
    res['segments'] = B
    print(json.dumps(res))


if __name__ == '__main__':
    main()
