"""Host<->device copy ceilings of the box next to the end-to-end pipeline rate (explains bench.py's e2e).

    python tools/pcie_probe.py [out.json]
"""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'pcie_probe.json')
    res = {}
    try:
        res['topo'] = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True, timeout=20).stdout[-1500:]
    except Exception as e:           # noqa: BLE001
        res['topo'] = repr(e)
    res['cpus_allowed'] = len(os.sched_getaffinity(0))
    nbytes = 903_168_000
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes // 2, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    d_out = torch.empty(nbytes // 2, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    ms = ev_time(lambda: d_in.copy_(h_in, non_blocking=True))
    res['h2d_GBps'] = nbytes / ms / 1e6
    ms = ev_time(lambda: h_out.copy_(d_out, non_blocking=True))
    res['d2h_GBps'] = nbytes / 2 / ms / 1e6

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    ms = ev_time(both)
    res['bidir'] = {'ms': ms, 'h2d_GBps': nbytes / ms / 1e6, 'd2h_GBps': nbytes / 2 / ms / 1e6}
    # chunked copies like the pipeline's sub-batches
    for k in (8, 16, 64):
        step = nbytes // k

        def chunks():
            for i in range(k):
                d_in[i * step:(i + 1) * step].copy_(h_in[i * step:(i + 1) * step], non_blocking=True)
        ms = ev_time(chunks)
        res['h2d_GBps_%d_chunks' % k] = nbytes / ms / 1e6
    del h_in, h_out, d_in, d_out

    # the pipeline itself with different sub-batch sizes
    import bench
    from deep_audio_mixer_b200.pipeline import SegmentBatchPipeline
    n = bench.SECONDS * bench.SR
    pipe = SegmentBatchPipeline(n_stems=bench.N_STEMS, group=bench.GROUP, hop=bench.HOP, chunk_samples=bench.CHUNK_SAMPLES)
    nc = pipe.n_chunks(n)
    raw_host = torch.empty((bench.N_SEG, bench.N_STEMS, bench.N_CH, n), dtype=torch.float32).pin_memory()
    raw_host.normal_(0, 0.1)
    m_host = torch.from_numpy(bench.synth_m(bench.N_SEG * nc)).pin_memory()
    out_host = torch.empty((bench.N_SEG, bench.N_CH, n), dtype=torch.float64).pin_memory()
    lufs_host = torch.empty(bench.N_SEG, dtype=torch.float64).pin_memory()
    streams = [torch.cuda.Stream() for _ in range(3)]
    stem_s = bench.N_SEG * bench.N_STEMS * bench.SECONDS
    res['pipeline'] = {}
    for sb in (2, 4, 8, 16, 32):
        t0 = time.perf_counter()
        ms = ev_time(lambda: pipe.run_host(raw_host, m_host, out_host, lufs_host, sub_batch=sb, streams=streams), reps=3, warm=2)
        res['pipeline']['sub_batch_%d' % sb] = {'ms': ms, 'stem_s_per_s': stem_s / ms * 1e3,
                                                'h2d_GBps': raw_host.numel() * 4 / ms / 1e6,
                                                'wall_s': time.perf_counter() - t0}
    with open(out_path, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
