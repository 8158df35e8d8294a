#!/usr/bin/env python
"""Headline benchmark: stem-audio-seconds per second through STFT -> dB -> gains -> mix (+loudness).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload segments|corpus|train]

--workload segments (default; what the driver runs) is BASELINE.json configs[1]; corpus = configs[2] (bulk
featurisation of a MedleyDB-shaped corpus, songs LPT-sharded over the ranks, statistics all-reduce) and train =
configs[3] (ResNet18 data-parallel training epoch, DDP gradient all-reduce) live in tools/bench_workloads.py.

Workload (BASELINE.json configs[1]): model_scalar_2s inference batch of 64 synthetic 8-stem 10 s
stereo 44.1 kHz segments per GPU: bus sum 8->4, dB-STFT features of every 2 s chunk (hop 512, T=173,
the only shape model_scalar_2s accepts), per-chunk gains 10**(0.5 m), zero-order-hold masks + stem sum
(float64, as in the reference), BS.1770 loudness of every mix and -20 LUFS normalisation.  One "step" =
one pass of that hot path over one batch = 64 x 8 x 10 = 5120 stem-audio-seconds per GPU.  The CNN
itself stays in PyTorch and is timed separately ("cnn" key), as BASELINE.md section 4 prescribes.

Prints ONE JSON line (rank 0).  Launched by the driver as plain python (N=1) or torchrun (N>1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
N_SEG, N_STEMS, N_CH, SECONDS = 64, 8, 2, 10
CHUNK_SAMPLES, HOP = 2 * SR, 512
GROUP = (0, 1, 1, 2, 2, 3, 3, 3)
METRIC = 'stem-audio-sec/sec (STFT->dB->gains->mix)'
UNIT = 'stem-audio-s/s'
WORKLOAD = ('config2: model_scalar_2s inference batch of 64 synthetic 8-stem 10 s stereo 44.1 kHz segments; '
            'bus-sum 8->4 (+ mono down-mix of each bus), 2 s chunks @ hop 512 (T=173), float64 mask/mix, '
            'BS.1770 loudness, -20 LUFS normalise')
# one dict for BOTH arms (ours and --impl reference): the driver compares them key by key
CONFIG = {'workload': WORKLOAD, 'segments_per_gpu': N_SEG, 'stems': N_STEMS, 'seconds': SECONDS, 'hop': HOP,
          'chunk_samples': CHUNK_SAMPLES,
          'l2': 'inputs (1.8 GB per step) are larger than the 126 MB L2; no explicit flush',
          'parallelism': 'segments sharded per GPU, no data-path collective'}


def synth_m(n_rows, seed=1234 + 2000):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-1.0, 0.5, size=(n_rows, 4)).astype(np.float32)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        # samples taken while the GPU was under this benchmark's load: from the start of the (long)
        # warm-up to the end of the timed region; the timed region itself can be shorter than one
        # nvidia-smi sampling period
        window = [ln for ts, ln in self.lines
                  if (self.t_begin is None or ts >= self.t_begin) and (self.t_end is None or ts <= self.t_end + 0.05)]
        for ln in window:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def cpu_reference_throughput(n_seg_sample, steps, warmup, variants=False):
    """The oracle port (oracle/cpu_reference.py) on the host cores, on a bounded sample of the workload.
    Segments are independent, so they are spread over a thread pool with one worker per core (torch and
    numpy release the GIL in their kernels; measured 2.4x faster than one segment at a time with
    torch's own intra-op threads).  ``variants`` adds the other three cells of SURVEY 8(d)'s table (one
    thread / all cores x float32 / float64 features) on small samples."""
    import concurrent.futures as cf
    import torch
    from oracle import cpu_reference as CR
    from deep_audio_mixer_b200 import synthetic
    cores = os.cpu_count() or 1
    n_seg_sample = max(n_seg_sample, min(cores, N_SEG))
    torch.set_num_threads(1)
    n = SECONDS * SR
    nc = n // CHUNK_SAMPLES
    raw = np.stack([synthetic.make_song(synthetic.config_seed(2, i), N_STEMS, n, kind='white', dtype=np.float32)
                    for i in range(n_seg_sample)])
    m = synth_m(n_seg_sample * nc)

    def one(b, dtype=None):
        return CR.segment_step(raw[b], GROUP, m[b * nc:(b + 1) * nc], CHUNK_SAMPLES, HOP, SR, feature_dtype=dtype)[2]

    with cf.ThreadPoolExecutor(max_workers=cores) as ex:
        for _ in range(warmup):
            list(ex.map(one, range(min(cores, n_seg_sample))))
        t0 = time.perf_counter()
        for _ in range(steps):
            list(ex.map(one, range(n_seg_sample)))
        dt = (time.perf_counter() - t0) / steps
        extra = None
        if variants:
            per_seg = N_STEMS * SECONDS
            extra = {}
            t0 = time.perf_counter()
            list(ex.map(lambda b: one(b, np.float64), range(n_seg_sample)))
            extra['float64_features_all_cores'] = n_seg_sample * per_seg / (time.perf_counter() - t0)
            for name, dt_ in (('float32_features_1_thread', None), ('float64_features_1_thread', np.float64)):
                one(0, dt_)
                t0 = time.perf_counter()
                for b in range(2):
                    one(b, dt_)
                extra[name] = 2 * per_seg / (time.perf_counter() - t0)
    stem_s = n_seg_sample * N_STEMS * SECONDS
    sample = ('%d of %d segments (%d stem-audio-s) per step, one worker thread per core, '
              'float32 features like the inference path') % (n_seg_sample, N_SEG, stem_s)
    if variants:
        return stem_s / dt, dt, cores, sample, extra
    return stem_s / dt, dt, cores, sample


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # every step is the WHOLE per-GPU batch (64 segments, ~0.7 s on 16 cores): same workload and config as our arm
    value, dt, cores, sample = cpu_reference_throughput(N_SEG, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(CONFIG),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def bind_to_gpu_numa_node(local):
    """Pin this rank (and so the pages of its pinned host buffers, by first touch) to the CPUs of the NUMA node its
    GPU hangs off -- sysfs numa_node of the PCI device, NVML's CPU affinity as a fall-back: with several ranks the
    end-to-end leg is bound by host memory / PCIe root placement."""
    allowed = os.sched_getaffinity(0)
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        addr = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % addr) as f:
            node = int(f.read().strip())
        if node >= 0:
            cpus = set()
            with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
                for part in f.read().strip().split(','):
                    lo, _, hi = part.partition('-')
                    cpus.update(range(int(lo), int(hi or lo) + 1))
            use = cpus & allowed
            if use and use != allowed:
                os.sched_setaffinity(0, use)
                return {'source': 'sysfs', 'numa_node': node, 'cpus': len(use)}
            return {'source': 'sysfs', 'numa_node': node, 'cpus': len(use), 'bound': False}
    except Exception:            # noqa: BLE001 -- best effort
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return {'source': 'nvml', 'cpus': len(use)}
    except Exception:            # noqa: BLE001 -- best effort; the bench runs unpinned otherwise
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from deep_audio_mixer_b200 import pipeline as P
    from deep_audio_mixer_b200 import synthetic, models, _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib.lib()

    n = SECONDS * SR
    nc = n // CHUNK_SAMPLES
    pipe = P.SegmentBatchPipeline(N_STEMS, GROUP, CHUNK_SAMPLES, HOP, device=dev)
    raw = synthetic.make_batch_torch(synthetic.config_seed(2, rank), (N_SEG, N_STEMS, N_CH, n), dev)
    m_host = torch.from_numpy(synth_m(N_SEG * nc)).pin_memory()
    m = m_host.to(dev)
    F, T = 1025, 1 + CHUNK_SAMPLES // HOP
    feats_buf = pipe.feature_buffer(N_SEG * nc, 4, dev)     # frame axis padded 173 -> 176 (sector-aligned rows)
    stat = torch.zeros(2, dtype=torch.float64, device=dev)

    def step():
        res = pipe.run(raw, m, feats_out=feats_buf)
        if world > 1:                      # the only collective of the path: a tiny statistics all-reduce
            stat[0] = res['lufs'].sum()
            stat[1] = float(N_SEG)
            dist.all_reduce(stat)
        return res

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib = _lib.lib()
    res = step()
    torch.cuda.synchronize()
    # kernels of OUR library per step, counted by the library's launch sites while one step runs eagerly (a graph
    # replay re-launches exactly these), and the K1 variant the dispatcher picked
    n0 = lib.dam_launch_count()
    res = step()
    torch.cuda.synchronize()
    launches_per_step = int(lib.dam_launch_count() - n0)
    k1_kernel = lib.dam_last_stft_kernel().decode()

    # The step is ~11 of our kernel launches plus a few tiny torch ops and takes ~2.4 ms on the device,
    # about what Python needs to enqueue it, so the whole step is captured ONCE in a CUDA graph and the
    # timed region replays it (with several ranks the 16-byte NCCL all-reduce follows each replay).
    side = torch.cuda.Stream(dev)
    graphs = {}

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fn()                           # warm on the capture stream
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                out = fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        return g, out

    use_graph = os.environ.get('DAM_BENCH_GRAPH', '1') != '0'
    if use_graph:
        def graph_body():
            r = pipe.run(raw, m, feats_out=feats_buf)
            if world > 1:
                stat[0:1].copy_(r['lufs'].sum().reshape(1))     # device-to-device (graph-capturable)
                stat[1:2].fill_(float(N_SEG))
            return r
        g_step, res = capture(graph_body)
        if world > 1:
            def run_step():                # graph replay, then the step's only collective (NCCL, 16 bytes)
                g_step.replay()
                dist.all_reduce(stat)
        else:
            run_step = g_step.replay
    else:
        run_step = step

    # warm-up: exactly the W steps asked for (>= 3), reported as `warmup`.  Before them, `warmup_extra_steps` keep
    # the GPU under this load for ~0.7 s so that clocks / power settle and nvidia-smi (50 ms period) has samples of
    # it; they are reported separately.  (With several ranks the step contains a collective, so every rank must run
    # the SAME number of steps: a fixed count instead of a time-based one.)
    sampler.mark_begin()
    n_warm = max(args.warmup, 3)
    n_extra = 0
    t_w = time.time()
    while (n_extra < 200) if world > 1 else (time.time() - t_w < 0.7):
        run_step()
        n_extra += 1
        if n_extra % 8 == 0:
            torch.cuda.synchronize()
    for _ in range(n_warm):
        run_step()
    torch.cuda.synchronize()

    # ---- timed region: CUDA events on the launching stream, barrier + synchronize on both sides
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.mark_end()
    elapsed_ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t_ms.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-stage device time: each stage captured in its own graph and replayed back to back
    # (CUDA events on the launching stream around 10 replays; no host gaps inside a graph)
    def time_graph(fn, reps=10):
        g, out = capture(fn)
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps, out

    stages = {}
    stages['bus_sum'], (buses, mono) = time_graph(lambda: pipe.stage_bus(raw, want_mono=True))
    stages['stft_db'], _ = time_graph(lambda: pipe.features(mono, feats_buf))
    gains = pipe.stage_gains(m, N_SEG, nc)
    stages['gain_mix'], mixres = time_graph(lambda: pipe.stage_mix(buses, gains))
    stages['loudness'], (lufs, factor) = time_graph(lambda: pipe.stage_loudness(mixres['mix'], mixres['kw_workspace'],
                                                                                with_gain=True))
    stages['scale'], _ = time_graph(lambda: pipe.stage_normalise(mixres['mix'], lufs, factor))
    mono_keep = mono
    del buses, mono, mixres
    stem_s_per_step = N_SEG * N_STEMS * SECONDS
    value = world * stem_s_per_step * args.steps / (elapsed_ms * 1e-3)

    # ---- roofline of the dominant kernel (K1, 4 launches per step), timed live above
    ab = P.algorithmic_bytes(N_SEG, N_STEMS, N_CH, n, nc, CHUNK_SAMPLES, HOP)
    peak, peak_src = peaks()
    k1_ms_per_launch = stages['stft_db'] / 4.0
    achieved = ab['stft_db'] / 4.0 / (k1_ms_per_launch * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': k1_kernel + ' (K1: 4 launches per step, mono rows from the bus-sum kernel)',
                'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': ab['stft_db'] / 4.0, 'avg_launch_ms': k1_ms_per_launch,
                'stage_ms': stages, 'stage_algorithmic_bytes': ab}
    # DRAM traffic needs a profiler: it is taken from the committed ncu capture of this launch, but only if that
    # capture is of the kernel that was dispatched just now (otherwise null rather than a stale number)
    traffic_path = os.path.join(ROOT, 'profiles', 'k1_traffic.json')
    if os.path.exists(traffic_path):
        with open(traffic_path) as f:
            tj = json.load(f)
        if tj.get('dispatched_kernel') == k1_kernel:
            roofline['traffic'] = tj.get('dram_bytes_per_launch')
            roofline['traffic_source'] = 'profiles/k1_traffic.json (ncu --set full, commit %s)' % tj.get('commit', '?')
        else:
            roofline['traffic_source'] = 'profiles/k1_traffic.json is of %r, not of the kernel dispatched here' % tj.get('dispatched_kernel')

    # ---- the same step with K1 in float64 arithmetic (meets north_star's 1e-4 dB on EVERY bin; the reference's
    # training-path dtype).  Same batch, same graph capture, same event timing.
    f64_leg = None
    parity = None
    if rank == 0:
        import torch as _t
        pipe.feature_compute_dtype = _t.float64
        g64, _ = capture(lambda: pipe.run(raw, m, feats_out=feats_buf))
        for _ in range(3):
            g64.replay()
        n64 = max(3, args.steps // 4)
        a, b = _t.cuda.Event(enable_timing=True), _t.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n64):
            g64.replay()
        b.record()
        _t.cuda.synchronize()
        ms64 = a.elapsed_time(b) / n64
        k1_64, _ = time_graph(lambda: pipe.features(mono_keep, feats_buf))
        k1_64_name = lib.dam_last_stft_kernel().decode()
        pipe.feature_compute_dtype = None
        f64_leg = {'value': stem_s_per_step / (ms64 * 1e-3), 'unit': UNIT, 'ms_per_step': ms64, 'steps': n64,
                   'k1_ms_per_step': k1_64, 'k1_kernel': k1_64_name,
                   'k1_roofline_frac': ab['stft_db'] / (k1_64 * 1e-3) / 1e9 / peak,
                   'note': 'per GPU; K1 computes in float64 (<= 1e-4 dB on every bin vs the reference float64 path), '
                           'output rounded to float32'}
        # parity of the benchmarked arithmetic against the reference's own outputs (tests/golden, unmodified
        # reference): measured here, asserted in tests/test_stft_gpu.py
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import parity_report
        parity = parity_report.k1_vs_reference_goldens(dev)
    del mono_keep

    # ---- end to end through the public API with HOST buffers (uploads/downloads inside the timed region)
    e2e = None
    cnn = None
    cpu_base = None
    if rank == 0 or world > 1:
        raw_host = torch.empty((N_SEG, N_STEMS, N_CH, n), dtype=torch.float32).pin_memory()
        raw_host.copy_(raw)
        out_host = torch.empty((N_SEG, N_CH, n), dtype=torch.float64).pin_memory()
        lufs_host = torch.empty(N_SEG, dtype=torch.float64).pin_memory()
        streams = [torch.cuda.Stream(dev) for _ in range(3)]
        e2e_steps = max(3, min(args.steps, 7))

        def time_host_steps(src_host, dst_host=None, copy_mode='overlap'):
            """Per-step CUDA-event times of run_host (upload + kernels + download each step); the rate is
            quoted on the MEDIAN step (one host-side hiccup -- page migration, another process on the PCIe
            root -- otherwise triples a 5-step sum) and every step time is listed beside it."""
            dst_host = out_host if dst_host is None else dst_host
            # warm-up: the caching allocator needs a few steps to grow the pools the three streams recycle
            # (every cudaMalloc on the way synchronises); fixed count so that all ranks stay in step
            for _ in range(6):
                pipe.run_host(src_host, m_host, dst_host, lufs_host, streams=streams, copy_mode=copy_mode)
                torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            times = []
            t_all0 = torch.cuda.Event(enable_timing=True)
            t_all0.record()
            for _ in range(e2e_steps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                pipe.run_host(src_host, m_host, dst_host, lufs_host, streams=streams, copy_mode=copy_mode)
                b.record()
                torch.cuda.synchronize()
                times.append(a.elapsed_time(b))
            total = t_all0.elapsed_time(b)
            t = torch.tensor(times + [total], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)          # slowest rank, step by step and over the K steps
                dist.barrier()
            t = [float(v) for v in t.cpu()]
            times, total = t[:-1], t[-1]
            if world > 1:
                # several ranks: the contract's statistic -- K steps between barriers, slowest rank.  (The ranks are
                # not in lock step: GPUs on the faster half of the host's PCIe fabric finish early and the others
                # then speed up, so a per-step max over ranks describes no real step.)
                return total / e2e_steps, times
            # one rank: the median step (one host-side hiccup otherwise triples a 7-step sum)
            return float(np.median(times)), times

        up_bytes = int(raw_host.numel() * 4 + m_host.numel() * 4)
        down_bytes = int(out_host.numel() * 8 + lufs_host.numel() * 8)
        med, step_ms = time_host_steps(raw_host)
        e2e = {'value': world * stem_s_per_step / (med * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': up_bytes, 'd2h_bytes_per_step': down_bytes,
               'ms_per_step': med, 'step_ms': [round(v, 3) for v in step_ms],
               'statistic': 'median step' if world == 1 else 'K steps between barriers / K, slowest rank (step_ms: per-step max over ranks)',
               'cpu_binding': numa, 'achieved_GBps_all_ranks': world * (up_bytes + down_bytes) / med / 1e6,
               'check': float(lufs_host[0])}
        # the measured ceiling of this box with all ranks copying at once (tools/pcie_probe_multi.py, committed)
        probe_path = os.path.join(ROOT, 'profiles', 'r02_pcie_probe_%d.json' % world)
        if os.path.exists(probe_path):
            with open(probe_path) as f:
                pj = json.load(f)
            pl = pj.get('numa_bound') if isinstance(pj.get('numa_bound'), dict) else pj['default_placement']
            h2d, bidir = pl['h2d']['aggregate_GBps'], pl['bidir']['aggregate_GBps']
            # lower bound of a step on this host: the uploads alone at the rate all ranks get together when nothing
            # else is on the links (the downloads can hide behind them at best); `bidir` is the rate of the same
            # traffic mix with both directions started together
            floor_ms = world * up_bytes / h2d / 1e6
            e2e['pcie_ceiling'] = {'h2d_only_aggregate_GBps': h2d, 'bidir_aggregate_GBps': bidir,
                                   'source': 'profiles/r02_pcie_probe_%d.json (tools/pcie_probe_multi.py, all ranks at once)' % world,
                                   'step_floor_ms': floor_ms, 'e2e_ms_over_floor': med / floor_ms,
                                   'e2e_GBps_over_bidir': e2e['achieved_GBps_all_ranks'] / bidir}
        # variants of the same call (not the headline): one copy direction at a time per GPU; PCM16 host buffers
        # (the format the reference's WAV files are in -- int16 -> x/32768 float32 happens inside the bus-sum
        # kernel, halving the upload); float32 mix out (halves the download; the reference's mix is float64)
        med_s, step_s = time_host_steps(raw_host, copy_mode='serial')
        e2e['serial_copies'] = {'value': world * stem_s_per_step / (med_s * 1e-3), 'unit': UNIT, 'ms_per_step': med_s,
                                'step_ms': [round(v, 3) for v in step_s]}
        raw16_host = torch.empty((N_SEG, N_STEMS, N_CH, n), dtype=torch.int16).pin_memory()
        raw16_host.copy_((raw * 32768.0).round().clamp(-32768, 32767).to(torch.int16))
        med16, step16 = time_host_steps(raw16_host)
        e2e['pcm16_input'] = {'value': world * stem_s_per_step / (med16 * 1e-3), 'unit': UNIT,
                              'h2d_bytes_per_step': int(raw16_host.numel() * 2 + m_host.numel() * 4),
                              'd2h_bytes_per_step': down_bytes,
                              'ms_per_step': med16, 'step_ms': [round(v, 3) for v in step16], 'check': float(lufs_host[0])}
        out32_host = torch.empty((N_SEG, N_CH, n), dtype=torch.float32).pin_memory()
        med32, step32 = time_host_steps(raw16_host, dst_host=out32_host)
        e2e['pcm16_input_float32_mix_out'] = {'value': world * stem_s_per_step / (med32 * 1e-3), 'unit': UNIT,
                                              'h2d_bytes_per_step': int(raw16_host.numel() * 2 + m_host.numel() * 4),
                                              'd2h_bytes_per_step': int(out32_host.numel() * 4 + lufs_host.numel() * 8),
                                              'ms_per_step': med32, 'step_ms': [round(v, 3) for v in step32]}
        del out32_host
        del raw_host, out_host, raw16_host

    if rank == 0:
        # the CNN between the two halves: PyTorch/cuDNN, random-init scalar-2s, fp32 without TF32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.manual_seed(0)
        net = models.MixingModelScalar2s().to(dev).eval()
        with torch.no_grad():
            for _ in range(2):
                for i in range(0, feats_buf.shape[0], 64):
                    net(feats_buf[i:i + 64, ..., :T])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(0, feats_buf.shape[0], 64):
                net(feats_buf[i:i + 64, ..., :T])
            b.record()
            torch.cuda.synchronize()
        cnn_ms = a.elapsed_time(b)
        hot_ms = elapsed_ms / args.steps
        cnn = {'model': 'MixingModelScalar2s (PyTorch/cuDNN fp32, TF32 off, random init)', 'ms_per_batch': cnn_ms,
               'value_with_cnn': stem_s_per_step / ((hot_ms + cnn_ms) * 1e-3), 'unit': UNIT}
        del net
        # CPU baseline (the oracle port) on this box's host cores, bounded sample; rank 0 at N = 1 only
        if world == 1:
            v, dt, cores, sample, extra = cpu_reference_throughput(N_SEG, 3, 1, variants=True)
            cpu_base = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                        'ms_per_step': dt * 1e3, 'other_cells_stem_s_per_s': extra}
        else:
            cpu_base = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                        'sample': 'not timed at N > 1 (the CPU leg runs on rank 0 of the N = 1 run only)'}

    if rank == 0:
        launches = launches_per_step * args.steps
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': n_warm, 'warmup_extra_steps': n_extra, 'cuda_graph': use_graph, 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(CONFIG),
            'roofline': roofline, 'cpu_baseline': cpu_base, 'e2e': e2e, 'gpu_launches': launches,
            'gpu_launches_per_step': launches_per_step, 'parity': parity, 'float64_compute': f64_leg,
            'clocks': clocks, 'cnn': cnn,
            'check': {'lufs0': float(res['lufs'][0]), 'peak0': float(res['peak'][0, 0])},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='segments', choices=['segments', 'corpus', 'train'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload != 'segments':
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import bench_workloads
        (bench_workloads.run_corpus if args.workload == 'corpus' else bench_workloads.run_train)(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
