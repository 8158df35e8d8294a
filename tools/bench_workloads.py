"""The two multi-GPU configurations of BASELINE.json that are not the driver's headline line
(bench.py --workload corpus | train).  Same launch contract as bench.py: plain python for one GPU, torchrun for
N > 1, one rank per GPU, CUDA-event timing bracketed by barrier + synchronize, max over ranks, ONE JSON line.

corpus (configs[2]): bulk dB-STFT featurisation of a synthetic MedleyDB-shaped corpus -- 122 songs x ~8 stems x
    ~4 min, mono float32 stems resident on the device (generated there), 5 s chunks at hop 1024
    (rows [1025, 216], the HEAD shapes of data/dataset.py:132), songs LPT-sharded over the ranks by stem-seconds
    (SURVEY H9), K1 accumulating (sum, sum of squares, count) of the dB features, plus the whole-file BS.1770
    loudness of every stem for compute_mean_loudness (data/dataset.py:115-130).  The ONLY collective is the
    final all-reduce of the few statistics.  Strong scaling: the corpus is fixed, ranks split it.
train (configs[3]): ResNet18 data-parallel training steps: per rank a batch of 48 raw five-track 5 s stereo
    chunks -> FeatureBatcher (one K1 launch, float64 arithmetic like the reference's training path) -> ResNet18
    forward with the fused tail K3 -> MSE -> backward -> Adam(weight_decay=1e-5), model wrapped in
    DistributedDataParallel (NCCL gradient all-reduce, 3 154 180 parameters = 12.6 MB per step).
    Weak scaling (fixed batch per GPU).  The hot path (K1 + K3) is timed next to the whole step.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 44100
UNIT = 'stem-audio-s/s'


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('needs a CUDA device: there is no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    return world, rank, local, dev


def _max_over_ranks(ms, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def corpus_spec(n_songs=122, seed=1234 + 3000):
    """Deterministic MedleyDB-shaped corpus: per song (n_stems, seconds); seconds a multiple of the 5 s chunk."""
    rng = np.random.Generator(np.random.PCG64(seed))
    stems = rng.integers(4, 13, size=n_songs)                   # ~8 stems on average
    secs = (rng.uniform(150.0, 330.0, size=n_songs) // 5 * 5).astype(np.int64)   # ~4 min on average
    return [(int(s), int(t)) for s, t in zip(stems, secs)]


def run_corpus(args):
    import torch
    import torch.distributed as dist
    from deep_audio_mixer_b200 import features as F, sharding, _lib
    from deep_audio_mixer_b200.loudness import Meter
    world, rank, local, dev = _dist_setup()
    lib = _lib.lib()
    spec = corpus_spec()
    costs = [s * t for s, t in spec]
    mine = sharding.assign_lpt(costs, world)[rank]
    chunk, hop = 5 * SR, 1024
    T = 1 + chunk // hop
    # resident shard: every stem of every song of this rank, mono float32, generated on the device into ONE buffer
    # (stem after stem, every stem a whole number of 5 s chunks), so the whole shard is one K1 launch
    g = torch.Generator(device=dev).manual_seed(1234 + 3000 + rank)
    total = sum(spec[i][0] * spec[i][1] * SR for i in mine)
    shard = torch.empty(max(total, chunk), dtype=torch.float32, device=dev)
    songs, off = [], 0
    for i in mine:
        s, t = spec[i]
        view = shard[off:off + s * t * SR].view(s, t * SR)
        view.normal_(0.0, 0.1, generator=g)
        songs.append(view)
        off += s * t * SR
    n_rows = total // chunk
    rows = shard[:n_rows * chunk].view(n_rows, chunk)
    out = torch.empty((max(n_rows, 1), 1025, T), dtype=torch.float32, device=dev)   # the feature cache writer reads this
    meter = Meter(SR)
    stats = torch.zeros(3, dtype=torch.float64, device=dev)
    loud = torch.zeros(2, dtype=torch.float64, device=dev)                         # (sum of LUFS, count)

    def device_pass():
        stats.zero_()
        loud.zero_()
        if n_rows:
            F.stft_db(rows, 2048, hop, out=out.reshape(-1), stats=stats)
        for x in songs:                        # whole-file loudness of every stem (rows of one song share a length)
            lufs = meter.integrated_loudness_device(x[:, None, :])
            loud[0] += lufs.sum()
            loud[1] += float(x.shape[0])

    n0 = lib.dam_launch_count()
    device_pass()
    torch.cuda.synchronize()
    launches_per_step = int(lib.dam_launch_count() - n0)
    k1_kernel = lib.dam_last_stft_kernel().decode() if n_rows else ''
    # the pass is static: captured once in a CUDA graph (a rank's ~600 launches cost more Python time than GPU time)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        device_pass()
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            device_pass()
    torch.cuda.current_stream(dev).wait_stream(side)

    def one_pass():
        graph.replay()
        if world > 1:                          # the only collectives: a few statistics
            dist.all_reduce(stats)
            dist.all_reduce(loud)

    for _ in range(max(args.warmup, 3)):
        one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = _max_over_ranks(e0.elapsed_time(e1), dev, world) / args.steps
    # K1 alone on this rank's shard (roofline of the dominant kernel)
    k1_ms = float('nan')
    if n_rows:
        for _ in range(2):
            F.stft_db(rows, 2048, hop, out=out.reshape(-1))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            F.stft_db(rows, 2048, hop, out=out.reshape(-1))
        b.record()
        torch.cuda.synchronize()
        k1_ms = a.elapsed_time(b) / 3
    my_stem_s = sum(costs[i] for i in mine)
    k1_bytes = n_rows * (chunk * 4 + 1025 * T * 4)
    total_stem_s = sum(costs)
    loads = [sum(costs[i] for i in part) for part in sharding.assign_lpt(costs, world)]
    mean, std = float('nan'), float('nan')
    sc = stats.cpu().numpy()
    if sc[2] > 0:
        mean = sc[0] / sc[2]
        std = float(np.sqrt(max(sc[1] / sc[2] - mean * mean, 0.0)))
    if rank == 0:
        peak = 6549.8
        pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(pk):
            peak = float(json.load(open(pk))['hbm_gbs'])
        line = {
            'metric': 'stem-audio-sec/sec (bulk dB-STFT featurisation + per-stem loudness, config 3)',
            'value': total_stem_s / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic (generated on the device)',
            'config': {'workload': 'config3: 122 songs x 4..12 stems x 150..330 s mono float32, 5 s chunks @ hop 1024 '
                                   '(T=216), feature statistics + whole-file BS.1770 loudness per stem',
                       'total_stem_seconds': total_stem_s, 'songs': len(spec),
                       'sharding': 'LPT by stem-seconds', 'rank_loads_stem_s': loads,
                       'imbalance': max(loads) / (sum(loads) / len(loads)),
                       'l2': 'every song (>= 0.1 GB) is larger than the 126 MB L2'},
            'roofline': {'bound': 'hbm', 'kernel': k1_kernel, 'achieved': k1_bytes / (k1_ms * 1e-3) / 1e9, 'peak': peak,
                         'unit': 'GB/s', 'frac': k1_bytes / (k1_ms * 1e-3) / 1e9 / peak, 'traffic': None,
                         'k1_ms_rank0': k1_ms, 'rank0_stem_s': my_stem_s},
            'gpu_launches': launches_per_step * args.steps,
            'collectives': 'two all-reduces of 3 and 2 float64 per pass (NCCL)',
            'check': {'feature_mean_db': mean, 'feature_std_db': std, 'frames_bins': float(sc[2]),
                      'mean_lufs': float(loud[0].item() / max(loud[1].item(), 1.0))},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    import torch
    import torch.distributed as dist
    from deep_audio_mixer_b200 import models, mixing, trainer as TR, _lib
    world, rank, local, dev = _dist_setup()
    lib = _lib.lib()
    torch.backends.cudnn.allow_tf32 = False              # the reference trains in plain float32
    torch.backends.cuda.matmul.allow_tf32 = False
    B, n = int(os.environ.get('DAM_TRAIN_BATCH', '48')), 5 * SR
    torch.manual_seed(0)
    model = models.ResNet18().to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    opt = torch.optim.Adam(model.parameters(), weight_decay=1e-5)
    tr = TR.ModelTrainer(model, torch.nn.MSELoss(), opt, dev, model_name='resnet', ddp=True, log_every=0)
    batcher = TR.FeatureBatcher(2048, 1024, compute_dtype=torch.float64)
    g = torch.Generator(device=dev).manual_seed(4000 + rank)
    n_batches = 2                                           # two resident raw batches, alternated
    raws = [torch.randn((B, 5, 2, n), device=dev, generator=g) * 0.1 for _ in range(n_batches)]
    ev = {k: [] for k in ('k1', 'step')}

    def step(i, timed=False):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        train, gt = batcher(raws[i % n_batches])
        b.record()
        loss = tr.train_step(train, gt)
        c.record()
        if timed:
            ev['k1'].append((a, b))
            ev['step'].append((a, c))
        return loss

    n0 = lib.dam_launch_count()
    loss = step(0)
    torch.cuda.synchronize()
    launches_per_step = int(lib.dam_launch_count() - n0)
    k1_kernel = lib.dam_last_stft_kernel().decode()
    for i in range(max(args.warmup, 3) - 1):
        step(i + 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i, timed=True)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = _max_over_ranks(e0.elapsed_time(e1), dev, world) / args.steps
    k1_ms = float(np.median([a.elapsed_time(b) for a, b in ev['k1']]))
    # K3 (model tail forward + backward) alone on the training shapes
    x = torch.rand(B, 4, 1025, 216, device=dev) * 100 - 80
    m = torch.randn(B, 4, device=dev, requires_grad=True)
    go = torch.randn(B, 1025, 216, device=dev)

    def tail():
        y = mixing.feature_mix(x, m)
        y.backward(go)
        m.grad = None
    for _ in range(3):
        tail()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        tail()
    b.record()
    torch.cuda.synchronize()
    k3_ms = a.elapsed_time(b) / 10
    stem_s = B * 5 * 5                                       # five tracks x 5 s per sample
    if rank == 0:
        line = {
            'metric': 'stem-audio-sec/sec (ResNet18 data-parallel training step: K1 features + CNN + K3 tail + Adam, config 4)',
            'value': world * stem_s / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32 CNN (TF32 off), f64 K1', 'data': 'synthetic (generated on the device)',
            'config': {'workload': 'config4: ResNet18, per-rank batch %d x (4 stems + mix) x 5 s stereo float32, hop 1024 '
                                   '(T=216), MSELoss, Adam(weight_decay=1e-5), DistributedDataParallel' % B,
                       'samples_per_s': world * B / (ms * 1e-3), 'parameters': n_params,
                       'gradient_allreduce_bytes_per_step': n_params * 4,
                       'l2': 'per-step inputs (%.2f GB raw audio) are larger than the 126 MB L2' % (B * 5 * 2 * n * 4 / 1e9)},
            'hot_path': {'k1_ms': k1_ms, 'k1_kernel': k1_kernel, 'k3_fwd_bwd_ms': k3_ms,
                         'cnn_and_optimizer_ms': ms - k1_ms - k3_ms, 'hot_path_share_of_step': (k1_ms + k3_ms) / ms},
            'gpu_launches': launches_per_step * args.steps, 'gpu_launches_per_step': launches_per_step,
            'collectives': 'DDP gradient all-reduce (NCCL), %.1f MB per step' % (n_params * 4 / 1e6),
            'check': {'loss': float(loss)},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
