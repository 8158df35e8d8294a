"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol the header
declares, WAV I/O, sharding (incl. a world_size-2 gloo all-reduce), dataset indexing."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from deep_audio_mixer_b200 import _lib, build
    build.build_library()
    header = open(os.path.join(ROOT, 'include', 'dam_b200.h')).read()
    declared = set(re.findall(r'\b(dam_[a-z0-9_]+)\s*\(', header))
    declared -= {'dam_status', 'dam_dtype'}
    assert len(declared) >= 12
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.lib()
    assert lib.dam_version().decode().count('.') == 2
    assert lib.dam_stft_num_frames(88200, 1024) == 87
    # 4 chunks of 256 samples per row -> one slot per scan thread (1024), 4 doubles each, transposed for the scan
    # plus two partial segment sums per chunk
    assert lib.dam_kweight_workspace_bytes(2, 1000) == 2 * 1024 * 4 * 8 + 2 * 4 * 2 * 8


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype in include/dam_b200.h has as many parameters as its ctypes signature, with the same
    pointer / integer / double kinds in the same order (an ABI drift between header, library and binding
    would otherwise only show up as garbage on the GPU)."""
    from deep_audio_mixer_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'dam_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    protos = dict(re.findall(r'\b(dam_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', header, flags=re.S))
    assert set(protos) == set(_lib.SIGNATURES)

    def kind_of_c(param):
        param = param.strip()
        if param in ('void', ''):
            return None
        if '*' in param:
            return 'ptr'
        if re.match(r'(const\s+)?double\b', param):
            return 'f64'
        if re.match(r'(const\s+)?(int64_t|size_t)\b', param):
            return 'i64'
        if re.match(r'(const\s+)?int\b', param):
            return 'i32'
        raise AssertionError('unclassified parameter: %r' % param)

    def kind_of_ctypes(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, 'contents') or getattr(t, '_type_', None) is not None and \
                isinstance(getattr(t, '_type_'), type):
            return 'ptr'
        return {ctypes.c_double: 'f64', ctypes.c_int64: 'i64', ctypes.c_size_t: 'i64', ctypes.c_int: 'i32',
                ctypes.c_int32: 'i32'}[t]

    for name, params in protos.items():
        c_kinds = [k for k in (kind_of_c(q) for q in params.split(',')) if k is not None]
        py_kinds = [kind_of_ctypes(t) for t in _lib.SIGNATURES[name][1]]
        assert c_kinds == py_kinds, (name, c_kinds, py_kinds)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'deep_audio_mixer_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), fn


def test_missing_library_fails_loudly(monkeypatch):
    from deep_audio_mixer_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libdam_b200.so')
    with pytest.raises(_lib.DamError):
        _lib.lib()


def test_cpu_tensor_is_rejected_not_computed_on_cpu():
    from deep_audio_mixer_b200 import features, mixing
    with pytest.raises(ValueError):
        features.stft_db(torch.zeros(1, 4096))
    with pytest.raises(ValueError):
        mixing.gain_mix(torch.zeros(4, 2, 100), [[1.0]] * 4)


def test_wavio_roundtrip(tmp_path):
    from deep_audio_mixer_b200 import wavio
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.9, 0.9, size=(5000, 2))
    p = str(tmp_path / 'a.wav')
    wavio.write(p, x, 44100)
    wi = wavio.info(p)
    assert (wi.samplerate, wi.channels, wi.frames) == (44100, 2, 5000) and abs(wi.duration - 5000 / 44100) < 1e-12
    y, sr = wavio.read(p)
    assert y.dtype == np.float64 and y.shape == (5000, 2) and sr == 44100
    assert np.max(np.abs(y - np.round(x * 32768) / 32768)) == 0
    part, _ = wavio.read(p, start=100, stop=300)
    assert np.array_equal(part, y[100:300])
    raw, _ = wavio.read_raw(p)
    assert raw.dtype == np.int16
    wavio.write(str(tmp_path / 'f.wav'), x.astype(np.float32), 48000, subtype='FLOAT')
    z, sr = wavio.read(str(tmp_path / 'f.wav'), dtype='float32')
    assert sr == 48000 and np.array_equal(z, x.astype(np.float32))


def test_lpt_and_ranges():
    from deep_audio_mixer_b200 import sharding
    costs = [600, 20, 240, 240, 240, 30, 500, 100, 90]
    parts = sharding.assign_lpt(costs, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(costs)))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= 600 and max(loads) - min(loads) <= 240
    assert [sharding.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    # 122 equal songs on 8 ranks: 16 vs 15
    assert sorted(len(p) for p in sharding.assign_lpt([1] * 122, 8)) == [15] * 6 + [16] * 2


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from deep_audio_mixer_b200 import sharding
    sums = np.array([1.0, 2.0, 3.0, 4.0, 5.0]) * (rank + 1)
    counts = np.full(5, rank + 2.0)
    s, c = sharding.allreduce_sum_count(sums, counts)
    stats = torch.tensor([10.0 * (rank + 1), 1000.0 * (rank + 1), 5.0], dtype=torch.float64)
    mean, std = sharding.feature_mean_std(stats)
    q.put((rank, s.tolist(), c.tolist(), mean, std))
    dist.destroy_process_group()


def test_gloo_world2_statistics_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, s, c, mean, std in res:
        assert s == [3.0, 6.0, 9.0, 12.0, 15.0] and c == [5.0] * 5
        assert abs(mean - 3.0) < 1e-12 and abs(std - np.sqrt(300.0 - 9.0)) < 1e-9


def test_dataset_indexing_without_gpu(tmp_path):
    from deep_audio_mixer_b200 import dataset as D, wavio
    sr = 8000
    for name, secs in (('A', 7.3), ('B', 4.0), ('C', 11.9)):
        os.makedirs(str(tmp_path / name / (name + '_STEMS_JOINED')))
        wavio.write(str(tmp_path / name / (name + '_MIX.wav')), np.zeros((int(secs * sr), 2)), sr)
    ds = D.MultitrackAudioDataset(str(tmp_path), songlist=['A', 'B', 'C'], chunk_length=2, sr=sr, seed=7)
    # reference semantics (data/dataset.py:56-113): durations trimmed to a multiple of the chunk length
    durs = dict(zip(ds.songlist, ds.get_song_durations()))
    assert durs == {'A': 6, 'B': 4, 'C': 10}
    assert len(ds) == 3 + 2 + 5
    seen = [ds._calculate_song_index(i) for i in range(len(ds))]
    per_song = [int(d / 2) for d in ds.get_song_durations()]
    expect = [(s, c) for s, n in enumerate(per_song) for c in range(n)]
    assert seen == expect


def _make_tiny_model():
    """Same forward contract as the reference models (forward(x[B,4,F,T]) -> (masked, (m1..m4))) with the tail
    written as the reference's torch expression (model_scalar_1s.py:260-275): the product's CUDA tail cannot run
    on the CPU, and what this test covers is the trainer / DDP harness, not the kernel."""
    import torch

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(4, 3, 3, padding=1)
            self.heads = torch.nn.ModuleList([torch.nn.Linear(3, 1) for _ in range(4)])

        def forward(self, x):
            h = torch.relu(self.conv(x)).mean(dim=(2, 3))
            ms = tuple(head(h) for head in self.heads)
            masked = sum(m.unsqueeze(2) * x[:, i] for i, m in enumerate(ms))
            return masked, ms

    return Tiny()


def _ddp_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from deep_audio_mixer_b200 import trainer as TR, sharding
    torch.manual_seed(0)
    model = _make_tiny_model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 4, 9, 7, generator=g)
    gt = torch.randn(8, 9, 7, generator=g)
    lo, hi = sharding.shard_range(8, rank, world)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    tr = TR.ModelTrainer(model, torch.nn.MSELoss(), opt, torch.device('cpu'), ddp=True, log_every=0)
    assert isinstance(tr.model, torch.nn.parallel.DistributedDataParallel)
    tr.optimizer.zero_grad()
    masked, _ = tr.model(x[lo:hi])
    tr.criterion(masked, gt[lo:hi]).backward()
    grads = [p.grad.clone() for p in model.parameters()]
    loss = tr.train_step(x[lo:hi], gt[lo:hi])
    q.put((rank, [g_.numpy() for g_ in grads], [p.detach().numpy() for p in model.parameters()], float(loss)))
    dist.destroy_process_group()


def test_gloo_world2_ddp_gradients_equal_single_rank():
    """Config 4 (row e2): two ranks on half batches, gradients averaged by the DDP all-reduce, equal the
    single-process gradients on the concatenated batch; after one Adam(weight_decay=1e-5) step both ranks hold
    the same parameters as the single-process run (model_trainer.py:25-44, training_ignite.ipynb:286-300)."""
    import torch
    import torch.multiprocessing as mp
    from deep_audio_mixer_b200 import trainer as TR
    torch.manual_seed(0)
    model = _make_tiny_model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 4, 9, 7, generator=g)
    gt = torch.randn(8, 9, 7, generator=g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    tr = TR.ModelTrainer(model, torch.nn.MSELoss(), opt, torch.device('cpu'), log_every=0)
    masked, _ = model(x)
    torch.nn.MSELoss()(masked, gt).backward()
    ref_grads = [p.grad.clone().numpy() for p in model.parameters()]
    tr.train_step(x, gt)
    ref_params = [p.detach().numpy() for p in model.parameters()]

    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, grads, params, loss in res:
        for a, b in zip(grads, ref_grads):
            assert np.allclose(a, b, rtol=1e-5, atol=1e-7), rank
        for a, b in zip(params, ref_params):
            assert np.allclose(a, b, rtol=1e-5, atol=1e-7), rank


def test_trainer_epoch_matches_reference_loop():
    """ModelTrainer._train_epoch / _validate_epoch return what the reference's loops return (mean of the
    per-batch losses) and step the optimiser once per batch."""
    import torch
    from deep_audio_mixer_b200 import trainer as TR
    torch.manual_seed(3)
    data = [(torch.randn(2, 4, 5, 6), torch.randn(2, 5, 6)) for _ in range(3)]

    def run(use_trainer):
        torch.manual_seed(0)
        model = _make_tiny_model()
        opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=1e-5)
        crit = torch.nn.MSELoss()
        if use_trainer:
            tr = TR.ModelTrainer(model, crit, opt, torch.device('cpu'), log_every=0)
            return tr._train_epoch(data), tr._validate_epoch(data)
        running = 0.0
        for xb, yb in data:                    # model_trainer.py:25-44
            opt.zero_grad()
            masked, _ = model(xb)
            loss = crit(masked, yb)
            loss.backward()
            opt.step()
            running += loss.item()
        val = 0.0
        with torch.no_grad():
            for xb, yb in data:
                val += crit(model(xb)[0], yb).item()
        return running / len(data), val / len(data)

    a, b = run(True), run(False)
    assert abs(a[0] - b[0]) < 1e-6 and abs(a[1] - b[1]) < 1e-6


def test_corpus_spec_and_lpt_balance():
    """bench.py --workload corpus (config 3): the synthetic MedleyDB-shaped corpus is deterministic, sized like
    BASELINE.json says (122 songs x ~8 stems x ~4 min) and LPT keeps eight ranks within 2 % of each other."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import bench_workloads as BW
    from deep_audio_mixer_b200 import sharding
    spec = BW.corpus_spec()
    assert spec == BW.corpus_spec() and len(spec) == 122
    stems = np.array([s for s, _ in spec])
    secs = np.array([t for _, t in spec])
    assert 7.0 <= stems.mean() <= 9.0 and 220 <= secs.mean() <= 260 and np.all(secs % 5 == 0)
    costs = [s * t for s, t in spec]
    assert abs(sum(costs) - 122 * 8 * 240) / (122 * 8 * 240) < 0.1
    for world in (2, 4, 8):
        parts = sharding.assign_lpt(costs, world)
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.02
