"""Drop-in surfaces end to end on the GPU: models, mix_song_smooth, dataset, evaluator."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import features as OF
from oracle import loudness as OL
from oracle import mix as OM
from deep_audio_mixer_b200 import dataset as D
from deep_audio_mixer_b200 import evaluation as E
from deep_audio_mixer_b200 import inference_utils as IU
from deep_audio_mixer_b200 import models as M
from deep_audio_mixer_b200 import mixing as mixing_mod
from deep_audio_mixer_b200 import synthetic, wavio

pytestmark = pytest.mark.gpu
W = golden('window.npz')['hann_2048']


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


class FakeGainModel(torch.nn.Module):                    # same as tools/gen_golden.py
    def forward(self, x):
        mu = x.double().mean(dim=(2, 3))
        wob = 0.3 * torch.sin(mu[:, :1])
        m = 0.02 * mu + 0.1 * torch.arange(4, dtype=torch.float64, device=x.device)[None, :] + wob
        return None, tuple(m[:, i:i + 1] for i in range(4))


@pytest.mark.parametrize('name,ctor,T', [('scalar1s', M.MixingModelScalar1s, 87),
                                         ('scalar2s', M.MixingModelScalar2s, 173),
                                         ('resnet18', M.ResNet18, 216)])
def test_models_match_reference_architecture(cuda_device, name, ctor, T):
    g = golden('models.npz')
    torch.manual_seed(0)
    net = ctor().eval()
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == int(g['nparams_' + name])
    assert sorted(net.state_dict().keys()) == list(g['keys_' + name])
    gen = torch.Generator().manual_seed(11)
    x = (torch.rand((1, 4, 1025, T), generator=gen) * 100.0 - 80.0)
    net = net.cuda()
    with torch.no_grad():
        masked, ms = net(x.cuda())
    m = np.array([float(v) for v in ms])
    # same seed -> same initial weights as the reference; cuDNN vs CPU conv differ in the last bits
    assert np.allclose(m, g['m_' + name], rtol=2e-3, atol=2e-4)
    assert masked.shape == (1, 1025, T)
    assert np.allclose(masked.cpu().numpy()[0, ::41, ::7], g['masked_dec_' + name], rtol=5e-3, atol=5e-2)
    ref_tail = sum(ms[i].unsqueeze(2) * x.cuda()[:, i] for i in range(4))
    assert torch.allclose(masked, ref_tail, rtol=0, atol=1e-4)
    with pytest.raises(Exception):
        net(x.cuda().double())                              # float64 input raises in the reference too


def test_model_trains_through_fused_tail(cuda_device):
    torch.manual_seed(1)
    net = M.MixingModelScalar1s().cuda().train()
    opt = torch.optim.Adam(net.parameters(), weight_decay=1e-5)
    x = torch.rand(2, 4, 1025, 87, device='cuda') * 60 - 80
    gt = torch.rand(2, 1025, 87, device='cuda') * 60 - 80
    masked, _ = net(x)
    loss = torch.nn.functional.mse_loss(masked, gt)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    opt.step()


def test_mix_song_smooth_matches_reference_golden(cuda_device):
    g = golden('mix_smooth.npz')
    sr, cl, N = int(g['sr']), int(g['chunk_length']), int(g['N'])
    song = synthetic.make_song(int(g['seed']), 4, N, sr=sr, kind='music', dtype=np.float32)
    assert hashlib.sha256(song.tobytes()).hexdigest() == str(g['input_sha'])
    tracks = {k: song[i] for i, k in enumerate(OM.STEMS)}
    mixed, raw, smooth, mix = IU.mix_song_smooth(D.Featuriser(), FakeGainModel(), tracks, chunk_length=cl, sr=sr,
                                                 want_mix=True)
    raw = np.array([raw[k] for k in OM.STEMS])
    smooth = np.array([smooth[k] for k in OM.STEMS])
    assert raw.shape == g['raw'].shape
    assert np.max(np.abs(raw / g['raw'] - 1)) < 5e-5            # fp32 feature noise through the fake model
    assert np.max(np.abs(smooth / g['smooth'] - 1)) < 5e-5
    arr = np.stack([mixed[k] for k in OM.STEMS])
    assert arr.dtype == np.float64 and isinstance(mixed['bass'], np.ndarray)
    assert np.allclose(arr[:, :, ::13], g['mixed_dec'], rtol=5e-5, atol=1e-12)
    assert np.allclose(mix[:, ::13], g['mix_sum_dec'], rtol=5e-5, atol=1e-6)
    assert np.array_equal(mix, OM.stem_sum(arr))


def test_config1_scalar1s_song_vs_oracle(cuda_device):
    """BASELINE config 1: scalar-1s inference on one synthetic 4-stem 30 s 44.1 kHz song
    (chunk_length 2, hop 1024 -> T = 87, 14 gains, savgol window 3, interp_coef 94 500)."""
    sr, N = 44100, 44100 * 30
    song = synthetic.make_song(synthetic.config_seed(1, 0), 4, N, kind='music', dtype=np.float32)
    tracks = {k: song[i] for i, k in enumerate(OM.STEMS)}
    torch.manual_seed(0)
    net = M.MixingModelScalar1s().cuda().eval()

    def model_fn(feat):                                      # the oracle's features through the same CNN
        with torch.no_grad():
            _, ms = net(torch.from_numpy(feat)[None].float().cuda())
        return [float(m) for m in ms]

    ref_mixed, ref_raw, ref_smooth = OM.mix_song_smooth(model_fn, tracks, chunk_length=2, sr=sr)
    mixed, raw, smooth = IU.mix_song_smooth(D.Featuriser(), net, tracks, chunk_length=2, sr=sr)
    assert len(raw['bass']) == 14
    for k in OM.STEMS:
        assert np.allclose(raw[k], ref_raw[k], rtol=1e-4)     # batch-14 vs batch-1 cuDNN + fp32 features
        assert np.allclose(smooth[k], ref_smooth[k], rtol=1e-4)
        assert mixed[k].dtype == np.float64 and mixed[k].shape == (2, N)
        assert np.allclose(mixed[k], ref_mixed[k], rtol=1e-4, atol=1e-9)
        # given identical gains the mix arithmetic is exact
        assert np.array_equal(mixed[k], tracks[k] * OM.interpolate_mask(np.array(smooth[k]), N))


def _write_dataset(root, n_songs=2, seconds=7, sr=44100):
    names = []
    for i in range(n_songs):
        name = 'Song%d' % i
        names.append(name)
        song = synthetic.make_song(500 + i, 5, sr * seconds + 1234 * i, kind='music', dtype=np.float32) * 0.5
        os.makedirs(os.path.join(root, name, name + '_STEMS_JOINED'))
        for j, t in enumerate(['bass', 'drums', 'vocals', 'other']):
            wavio.write(os.path.join(root, name, name + '_STEMS_JOINED', '%s_STEM_%s.wav' % (name, t.upper())),
                        song[j].T, sr)
        wavio.write(os.path.join(root, name, name + '_MIX.wav'), song[4].T, sr)
    return names


def test_dataset_dropin(cuda_device, tmp_path):
    root = str(tmp_path)
    names = _write_dataset(root)
    ds = D.MultitrackAudioDataset(root, songlist=list(names), chunk_length=2, seed=3)
    assert len(ds) == 3 + 3 and ds.get_num_songs() == 2 and ds.get_tracklist()[-1] == 'mix'
    train, gt = ds[4]
    assert train.shape == (4, 1025, 87) and gt.shape == (1025, 87)
    assert train.dtype == torch.float64 and train.device.type == 'cpu'
    song_i, chunk_i = ds._calculate_song_index(4)
    audio = ds._read_chunk(song_i, chunk_i)                              # [5, n, 2] float64
    ref_train, ref_gt = OF.process_chunk(list(audio[:4]), audio[4], channel_axis=1)
    # the oracle is fed the torch window the product uses (1-ulp cos differences matter at -120 dB)
    ref = np.stack([OF.compute_features(np.mean(a, axis=1), window=W) for a in audio])
    assert np.max(np.abs(train.numpy() - ref[:4])) <= 1e-4 and np.max(np.abs(gt.numpy() - ref[4])) <= 1e-4
    assert np.max(np.abs(ref_train - ref[:4])) < 1e-2
    tb, gb = ds.get_batch([0, 4, 5])
    assert tb.shape == (3, 4, 1025, 87) and gb.shape == (3, 1025, 87) and tb.is_cuda
    assert torch.equal(tb[1].cpu(), train)
    ml = ds.compute_mean_loudness()
    name0 = ds.songlist[0]
    tr, _ = wavio.read(ds._get_track_path(name0, 'bass'))
    l0 = OL.integrated_loudness(tr, 44100)
    tr1, _ = wavio.read(ds._get_track_path(ds.songlist[1], 'bass'))
    assert abs(ml['bass'] - (l0 + OL.integrated_loudness(tr1, 44100)) / 2) < 1e-7


def test_evaluator_and_baselines(cuda_device):
    sr, N = 44100, 44100 * 17
    ref_song = synthetic.make_song(700, 4, N, kind='music', dtype=np.float32)
    test_song = synthetic.make_song(701, 4, N, kind='music', dtype=np.float32)
    ref_tracks = {k: ref_song[i] for i, k in enumerate(OM.STEMS)}
    test_tracks = {k: test_song[i] for i, k in enumerate(OM.STEMS)}
    mean_loud = {'bass': -25.0, 'drums': -22.0, 'vocals': -21.0, 'other': -24.0}
    ev = E.LoudnessEvaluator(D.Featuriser(), mean_loud, FakeGainModel(), sr=sr, seed=5)
    rel = ev.evaluate_loudness(test_tracks)
    assert np.allclose(rel, OL.evaluate_loudness(test_tracks, sr), atol=1e-8)
    norm, loud = ev.sum_and_normalize(test_tracks)
    s = OM.stem_sum(test_tracks)
    assert abs(loud - OL.integrated_loudness(s.T, sr)) < 1e-6
    assert abs(OL.integrated_loudness(norm.cpu().numpy().T.astype(np.float64), sr) - (-20.0)) < 1e-4
    ln = M.MeanLoudnessModel(mean_loud, sr).forward(test_tracks)
    for k in OM.STEMS:
        assert abs(OL.integrated_loudness(ln[k].T, sr) - mean_loud[k]) < 1e-6
    stats = ev.process_tracks(ref_tracks, test_tracks, n_random_samples=2)
    assert set(stats) == {'song_name', 'sum_error', 'loudnorm_error', 'mix_error', 'random_error'}
    ref_rel = OL.evaluate_loudness(ref_tracks, sr)
    assert abs(stats['sum_error'] - OL.loudness_dict_diff(OL.evaluate_loudness(test_tracks, sr), ref_rel)) < 1e-7
    mix = E.produce_mixture(test_tracks, sr=sr)
    assert torch.equal(mix, norm)


def test_integration_md_binding_stub_runs(cuda_device):
    """The ctypes stub INTEGRATION.md shows a reference maintainer (section B) is executed verbatim against the
    built library, so the document cannot drift from the C ABI."""
    import os
    import re
    from deep_audio_mixer_b200 import _lib, features as F
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    block = re.search(r"```python\n# data/dam_binding.py.*?```", text, re.S).group(0)
    code = block[len('```python\n'):-3].replace("'/path/to/libdam_b200.so'", repr(_lib.LIB_PATH))
    ns = {}
    exec(compile(code, 'INTEGRATION.md', 'exec'), ns)
    x = torch.randn(3, 44100, device='cuda', dtype=torch.float64)
    win = torch.hann_window(2048).cuda()
    got = ns['stft_db'](x, win)
    assert got.dtype == torch.float64 and torch.equal(got, F.stft_db(x, window=win))
    x32 = torch.randn(2, 30000, device='cuda')
    assert torch.equal(ns['stft_db'](x32, win, 2048, 512), F.stft_db(x32, hop=512, window=win))
    q = (x32 * 3000).to(torch.int16)
    assert torch.equal(ns['stft_db'](q, win), F.stft_db(q, window=win))


def test_bus_sum_mono_output_feeds_k1_with_the_same_bits(cuda_device):
    """The bus-sum kernel's optional channel-mean output: equal to the float32 mean of the stereo bus, and the
    features K1 computes from it (asynchronous mono path) equal those it computes from the stereo bus."""
    from deep_audio_mixer_b200 import features as F, mixing, pipeline as P
    group = (0, 1, 1, 2, 2, 3, 3, 3)
    g = torch.Generator(device='cuda').manual_seed(21)
    for dt, n in ((torch.float32, 88200 * 2 + 1000), (torch.float32, 88200 + 3), (torch.int16, 88200 * 2)):
        if dt == torch.int16:
            raw = torch.randint(-8000, 8000, (3, 8, 2, n), device='cuda', generator=g, dtype=torch.int32).to(dt)
        else:
            raw = torch.randn((3, 8, 2, n), device='cuda', generator=g) * 0.2
        buses = mixing.bus_sum(raw, group, 4, batch_dims=1)
        b2, mono = mixing.bus_sum(raw, group, 4, batch_dims=1, want_mono=True)
        assert torch.equal(b2, buses)
        assert torch.equal(mono, (buses[:, :, 0] + buses[:, :, 1]) * 0.5)
        pipe = P.SegmentBatchPipeline(8, group, 88200, 512)
        assert torch.equal(pipe.features(mono), pipe.features(buses))
    pipe = P.SegmentBatchPipeline(8, group, 88200, 512)
    raw = torch.randn((2, 8, 2, 88200 * 3), device='cuda', generator=g) * 0.2
    m = torch.rand((6, 4), device='cuda', generator=g) - 0.7
    a = pipe.run(raw, m)
    pipe.mono_features = False
    b = pipe.run(raw, m)
    for k in ('mix', 'lufs', 'peak', 'features'):
        assert torch.equal(a[k], b[k]), k
    # three channels: the divide form of the mean
    raw3 = torch.randn((2, 4, 3, 5000), device='cuda', generator=g)
    b3, m3 = mixing.bus_sum(raw3, [0, 0, 1, 1], 2, batch_dims=1, want_mono=True)
    b3h = b3.cpu().numpy()                                       # numpy divides; torch multiplies by 1/3
    assert np.array_equal(m3.cpu().numpy(), ((b3h[:, :, 0] + b3h[:, :, 1]) + b3h[:, :, 2]) / np.float32(3.0))


# ---- N3: gain post-processing on the device ------------------------------------------------------------------
@pytest.mark.parametrize('K', [14, 29, 149, 299])
def test_savgol_operator_matches_scipy(cuda_device, K):
    """inference_utils.py:136-141: savgol_filter(raw, odd(int(num_chunks/4)), 2) as a precomputed K x K operator
    applied on the device.  scipy's own result (interior convolution and the mode='interp' edge polynomials) is
    reproduced to rounding; bit equality is not attainable with a different summation order."""
    from scipy.signal import savgol_filter
    num_chunks = K + 1
    rng = np.random.default_rng(K)
    raw = rng.uniform(0.3, 2.0, size=(K, 4))
    win = int(num_chunks / 4)
    win = win if win % 2 else win + 1
    ref = savgol_filter(raw, win, 2, axis=0)
    got = IU.smooth_gains_device(torch.from_numpy(raw).cuda(), num_chunks).cpu().numpy()
    assert got.dtype == np.float64 and got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-13
    assert np.array_equal(IU.smooth_gains(raw[:, 0], num_chunks), savgol_filter(raw[:, 0], win, 2))
    with pytest.raises(ValueError):                      # num_chunks <= 7: scipy refuses win <= polyorder, like the reference
        IU.smooth_gains_device(torch.ones(6, 4, device='cuda'), 7)


def test_gains_from_heads_equals_numpy_float32_power(cuda_device):
    """scalar_dB_to_amplitude (data/dataset_utils.py:46-50) on float32 head outputs, device vs numpy."""
    rng = np.random.default_rng(5)
    m = rng.uniform(-3.0, 1.5, size=(4096, 4)).astype(np.float32)
    ref = np.power(10.0, 0.5 * m)                        # numpy >= 2: float32 in, float32 out
    assert ref.dtype == np.float32
    got = IU.gains_from_heads(torch.from_numpy(m).cuda()).cpu().numpy()
    assert got.dtype == np.float32
    ulp = np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    # numpy's float32 power is itself only faithful (<= 1 ulp): the device value, the correctly rounded one, is
    # within one float32 ulp (6e-8 relative) of it on every element
    assert ulp.max() <= 1
    exact = np.power(10.0, 0.5 * m.astype(np.float64)).astype(np.float32)
    assert (got == exact).mean() > 0.999


def test_mix_song_smooth_device_resident_has_no_host_round_trip(cuda_device):
    """return_device=True: tracks in, tracks / gains out, all on the device; equal to the host-facing call."""
    sr, N = 44100, 44100 * 12
    song = synthetic.make_song(41, 4, N, kind='music', dtype=np.float32)
    host_tracks = {k: song[i] for i, k in enumerate(OM.STEMS)}
    dev_tracks = {k: torch.from_numpy(v).cuda() for k, v in host_tracks.items()}
    a = IU.mix_song_smooth(D.Featuriser(), FakeGainModel(), host_tracks, chunk_length=1, sr=sr)
    b = IU.mix_song_smooth(D.Featuriser(), FakeGainModel(), dev_tracks, chunk_length=1, sr=sr, return_device=True)
    for k in OM.STEMS:
        assert b[0][k].is_cuda and b[1][k].is_cuda and b[2][k].is_cuda
        assert np.array_equal(a[0][k], b[0][k].cpu().numpy())
        assert np.array_equal(np.array(a[1][k]), b[1][k].cpu().numpy())
        assert np.array_equal(np.array(a[2][k]), b[2][k].cpu().numpy())


def test_mix_song_legacy(cuda_device):
    """inference_utils.py:44-102 (legacy mix_song): mono tracks, chunk k of every stem multiplied by the RAW head
    output m_k, stems summed, samples past the analysed chunks left at zero; mask_history holds the raw outputs."""
    sr, N = 8000, 8000 * 7 + 123
    song = synthetic.make_song(42, 4, N, sr=sr, kind='music', dtype=np.float32)[:, 0]          # [4, N] mono
    tracks = {k: song[i] for i, k in enumerate(OM.STEMS)}
    model = FakeGainModel()
    mixed, hist = IU.mix_song(D.Featuriser(), model, tracks, chunk_length=1, sr=sr)
    assert mixed.dtype == np.float64 and mixed.shape == (N,)
    n_an = N // sr - 1
    assert all(len(hist[k]) == n_an for k in OM.STEMS)
    # oracle: the reference's loop with the model called per chunk on the oracle's features
    expect = np.zeros(N)
    for k in range(n_an):
        seg = slice(k * sr, (k + 1) * sr)
        feats = np.stack([OF.compute_features(tracks[t][seg].astype(np.float32), window=W) for t in OM.STEMS])
        with torch.no_grad():
            _, ms = model(torch.from_numpy(feats)[None].float())
        for i, t in enumerate(OM.STEMS):
            assert abs(hist[t][k] - float(ms[i])) <= 5e-5 * max(1.0, abs(float(ms[i])))
            expect[seg] += tracks[t][seg].astype(np.float64) * hist[t][k]
    assert np.all(mixed[n_an * sr:] == 0.0)
    assert np.allclose(mixed, expect, rtol=1e-12, atol=1e-15)


# ---- N4: feature cache round trip ------------------------------------------------------------------------------
def test_feature_cache_round_trip(cuda_device, tmp_path):
    """data/dataset.py:213-268: precompute_features writes {chunk}_train.npy / {chunk}_gt.npy straight from K1
    output; a dataset opened with compute_features=False reads back exactly what __getitem__ computes on the fly,
    under the names the reference's reader uses (its own writer uses other names: names='write' produces those)."""
    root = str(tmp_path)
    names = _write_dataset(root)
    ds = D.MultitrackAudioDataset(root, songlist=list(names), chunk_length=2, seed=3)
    ds.precompute_features()
    cached = D.MultitrackAudioDataset(root, songlist=list(ds.songlist), chunk_length=2, compute_features=False)
    cached.songlist, cached.song_durations = list(ds.songlist), list(ds.song_durations)
    assert len(cached) == len(ds)
    for i in range(len(ds)):
        train, gt = ds[i]
        ctrain, cgt = cached[i]
        assert ctrain.shape == (4, 1025, 87) and cgt.shape == (1025, 87) and ctrain.dtype == np.float64
        assert np.array_equal(train.numpy(), ctrain) and np.array_equal(gt.numpy(), cgt)
    song0 = ds.songlist[0]
    fdir = os.path.join(root, song0, song0 + '_FEATURES')
    assert sorted(os.listdir(fdir))[:2] == ['0_gt.npy', '0_train.npy']
    ds.precompute_features(names='write')
    assert os.path.exists(os.path.join(fdir, '0_train_2s.npy')) and os.path.exists(os.path.join(fdir, '0_gt_2s.npy'))
    assert np.array_equal(np.load(os.path.join(fdir, '0_train_2s.npy')), np.load(os.path.join(fdir, '0_train.npy')))
    # augmentation on cached features adds 20 log10(g_s) per stem (data/dataset.py:170-179)
    aug = D.MultitrackAudioDataset(root, songlist=list(ds.songlist), chunk_length=2, compute_features=False, augment_data=True)
    aug.songlist, aug.song_durations = list(ds.songlist), list(ds.song_durations)
    np.random.seed(11)
    atrain, _ = aug[0]
    np.random.seed(11)
    g = np.random.uniform(0.6, 1.4, size=4)
    assert np.allclose(atrain, cached[0][0] + 20 * np.log10(g)[:, None, None])


# ---- N1: the dataset uploads the stored samples, not float64 --------------------------------------------------
def test_dataset_pcm_ingest_is_bit_identical_to_float64_upload(cuda_device, tmp_path):
    """PCM16 files go up as int16 and PCM24 files as their 3-byte samples (decoded by dam_pcm_unpack); the features
    are the bits the float64 upload of soundfile-style samples gives, with and without the augmentation gain."""
    sr = 44100
    for subtype, dtype in (('PCM_16', torch.int16), ('PCM_24', torch.float64)):
        root = str(tmp_path / subtype)
        os.makedirs(root)
        name = 'S'
        song = synthetic.make_song(91, 5, sr * 5 + 777, kind='music', dtype=np.float32) * 0.5
        os.makedirs(os.path.join(root, name, name + '_STEMS_JOINED'))
        for j, t in enumerate(['bass', 'drums', 'vocals', 'other']):
            wavio.write(os.path.join(root, name, name + '_STEMS_JOINED', '%s_STEM_%s.wav' % (name, t.upper())), song[j].T, sr, subtype)
        wavio.write(os.path.join(root, name, name + '_MIX.wav'), song[4].T, sr, subtype)
        ds = D.MultitrackAudioDataset(root, songlist=[name], chunk_length=2)
        dev_audio = ds._read_chunk_device(0, 1)
        assert dev_audio.dtype == dtype and dev_audio.shape == (5, 2 * sr, 2)
        host64 = ds._read_chunk(0, 1)                                        # float64 [5, n, 2], as soundfile returns it
        if subtype == 'PCM_24':
            assert np.array_equal(dev_audio.cpu().numpy(), host64)
        gains = [0.61, 0.97, 1.0, 1.23, 1.39]
        for g in (None, gains):
            a = ds._featurise(dev_audio, g)
            b = ds._featurise(torch.from_numpy(host64).cuda(), g)
            assert a.dtype == torch.float64 and torch.equal(a, b)
        train, gt = ds[1]
        ref = np.stack([OF.compute_features(np.mean(x, axis=1), window=W) for x in host64])
        assert np.max(np.abs(train.numpy() - ref[:4])) <= 1e-4 and np.max(np.abs(gt.numpy() - ref[4])) <= 1e-4


def test_host_gains_upload_ring(cuda_device):
    """gain_mix with host gains: many calls through the 8-slot pinned ring give the same result as device gains."""
    stems = torch.randn(3, 2, 5000, device='cuda')
    rng = np.random.default_rng(1)
    for _ in range(20):
        g = rng.uniform(0.5, 1.5, size=(3, 7))
        a = mixing_mod.gain_mix(stems, g)['mix']
        b = mixing_mod.gain_mix(stems, torch.from_numpy(g).cuda())['mix']
        assert torch.equal(a, b)


def test_song_batch_mixdown_matches_oracle(cuda_device):
    """Config 5 (data/listening_test_data_preparation.py:19-53 with known per-chunk gains): a batch of songs through
    one launch per stage equals the oracle song by song -- features, float64 mix (bit exact before normalisation),
    loudness, -20 LUFS normalisation."""
    from deep_audio_mixer_b200.pipeline import SongBatchMixdown
    sr, N, S, B = 44100, 44100 * 5 + 100, 3, 2
    songs = np.stack([synthetic.make_song(300 + b, S, N, kind='music', dtype=np.float32) for b in range(B)])   # [B,S,2,N]
    gains = np.random.default_rng(3).uniform(0.5, 1.5, size=(B, S, 4))
    mixer = SongBatchMixdown(chunk_samples=sr, hop=1024, sr=sr)
    res = mixer.run(torch.from_numpy(songs).cuda(), torch.from_numpy(gains).cuda())
    nc = N // sr
    feats = res['features'].cpu().numpy().reshape(B, S, nc, 1025, -1)
    for b in range(B):
        st = OM.apply_gains(songs[b], gains[b])
        mix = OM.stem_sum(st)
        L = OL.integrated_loudness(mix.T, sr)
        assert abs(float(res['lufs'][b]) - L) < 1e-8
        assert np.allclose(res['mix'][b].cpu().numpy(), OM.normalize_loudness(mix, L, -20.0), rtol=1e-9, atol=0)     # the gain follows the device loudness (1e-9 LU apart)
        for s in range(S):
            for k in range(nc):
                mono = np.mean(songs[b, s][:, k * sr:(k + 1) * sr], axis=0)
                truth = OF.compute_features(mono.astype(np.float64), window=W)
                amp_t, amp_g = 10.0 ** (truth / 20), 10.0 ** (feats[b, s, k].astype(np.float64) / 20)
                assert np.max(np.abs(amp_g - amp_t) / amp_t.max(axis=0, keepdims=True)) <= 1e-6
