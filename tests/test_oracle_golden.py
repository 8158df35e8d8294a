"""The oracle (oracle/*.py) against the reference's own outputs (tests/golden, produced by
tools/gen_golden.py from the unmodified reference) and against the known-answer values recorded
in the reference's notebooks.  CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import golden
from oracle import features as OF
from oracle import mix as OM
from oracle import loudness as OL
from deep_audio_mixer_b200 import synthetic


def test_window_matches_torch_hann():
    g = golden('window.npz')
    for n in (512, 1024, 2048):
        w = OF.hann_window_fp32(n)
        assert w.dtype == np.float32
        # fp32 cos implementations may differ in the last ulp
        assert np.max(np.abs(w - g['hann_%d' % n])) <= 1.2e-7


@pytest.mark.parametrize('kind', ['white', 'music', 'tonal'])
def test_features_fp64_match_reference(kind):
    g = golden('features.npz')
    w = golden('window.npz')['hann_2048']
    a = g['in_%s_float64' % kind]
    out = OF.compute_features(a, window=w)
    ref = g['out_%s_float64' % kind]
    assert out.shape == ref.shape == (1025, 1 + len(a) // 1024)
    assert out.dtype == np.float64
    # fp64 pocketfft vs fp64 MKL: ~1e-9 dB except bins that sit on the -100 dB clamp
    assert np.max(np.abs(out - ref)) < 1e-6


@pytest.mark.parametrize('kind', ['white', 'music'])
def test_features_fp32_same_error_class_as_reference(kind):
    """float32 in -> float32 compute in both; two fp32 FFTs differ by O(eps * frame norm),
    so compare in the amplitude domain relative to the frame's largest bin."""
    g = golden('features.npz')
    w = golden('window.npz')['hann_2048']
    a = g['in_%s_float32' % kind]
    out = OF.compute_features(a, window=w)
    ref = g['out_%s_float32' % kind]
    assert out.dtype == np.float32 and out.shape == ref.shape
    amp_o, amp_r = 10.0 ** (out.astype(np.float64) / 20), 10.0 ** (ref.astype(np.float64) / 20)
    peak = amp_r.max(axis=0, keepdims=True)
    assert np.max(np.abs(amp_o - amp_r) / peak) < 2e-6
    truth = g['out_%s_float64' % kind]
    # neither fp32 path is closer than ~1e-4 dB to the fp64 truth on every bin
    assert np.median(np.abs(out - truth)) < 1e-4


def test_features_variants():
    g = golden('features.npz')
    out = OF.compute_features(g['in_hop512'], 2048, 512, golden('window.npz')['hann_2048'])
    assert out.shape == (1025, 1 + 9000 // 512)
    assert np.max(np.abs(out - g['out_hop512'])) < 1e-6
    out = OF.compute_features(g['in_batch'], window=golden('window.npz')['hann_2048'])
    assert out.shape == g['out_batch'].shape == (2, 1025, 9)
    out = OF.compute_features(g['in_w1024'], 1024, 256, golden('window.npz')['hann_1024'])
    assert out.shape == g['out_w1024'].shape == (513, 1 + 9000 // 256)
    assert np.max(np.abs(out - g['out_w1024'])) < 1e-6
    sil = OF.compute_features(np.zeros(4096))
    assert np.all(sil == -100.0) and np.all(g['out_silence'] == -100.0)
    assert np.array_equal(OF.stereo_to_mono(g['in_stereo_interleaved']), g['out_stereo_mono'])


def test_recorded_shapes():
    """training.ipynb:137, experiments.ipynb:845, models/model_resnet.py:155."""
    assert OF.num_frames(44100, 512) == 87
    assert OF.num_frames(88200, 1024) == 87
    assert OF.num_frames(88200, 512) == 173
    assert OF.num_frames(220500, 1024) == 216


def test_gain_equals_db_offset_property():
    """experiments.ipynb:300-353: feat(g*x) == feat(x) + 20 log10 g (MSE 1.8e-10 there)."""
    x = synthetic.make_song(3, 1, 30000, kind='white', dtype=np.float64)[0, 0]
    f1 = OF.compute_features(0.7 * x)
    f0 = OF.compute_features(x)
    assert np.mean((f1 - (f0 + 20 * np.log10(0.7))) ** 2) < 1e-9


def test_interpolate_mask_matches_reference():
    g = golden('mask.npz')
    for key in g.files:
        if key.startswith('g_'):
            _, K, L = key.split('_')
            m = OM.interpolate_mask(g[key], int(L))
            assert m.dtype == np.float64
            assert np.array_equal(m, g['m_%s_%s' % (K, L)])
    assert np.all(OM.interpolate_mask(np.array([0.5]), 10) == 0.0)      # K == 1 quirk


def test_scalar_conversions():
    g = golden('scalar.npz')
    assert np.array_equal(OM.scalar_dB_to_amplitude(g['x']), g['db2amp'])
    assert np.array_equal(OM.scalar_amplitude_to_dB(np.abs(g['x']) + 0.1), g['amp2db'])


def test_mix_song_smooth_matches_reference():
    g = golden('mix_smooth.npz')
    sr, cl, N = int(g['sr']), int(g['chunk_length']), int(g['N'])
    song = synthetic.make_song(int(g['seed']), 4, N, sr=sr, kind='music', dtype=np.float32)
    assert hashlib.sha256(song.tobytes()).hexdigest() == str(g['input_sha'])
    tracks = {k: song[i] for i, k in enumerate(OM.STEMS)}

    def model_fn(feat):                       # tools/gen_golden.py:FakeGainModel
        mu = feat.astype(np.float64).mean(axis=(1, 2))
        return 0.02 * mu + 0.1 * np.arange(4) + 0.3 * np.sin(mu[0])

    mixed, raw, smooth = OM.mix_song_smooth(model_fn, tracks, chunk_length=cl, sr=sr)
    raw = np.array([raw[k] for k in OM.STEMS])
    smooth = np.array([smooth[k] for k in OM.STEMS])
    assert raw.shape == g['raw'].shape == (4, N // (cl * sr) - 1)
    # the gains are a function of fp32 feature means: two fp32 FFTs (pocketfft vs MKL) differ by
    # O(1e-3 dB) on bins near the floor, which the fake model turns into O(1e-5) relative gain noise
    assert np.max(np.abs(raw / g['raw'] - 1)) < 5e-5
    assert np.max(np.abs(smooth / g['smooth'] - 1)) < 5e-5
    arr = np.stack([mixed[k] for k in OM.STEMS])
    assert str(arr.dtype) == str(g['mixed_dtype']) == 'float64'
    assert np.allclose(arr[:, :, ::13], g['mixed_dec'], rtol=5e-5, atol=1e-12)
    assert np.allclose(OM.stem_sum(mixed)[:, ::13], g['mix_sum_dec'], rtol=5e-5, atol=1e-6)
    # exact arithmetic given the reference's own gains
    arr2 = OM.apply_gains(song, g['smooth'])
    assert np.array_equal(arr2[:, :, ::13], g['mixed_dec'])
    assert np.array_equal(np.abs(arr2).sum(axis=-1), g['mixed_abs_sum'])


def test_dummy_model_matches_reference():
    g = golden('dummy.npz')
    y = OM.db_sum(g['x'])
    assert y.dtype == np.float32
    assert np.max(np.abs(y - g['y'])) < 5e-6


def test_feature_mix_and_grad():
    rng = np.random.default_rng(0)
    x = rng.uniform(-100, 40, size=(3, 4, 17, 5)).astype(np.float32)
    m = rng.normal(size=(3, 4)).astype(np.float32)
    y = OM.feature_mix(x, m)
    assert np.allclose(y, np.einsum('bi,bift->bft', m, x), rtol=1e-5, atol=1e-4)
    go = rng.normal(size=y.shape).astype(np.float32)
    gm, gx = OM.feature_mix_grad(x, m, go)
    eps = 1e-2
    m2 = m.copy(); m2[1, 2] += eps
    num = ((OM.feature_mix(x, m2).astype(np.float64) - y) * go).sum() / eps
    assert abs(num - gm[1, 2]) < 1e-2 * max(1.0, abs(gm[1, 2]))


def test_loudness_linearity_pin():
    """experiments.ipynb:649-678: L(0.8x)-L(x) = -1.938, x1.4 -> +2.923, x0.5 -> -6.021."""
    x = synthetic.make_song(9, 1, 44100 * 3, kind='music', dtype=np.float64)[0].T   # [N, 2]
    L = OL.integrated_loudness(x)
    for gain, delta in ((0.8, -1.938), (1.4, 2.923), (0.5, -6.021)):
        assert abs(OL.integrated_loudness(gain * x) - L - delta) < 1e-3


def test_loudness_bs1770_sine():
    """997 Hz 0 dBFS sine reads -3.01 LKFS (RBJ-form high-pass adds ~-0.04 LU)."""
    sr = 48000
    t = np.arange(sr * 5) / sr
    x = np.sin(2 * np.pi * 997 * t)
    assert abs(OL.integrated_loudness(x, sr) - (-3.01)) < 0.06
    b, a = OL.k_weighting_filters(48000)[0]
    assert np.allclose(b, [1.53512485958697, -2.69169618940638, 1.19839281085285], atol=2e-4)
    assert np.allclose(a, [1.0, -1.69065929318241, 0.73248077421585], atol=2e-4)


def test_bus_sum_and_normalise():
    rng = np.random.default_rng(1)
    st = rng.normal(size=(8, 2, 50)).astype(np.float32)
    grp = [0, 1, 1, 2, 3, 3, 3, -1]
    b = OM.bus_sum(st, grp)
    assert b.dtype == np.float32 and np.array_equal(b[1], st[1] + st[2])
    assert np.array_equal(b[3], (st[4] + st[5]) + st[6])
    y = OM.normalize_loudness(st[0].astype(np.float64), -23.0, -20.0)
    assert np.allclose(y, st[0] * 10 ** (3 / 20))
    p = OM.peak_normalize(st[0])
    assert np.allclose(np.abs(p).max(axis=1), 1.0)


def _loudness_cases(sr=44100):
    cases = {}
    for kind in ('white', 'music', 'tonal', 'pink'):
        cases[kind] = synthetic.make_song(7, 1, sr * 6, kind=kind, dtype=np.float64)[0].T      # [N, 2]
    g = cases['music'].copy()
    g[sr * 2:sr * 4] *= 1e-5                     # a near-silent stretch: blocks under the -70 LUFS absolute gate
    cases['abs_gated'] = g
    q = cases['white'].copy() * 0.003
    q[:sr] *= 300                                # 50 dB level step: the quiet part falls under the relative gate
    cases['rel_gated'] = q
    cases['mono'] = cases['pink'][:, :1]
    return cases


def test_loudness_matches_torchaudio():
    """Independent pin for oracle/loudness.py: torchaudio.functional.loudness (installed here; its docstring
    states it follows pyloudnorm's implementation of ITU-R BS.1770-4) on seven signal classes, two of
    which exercise the absolute and the relative gate.  Agreement is at rounding level (~1e-14 LU)."""
    torch = pytest.importorskip('torch')
    torchaudio = pytest.importorskip('torchaudio')
    sr = 44100
    for name, x in _loudness_cases(sr).items():
        ours = OL.integrated_loudness(x, sr)
        theirs = float(torchaudio.functional.loudness(torch.from_numpy(np.ascontiguousarray(x.T)), sr))
        assert abs(ours - theirs) <= 1e-9, (name, ours, theirs)
    # the gated cases really gate: the ungated mean of the block energies reads differently
    for name in ('abs_gated', 'rel_gated'):
        z = OL.block_energies(_loudness_cases(sr)[name], sr)
        ungated = -0.691 + 10 * np.log10(np.sum(z.mean(axis=1)))
        assert abs(ungated - OL.gated_loudness(z)) > 0.05, name
