"""K2 / N2 / K3 / K5 / K6 against the oracle and the reference goldens.  GPU only.
Integer/index work and float64 mask arithmetic are checked bit-exactly."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import mix as OM
from deep_audio_mixer_b200 import mixing, synthetic, inference_utils as IU

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('S,C,N,K', [(4, 2, 10000, 14), (4, 2, 10001, 7), (3, 1, 999, 999), (2, 2, 4096, 1),
                                     (8, 2, 12345, 29), (4, 2, 64, 63), (5, 3, 1000, 2), (32, 2, 20000, 9)])
def test_gain_mix_float64_mask_path_bit_exact(cuda_device, S, C, N, K):
    rng = np.random.default_rng(S * 1000 + K)
    stems = rng.normal(size=(S, C, N)).astype(np.float32)
    gains = rng.uniform(0.2, 1.8, size=(S, K))
    ref_stems = OM.apply_gains(stems, gains)                      # float64, K == 1 -> zeros
    ref_mix = OM.stem_sum(ref_stems)
    res = mixing.gain_mix(cuda(stems), gains, want_mix=True, want_stems=True, want_stats=True)
    assert res['stems'].dtype == torch.float64 and res['mix'].dtype == torch.float64
    assert np.array_equal(res['stems'].cpu().numpy(), ref_stems)
    assert np.array_equal(res['mix'].cpu().numpy(), ref_mix)
    assert np.allclose(res['peak'].cpu().numpy(), np.abs(ref_mix).max(axis=1), rtol=0, atol=0)
    assert np.allclose(res['sumsq'].cpu().numpy(), (ref_mix ** 2).sum(axis=1), rtol=1e-12)
    # mix only (the pipeline's call): without the per-stem outputs the compiler is free to contract the product
    # into the sum -- the kernel must still round the product first, like numpy's `track * mask`
    only = mixing.gain_mix(cuda(stems), gains, want_mix=True)
    assert np.array_equal(only['mix'].cpu().numpy(), ref_mix)
    only32 = mixing.gain_mix(cuda(stems), gains[:, :1], hold_single=True, compute_dtype=torch.float32, want_mix=True)
    ref32 = OM.stem_sum(stems * gains[:, :1, None].astype(np.float32))
    assert np.array_equal(only32['mix'].cpu().numpy(), ref32)


def test_gain_mix_float64_stems_and_unaligned_views(cuda_device):
    rng = np.random.default_rng(3)
    stems = rng.normal(size=(4, 2, 5003))
    gains = rng.uniform(0.5, 1.5, size=(4, 11))
    t = cuda(stems)[:, :, 1:5002]                                 # odd offset: scalar path
    ref = OM.apply_gains(stems[:, :, 1:5002], gains)
    res = mixing.gain_mix(t, gains, want_stems=True)
    assert np.array_equal(res['stems'].cpu().numpy(), ref)
    assert np.array_equal(res['mix'].cpu().numpy(), OM.stem_sum(ref))


def test_gain_mix_matches_reference_golden_mix(cuda_device):
    g = golden('mix_smooth.npz')
    N = int(g['N'])
    song = synthetic.make_song(int(g['seed']), 4, N, sr=int(g['sr']), kind='music', dtype=np.float32)
    assert hashlib.sha256(song.tobytes()).hexdigest() == str(g['input_sha'])
    res = mixing.gain_mix(cuda(song), g['smooth'], want_stems=True)
    st = res['stems'].cpu().numpy()
    assert np.array_equal(st[:, :, ::13], g['mixed_dec'])
    assert np.array_equal(np.abs(st).sum(axis=-1), g['mixed_abs_sum'])
    assert np.array_equal(res['mix'].cpu().numpy()[:, ::13], g['mix_sum_dec'])


def test_scalar_gain_float32_path_is_random_model(cuda_device):
    """RandomModel: python float x float32 array -> float32; np.sum of the float32 stack."""
    rng = np.random.default_rng(4)
    stems = rng.normal(size=(4, 2, 30000)).astype(np.float32)
    gains = [float(rng.uniform(0.5, 1.5)) for _ in range(4)]
    ref = np.stack([g * stems[i] for i, g in enumerate(gains)])
    assert ref.dtype == np.float32
    res = mixing.gain_mix(cuda(stems), [[g] for g in gains], hold_single=True, compute_dtype=torch.float32,
                          want_stems=True)
    assert res['stems'].dtype == torch.float32
    assert np.array_equal(res['stems'].cpu().numpy(), ref)
    assert np.array_equal(res['mix'].cpu().numpy(), np.sum(ref, axis=0))
    assert np.array_equal(mixing.stem_sum(cuda(stems)).cpu().numpy(), np.sum(stems, axis=0))


def test_interpolate_mask_dropin_matches_reference(cuda_device):
    g = golden('mask.npz')
    for key in g.files:
        if key.startswith('g_'):
            _, K, L = key.split('_')
            m = IU.interpolate_mask(g[key], int(L))
            assert m.dtype == np.float64 and np.array_equal(m, g['m_%s_%s' % (K, L)])
    with pytest.raises(AssertionError):
        IU.interpolate_mask(np.ones(5), 4)


def test_bus_sum(cuda_device):
    rng = np.random.default_rng(5)
    st = rng.normal(size=(8, 2, 10001)).astype(np.float32)
    grp = [0, 1, 1, 2, 3, 3, 3, -1]
    out = mixing.bus_sum(cuda(st), grp, 4)
    assert np.array_equal(out.cpu().numpy(), OM.bus_sum(st, grp, 4))
    # PCM16 ingest: x/32768 in float32 (what librosa.load yields), then the same float32 bus sums
    q = synthetic.quantise_int16(st * 0.2)
    as_f32 = q.astype(np.float32) * np.float32(1.0 / 32768.0)
    out16 = mixing.bus_sum(cuda(q), grp, 4)
    assert out16.dtype == torch.float32 and np.array_equal(out16.cpu().numpy(), OM.bus_sum(as_f32, grp, 4))
    qb = np.stack([q, q[::-1]])[:, :, :, :10000]                   # batched + multiple-of-8 vector path
    outb = mixing.bus_sum(cuda(qb), grp, 4, batch_dims=1).cpu().numpy()
    for b in range(2):
        assert np.array_equal(outb[b], OM.bus_sum(qb[b].astype(np.float32) * np.float32(1.0 / 32768.0), grp, 4))
    st64 = rng.normal(size=(5, 2, 4096))
    out = mixing.bus_sum(cuda(st64), [3, 3, 0, 0, 0], 4)
    assert np.array_equal(out.cpu().numpy(), OM.bus_sum(st64, [3, 3, 0, 0, 0], 4))


@pytest.mark.parametrize('shape', [(3, 4, 1025, 87), (2, 4, 1025, 173), (1, 4, 7, 5)])
def test_feature_mix_forward_bit_exact_and_backward(cuda_device, shape):
    rng = np.random.default_rng(6)
    x = rng.uniform(-100, 40, size=shape).astype(np.float32)
    m = rng.normal(size=shape[:2]).astype(np.float32)
    xt, mt = cuda(x), cuda(m).requires_grad_(True)
    y = mixing.feature_mix(xt, mt)
    assert np.array_equal(y.detach().cpu().numpy(), OM.feature_mix(x, m))
    go = rng.normal(size=y.shape).astype(np.float32)
    y.backward(cuda(go))
    gm, gx = OM.feature_mix_grad(x, m, go)
    assert np.allclose(mt.grad.cpu().numpy(), gm, rtol=2e-5, atol=1e-3 * np.abs(gm).max() * 1e-2)
    xt2 = cuda(x).requires_grad_(True)
    mixing.feature_mix(xt2, cuda(m)).backward(cuda(go))
    assert np.array_equal(xt2.grad.cpu().numpy(), gx.astype(np.float32))
    # the reference formulation in torch ops gives the same forward bits
    ref = torch.zeros_like(xt[:, 0])
    for i in range(4):
        ref += mt.detach()[:, i:i + 1].unsqueeze(2) * xt[:, i]
    assert torch.equal(ref, y.detach())


def test_db_sum_matches_dummy_model_golden(cuda_device):
    g = golden('dummy.npz')
    y = mixing.db_sum(cuda(g['x'])).cpu().numpy()
    assert y.shape == g['y'].shape
    assert np.max(np.abs(y - g['y'])) < 2e-5
    assert np.max(np.abs(y - OM.db_sum(g['x']))) < 2e-5


def test_scale_and_normalise(cuda_device):
    rng = np.random.default_rng(7)
    x = rng.normal(size=(2, 30001))
    y = mixing.normalize_loudness(cuda(x), -23.5, -20.0).cpu().numpy()
    assert np.array_equal(y, OM.normalize_loudness(x, -23.5, -20.0))
    x32 = x.astype(np.float32)
    p = mixing.peak_normalize(cuda(x32)).cpu().numpy()
    ref = OM.peak_normalize(x32)
    assert p.dtype == np.float32 and np.max(np.abs(p - ref)) <= 6e-8
    am, sq = mixing.row_reduce(cuda(x32))
    assert np.array_equal(am.cpu().numpy(), np.abs(x32).max(axis=1).astype(np.float64))
    assert np.allclose(sq.cpu().numpy(), (x32.astype(np.float64) ** 2).sum(axis=1), rtol=1e-12)


def test_no_writes_outside_outputs(cuda_device):
    """Guard-band check (compute-sanitizer is closed on this pool) with odd sizes and offsets."""
    from deep_audio_mixer_b200 import _lib
    rng = np.random.default_rng(11)
    for N in (4096, 4099, 1):
        stems = cuda(rng.normal(size=(3, 2, N)).astype(np.float32))
        res = mixing.gain_mix(stems, rng.uniform(0.5, 1.5, size=(3, max(1, min(5, N)))), want_stems=True)
        ref = mixing.gain_mix(stems.clone(), rng.uniform(0.5, 1.5, size=(3, 1)), hold_single=True)
        assert res['mix'].shape == (2, N) and ref['mix'].shape == (2, N)
        assert torch.isfinite(res['mix']).all() and torch.isfinite(res['stems']).all()
    # scale_rows / row_reduce on an offset, odd-length view
    base = cuda(rng.normal(size=(4, 1003)))
    view = base[:, 3:1000]
    y = mixing.scale_rows(view, [2.0, 3.0, 4.0, 5.0])
    assert np.array_equal(y.cpu().numpy(), view.cpu().numpy() * np.array([[2.0], [3.0], [4.0], [5.0]]))
    am, sq = mixing.row_reduce(view)
    assert np.array_equal(am.cpu().numpy(), np.abs(view.cpu().numpy()).max(axis=1))


def test_scale_scalar_entry_point(cuda_device):
    """dam_scale: host scalar gain, float64 product rounded to the output type (odd sizes and views too)."""
    for n in (1, 7, 4096, 100003):
        for dt in (torch.float32, torch.float64):
            x = torch.randn(n, device='cuda', dtype=dt)
            got = mixing.scale(x, 0.3183098861837907)
            ref = (x.double() * 0.3183098861837907).to(dt)
            assert torch.equal(got, ref)
        x = torch.randn(2, n, device='cuda')
        got = mixing.scale(x, -2.5, out_dtype=torch.float64)
        assert got.dtype == torch.float64 and torch.equal(got, x.double() * -2.5)
    y = mixing.normalize_loudness(torch.ones(2, 1000, device='cuda', dtype=torch.float64), -23.0, -20.0)
    assert torch.allclose(y, torch.full_like(y, 10 ** (3 / 20)), rtol=0, atol=1e-15)


@pytest.mark.parametrize('B,S,C,N,K,st_dtype,out_dtype', [
    (3, 4, 2, 100000, 5, torch.float32, torch.float64),      # last chunk partial (100000 = 390 * 256 + 160)
    (2, 4, 2, 65536, 7, torch.float32, torch.float64),       # whole chunks only
    (2, 3, 1, 88201, 3, torch.float32, torch.float64),       # n % 4 != 0: plain mixdown + the meter's own first pass
    (2, 4, 2, 88200, 4, torch.float32, torch.float32),       # float32 mix: the states follow the STORED values
    (1, 2, 2, 44100, 1, torch.float64, torch.float64),       # float64 stems, one held gain
    (1, 2, 2, 2400000, 9, torch.float32, torch.float64),     # 9 375 chunks per row: scanned by a cluster of two CTAs
    (64, 4, 2, 441000, 5, torch.float32, torch.float64),     # the bench batch
])
def test_gain_mix_with_meter_states_is_bit_identical_to_mixing_then_measuring(cuda_device, B, S, C, N, K, st_dtype, out_dtype):
    """dam_gain_mix_kweight + dam_kweight_segments_resume (the mix read once by the meter) against dam_gain_mix +
    dam_kweight_segments (read twice): same mix, same segment sums, same loudness, to the bit."""
    from deep_audio_mixer_b200.loudness import Meter
    gen = torch.Generator(device='cuda').manual_seed(B * 7 + N)
    stems = (torch.rand((B, S, C, N), generator=gen, device='cuda', dtype=torch.float32) - 0.5).to(st_dtype)
    stems[:, :, :, N // 3:N // 3 + 30000] *= 1e-5                            # a stretch the gate drops
    gains = torch.rand((B, S, K), generator=gen, device='cuda', dtype=torch.float64) + 0.5
    meter = Meter(44100)
    hold = K == 1
    plain = mixing.gain_mix(stems, gains, hold_single=hold, out_dtype=out_dtype, want_stats=True)
    seg_plain = meter._segment_sums(plain['mix'].reshape(B * C, N))[0].clone()
    lufs_plain = meter.integrated_loudness_device(plain['mix']).clone()
    fused = mixing.gain_mix(stems, gains, hold_single=hold, out_dtype=out_dtype, want_stats=True, meter=meter)
    assert plain['kw_workspace'] is None and fused['kw_workspace'] is not None
    assert torch.equal(fused['mix'], plain['mix'])
    assert torch.equal(fused['peak'], plain['peak'])
    assert torch.allclose(fused['sumsq'], plain['sumsq'], rtol=1e-12, atol=0)
    seg_fused = meter._segment_sums(fused['mix'].reshape(B * C, N), fused['kw_workspace'])[0].clone()
    assert torch.equal(seg_fused, seg_plain)
    fused = mixing.gain_mix(stems, gains, hold_single=hold, out_dtype=out_dtype, meter=meter)
    lufs_fused = meter.integrated_loudness_device(fused['mix'], fused['kw_workspace'])
    assert torch.equal(lufs_fused, lufs_plain)
    assert torch.isfinite(lufs_plain).all()
    with pytest.raises(ValueError):
        meter.integrated_loudness_device(fused['mix'][:, :, :N - 4], fused['kw_workspace'])
    # evaluation.py:59-63 as evaluation.sum_and_normalize runs it: stem sum in the audio dtype, host-side gating
    song = stems[0]
    total, kw = mixing.stem_sum(song, meter=meter)
    assert torch.equal(total, mixing.stem_sum(song))
    assert meter.integrated_loudness(total.t(), kw) == meter.integrated_loudness(mixing.stem_sum(song).t())


def test_head_gains_kernel_equals_the_torch_expression_it_replaces(cuda_device):
    """dam_head_gains against torch.pow(10.0, 0.5 * m).double() transposed (what the pipeline ran before): same bits,
    on random head outputs, on a dense sweep of float32 values and on the edge values of the power function."""
    gen = torch.Generator(device='cuda').manual_seed(5)
    B, nc, S = 7, 13, 4
    m = torch.randn((B * nc, S), generator=gen, device='cuda') * 2.0
    ref = torch.pow(10.0, 0.5 * m.float()).double().reshape(B, nc, S).permute(0, 2, 1).contiguous()
    assert torch.equal(mixing.head_gains(m, B, nc), ref)
    sweep = torch.linspace(-12.0, 12.0, 4_000_000, device='cuda').reshape(-1, 4)
    edge = torch.tensor([0.0, -0.0, 1.0, -1.0, 2.0, 80.0, -80.0, 200.0, -200.0, float('inf'), -float('inf'), 1e-30],
                        device='cuda').reshape(-1, 4)
    for mm in (sweep, edge):
        n = mm.shape[0]
        got = mixing.head_gains(mm, n, 1)
        want = torch.pow(10.0, 0.5 * mm).double().reshape(n, 1, 4).permute(0, 2, 1).contiguous()
        assert torch.equal(got, want)
    nan = mixing.head_gains(torch.full((1, 4), float('nan'), device='cuda'), 1, 1)
    assert torch.isnan(nan).all()


def test_chunk_kernel_random_shapes_bit_exact(cuda_device):
    """The persistent chunk kernel (aligned planes, n % 4 == 0) on 40 random shapes: stem counts around the
    four-stems-in-flight grouping, chunk counts around one, gain runs shorter and longer than a chunk, float32 and
    float64 stems -- masked stems and mix equal to the oracle's numpy arithmetic, bit for bit."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        S = int(rng.integers(1, 10))
        C = int(rng.integers(1, 4))
        N = 4 * int(rng.integers(1, 1300))
        K = int(rng.integers(1, min(N, 40) + 1))
        dt = np.float32 if case % 3 else np.float64
        stems = rng.normal(size=(S, C, N)).astype(dt)
        stems[:, :, ::7] *= -1.0
        gains = rng.uniform(0.2, 1.8, size=(S, K))
        ref_stems = OM.apply_gains(stems, gains)
        ref_mix = OM.stem_sum(ref_stems)
        res = mixing.gain_mix(cuda(stems), gains, want_mix=True, want_stems=bool(case & 1), want_stats=True)
        assert np.array_equal(res['mix'].cpu().numpy(), ref_mix), (case, S, C, N, K)
        if case & 1:
            assert np.array_equal(res['stems'].cpu().numpy(), ref_stems), (case, S, C, N, K)
        assert np.array_equal(res['peak'].cpu().numpy(), np.abs(ref_mix).max(axis=-1)), (case, S, C, N, K)
        np.testing.assert_allclose(res['sumsq'].cpu().numpy(), (ref_mix ** 2).sum(axis=-1), rtol=1e-12)
