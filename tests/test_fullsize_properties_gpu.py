"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run
1.8 GB batches in seconds): linearity, exact power-of-two scaling, batch independence against the
oracle on sampled rows, idempotence, shard independence.  GPU only."""
import numpy as np
import pytest
import torch

import bench
from oracle import features as OF
from oracle import loudness as OL
from deep_audio_mixer_b200 import features as F, mixing, synthetic
from deep_audio_mixer_b200 import pipeline as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def batch():
    dev = torch.device('cuda', 0)
    n = bench.SECONDS * bench.SR
    raw = synthetic.make_batch_torch(4242, (bench.N_SEG, bench.N_STEMS, bench.N_CH, n), dev)
    m = torch.from_numpy(bench.synth_m(bench.N_SEG * (n // bench.CHUNK_SAMPLES))).to(dev)
    pipe = P.SegmentBatchPipeline(bench.N_STEMS, bench.GROUP, bench.CHUNK_SAMPLES, bench.HOP, device=dev)
    return pipe, raw, m


def test_config2_full_batch_pipeline_properties(cuda_device, batch):
    pipe, raw, m = batch
    res = pipe.run(raw, m)
    torch.cuda.synchronize()
    B, S, C, N = raw.shape
    assert res['features'].shape == (B * 5, 4, 1025, 173) and res['mix'].shape == (B, C, N)
    assert torch.isfinite(res['features']).all() and torch.isfinite(res['mix']).all()
    # idempotence: a second run gives the same bits for every stage output that has a fixed summation order
    res2 = pipe.run(raw, m)
    assert torch.equal(res['features'], res2['features'])
    assert torch.equal(res['peak'], res2['peak'])
    assert torch.allclose(res['lufs'], res2['lufs'], rtol=0, atol=1e-9)          # atomics: order may differ
    # every mix is normalised to -20 LUFS (checked with the oracle on two sampled segments)
    for b in (0, B - 1):
        L = OL.integrated_loudness(res['mix'][b].cpu().numpy().T, bench.SR)
        assert abs(L - (-20.0)) < 1e-6
    # shard independence: running half of the batch gives the same bits for those segments (N-GPU == 1-GPU)
    half = pipe.run(raw[B // 2:], m[(B // 2) * 5:])
    assert torch.equal(half['features'], res['features'][(B // 2) * 5:])
    assert torch.equal(half['peak'], res['peak'][B // 2:])
    assert torch.allclose(half['mix'], res['mix'][B // 2:], rtol=1e-12, atol=0)
    # sampled rows against the oracle (features of segment b, chunk k, bus s)
    buses = mixing.bus_sum(raw, bench.GROUP, 4, batch_dims=1)
    win = torch.hann_window(2048)
    for (b, k, s) in ((0, 0, 0), (17, 3, 2), (B - 1, 4, 3)):
        chunk = buses[b, s, :, k * bench.CHUNK_SAMPLES:(k + 1) * bench.CHUNK_SAMPLES].cpu().numpy()
        truth = OF.compute_features(np.mean(chunk, axis=0).astype(np.float64), 2048, bench.HOP, window=win.numpy())
        got = res['features'][b * 5 + k, s].cpu().numpy().astype(np.float64)
        amp_t, amp_g = 10.0 ** (truth / 20), 10.0 ** (got / 20)
        assert np.max(np.abs(amp_g - amp_t) / amp_t.max(axis=0, keepdims=True)) <= 1e-6


def test_full_size_linearity(cuda_device, batch):
    pipe, raw, m = batch
    B, S, C, N = raw.shape
    buses = mixing.bus_sum(raw, bench.GROUP, 4, batch_dims=1)
    # bus sum is exactly linear under power-of-two scaling
    assert torch.equal(mixing.bus_sum(raw * 0.5, bench.GROUP, 4, batch_dims=1), buses * 0.5)
    # features: feat(g x) = feat(x) + 20 log10 g (experiments.ipynb:300-353), exact for g = 2^-k in fp32
    f0 = pipe.features(buses)
    f1 = pipe.features(buses * 0.25)
    above = f1 > -99.0                                             # away from the -100 dB clamp
    d = (f0 - f1)[above]
    assert torch.max(torch.abs(d - 20 * np.log10(4.0))) < 2e-5
    del f0, f1, d, above
    # mix: doubling every gain doubles the mix exactly (float64 products, power of two)
    g = torch.pow(10.0, 0.5 * m.float()).double().reshape(B, 5, 4).permute(0, 2, 1).contiguous()
    r1 = mixing.gain_mix(buses, g, want_stats=True)
    r2 = mixing.gain_mix(buses, 2.0 * g, want_stats=True)
    assert torch.equal(r2['mix'], 2.0 * r1['mix']) and torch.equal(r2['peak'], 2.0 * r1['peak'])
    # loudness: L(g x) = L(x) + 20 log10 g (experiments.ipynb:649-678) on all 64 mixes at once
    L1 = pipe.meter.integrated_loudness_device(r1['mix'])
    L2 = pipe.meter.integrated_loudness_device(r2['mix'])
    assert torch.max(torch.abs(L2 - L1 - 20 * np.log10(2.0))) < 1e-9
    # checksum of checksums: sum of squares reported by K2 equals a float64 torch reduction
    ss = (r1['mix'] ** 2).sum(dim=-1)
    assert torch.allclose(r1['sumsq'], ss, rtol=1e-10)


def test_config5_largest_song_mixdown_properties(cuda_device):
    """BASELINE config 5 at its largest point: 32 stems x 600 s stereo float32 (6.8 GB) through K2 in float64.
    Unit gains reproduce the plain stem sum, doubling every gain doubles the mix bit for bit, a gain of zero
    on all stems but one returns that stem, and peak / sum of squares agree with the written mix; sampled
    windows are compared with the oracle's mask + multiply + sum."""
    from oracle import mix as OM
    S, C, N = 32, 2, 600 * 44100
    free, _ = torch.cuda.mem_get_info()
    if free < 12e9:
        pytest.skip('needs 12 GB of free device memory')
    g = torch.Generator(device='cuda').manual_seed(5)
    stems = torch.empty((S, C, N), device='cuda', dtype=torch.float32)
    for s in range(S):
        stems[s].normal_(0, 0.05, generator=g)
    K = 299                                                     # floor(L / 2) - 1 chunks
    gains = torch.rand((S, K), device='cuda', dtype=torch.float64, generator=g) + 0.5
    r1 = mixing.gain_mix(stems, gains, want_mix=True, want_stats=True)
    r2 = mixing.gain_mix(stems, gains * 2.0, want_mix=True, want_stats=True)
    assert torch.equal(r2['mix'], r1['mix'] * 2.0)
    assert torch.equal(r2['peak'], r1['peak'] * 2.0)
    assert torch.equal(r1['peak'], r1['mix'].abs().amax(dim=1))
    assert torch.allclose(r1['sumsq'], (r1['mix'] ** 2).sum(dim=1), rtol=1e-10)
    ones = torch.ones((S, 1), device='cuda', dtype=torch.float64)
    plain = mixing.gain_mix(stems, ones, hold_single=True, want_mix=True)['mix']
    acc = stems[0].double()
    for s_ in range(1, S):
        acc = acc + stems[s_].double()                          # the reference's order: stems added one by one
    assert torch.equal(plain, acc)
    del acc
    assert torch.equal(mixing.stem_sum(stems[:4]), ((stems[0] + stems[1]) + stems[2]) + stems[3])   # float32 path
    one_hot = torch.zeros((S, 1), device='cuda', dtype=torch.float64)
    one_hot[7] = 1.0
    assert torch.equal(mixing.gain_mix(stems, one_hot, hold_single=True, want_mix=True)['mix'], stems[7].double())
    # oracle on sampled windows (the gain boundaries of the first / a middle / the last chunk included)
    coef = int(N / K)
    gh = gains.cpu().numpy()
    for lo in (0, 17 * coef - 1000, (K - 1) * coef - 500, N - 4000):
        hi = min(N, lo + 4000)
        ref = np.zeros((C, hi - lo))
        for s in range(S):
            idx = np.minimum(np.arange(lo, hi) // coef, K - 1)
            ref = ref + stems[s, :, lo:hi].cpu().numpy().astype(np.float64) * gh[s][idx] if s else \
                stems[s, :, lo:hi].cpu().numpy().astype(np.float64) * gh[s][idx]
        assert np.array_equal(r1['mix'][:, lo:hi].cpu().numpy(), ref), lo
    full_mask = OM.interpolate_mask(gh[3], N)                  # the reference's mask over the whole song, one stem
    assert np.array_equal(full_mask[::9973], gh[3][np.minimum(np.arange(0, N, 9973) // coef, K - 1)])


def test_config3_shard_featurisation_properties(cuda_device):
    """BASELINE config 3, one GPU's shard of the bulk featurisation (2928 five-second rows, T = 216): the
    gain <-> dB-offset property on every row, batch independence (a row alone == the row inside the batch),
    the statistics the all-reduce carries, and sampled rows against the oracle."""
    rows, n, hop = 122 * 8 * 240 // 5 // 8, 220500, 1024
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.empty((rows, n), device='cuda', dtype=torch.float32)
    x.normal_(0, 0.1, generator=g)
    stats = torch.zeros(3, dtype=torch.float64, device='cuda')
    f0 = F.stft_db(x, hop=hop, stats=stats)
    assert f0.shape == (rows, 1025, 216)
    f1 = F.stft_db(x * 0.125, hop=hop)
    sel = f1 > -99.0
    assert torch.max(torch.abs((f0 - f1)[sel] - 20 * np.log10(8.0))) < 2e-5
    s = stats.cpu().numpy()
    assert s[2] == f0.numel()
    assert abs(s[0] / s[2] - float(f0.double().mean())) < 1e-6
    for r in (0, 1234, rows - 1):
        alone = F.stft_db(x[r:r + 1], hop=hop)
        assert torch.equal(alone[0], f0[r])
        truth = OF.compute_features(x[r].cpu().numpy().astype(np.float64), 2048, hop,
                                    window=torch.hann_window(2048).numpy())
        amp_t, amp_g = 10.0 ** (truth / 20), 10.0 ** (f0[r].cpu().numpy().astype(np.float64) / 20)
        assert np.max(np.abs(amp_g - amp_t) / amp_t.max(axis=0, keepdims=True)) <= 1e-6
    # PCM16 ingest of the same shard: x/32768 happens in the kernel, features equal those of the float copy
    q = (x[:64] * 32768.0).round().clamp(-32768, 32767).to(torch.int16)
    assert torch.equal(F.stft_db(q, hop=hop), F.stft_db(q.float() / 32768.0, hop=hop))
