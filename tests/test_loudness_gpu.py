"""K4 (dam_kweight_segments + host gating) against the oracle's pyloudnorm restatement.  GPU only.
Tolerance: 1e-8 LU on integrated loudness, 1e-9 relative on block energies (float64 filters with
an exact chunked state hand-off; only the summation order differs from scipy.lfilter + np.sum)."""
import numpy as np
import pytest
import torch

from oracle import loudness as OL
from deep_audio_mixer_b200 import synthetic
from deep_audio_mixer_b200.loudness import Meter

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('kind,n,dtype', [('music', 44100 * 5, np.float32), ('white', 44100 * 3 + 17, np.float64),
                                          ('tonal', 44100 * 7 + 255, np.float32), ('music', 17641, np.float32),
                                          ('pink', 44100 * 30, np.float32)])
def test_block_energies_and_loudness_match_oracle(cuda_device, kind, n, dtype):
    x = synthetic.make_song(31, 1, n, kind=kind, dtype=dtype)[0]          # [2, n]
    meter = Meter(44100)
    z = meter.block_energies(torch.from_numpy(x).cuda())
    z_ref = OL.block_energies(x.T, 44100)
    assert z.shape == z_ref.shape
    assert np.max(np.abs(z - z_ref) / (np.abs(z_ref) + 1e-300)) < 1e-9
    L = meter.integrated_loudness(x.T)
    assert abs(L - OL.integrated_loudness(x.T, 44100)) < 1e-8


def test_mono_and_48k_and_interleaved_device_input(cuda_device):
    sr = 48000
    t = np.arange(sr * 5) / sr
    x = np.sin(2 * np.pi * 997 * t)
    meter = Meter(sr)
    L = meter.integrated_loudness(x)
    assert abs(L - OL.integrated_loudness(x, sr)) < 1e-8
    assert abs(L - (-3.01)) < 0.06                                        # BS.1770 conformance tone
    st = synthetic.make_song(32, 1, 44100 * 2, kind='music', dtype=np.float64)[0]
    inter = torch.from_numpy(np.ascontiguousarray(st.T)).cuda()           # [N, 2] contiguous: strided rows
    assert abs(Meter(44100).integrated_loudness(inter) - OL.integrated_loudness(st.T, 44100)) < 1e-8


def test_loudness_linearity_pin(cuda_device):
    """experiments.ipynb:649-678: x0.8 -> -1.938 LU, x1.4 -> +2.923 LU, x0.5 -> -6.021 LU."""
    x = synthetic.make_song(9, 1, 44100 * 3, kind='music', dtype=np.float64)[0].T
    meter = Meter(44100)
    L = meter.integrated_loudness(x)
    for gain, delta in ((0.8, -1.938), (1.4, 2.923), (0.5, -6.021)):
        assert abs(meter.integrated_loudness(gain * x) - L - delta) < 1e-3


def test_batch_of_stems_and_gating(cuda_device):
    song = synthetic.make_song(33, 4, 44100 * 6, kind='music', dtype=np.float32)
    song[2, :, 44100:3 * 44100] = 0.0                                      # a silent stretch: gating acts
    meter = Meter(44100)
    got = meter.integrated_loudness_batch(torch.from_numpy(song).cuda())
    ref = [OL.integrated_loudness(song[i].T, 44100) for i in range(4)]
    assert np.max(np.abs(np.array(got) - np.array(ref))) < 1e-8
    with pytest.raises(ValueError):
        meter.integrated_loudness(np.zeros((1000, 2)))
    assert meter.integrated_loudness(np.zeros((44100, 2))) == -np.inf


@pytest.mark.parametrize('n,rate', [(44100 * 3 + 17, 44100), (48000 * 2, 48000), (17640, 44100), (44100 * 11, 44100)])
def test_kweight_blocks_one_call_entry_point(cuda_device, n, rate):
    """dam_kweight_blocks (block geometry evaluated inside the library) == the oracle's block energies."""
    from oracle import loudness as OL
    from deep_audio_mixer_b200.loudness import Meter
    rng = np.random.default_rng(n)
    x = rng.normal(scale=0.2, size=(3, n))
    z = Meter(rate).block_energies_device(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = np.stack([OL.block_energies(x[r][:, None], rate)[0] for r in range(3)])
    assert z.shape == ref.shape
    assert np.allclose(z, ref, rtol=1e-9, atol=1e-18), np.abs(z / ref - 1).max()
    z2 = Meter(rate).block_energies(torch.from_numpy(x).cuda())
    assert np.allclose(z, z2, rtol=1e-12, atol=0)
    with pytest.raises(Exception):
        Meter(rate).block_energies_device(torch.zeros(1, int(0.3 * rate), device='cuda', dtype=torch.float64))


@pytest.mark.parametrize('n', [256 * 8192 + 300, 256 * 8193 * 2 + 5, 256 * 41344 - 77])
def test_long_rows_scan_over_a_cluster_of_ctas(cuda_device, n):
    """Rows of more than 8 192 chunks (47 s) are scanned by a thread-block cluster of 2 / 4 / 8 CTAs that hand the
    parts' end states over through distributed shared memory: a 4-minute stem (41 344 chunks, 8 parts, the last two
    partial / empty) must give the oracle's loudness like a short one."""
    rng = np.random.default_rng(n % 1000)
    x = (rng.normal(size=(n, 1)) * 0.1 * (1.0 + 0.5 * np.sin(np.arange(n) / 3.0e5))[:, None])
    meter = Meter(44100)
    assert abs(meter.integrated_loudness(x) - OL.integrated_loudness(x, 44100)) < 1e-8
    z = meter.block_energies(torch.from_numpy(np.ascontiguousarray(x.T)).cuda())
    z_ref = OL.block_energies(x, 44100)
    assert np.max(np.abs(z - z_ref) / (np.abs(z_ref) + 1e-300)) < 1e-9


def test_gating_launch_also_returns_the_normalisation_gain(cuda_device):
    """dam_gated_loudness_gain: the factor of evaluation.py:64 from the gating launch equals 10 ** ((target - L) / 20)
    evaluated by torch on the loudness the same launch returned."""
    gen = torch.Generator(device='cuda').manual_seed(11)
    x = (torch.rand((5, 2, 3 * 44100), generator=gen, device='cuda', dtype=torch.float64) - 0.5)
    x *= torch.tensor([1.0, 0.3, 0.05, 2.0, 1e-3], device='cuda', dtype=torch.float64)[:, None, None]
    meter = Meter(44100)
    lufs = meter.integrated_loudness_device(x).clone()
    lufs2, gain = meter.integrated_loudness_device(x, gain_target=-20.0)
    assert torch.equal(lufs2, lufs)
    assert torch.equal(gain, torch.pow(10.0, (-20.0 - lufs) / 20.0))
