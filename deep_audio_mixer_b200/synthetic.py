"""Deterministic synthetic multitracks; no real MedleyDB/MUSDB audio
is available, so every test and benchmark runs on these.

numpy generator (host, small sizes, parity tests): PCG64(seed), each stem = pink-ish noise
(white noise shaped 1/sqrt(f)) + 8 harmonics of a per-stem f0 with a 2 Hz amplitude envelope,
scaled to -20 dBFS RMS +- U(-6, +6) dB; stereo = two decorrelated noise draws sharing the tonal
part; float32 planar [S, 2, N].  ``kind='white'`` is the equal-level white-noise variant.

torch generator (device, benchmark sizes): same recipe with torch.Generator on the target
device so multi-GB inputs never cross PCIe.
"""
import numpy as np

F0 = (55.0, 110.0, 220.0, 440.0, 82.5, 165.0, 330.0, 660.0)


def _pink(rng, n):
    w = rng.standard_normal(n)
    spec = np.fft.rfft(w)
    f = np.arange(spec.shape[0], dtype=np.float64)
    f[0] = 1.0
    spec /= np.sqrt(f)
    y = np.fft.irfft(spec, n)
    return y / np.sqrt(np.mean(y * y))


def make_song(seed: int, n_stems: int = 4, n_samples: int = 44100 * 4, sr: int = 44100,
              kind: str = 'music', dtype=np.float32) -> np.ndarray:
    """Returns planar [S, 2, N] audio in ``dtype``.  kind in {'music', 'white', 'tonal', 'pink'}."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(n_samples, dtype=np.float64) / sr
    out = np.empty((n_stems, 2, n_samples), dtype=np.float64)
    for s in range(n_stems):
        level_db = -20.0 if kind == 'white' else -20.0 + rng.uniform(-6.0, 6.0)
        tonal = np.zeros(n_samples)
        if kind in ('music', 'tonal'):
            f0 = F0[s % len(F0)] * (1.0 + 0.01 * (s // len(F0)))
            env = 0.5 * (1.0 + np.sin(2 * np.pi * 2.0 * t + rng.uniform(0, 2 * np.pi)))
            for h in range(1, 9):
                tonal += np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi)) / h
            tonal *= env
            tonal /= np.sqrt(np.mean(tonal * tonal)) + 1e-30
        for c in range(2):
            if kind == 'white':
                x = rng.standard_normal(n_samples)
            elif kind == 'tonal':
                x = tonal.copy()
            elif kind == 'pink':
                x = _pink(rng, n_samples)
            else:
                x = 0.7 * _pink(rng, n_samples) + 0.7 * tonal
            x = x / (np.sqrt(np.mean(x * x)) + 1e-30) * 10.0 ** (level_db / 20.0)
            out[s, c] = x
    return out.astype(dtype)


def quantise_int16(x: np.ndarray) -> np.ndarray:
    """round(x * 32768) clipped to int16: the int16 ingest path's input."""
    return np.clip(np.round(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)


def config_seed(config: int, song_index: int) -> int:
    return 1234 + 1000 * config + song_index


def make_batch_torch(seed: int, shape, device, sr: int = 44100, dtype=None):
    """Device-side generator for benchmark-sized inputs: [..., N] float32 'music-like' audio
    (white noise low-passed by a one-pole cumulative filter approximated with a short FIR, plus
    harmonics), -20 dBFS RMS.  Deterministic for a given (seed, shape, device type)."""
    import torch
    dtype = dtype or torch.float32
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = shape[-1]
    lead = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
    out = torch.empty((lead, n), device=device, dtype=dtype)
    t = torch.arange(n, device=device, dtype=torch.float32) / sr
    step = max(1, min(lead, (1 << 26) // max(n, 1)))
    for i0 in range(0, lead, step):
        i1 = min(lead, i0 + step)
        x = torch.randn((i1 - i0, n), device=device, dtype=torch.float32, generator=g)
        # cheap spectral tilt: x[n] + 0.9 x[n-1] + 0.7 x[n-2] + 0.5 x[n-3]
        y = x.clone()
        y[:, 1:] += 0.9 * x[:, :-1]
        y[:, 2:] += 0.7 * x[:, :-2]
        y[:, 3:] += 0.5 * x[:, :-3]
        rows = torch.arange(i0, i1, device=device, dtype=torch.float32)
        f0 = 55.0 * torch.pow(2.0, (rows % 4))
        ph = 2 * np.pi * f0[:, None] * t[None, :]
        tone = torch.sin(ph) + 0.5 * torch.sin(2 * ph) + 0.33 * torch.sin(3 * ph)
        y = 0.05 * y + 0.05 * tone * (0.5 + 0.5 * torch.sin(2 * np.pi * 2.0 * t))[None, :]
        out[i0:i1] = y.to(dtype)
    return out.reshape(shape)
