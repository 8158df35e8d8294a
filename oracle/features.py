"""Oracle: STFT -> magnitude -> dB featurisation (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, in numpy, what the reference computes with library calls:

* ``MultitrackAudioDataset.compute_features``   /root/reference/data/dataset.py:132-162
    torch.stft(audio, n_fft, hop, window=torch.hann_window(n_fft), return_complex=True)
    -> torch.abs -> torchaudio.functional.amplitude_to_DB(multiplier=20, amin=1e-5, db_multiplier=0)
* ``_stereo_to_mono``                            /root/reference/data/dataset.py:181-183
* ``_augment_audio`` / ``_augment_features``     /root/reference/data/dataset.py:164-179
* ``_process_on_the_fly`` (stacking order)       /root/reference/data/dataset.py:185-210

torch.stft defaults that matter (torch 1.6 .. 2.11 identical): center=True,
pad_mode='reflect', onesided=True, normalized=False, win_length=n_fft.
The compute dtype follows the input dtype exactly as in the reference: float64 in
(training path, sf.read) -> complex128 -> float64 out; float32 in (inference path,
librosa.load) -> complex64 -> float32 out.  numpy >= 2.0 runs rfft in single
precision for float32 input, i.e. the same error class as MKL's float32 r2c.
"""
import numpy as np

TRACKLIST = ['bass', 'drums', 'vocals', 'other', 'mix']      # data/dataset.py:42


def hann_window_fp32(n_fft: int) -> np.ndarray:
    """torch.hann_window(n_fft) (periodic, float32) restated op-for-op.

    ATen builds it as arange(n_fft+1, float32).mul_(2*pi/n_fft).cos_().mul_(-0.5).add_(0.5)
    and drops the last element; every step rounds to float32 (data/dataset.py:148 passes this
    float32 table to torch.stft, which promotes it to the signal dtype).
    """
    k = np.arange(n_fft, dtype=np.float32)
    step = np.float32(2.0 * np.pi / n_fft)
    w = np.cos(k * step, dtype=np.float32)
    return (w * np.float32(-0.5) + np.float32(0.5)).astype(np.float32)


def num_frames(n_samples: int, hop: int) -> int:
    """T = 1 + floor(n / hop) for center=True."""
    return 1 + n_samples // hop


def frame_signal(audio: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """Reflect-pad n_fft/2 on both sides and cut frames; returns [..., T, n_fft]."""
    pad = n_fft // 2
    if audio.shape[-1] <= pad:
        raise ValueError('reflect padding needs more than n_fft/2 samples')
    width = [(0, 0)] * (audio.ndim - 1) + [(pad, pad)]
    padded = np.pad(audio, width, mode='reflect')
    T = num_frames(audio.shape[-1], hop)
    idx = np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]
    return padded[..., idx]


def stft(audio: np.ndarray, window: np.ndarray = None, n_fft: int = 2048,
         hop: int = 1024) -> np.ndarray:
    """Complex one-sided STFT, layout [..., F, T] like torch.stft."""
    audio = np.asarray(audio)
    if audio.dtype not in (np.float32, np.float64):
        raise TypeError('audio must be float32 or float64')
    if window is None:
        window = hann_window_fp32(n_fft)
    w = np.asarray(window).astype(audio.dtype)        # fp32 table promoted to signal dtype
    frames = frame_signal(audio, n_fft, hop) * w
    spec = np.fft.rfft(frames, axis=-1)               # [..., T, F]
    return np.swapaxes(spec, -1, -2)


def amplitude_to_db(mag: np.ndarray, multiplier: float = 20.0, amin: float = 1e-5,
                    db_multiplier: float = 0.0) -> np.ndarray:
    """torchaudio.functional.amplitude_to_DB without top_db (functional.py:356 in 2.11)."""
    dt = mag.dtype
    x = np.log10(np.maximum(mag, dt.type(amin)))
    return (x * dt.type(multiplier) - dt.type(multiplier * db_multiplier)).astype(dt)


def compute_features(audio: np.ndarray, window_size: int = 2048, hop_length: int = 1024,
                     window: np.ndarray = None) -> np.ndarray:
    """data/dataset.py:132-162.  [..., n] -> [..., window_size/2+1, T], dtype follows input."""
    spec = stft(audio, window, window_size, hop_length)
    mag = np.abs(spec)
    return np.ascontiguousarray(amplitude_to_db(mag, 20.0, 1e-5, 0.0))


def stereo_to_mono(audio: np.ndarray, channel_axis: int = 1) -> np.ndarray:
    """data/dataset.py:181-183: np.mean over the channel axis ([n, 2] interleaved there)."""
    return np.mean(audio, axis=channel_axis)


def augment_audio(audio: np.ndarray, gain: float) -> np.ndarray:
    """data/dataset.py:164-168 with the random draw made by the caller."""
    return gain * audio


def augment_features(features: np.ndarray, gains: np.ndarray) -> np.ndarray:
    """data/dataset.py:170-179: + 20*log10(g_s) per stem on [S, F, T] dB maps."""
    gains_db = 20 * np.log10(np.asarray(gains))
    return np.add(features, gains_db[:, np.newaxis, np.newaxis])


def process_chunk(stems: np.ndarray, mix: np.ndarray, channel_axis: int = 1,
                  gains=None, window_size: int = 2048, hop_length: int = 1024):
    """data/dataset.py:185-210 on in-memory audio.

    stems: 4 arrays (bass, drums, vocals, other) of stereo audio for one chunk, ``mix`` the
    mixture chunk.  Returns (train[4, F, T], gt[F, T]).
    """
    feats = []
    for i, a in enumerate(list(stems) + [mix]):
        a = np.asarray(a)
        if a.ndim == 2:
            a = stereo_to_mono(a, channel_axis)
        if gains is not None:
            a = augment_audio(a, gains[i])
        feats.append(compute_features(a, window_size, hop_length))
    return np.stack(feats[:4]), feats[4]
