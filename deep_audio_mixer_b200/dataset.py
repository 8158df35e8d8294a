"""Drop-in for the reference's ``MultitrackAudioDataset`` (data/dataset.py:16-304) with the feature
computation on the GPU.

Same constructor, same directory layout ({song}/{song}_MIX.wav, {song}/{song}_STEMS_JOINED/
{song}_STEM_{TRACK}.wav), same chunk indexing, same ``__getitem__`` result -- a tuple
``(train_features [4, 1025, T], gt_features [1025, T])`` in float64 (sf.read yields float64 and the
features follow the input dtype) -- and the same ``compute_features`` / ``compute_mean_loudness``
methods.  What differs is where the arithmetic runs: the five tracks of a chunk are featurised by
one K1 launch (stereo -> mono and the augmentation gain are fused into its staging), and
``get_batch`` featurises many chunks per launch for the training loop.

CUDA cannot be used in forked DataLoader workers: use ``num_workers=0`` with
``get_batch`` / ``batches()``, or a 'spawn' context.
"""
import os
import random
from statistics import mean

import numpy as np
import torch
from torch.utils import data

from . import features, wavio
from .loudness import Meter


class MultitrackAudioDataset(data.Dataset):
    def __init__(self, base_path: str, songlist: list = None, chunk_length: int = 5,
                 sr: int = 44100, seed: int = None, normalize: bool = False,
                 compute_features: bool = True, augment_data: bool = False,
                 device=None, return_device: bool = False):
        self._base_path = base_path
        self._chunk_length = chunk_length
        self._normalize = normalize
        self._compute_features = compute_features
        self._augment = augment_data
        self._sr = sr
        self._tracklist = ['bass', 'drums', 'vocals', 'other', 'mix']
        self._device = device
        self._return_device = return_device

        if not songlist:
            songlist = [song_name for song_name in os.listdir(self._base_path) if
                        os.path.isdir(os.path.join(self._base_path, song_name))]
        self.songlist = songlist
        if seed:
            random.seed(seed)
        random.shuffle(songlist)
        self._len, self.song_durations = self._calculate_dataset_length()

    # ---- bookkeeping, identical to the reference (data/dataset.py:56-113) ---------------------
    def _calculate_dataset_length(self) -> tuple:
        dataset_len = 0
        song_durations = []
        for song_name in self.songlist:
            track_path = os.path.join(self._base_path, song_name, '{}_MIX.wav'.format(song_name))
            song_duration = int(wavio.info(track_path).duration)
            song_durations.append(song_duration - (song_duration % self._chunk_length))
            dataset_len += int(song_duration / self._chunk_length)
        return dataset_len, song_durations

    def _get_track_path(self, song_name, track_name):
        song_path = os.path.join(self._base_path, song_name)
        if track_name == 'mix':
            return os.path.join(song_path, '{}_MIX.wav'.format(song_name))
        return os.path.join(song_path, '{}_STEMS_JOINED'.format(song_name),
                            '{}_STEM_{}.wav'.format(song_name, track_name.upper()))

    def _calculate_song_index(self, chunk_i: int) -> tuple:
        song_i = 0
        num_chunks_in_song = int(self.song_durations[song_i] / self._chunk_length)
        while chunk_i >= num_chunks_in_song and song_i < len(self.songlist) - 1:
            chunk_i -= num_chunks_in_song
            song_i += 1
            num_chunks_in_song = int(self.song_durations[song_i] / self._chunk_length)
        return song_i, chunk_i

    def _dev(self):
        return torch.device(self._device) if self._device is not None else \
            torch.device('cuda', torch.cuda.current_device())

    # ---- features ---------------------------------------------------------------------------
    def compute_features(self, audio, window_size: int = 2048, hop_length: int = 1024):
        """data/dataset.py:132-162 on the GPU (dtype follows the input; host in -> host out)."""
        return features.compute_features(audio, window_size, hop_length, device=self._device)

    @staticmethod
    def _augment_audio(audio, gain_from: float = 0.6, gain_to: float = 1.4):
        return np.random.uniform(gain_from, gain_to) * audio

    @staticmethod
    def _augment_features(feats, gain_from: float = 0.6, gain_to: float = 1.4):
        random.seed(None)
        gains = np.random.uniform(gain_from, gain_to, size=len(feats))
        gains_db = 20 * np.log10(gains)
        if isinstance(feats, torch.Tensor):
            return feats + torch.as_tensor(gains_db, dtype=feats.dtype, device=feats.device)[:, None, None]
        return np.add(feats, gains_db[:, np.newaxis, np.newaxis])

    @staticmethod
    def _stereo_to_mono(audio: np.ndarray) -> np.array:
        return np.mean(audio, axis=1)

    def _read_chunk(self, song_i: int, chunk_i: int) -> np.ndarray:
        """[5, n, C] float64: the chunk of every track, as soundfile would return it."""
        song_name = self.songlist[song_i]
        a = chunk_i * self._chunk_length * self._sr
        b = (chunk_i + 1) * self._chunk_length * self._sr
        chunks = []
        for track_name in self._tracklist:
            audio, _ = wavio.read(self._get_track_path(song_name, track_name), start=a, stop=b)
            if audio.ndim == 1:
                raise IndexError('tracks are expected to be stereo (data/dataset.py:195)')
            chunks.append(audio)
        return np.stack(chunks)

    def _read_chunk_device(self, song_i: int, chunk_i: int) -> torch.Tensor:
        """The chunk of every track on the device, [5, n, C], in the narrowest form that still yields the bits of
        soundfile's float64 samples (N1): PCM16 files are uploaded as int16 (2 bytes per sample instead of 8; K1
        scales by 1/32768 and averages the channels exactly), PCM24 files as their 3-byte samples and decoded by
        dam_pcm_unpack; anything else goes up as the float64 array ``_read_chunk`` returns."""
        song_name = self.songlist[song_i]
        a = chunk_i * self._chunk_length * self._sr
        b = (chunk_i + 1) * self._chunk_length * self._sr
        paths = [self._get_track_path(song_name, t) for t in self._tracklist]
        infos = [wavio.info(p) for p in paths]
        widths = {(wi.sampwidth, wi.is_float, wi.channels) for wi in infos}
        if len(widths) == 1 and not infos[0].is_float and infos[0].sampwidth in (2, 3):
            sw, _, ch = next(iter(widths))
            if ch == 1:
                raise IndexError('tracks are expected to be stereo (data/dataset.py:195)')
            raw = np.stack([wavio.read_bytes(p, a, b)[0] for p in paths])          # [5, n * C * sw] uint8
            dev_raw = torch.from_numpy(raw).to(self._dev(), non_blocking=True)
            n = raw.shape[1] // (sw * ch)
            if sw == 2:
                return dev_raw.view(torch.int16).reshape(5, n, ch)
            from . import mixing
            return mixing.pcm_unpack(dev_raw.reshape(-1), 24, torch.float64, shape=(5, n, ch))
        return torch.from_numpy(self._read_chunk(song_i, chunk_i)).to(self._dev(), non_blocking=True)

    def _featurise(self, audio: torch.Tensor, gains=None) -> torch.Tensor:
        """audio [R, n, C] interleaved on the device (float64, or int16 PCM) -> [R, 1025, T] float64."""
        g = None if gains is None else torch.as_tensor(gains, dtype=torch.float64)
        return features.stft_db(audio, channel_axis=-1, row_gain=g, compute_dtype=torch.float64, out_dtype=torch.float64)

    def _process_on_the_fly(self, song_i: int, chunk_i: int) -> tuple:
        """data/dataset.py:185-210: five tracks, mono down-mix, optional gain, features; one launch."""
        audio = self._read_chunk_device(song_i, chunk_i)
        gains = [np.random.uniform(0.6, 1.4) for _ in self._tracklist] if self._augment else None
        feats = self._featurise(audio, gains)
        train, gt = feats[:4], feats[4]
        if not self._return_device:
            train, gt = train.cpu(), gt.cpu()
        return train, gt

    def get_batch(self, indices) -> tuple:
        """Featurise many chunks in one launch: returns (train [B, 4, 1025, T], gt [B, 1025, T]) on
        the device.  This is the path a training loop should use instead of forked workers."""
        chunks = [self._read_chunk_device(*self._calculate_song_index(i)) for i in indices]    # B x [5,n,C]
        B = len(chunks)
        if len({c.dtype for c in chunks}) > 1:
            chunks = [c.double() / 32768.0 if c.dtype == torch.int16 else c for c in chunks]
        dev_audio = torch.stack(chunks).reshape(B * 5, *chunks[0].shape[1:])
        gains = np.random.uniform(0.6, 1.4, size=B * 5) if self._augment else None
        feats = self._featurise(dev_audio, gains)                                             # [B*5, F, T]
        feats = feats.reshape(B, 5, feats.shape[-2], feats.shape[-1])
        return feats[:, :4].contiguous(), feats[:, 4].contiguous()

    def batches(self, batch_size: int, drop_last: bool = False):
        for i in range(0, len(self), batch_size):
            idx = list(range(i, min(i + batch_size, len(self))))
            if drop_last and len(idx) < batch_size:
                return
            yield self.get_batch(idx)

    def _process_precomputed(self, song_i, chunk_i) -> tuple:
        """data/dataset.py:253-268 (reads the names the reference READS; its writer uses others)."""
        song_name = self.songlist[song_i]
        features_dir = os.path.join(self._base_path, song_name, '{}_FEATURES'.format(song_name))
        suffix = '_norm' if self._normalize else ''
        gt = np.load(os.path.join(features_dir, '{}_gt{}.npy'.format(chunk_i, suffix)))
        train = np.load(os.path.join(features_dir, '{}_train{}.npy'.format(chunk_i, suffix)))
        if self._augment:
            train = MultitrackAudioDataset._augment_features(train)
        return train, gt

    def precompute_features(self, names: str = 'read'):
        """N4: write the feature cache straight from K1 output.  names='read' uses the file names
        _process_precomputed loads ({chunk}_train.npy / {chunk}_gt.npy); names='write' uses the ones
        the reference's (broken) writer produces ({chunk}_train_{len}s.npy, data/dataset.py:244-249)."""
        for song_i, song_name in enumerate(self.songlist):
            features_dir = os.path.join(self._base_path, song_name, '{}_FEATURES'.format(song_name))
            os.makedirs(features_dir, exist_ok=True)
            suffix = '_norm' if self._normalize else ''
            tag = '' if names == 'read' else '_{}s'.format(self._chunk_length)
            for chunk_i in range(int(self.song_durations[song_i] / self._chunk_length)):
                feats = self._featurise(self._read_chunk_device(song_i, chunk_i)).cpu().numpy()
                np.save(os.path.join(features_dir, '{}_gt{}{}.npy'.format(chunk_i, tag, suffix)), feats[4])
                np.save(os.path.join(features_dir, '{}_train{}{}.npy'.format(chunk_i, tag, suffix)), feats[:4])

    def compute_mean_loudness(self, rank: int = 0, world_size: int = 1) -> dict:
        """data/dataset.py:115-130: mean integrated loudness per track name over all songs.  With
        world_size > 1 songs are sharded over ranks and the per-track (sum, count) is all-reduced."""
        from . import sharding
        meter = Meter(self._sr)
        sums = np.zeros(len(self._tracklist))
        counts = np.zeros(len(self._tracklist))
        durations = [wavio.info(self._get_track_path(s, 'mix')).frames for s in self.songlist]
        mine = sharding.assign_lpt(durations, world_size)[rank]
        for song_i in mine:
            song_name = self.songlist[song_i]
            for ti, track_name in enumerate(self._tracklist):
                track, _ = wavio.read(self._get_track_path(song_name, track_name))
                sums[ti] += meter.integrated_loudness(track)
                counts[ti] += 1
        sums, counts = sharding.allreduce_sum_count(sums, counts, self._dev() if world_size > 1 else None)
        return {t: float(sums[i] / counts[i]) for i, t in enumerate(self._tracklist)}

    def __getitem__(self, index: int) -> tuple:
        song_i, chunk_i = self._calculate_song_index(index)
        if self._compute_features:
            return self._process_on_the_fly(song_i, chunk_i)
        return self._process_precomputed(song_i, chunk_i)

    def __len__(self) -> int:
        return self._len

    def get_num_songs(self) -> int:
        return len(self.songlist)

    def get_song_durations(self) -> list:
        return self.song_durations

    def get_tracklist(self) -> list:
        return self._tracklist


class Featuriser:
    """Dataset-less stand-in for the two methods the mix path uses (inference_utils.py:75-77,116-118)."""

    def __init__(self, device=None):
        self._device = device
        self._tracklist = ['bass', 'drums', 'vocals', 'other', 'mix']

    def compute_features(self, audio, window_size: int = 2048, hop_length: int = 1024):
        return features.compute_features(audio, window_size, hop_length, device=self._device)

    def get_tracklist(self):
        return self._tracklist
