"""Drop-in for the reference's ``LoudnessEvaluator`` (evaluation.py:21-144) on the GPU.

Stem sum, integrated loudness of every stem, -20 LUFS normalisation and the loudness-difference
metric keep the reference's semantics; loudness runs through kernel K4 (all four stems of a song in
one pass), sums and normalisations through K2 / K5.  The xlsx report is written only if openpyxl is
installed (it is reporting I/O, out of the hot-path scope); wavs are written with the built-in
PCM16 writer.
"""
import os
from collections import OrderedDict
from statistics import mean

import numpy as np
import torch

from . import mixing, wavio
from .inference_utils import mix_song_smooth, stack_tracks
from .loudness import Meter
from .models import MeanLoudnessModel, RandomModel


class LoudnessEvaluator:
    def __init__(self, dataset, d_mean_loudness, mix_model, sr=44100, seed=None, results_dir='./experiment'):
        if seed:
            np.random.seed(seed)
        self.sr = sr
        self.d = dataset
        self.mix_model = mix_model
        self.mean_loudness_model = MeanLoudnessModel(d_mean_loudness, sr)
        self.random_model = RandomModel()
        self.meter = Meter(sr)
        self.keys = ('bass', 'drums', 'vocals', 'other')
        self.results_dir = results_dir

    def evaluate_loudness(self, tracks: dict) -> list:
        """evaluation.py:39-46: loudness of each stem minus the mean of the four."""
        stems, _ = stack_tracks(tracks, self.keys)
        per_track = self.meter.integrated_loudness_batch(stems)
        avg = mean(per_track)
        return [l - avg for l in per_track]

    @staticmethod
    def _calculate_diff_between_loudness_dicts(l_dict1: OrderedDict, l_dict2: OrderedDict):
        a1 = np.array(list(l_dict1.values()))
        a2 = np.array(list(l_dict2.values()))
        return float(np.mean(np.abs(a1 - a2)))

    def sum_and_normalize(self, track_dict: dict, target: float = -20.0):
        """evaluation.py:59-64: stem sum -> integrated loudness -> scale to `target` LUFS.
        Returns (normalised mix [C, N] on the device, loudness of the raw sum)."""
        stems, _ = stack_tracks(track_dict, list(track_dict.keys()))
        track_sum, kw = mixing.stem_sum(stems, meter=self.meter)      # the sum kernel also runs the meter's first pass
        loudness = self.meter.integrated_loudness(track_sum.t(), kw)
        return mixing.normalize_loudness(track_sum, loudness, target), loudness

    def _sum_and_evaluate_tracks(self, track_dict, reference_dict, song_name, identifier, write_to_disk=True):
        if write_to_disk:
            norm, _ = self.sum_and_normalize(track_dict)
            os.makedirs(self.results_dir, exist_ok=True)
            wavio.write(os.path.join(self.results_dir, '{}_{}.wav'.format(song_name, identifier)),
                        norm.t().cpu().numpy(), self.sr)
        per_track_loudness = self.evaluate_loudness(track_dict)
        loudness_dict = OrderedDict(zip(self.keys, per_track_loudness))
        if reference_dict:
            return loudness_dict, self._calculate_diff_between_loudness_dicts(loudness_dict, reference_dict)
        return loudness_dict, None

    def process_tracks(self, reference_tracks: dict, test_tracks: dict, song_name: str = 'song',
                       n_random_samples: int = 5, write_wavs_to_disk=False, chunk_length=2) -> dict:
        """evaluation.py:77-116 on already-loaded track dicts ({name: [2, N]})."""
        stats = {'song_name': song_name}
        reference, _ = self._sum_and_evaluate_tracks(reference_tracks, None, song_name, 'reference',
                                                     write_wavs_to_disk)
        _, stats['sum_error'] = self._sum_and_evaluate_tracks(test_tracks, reference, song_name, 'sum',
                                                              write_wavs_to_disk)
        loudnorm = self.mean_loudness_model.forward(test_tracks)
        _, stats['loudnorm_error'] = self._sum_and_evaluate_tracks(loudnorm, reference, song_name,
                                                                   'loudnorm', write_wavs_to_disk)
        mixed, _, _ = mix_song_smooth(self.d, self.mix_model, test_tracks, chunk_length=chunk_length,
                                      sr=self.sr, return_device=True)
        _, stats['mix_error'] = self._sum_and_evaluate_tracks(mixed, reference, song_name, 'mix',
                                                              write_wavs_to_disk)
        random_errors = []
        for exp_i in range(n_random_samples):
            rnd = self.random_model.forward(test_tracks)
            _, err = self._sum_and_evaluate_tracks(rnd, reference, song_name, 'random_{}'.format(exp_i),
                                                   write_wavs_to_disk)
            random_errors.append(err)
        stats['random_error'] = mean(random_errors)
        return stats

    def process_song(self, base_dir: str, song_name: str, n_random_samples: int = 5,
                     write_wavs_to_disk=False) -> dict:
        ref = load_tracks_musdb18(os.path.join(base_dir, 'manual_gain_mixes'), song_name,
                                  tracklist=self.keys, sr=self.sr)
        test = load_tracks_musdb18(os.path.join(base_dir, 'test'), song_name, tracklist=self.keys, sr=self.sr)
        return self.process_tracks(ref, test, song_name, n_random_samples, write_wavs_to_disk)

    def process_songlist(self, base_dir, songlist, write_to_disk=False, xlsx_path='./stats.xlsx'):
        keys = ['song_name', 'sum_error', 'random_error', 'loudnorm_error', 'mix_error']
        rows = [self.process_song(base_dir, s, write_wavs_to_disk=write_to_disk) for s in songlist]
        means = {k: mean(r[k] for r in rows) for k in keys[1:]}
        try:
            from openpyxl import Workbook
        except ImportError:
            return rows, means
        book = Workbook()
        sheet = book.active
        for key_i, key in enumerate(keys):
            sheet.cell(column=key_i + 1, row=1, value=key)
        for i, r in enumerate(rows):
            for key_i, key in enumerate(keys):
                sheet.cell(column=key_i + 1, row=i + 2,
                           value=r[key] if key == 'song_name' else '{:.4f}'.format(r[key]))
        for key_i, key in enumerate(keys):
            sheet.cell(column=key_i + 1, row=len(rows) + 2,
                       value='Mean' if key == 'song_name' else '{:.2f}'.format(means[key]))
        book.save(xlsx_path)
        return rows, means


def load_tracks_musdb18(base_dir, song_name, tracklist=('bass', 'drums', 'vocals', 'other', 'mix'), sr=44100):
    """data/dataset_utils.py:71-83 with the WAV reader: float32 planar [C, N] like librosa.load."""
    loaded = {}
    for track in tracklist:
        name = 'mixture' if track == 'mix' else track
        audio, file_sr = wavio.read(os.path.join(base_dir, song_name, '{}.wav'.format(name)), dtype='float32')
        if file_sr != sr:
            raise ValueError('resampling is not implemented (file %d Hz, requested %d Hz)' % (file_sr, sr))
        loaded[track] = np.ascontiguousarray(audio.T if audio.ndim == 2 else audio[None])
    return loaded


def produce_mixture(track_dict: dict, meter: Meter = None, sr=44100, target=-20.0):
    """data/listening_test_data_preparation.py:19-25 without the file write: stem sum scaled to
    -20 LUFS.  Returns the [C, N] device tensor."""
    meter = meter or Meter(sr)
    stems, _ = stack_tracks(track_dict, list(track_dict.keys()))
    track_sum, kw = mixing.stem_sum(stems, meter=meter)
    loudness = meter.integrated_loudness(track_sum.t(), kw)
    return mixing.normalize_loudness(track_sum, loudness, target)
