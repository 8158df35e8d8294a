"""Batched segment pipeline: S raw stems -> 4 buses -> dB-STFT features -> (CNN gains) -> gain masks ->
mixdown -> BS.1770 loudness -> -20 LUFS normalisation, for a batch of independent segments.

This is BASELINE.json config 2 ("model_scalar_2s inference batch of 64 synthetic 8-stem 10 s
segments") expressed with the kernels of this package; every stage is asynchronous on one CUDA
stream and nothing synchronises with the host.  Reference semantics per stage:

  bus sum        data/medleydb_data_surgery.ipynb:305-332   (float32, stem order)
  features       data/dataset.py:132-162 + mono mean :181-183, per chunk of chunk_samples
  gains          10**(0.5*m) per chunk (data/dataset_utils.py:46-50); segments shorter than
                 8 chunks cannot go through mix_song_smooth's Savitzky-Golay filter,
                 so the per-chunk gains are held directly
  mask + mix     inference_utils.py:12-41,143 and evaluation.py:59-60 (float64)
  loudness       pyloudnorm integrated loudness of the mix (evaluation.py:63)
  normalise      x * 10**((-20 - L)/20) (evaluation.py:64)
"""
import numpy as np
import torch

from . import features, mixing
from .loudness import Meter

class SegmentBatchPipeline:
    def __init__(self, n_stems=8, group=(0, 1, 1, 2, 2, 3, 3, 3), chunk_samples=88200, hop=512, n_fft=2048,
                 sr=44100, target_lufs=-20.0, device=None):
        self.n_stems, self.group = n_stems, list(group)
        self.chunk_samples, self.hop, self.n_fft, self.sr = chunk_samples, hop, n_fft, sr
        self.target = target_lufs
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.meter = Meter(sr)
        self.timings = None          # optional dict of CUDA event pairs per stage
        self.mono_features = True    # K1 reads the bus-sum kernel's mono down-mix instead of averaging channels
        self.fuse_meter = True       # K2 also computes the meter's first pass over the mix it writes (dam_gain_mix_kweight)
        self.feature_compute_dtype = None     # None: follows the audio like the reference (float32); torch.float64
                                              # meets 1e-4 dB on every bin (the reference's training-path dtype)

    def n_chunks(self, n_samples: int) -> int:
        return n_samples // self.chunk_samples

    def feature_buffer(self, n_rows: int, n_bus: int = 4, device=None) -> torch.Tensor:
        """Destination of ``features``: [n_rows, n_bus, F, pitch] float32 with the frame axis padded to a
        multiple of 8 (32 bytes).  With T = 173 -- the only length model_scalar_2s accepts -- a dense row is
        692 bytes and 7 of 8 of K1's 8-frame output segments straddle two DRAM sectors (12 % slower, ncu in
        profiles/); consumers read the [..., :T] view, as non-contiguous as the reference's own features."""
        F = self.n_fft // 2 + 1
        T = features.num_frames(self.chunk_samples, self.hop)
        return torch.empty((n_rows, n_bus, F, features.aligned_pitch(T)), dtype=torch.float32,
                           device=device if device is not None else self.device)

    def features(self, buses: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """buses [B, 4, C, N] float32 (channels averaged inside K1) or their mono down-mix [B, 4, N] from
        ``stage_bus`` -> [B * n_chunks, 4, F, T] float32 (model input layout; a view of a ``feature_buffer``
        whose frame axis is padded to the next multiple of 8).  Both forms give the same bits."""
        mono = buses.dim() == 3
        B, S = buses.shape[0], buses.shape[1]
        N = buses.shape[-1]
        nc = self.n_chunks(N)
        F = self.n_fft // 2 + 1
        T = features.num_frames(self.chunk_samples, self.hop)
        if out is None:
            out = self.feature_buffer(B * nc, S, buses.device)
        pitch = out.shape[-1]
        if tuple(out.shape[:3]) != (B * nc, S, F) or pitch < T or not out.is_contiguous():
            raise ValueError('features: out must be a contiguous [%d, %d, %d, >=%d] buffer' % (B * nc, S, F, T))
        for s in range(S):
            # rows = (segment, chunk) of bus s: two-level addressing
            src = buses[:, s]
            if mono:
                view = src.as_strided((B, nc, self.chunk_samples), (buses.stride(0), self.chunk_samples, 1),
                                      src.storage_offset())
                features.stft_db(view, self.n_fft, self.hop, out=out[:, s], out_row_stride=S * F * pitch,
                                 frame_pitch=pitch, compute_dtype=self.feature_compute_dtype, out_dtype=torch.float32)
            else:
                C = buses.shape[2]
                view = src.as_strided((B, nc, C, self.chunk_samples),
                                      (buses.stride(0), self.chunk_samples, buses.stride(2), 1), src.storage_offset())
                features.stft_db(view, self.n_fft, self.hop, channel_axis=-2, out=out[:, s],
                                 out_row_stride=S * F * pitch, frame_pitch=pitch,
                                 compute_dtype=self.feature_compute_dtype, out_dtype=torch.float32)
        return out[..., :T]

    def _ev(self, name):
        if self.timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.timings.setdefault(name, []).append(e)

    # ---- the five stages, separately callable (bench.py captures each in its own CUDA graph) --------
    def stage_bus(self, raw: torch.Tensor, want_mono: bool = False):
        """raw [B, S, C, N] -> buses [B, 4, C, N] (kernel N2); with ``want_mono`` also their channel mean
        [B, 4, N], which K1 featurises through its asynchronous mono path (same bits, 20 % less K1 time for
        0.45 GB more bus-sum writes per bench batch)."""
        return mixing.bus_sum(raw, self.group, 4, batch_dims=1, want_mono=want_mono)

    def stage_gains(self, m: torch.Tensor, B: int, nc: int) -> torch.Tensor:
        """head outputs m [B*nc, 4] -> amplitude gains [B, 4, nc]: 10**(0.5 m) in float32 like the
        reference (data/dataset_utils.py:46-50), widened to float64 for the mask (one launch, dam_head_gains)."""
        return mixing.head_gains(m, B, nc)

    def stage_mix(self, buses: torch.Tensor, gains: torch.Tensor, for_meter: bool = True):
        """zero-order-hold masks, per-bus products, stem sum, peak and sum of squares (kernel K2).  With
        ``for_meter`` the kernel also leaves the meter's first pass over the mix (the K-weighting chunk states) in
        ``kw_workspace`` for ``stage_loudness``."""
        N = buses.shape[-1]
        return mixing.gain_mix(buses, gains, interp_coef=int(N / gains.shape[-1]), compute_dtype=torch.float64,
                               want_mix=True, want_stats=True, meter=self.meter if for_meter else None)

    def stage_loudness(self, mix: torch.Tensor, kw_workspace=None, with_gain: bool = False):
        """BS.1770 integrated loudness of every mix, on the device (kernel K4 + gating); with the ``kw_workspace``
        of ``stage_mix`` it reads the mix once.  ``with_gain``: returns (lufs, factor) with the normalisation factor
        10**((target - L)/20) from the gating launch."""
        return self.meter.integrated_loudness_device(mix, kw_workspace, gain_target=self.target if with_gain else None)

    def stage_normalise(self, mix: torch.Tensor, lufs: torch.Tensor, factor: torch.Tensor = None) -> torch.Tensor:
        """x * 10**((target - L)/20) (kernel K5; the factors never leave the device)."""
        B, C, N = mix.shape
        if factor is None:
            factor = torch.pow(10.0, (self.target - lufs) / 20.0)
        return mixing.scale_rows(mix.reshape(B, C * N), factor).reshape(B, C, N)

    def run(self, raw: torch.Tensor, m: torch.Tensor = None, model=None, feats_out: torch.Tensor = None):
        """raw [B, S, C, N] float32 on the device.  ``m`` [B*n_chunks, 4] are the CNN head outputs; if
        None they are computed with ``model`` from the features.  Returns a dict with the normalised
        mixes [B, C, N] (float64), their loudness before normalisation [B], the features and m."""
        B, S, C, N = raw.shape
        nc = self.n_chunks(N)
        self._ev('bus0')
        buses, mono = self.stage_bus(raw, want_mono=True)                           # [B, 4, C, N], [B, 4, N]
        self._ev('bus1')
        feats = self.features(mono if self.mono_features else buses, feats_out)
        self._ev('stft1')
        if m is None:
            with torch.no_grad():
                _, ms = model(feats)
            m = torch.cat([v.reshape(-1, 1) for v in ms], dim=1)
        gains = self.stage_gains(m, B, nc)
        self._ev('gain0')
        res = self.stage_mix(buses, gains, for_meter=self.fuse_meter)
        self._ev('mix1')
        lufs, factor = self.stage_loudness(res['mix'], res['kw_workspace'], with_gain=True)    # [B] on the device
        self._ev('loud1')
        norm = self.stage_normalise(res['mix'], lufs, factor)
        self._ev('scale1')
        return dict(mix=norm, lufs=lufs, peak=res['peak'], sumsq=res['sumsq'], features=feats, m=m)

    def run_host(self, raw_host: torch.Tensor, m_host: torch.Tensor, out_host: torch.Tensor,
                 lufs_host: torch.Tensor, sub_batch: int = 4, streams=None, copy_mode: str = 'overlap'):
        """End-to-end form with HOST (pinned) buffers: raw_host [B, S, C, N] float32 or int16 (PCM16) -> out_host
        [B, C, N] (float64 like the reference's mix, or float32: the dtype of ``out_host`` decides) and lufs_host
        [B].  The batch is cut into sub-batches; uploads, kernels and downloads of consecutive sub-batches overlap
        on three streams.  copy_mode='serial' queues the downloads on the upload stream, one direction at a time
        per GPU (with many GPUs behind one host, concurrent up- and downloads interfere: see DESIGN.md section 6)."""
        B = raw_host.shape[0]
        nc = self.n_chunks(raw_host.shape[-1])
        dev = self.device
        if streams is None:
            streams = [torch.cuda.Stream(dev) for _ in range(3)]
        s_in, s_cmp, s_out = streams
        if copy_mode == 'serial':
            s_out = s_in
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)
        keep = []
        for b0 in range(0, B, sub_batch):
            b1 = min(B, b0 + sub_batch)
            with torch.cuda.stream(s_in):
                raw = raw_host[b0:b1].to(dev, non_blocking=True)
                m = m_host[b0 * nc:b1 * nc].to(dev, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in)
                res = self.run(raw, m)
                raw.record_stream(s_cmp)
                m.record_stream(s_cmp)
                ev_cmp = torch.cuda.Event()
                ev_cmp.record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp)
                out_host[b0:b1].copy_(res['mix'], non_blocking=True)      # converts on the device if out_host is float32
                lufs_host[b0:b1].copy_(res['lufs'], non_blocking=True)
                res['mix'].record_stream(s_out)
                res['lufs'].record_stream(s_out)
            keep.append(res)
        for st in streams:
            cur.wait_stream(st)
        return keep


class SongBatchMixdown:
    """Known-gain single-pass mixdown of a BATCH of songs of one shape (BASELINE.json config 5, the listening-test
    flow of data/listening_test_data_preparation.py:19-53 with per-chunk gains already known): dB-STFT features of
    every stem chunk, zero-order-hold gain masks + stem sum, BS.1770 loudness of every mix, -20 LUFS normalisation.
    One launch per stage for the whole batch (K1 x 1, K2 x 1, K4 x 3 + gating, K5 x 1), nothing synchronises with
    the host and all shapes are static, so a batch is CUDA-graph capturable -- a 30 s song alone sits on the launch
    floor (a few hundred microseconds of launches for ~50 us of memory traffic), sixteen of them do not."""

    def __init__(self, chunk_samples=88200, hop=1024, n_fft=2048, sr=44100, target_lufs=-20.0):
        self.chunk_samples, self.hop, self.n_fft, self.sr, self.target = chunk_samples, hop, n_fft, sr, target_lufs
        self.meter = Meter(sr)

    def feature_buffer(self, n_rows: int, device=None) -> torch.Tensor:
        """[n_rows, F, pitch] float32 with the frame axis padded to a multiple of 8 (sector-aligned rows, see
        SegmentBatchPipeline.feature_buffer); ``features`` returns its [..., :T] view."""
        T = features.num_frames(self.chunk_samples, self.hop)
        return torch.empty((n_rows, self.n_fft // 2 + 1, features.aligned_pitch(T)), dtype=torch.float32, device=device)

    def features(self, stems: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """stems [B, S, C, N] float32 / int16 -> [B * S * n_chunks, F, T] float32 (rows ordered song, stem, chunk),
        channels averaged inside K1, ONE launch: rows are addressed on two levels (stem plane, chunk)."""
        B, S, C, N = stems.shape
        nc = N // self.chunk_samples
        if not stems.is_contiguous():
            raise ValueError('stems must be contiguous [B, S, C, N]')
        view = stems.as_strided((B * S, nc, C, self.chunk_samples), (stems.stride(1), self.chunk_samples, stems.stride(2), 1),
                                stems.storage_offset())
        if out is None:
            out = self.feature_buffer(B * S * nc, stems.device)
        T = features.num_frames(self.chunk_samples, self.hop)
        F = self.n_fft // 2 + 1
        pitch = out.shape[-1]
        if tuple(out.shape[:2]) != (B * S * nc, F) or pitch < T or not out.is_contiguous():
            raise ValueError('features: out must be a contiguous [%d, %d, >=%d] buffer' % (B * S * nc, F, T))
        if C == 1:
            features.stft_db(view[:, :, 0], self.n_fft, self.hop, out=out, out_row_stride=F * pitch, frame_pitch=pitch)
        else:
            features.stft_db(view, self.n_fft, self.hop, channel_axis=-2, out=out, out_row_stride=F * pitch, frame_pitch=pitch)
        return out[..., :T]

    def run(self, stems: torch.Tensor, gains: torch.Tensor, want_features: bool = True, feats_out: torch.Tensor = None):
        """stems [B, S, C, N], gains [B, S, K] (device).  Returns dict(features, mix [B, C, N] float64 normalised to
        the target, lufs [B] before normalisation, peak, sumsq)."""
        B, S, C, N = stems.shape
        feats = self.features(stems, feats_out) if want_features else None
        res = mixing.gain_mix(stems, gains, compute_dtype=torch.float64, want_mix=True, want_stats=True, meter=self.meter)
        lufs, factor = self.meter.integrated_loudness_device(res['mix'], res['kw_workspace'], gain_target=self.target)
        norm = mixing.scale_rows(res['mix'].reshape(B, C * N), factor).reshape(B, C, N)
        return dict(features=feats, mix=norm, lufs=lufs, peak=res['peak'], sumsq=res['sumsq'])

    @staticmethod
    def algorithmic_bytes(B, S, C, N, chunk_samples, hop, n_fft=2048, want_features=True):
        """Single-pass minimum: every stem sample read once, features and the normalised mix written
        once (the loudness reduction rides on the mix pass)."""
        nc = N // chunk_samples
        feats = B * S * nc * (n_fft // 2 + 1) * (1 + chunk_samples // hop) * 4 if want_features else 0
        return B * S * C * N * 4 + feats + B * C * N * 8


def algorithmic_bytes(B, S, C, N, n_chunks, chunk_samples, hop, n_fft=2048, n_bus=4):
    """Minimum HBM bytes per stage for one batch (each operand read/written once)."""
    F = n_fft // 2 + 1
    T = 1 + chunk_samples // hop
    raw = B * S * C * N * 4
    bus = B * n_bus * C * N * 4
    feats = B * n_chunks * n_bus * F * T * 4
    mix = B * C * N * 8
    mono = B * n_bus * N * 4                       # channel mean of every bus, written by the bus sum, read by K1
    return {'bus_sum': raw + bus + mono, 'stft_db': B * n_bus * n_chunks * chunk_samples * 4 + feats,
            'gain_mix': bus + mix, 'kweight': mix, 'scale': 2 * mix}
