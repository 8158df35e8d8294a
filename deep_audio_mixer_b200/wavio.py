"""Minimal RIFF/WAVE reader and writer (N1: the ingest step in front of K1).

The reference reads audio with ``soundfile`` (data/dataset.py:70,125,194) and ``librosa.load``
(data/dataset_utils.py:65,80); neither is installed here and both are I/O only.  This module covers
what those calls are used for: ``info(path).duration``, ``read(path, start, stop)`` returning
float64 [n, C] scaled like soundfile (PCM / 2**(bits-1)), and ``read_raw`` returning the stored
samples untouched (int16 stays int16, so K1 can ingest PCM16 directly at half the bytes).
PCM 16/24/32-bit and IEEE float 32/64 are supported.
"""
import struct
from collections import namedtuple

import numpy as np

WavInfo = namedtuple('WavInfo', 'samplerate channels frames duration sampwidth is_float data_offset')


def info(path) -> WavInfo:
    with open(path, 'rb') as f:
        riff, _, wave = struct.unpack('<4sI4s', f.read(12))
        if riff != b'RIFF' or wave != b'WAVE':
            raise ValueError('%s is not a RIFF/WAVE file' % path)
        fmt = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise ValueError('no data chunk in %s' % path)
            cid, size = struct.unpack('<4sI', hdr)
            if cid == b'fmt ':
                raw = f.read(size + (size & 1))
                tag, ch, sr, _, _, bits = struct.unpack('<HHIIHH', raw[:16])
                if tag == 0xFFFE and size >= 26:                     # WAVE_FORMAT_EXTENSIBLE
                    tag = struct.unpack('<H', raw[24:26])[0]
                fmt = (tag, ch, sr, bits)
            elif cid == b'data':
                if fmt is None:
                    raise ValueError('data chunk before fmt chunk in %s' % path)
                tag, ch, sr, bits = fmt
                if tag not in (1, 3):
                    raise ValueError('unsupported WAVE format tag %d' % tag)
                sw = bits // 8
                frames = size // (sw * ch)
                return WavInfo(sr, ch, frames, frames / sr, sw, tag == 3, f.tell())
            else:
                f.seek(size + (size & 1), 1)


def read_raw(path, start=0, stop=None):
    """Stored samples [n, C] without scaling (int16 / int32 (24-bit sign-extended) / float32 /
    float64) and the sample rate."""
    wi = info(path)
    stop = wi.frames if stop is None else min(stop, wi.frames)
    start = max(0, min(start, stop))
    n = stop - start
    with open(path, 'rb') as f:
        f.seek(wi.data_offset + start * wi.sampwidth * wi.channels)
        buf = f.read(n * wi.sampwidth * wi.channels)
    if wi.is_float:
        data = np.frombuffer(buf, dtype='<f4' if wi.sampwidth == 4 else '<f8')
    elif wi.sampwidth == 2:
        data = np.frombuffer(buf, dtype='<i2')
    elif wi.sampwidth == 4:
        data = np.frombuffer(buf, dtype='<i4')
    elif wi.sampwidth == 3:
        b = np.frombuffer(buf, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        data = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
        data = np.where(data & 0x800000, data - 0x1000000, data).astype(np.int32)
    else:
        raise ValueError('unsupported sample width %d' % wi.sampwidth)
    return data.reshape(n, wi.channels), wi.samplerate


def read_bytes(path, start=0, stop=None):
    """The stored bytes of frames [start, stop) untouched (uint8 array) plus the file's WavInfo: what the device
    ingest uploads -- 2 or 3 bytes per sample instead of the 8 of a float64 array."""
    wi = info(path)
    stop = wi.frames if stop is None else min(stop, wi.frames)
    start = max(0, min(start, stop))
    n = stop - start
    with open(path, 'rb') as f:
        f.seek(wi.data_offset + start * wi.sampwidth * wi.channels)
        buf = f.read(n * wi.sampwidth * wi.channels)
    return np.frombuffer(buf, dtype=np.uint8), wi


def read(path, start=0, stop=None, dtype='float64'):
    """soundfile.read(path, start=, stop=): float [n, C] (or [n] for mono), PCM scaled by
    2**-(bits-1)."""
    wi = info(path)
    raw, sr = read_raw(path, start, stop)
    if wi.is_float:
        data = raw.astype(dtype)
    else:
        data = raw.astype(dtype) / float(1 << (8 * wi.sampwidth - 1))
    if wi.channels == 1:
        data = data[:, 0]
    return data, sr


def write(path, data, samplerate, subtype='PCM_16'):
    """soundfile.write for [n, C] / [n] arrays: PCM_16 (default, clipped), PCM_24 or FLOAT."""
    data = np.asarray(data)
    if data.ndim == 1:
        data = data[:, None]
    n, ch = data.shape
    if subtype == 'PCM_16':
        pcm = np.clip(np.round(data.astype(np.float64) * 32768.0), -32768, 32767).astype('<i2')
        tag, bits, payload = 1, 16, pcm.tobytes()
    elif subtype == 'PCM_24':
        v = np.clip(np.round(data.astype(np.float64) * 8388608.0), -8388608, 8388607).astype('<i4')
        tag, bits, payload = 1, 24, np.ascontiguousarray(v).view(np.uint8).reshape(-1, 4)[:, :3].tobytes()
    elif subtype == 'FLOAT':
        tag, bits, payload = 3, 32, data.astype('<f4').tobytes()
    else:
        raise ValueError('subtype must be PCM_16, PCM_24 or FLOAT')
    sw = bits // 8
    with open(path, 'wb') as f:
        f.write(struct.pack('<4sI4s', b'RIFF', 36 + len(payload), b'WAVE'))
        f.write(struct.pack('<4sIHHIIHH', b'fmt ', 16, tag, ch, samplerate, samplerate * ch * sw, ch * sw, bits))
        f.write(struct.pack('<4sI', b'data', len(payload)))
        f.write(payload)
