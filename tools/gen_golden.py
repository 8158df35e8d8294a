"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tools/gen_golden.py

Every array written here is an output of code imported from /root/reference (through
tools/ref_harness.py), computed on seeded synthetic inputs.  The fixtures are committed; the GPU
box has no /root/reference and only reads the .npz files.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

import ref_harness  # noqa: E402
from deep_audio_mixer_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class FakeGainModel(torch.nn.Module):
    """Deterministic stand-in for the CNN: m_i = 0.02*mean(x_i) + 0.1*i + 0.3*sin(mean(x_0)).
    Keeps the mix_song_smooth golden independent of conv implementations."""

    def forward(self, x):
        mu = x.double().mean(dim=(2, 3))                     # [B, 4]
        wob = 0.3 * torch.sin(mu[:, :1])
        m = 0.02 * mu + 0.1 * torch.arange(4, dtype=torch.float64)[None, :] + wob
        return None, tuple(m[:, i:i + 1] for i in range(4))


def main():
    torch.set_num_threads(1)
    torch.manual_seed(0)
    ref = ref_harness.load()
    DS = ref.dataset.MultitrackAudioDataset
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- windows -----------------------------------------------------------------------
    np.savez(os.path.join(GOLDEN, 'window.npz'),
             **{'hann_%d' % n: torch.hann_window(n).numpy() for n in (512, 1024, 2048)})

    # ---- compute_features (data/dataset.py:132-162) --------------------------------------
    feats = {}
    n = 12365
    for kind in ('white', 'music', 'tonal'):
        song = synthetic.make_song(77, n_stems=1, n_samples=n, kind=kind, dtype=np.float64)
        mono64 = song[0].mean(axis=0)
        for dt in (np.float64, np.float32):
            a = mono64.astype(dt)
            out = DS.compute_features(None, a).numpy()
            tag = '%s_%s' % (kind, np.dtype(dt).name)
            feats['in_' + tag] = a
            feats['out_' + tag] = np.ascontiguousarray(out)
    # hop 512 (historic configuration) and a [2, n] batch
    a = synthetic.make_song(78, 1, 9000, kind='music', dtype=np.float64)[0]
    feats['in_hop512'] = a[0]
    feats['out_hop512'] = np.ascontiguousarray(DS.compute_features(None, a[0], 2048, 512).numpy())
    feats['in_batch'] = a.astype(np.float32)
    feats['out_batch'] = np.ascontiguousarray(DS.compute_features(None, a.astype(np.float32)).numpy())
    # smaller window
    feats['in_w1024'] = a[1]
    feats['out_w1024'] = np.ascontiguousarray(DS.compute_features(None, a[1], 1024, 256).numpy())
    # silence -> -100 dB floor
    feats['out_silence'] = np.ascontiguousarray(DS.compute_features(None, np.zeros(4096)).numpy())
    # stereo -> mono (data/dataset.py:181-183), interleaved [n, 2]
    st = synthetic.make_song(79, 1, 5000, kind='music', dtype=np.float64)[0].T.copy()
    feats['in_stereo_interleaved'] = st
    feats['out_stereo_mono'] = DS._stereo_to_mono(st)
    np.savez_compressed(os.path.join(GOLDEN, 'features.npz'), **feats)

    # ---- interpolate_mask (inference_utils.py:12-41) --------------------------------------
    masks = {}
    rng = np.random.default_rng(5)
    for K, L in ((14, 1000), (1, 10), (7, 7), (3, 100), (29, 1323), (2, 5)):
        g = rng.uniform(0.2, 1.8, size=K)
        masks['g_%d_%d' % (K, L)] = g
        masks['m_%d_%d' % (K, L)] = ref.inference.interpolate_mask(g, L)
    np.savez_compressed(os.path.join(GOLDEN, 'mask.npz'), **masks)

    # ---- scalar dB <-> amplitude (data/dataset_utils.py:39-50) ----------------------------
    xs = np.linspace(-3, 3, 13)
    np.savez(os.path.join(GOLDEN, 'scalar.npz'), x=xs,
             db2amp=ref.dataset_utils.scalar_dB_to_amplitude(xs),
             amp2db=ref.dataset_utils.scalar_amplitude_to_dB(np.abs(xs) + 0.1))

    # ---- mix_song_smooth (inference_utils.py:105-145) with the D4 mono shim -----------------
    class ShimDataset:
        """the checked-in mix_song_smooth passes [2, n] slices; the mono down-mix
        the rest of the code base applies (data/dataset.py:181-183) is inserted here, outside the
        reference, so the reference itself runs unmodified."""

        def get_tracklist(self):
            return ['bass', 'drums', 'vocals', 'other', 'mix']

        def compute_features(self, audio):
            return DS.compute_features(None, np.mean(audio, axis=0))

    sr, chunk_length, N = 2000, 1, 27013
    song = synthetic.make_song(81, 4, N, sr=sr, kind='music', dtype=np.float32)
    tracks = {k: song[i] for i, k in enumerate(('bass', 'drums', 'vocals', 'other'))}
    ref.inference.device = torch.device('cpu')
    mixed, raw, smooth = ref.inference.mix_song_smooth(ShimDataset(), FakeGainModel(), tracks,
                                                       chunk_length=chunk_length, sr=sr)
    keys = ('bass', 'drums', 'vocals', 'other')
    mixed_arr = np.stack([mixed[k] for k in keys])            # [4, 2, N] float64
    np.savez_compressed(
        os.path.join(GOLDEN, 'mix_smooth.npz'),
        sr=sr, chunk_length=chunk_length, N=N, seed=81, input_sha=sha(song),
        raw=np.array([raw[k] for k in keys]), smooth=np.array([smooth[k] for k in keys]),
        mixed_dec=mixed_arr[:, :, ::13], mixed_abs_sum=np.abs(mixed_arr).sum(axis=-1),
        mix_sum_dec=mixed_arr.sum(axis=0)[:, ::13], mixed_dtype=str(mixed_arr.dtype))

    # ---- models: parameter counts, seeded forward, tail (model_scalar_1s.py:207-275 etc.) ---
    models = {}
    feat_in = {}
    for name, ctor, T in (('scalar1s', ref.m1.MixingModelScalar1s, 87),
                          ('scalar2s', ref.m2.MixingModelScalar2s, 173),
                          ('resnet18', ref.resnet.ResNet18, 216)):
        torch.manual_seed(0)
        net = ctor().eval()
        models['nparams_' + name] = sum(p.numel() for p in net.parameters() if p.requires_grad)
        g = torch.Generator().manual_seed(11)
        x = (torch.rand((1, 4, 1025, T), generator=g) * 100.0 - 80.0)
        with torch.no_grad():
            masked, ms = net(x)
        models['m_' + name] = np.array([float(m) for m in ms])
        models['masked_dec_' + name] = masked.numpy()[0, ::41, ::7]
        models['keys_' + name] = np.array(sorted(net.state_dict().keys()))
    np.savez_compressed(os.path.join(GOLDEN, 'models.npz'), **models)

    # ---- DummyModel (models/baselines/dummy_model.py:23-34) --------------------------------
    g = torch.Generator().manual_seed(12)
    x = torch.rand((2, 4, 1025, 3), generator=g) * 2.0 - 3.0
    with torch.no_grad():
        y = ref.dummy.DummyModel()(x)
    np.savez_compressed(os.path.join(GOLDEN, 'dummy.npz'), x=x.numpy(), y=y.numpy())

    total = sum(os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN))
    print('golden fixtures written to %s (%.1f KB)' % (GOLDEN, total / 1024))


if __name__ == '__main__':
    main()
