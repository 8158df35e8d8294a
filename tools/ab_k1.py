"""A/B of K1's kernels (DAM_K1_ASYNC = 0 phase kernel, 1 = default: the TMA-staged kernel of stft_tma.cu wherever
the shape qualifies, 3 = the same with the tables in global memory) on the
shapes that matter; every mode must give bit-identical output.  Each mode runs in its own process because the
knob is read once.

    python tools/ab_k1.py [out.json]           # parent: runs the modes, prints a table
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [  # label, rows, channels, n, hop, dtype, padded pitch
    ('bench K1: mono f32 hop512 T=173 pitch176', 1280, 1, 88200, 512, 'f32', True),
    ('mono f32 hop512 T=173 dense', 1280, 1, 88200, 512, 'f32', False),
    ('mono i16 hop512 T=173 pitch176', 1280, 1, 88200, 512, 'i16', True),
    ('stereo f32 hop512 T=173 pitch176', 1280, 2, 88200, 512, 'f32', True),
    ('stereo i16 hop512 T=173 pitch176', 1280, 2, 88200, 512, 'i16', True),
    ('HEAD: mono f32 hop1024 T=216', 2048, 1, 220500, 1024, 'f32', False),
    ('mono i16 hop1024 T=216', 2048, 1, 220500, 1024, 'i16', False),
    ('stereo f32 hop1024 T=216', 1024, 2, 220500, 1024, 'f32', False),
    ('stereo i16 interleaved hop1024 T=216', 1024, -2, 220500, 1024, 'i16', False),
    ('mono f32 hop1024 T=87 pitch88', 4096, 1, 88200, 1024, 'f32', True),
    ('mono f32 hop1024 T=173 pitch176', 2560, 1, 176400, 1024, 'f32', True),
    ('mono f32 hop256 T=345', 640, 1, 88200, 256, 'f32', False),
]


def child():
    import torch
    from deep_audio_mixer_b200 import features as F
    res = []
    for label, rows, ch, n, hop, dt, padded in SHAPES:
        g = torch.Generator(device='cuda').manual_seed(7)
        inter = ch < 0
        ch = abs(ch)
        shape = (rows, ch, n) if ch > 1 else (rows, n)
        x = torch.randn(shape, device='cuda', generator=g) * 0.1
        if dt == 'i16':
            x = (x * 32768.0).round().clamp(-32768, 32767).to(torch.int16)
        kw = dict(hop=hop)
        if ch > 1:
            kw['channel_axis'] = -2
            if inter:
                x = x.transpose(1, 2).contiguous()
                kw['channel_axis'] = -1
        T = 1 + n // hop
        pitch = F.aligned_pitch(T) if padded else T
        out = torch.empty(rows, 1025, pitch, device='cuda')
        kw.update(out=out, frame_pitch=pitch)
        for _ in range(3):
            F.stft_db(x, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            F.stft_db(x, **kw)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        digest = hashlib.sha1(out[..., :T].contiguous().cpu().numpy().tobytes()).hexdigest()[:16]
        byts = x.numel() * x.element_size() + rows * 1025 * T * 4
        ms = ts[len(ts) // 2]
        res.append({'label': label, 'ms': ms, 'GBps': byts / ms / 1e6, 'sha': digest,
                    'stem_s_per_s': rows * n / 44100.0 / (ms * 1e-3)})
        del x, out
    print('RESULT ' + json.dumps(res))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        return child()
    table = {}
    modes = [('phase', {'DAM_K1_ASYNC': '0'}), ('tma (default)', {}),
             ('tma tab-global', {'DAM_K1_ASYNC': '3'})]
    # library variants (build.py --variant NAME ...): the default dispatch on each
    for v in [v for v in os.environ.get('AB_VARIANTS', '').split(',') if v]:
        modes.append(('lib:' + v, {'DAM_LIB_VARIANT': v}))
    if os.environ.get('AB_MODES'):
        keep = os.environ['AB_MODES'].split(',')
        modes = [m for m in modes if m[0].split()[0] in keep or m[0].startswith('lib:')]
    for name, env in modes:
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], env=e, capture_output=True, text=True,
                           timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
        if not line:
            print(name, 'FAILED', r.stderr[-1500:])
            continue
        table[name] = json.loads(line[0][7:])
    ref = table.get('phase')
    for i, (label, *_rest) in enumerate(SHAPES):
        print(label)
        for name, _ in modes:
            if name not in table:
                continue
            c = table[name][i]
            same = '' if ref is None else ('  identical' if c['sha'] == ref[i]['sha'] else '  DIFFERENT OUTPUT')
            print('   %-20s %7.3f ms  %6.0f GB/s (%.3f of 6549.8)  %.2f M stem-s/s%s' %
                  (name, c['ms'], c['GBps'], c['GBps'] / 6549.8, c['stem_s_per_s'] / 1e6, same))
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'ab_k1.json')
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(table, open(out, 'w'), indent=1)


if __name__ == '__main__':
    main()
