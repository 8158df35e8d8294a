"""A/B of K1's staging modes (DAM_K1_ASYNC = 0 phase kernel, 2 TMA-staged + tables in shared memory,
3 TMA-staged + tables in global memory; no knob = the library's own choice) on the shapes that matter; every mode must give bit-identical
output.  Each mode runs in its own process because the knob is read once.

    python tools/ab_async.py            # parent: runs the modes, prints a table
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [  # label, rows, channels, n, hop
    ('mono hop512 T=173', 1280, 1, 88200, 512),
    ('stereo hop512 T=173 (bench K1 x4)', 1280, 2, 88200, 512),
    ('mono hop1024 T=173', 2560, 1, 176400, 1024),
    ('mono hop1024 T=216', 2048, 1, 220500, 1024),
    ('mono hop1024 T=87', 4096, 1, 88200, 1024),
    ('stereo hop1024 T=216', 1024, 2, 220500, 1024),
    ('mono hop256 T=345', 640, 1, 88200, 256),
    ('mono odd offset hop512', 640, 1, 88201, 512),
    ('mono hop512 T=176', 1280, 1, 89600, 512),
    ('stereo hop512 T=176', 1280, 2, 89600, 512),
]


def child():
    import torch
    from deep_audio_mixer_b200 import features as F
    res = []
    for label, rows, ch, n, hop in SHAPES:
        g = torch.Generator(device='cuda').manual_seed(7)
        shape = (rows, ch, n) if ch > 1 else (rows, n)
        x = torch.randn(shape, device='cuda', generator=g) * 0.1
        T = 1 + n // hop
        out = torch.empty(rows, 1025, T, device='cuda')
        kw = dict(hop=hop, out=out)
        if ch > 1:
            kw['channel_axis'] = -2
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
        digest = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:16]
        byts = x.numel() * 4 + out.numel() * 4
        res.append({'label': label, 'ms': ts[len(ts) // 2], 'GBps': byts / ts[len(ts) // 2] / 1e6, 'sha': digest})
        del x, out
    print('RESULT ' + json.dumps(res))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        return child()
    table = {}
    modes = [('phase', {'DAM_K1_ASYNC': '0'}), ('phase ft8', {'DAM_K1_ASYNC': '0', 'DAM_FT_F32': '8'}),
             ('phase ft16', {'DAM_K1_ASYNC': '0', 'DAM_FT_F32': '16'}),
             ('default', {}),
             ('async tab-smem ft8', {'DAM_K1_ASYNC': '2', 'DAM_FT_F32': '8'}),
             ('async tab-global ft8', {'DAM_K1_ASYNC': '3', 'DAM_FT_F32': '8'})]
    for name, env in modes:
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], env=e, capture_output=True, text=True,
                           timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
        if not line:
            print(name, 'FAILED', r.stderr[-800:])
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
            print('   %-22s %7.3f ms  %6.0f GB/s (%.1f %% of 6549.8)%s' % (name, c['ms'], c['GBps'], c['GBps'] / 65.498, same))
    out = os.path.join(ROOT, 'gpurun_out', 'ab_async.json')
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(table, open(out, 'w'), indent=1)


if __name__ == '__main__':
    main()
