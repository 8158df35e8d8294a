"""Host<->device copy ceilings with ALL ranks copying at once (explains the end-to-end scaling curve of bench.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/pcie_probe_multi.py [out.json]

Every rank moves the per-step traffic of bench.py's end-to-end leg (1.81 GB up, 0.45 GB down) between pinned host
memory and its GPU, all ranks released together by a barrier; reported per rank (CUDA events) and as the aggregate
(bytes of all ranks / slowest rank's time).  Variants: H2D only, D2H only, both directions at once, both directions
one after the other, and the same with the process bound to the CPUs of the NUMA node its GPU hangs off (sysfs)
BEFORE the pinned buffers are allocated and first touched.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

UP, DOWN = 1_806_336_000, 451_584_000


def sysfs_numa(local):
    """(pci bus id, numa node of the GPU's PCI device, cpus of that node) from sysfs; node -1 = unknown."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        addr = '%04x:%02x:%02x.0' % (dom, bus, dev)
        with open('/sys/bus/pci/devices/%s/numa_node' % addr) as f:
            node = int(f.read().strip())
        cpus = set()
        if node >= 0:
            with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
                for part in f.read().strip().split(','):
                    lo, _, hi = part.partition('-')
                    cpus.update(range(int(lo), int(hi or lo) + 1))
        return addr, node, cpus
    except Exception as e:                                   # noqa: BLE001
        return repr(e), -1, set()


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'pcie_probe_%d.json' % world)

    d_in = torch.empty(UP, dtype=torch.uint8, device=dev)
    d_out = torch.empty(DOWN, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    addr, node, cpus = sysfs_numa(local)
    allowed = os.sched_getaffinity(0)

    def measure(h_in, h_out):
        def timed(fn, reps=4):
            fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                ts = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(ts, t)
                return [float(x.item()) for x in ts]
            return [float(t.item())]

        def up():
            d_in.copy_(h_in, non_blocking=True)

        def down():
            h_out.copy_(d_out, non_blocking=True)

        def both():
            cur = torch.cuda.current_stream()
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
            cur.wait_stream(s1)
            cur.wait_stream(s2)

        def serial():
            d_in.copy_(h_in, non_blocking=True)
            h_out.copy_(d_out, non_blocking=True)

        res = {}
        for name, fn, nbytes in (('h2d', up, UP), ('d2h', down, DOWN), ('bidir', both, UP + DOWN), ('serial', serial, UP + DOWN)):
            ms = timed(fn)
            res[name] = {'ms_per_rank': [round(v, 3) for v in ms], 'GBps_per_rank': [round(nbytes / v / 1e6, 2) for v in ms],
                         'aggregate_GBps': world * nbytes / max(ms) / 1e6}
        return res

    out = {'world': world, 'bytes_up': UP, 'bytes_down': DOWN}
    h_in = torch.empty(UP, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(1)
    out['default_placement'] = measure(h_in, h_out)
    del h_in, h_out
    bound = False
    if cpus & allowed and (cpus & allowed) != allowed:
        os.sched_setaffinity(0, cpus & allowed)
        bound = True
    h_in = torch.empty(UP, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)                                        # first touch on the bound CPUs
    h_out.fill_(1)
    out['numa_bound'] = measure(h_in, h_out) if bound else 'not applicable: %s' % (
        'GPU numa node unknown (%s)' % node if node < 0 else 'node cpus == allowed cpus')
    info = [None] * world
    mine = {'rank': rank, 'pci': addr, 'numa_node': node, 'node_cpus': len(cpus), 'allowed_cpus': len(allowed), 'bound': bound}
    if world > 1:
        dist.all_gather_object(info, mine)
    else:
        info = [mine]
    out['ranks'] = info
    if rank == 0:
        try:
            with open('/sys/devices/system/node/online') as f:
                out['numa_nodes_online'] = f.read().strip()
        except OSError:
            out['numa_nodes_online'] = 'unknown'
        with open(out_path, 'w') as f:
            json.dump(out, f, indent=1)
        for k in ('default_placement', 'numa_bound'):
            if isinstance(out[k], dict):
                print(k, {n: round(v['aggregate_GBps'], 1) for n, v in out[k].items()})
            else:
                print(k, out[k])
        print(info)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
