"""Summarise an `ncu -i report --page raw --csv` dump into a small JSON: per launch, the kernel name, duration, DRAM
bytes and the handful of counters DESIGN.md quotes.   python tools/ncu_summarize.py raw.csv out.json [note]"""
import csv
import json
import sys

KEYS = {
    'gpu__time_duration.sum': 'time',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'smsp__inst_executed.sum': 'warp_instructions',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed': 'lsu_data_pipe_pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'shared_wavefronts',
    'launch__registers_per_thread': 'registers',
    'launch__shared_mem_per_block_dynamic': 'dynamic_smem',
    'launch__grid_size': 'grid',
    'smsp__sass_inst_executed_op_local_ld.sum': 'local_loads',
}
UNIT_SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0,
              'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        rec = {'kernel': r[idx['Kernel Name']]}
        for k, name in KEYS.items():
            if k in idx and r[idx[k]] != '':
                v = float(r[idx[k]].replace(',', ''))
                u = units[idx[k]]
                if name in ('time', 'dram_read', 'dram_write', 'dynamic_smem') and u in UNIT_SCALE:
                    v *= UNIT_SCALE[u]
                rec[name] = v
        if 'dram_read' in rec and 'dram_write' in rec:
            rec['dram_bytes'] = rec['dram_read'] + rec['dram_write']
        out.append(rec)
    res = {'note': sys.argv[3] if len(sys.argv) > 3 else '', 'units': 'time in seconds, bytes in bytes', 'launches': out}
    json.dump(res, open(sys.argv[2], 'w'), indent=1)
    for rec in out:
        print('%-60s %8.1f us  dram %7.1f MB' % (rec['kernel'][:60], rec.get('time', 0) * 1e6, rec.get('dram_bytes', 0) / 1e6))


if __name__ == '__main__':
    main()
