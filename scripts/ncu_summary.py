#!/usr/bin/env python
"""Turns an `ncu --set full` report into the per-launch summary kept under profiles/ (and read by bench.py
for `roofline.traffic`).   python scripts/ncu_summary.py gpurun_out/step_r2.ncu-rep profiles/r2_step_ncu_summary.json"""
import csv
import json
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'sm__cycles_active.avg', 'l1tex__m_xbar2l1tex_read_bytes.sum']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, check=True).stdout.decode()
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[ix['Kernel Name']]
        short = name.split('(')[0].replace('void ', '').replace('<unnamed>::', '').replace('ttl_mlp::', '')
        e = {'kernel': short}
        for k in KEEP:
            if k in ix:
                e[k] = ('%s %s' % (r[ix[k]], units[ix[k]])).strip()
        res.append(e)
    json.dump(res, open(out, 'w'), indent=1)
    for e in res:
        print(e['kernel'][:60], e.get('gpu__time_duration.sum'), e.get('dram__bytes_read.sum'), e.get('dram__bytes_write.sum'),
              e.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'))


if __name__ == '__main__':
    main()
