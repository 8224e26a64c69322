#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into profiles/<name>.csv and
update profiles/ncu_traffic.json (dram bytes per launch, used by bench.py's roofline.traffic)."""
import csv, io, json, os, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_ltcfabric.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']

rep, name, title, key = sys.argv[1:5]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, 'profiles', name + '.csv'), 'w') as f:
    f.write('# %s\n# kernel: %s\n# source: ncu --set full --clock-control none --import-source on (one launch), ncu -i --page raw --csv\n' % (title, d.get('Kernel Name', '')))
    f.write('metric,unit,value\n')
    for h in hdr:
        if h in KEEP or h.startswith('smsp__average_warps_issue_stalled'):
            f.write('%s,%s,%s\n' % (h, u[h], d[h]))
def to_bytes(h):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[u[h]]
    return float(d[h]) * scale
tj = os.path.join(root, 'profiles', 'ncu_traffic.json')
t = json.load(open(tj)) if os.path.exists(tj) else {}
t[key] = {'dram_bytes_read': to_bytes('dram__bytes_read.sum'), 'dram_bytes_write': to_bytes('dram__bytes_write.sum'),
          'duration_ms_under_ncu': float(d['gpu__time_duration.sum']) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}[u['gpu__time_duration.sum']],
          'capture': name + '.csv', 'title': title}
json.dump(t, open(tj, 'w'), indent=1, sort_keys=True)
print(key, t[key])
