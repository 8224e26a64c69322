#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into profiles/<name>.csv and
update profiles/ncu_traffic.json (dram bytes per launch, used by bench.py's roofline.traffic)."""
import csv, io, json, os, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'lts__t_sectors.sum.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sectors.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_ltcfabric.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']

rep, name, title, key = sys.argv[1:5]
sass_sha = sys.argv[5] if len(sys.argv) > 5 else None      # bench.py's kernel_sass_sha16 of the build that was profiled
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
def first(*names):
    for h in names:
        if h in d and d[h] not in ('', 'n/a'):
            return h
    return None
l2h = first('lts__t_bytes.sum', 'lts__t_sectors.sum')
l2_bytes = None
if l2h:
    l2_bytes = to_bytes(l2h) if 'bytes' in l2h else float(d[l2h]) * 32.0
t[key] = {'kernel_sass_sha16': sass_sha, 'l2_bytes': l2_bytes,
          'lts_throughput_pct': float(d['lts__throughput.avg.pct_of_peak_sustained_elapsed']) if 'lts__throughput.avg.pct_of_peak_sustained_elapsed' in d else None,
          'lts_sectors_pct': float(d['lts__t_sectors.sum.pct_of_peak_sustained_elapsed']) if 'lts__t_sectors.sum.pct_of_peak_sustained_elapsed' in d else None,
          'lts_hit_rate_pct': float(d['lts__t_sector_hit_rate.pct']) if 'lts__t_sector_hit_rate.pct' in d else None,
          'dram_bytes_read': to_bytes('dram__bytes_read.sum'), 'dram_bytes_write': to_bytes('dram__bytes_write.sum'),
          'duration_ms_under_ncu': float(d['gpu__time_duration.sum']) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}[u['gpu__time_duration.sum']],
          'capture': name + '.csv', 'title': title}
json.dump(t, open(tj, 'w'), indent=1, sort_keys=True)
print(key, t[key])
