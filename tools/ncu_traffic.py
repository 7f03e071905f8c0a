"""profiles/ncu_traffic.json from an `ncu --set full` capture of one frame's cost-volume kernels (bench.py reads it
for roofline.traffic / roofline.bound):  python tools/ncu_traffic.py gpurun_out/prof_costvol.ncu-rep
Kernels are recognised by name: k_line2<0,..> = ci_h1, k_line_vv = v2_v3_fused, k_line2<3,..> = v2 then v3 (when the
vertical passes run as two launches), k_line2<2,..> = h4_wta; the first launch of each in the capture is used."""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
iK = col['Kernel Name']
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
tscale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'msecond': 1.0, 'nsecond': 1e-6}


def val(r, name, sc=None):
    i = col[name]
    v = float(r[i])
    return v * (sc or {}).get(units[i], 1.0)


res = {}
nv = 0
for r in rows[2:]:
    k = r[iK]
    if 'k_line_vv' in k:
        name = 'v2_v3_fused'
    else:
        m = re.search(r'k_line2?<\(?(?:int\))?(\d)', k)
        if not m:
            continue
        mode = int(m.group(1))
        if mode == 0:
            name = 'ci_h1'
        elif mode == 2:
            name = 'h4_wta'
        elif mode == 3:
            nv += 1
            name = 'v2' if nv == 1 else 'v3'
        else:
            continue
    if name in res:
        continue
    rd = val(r, 'dram__bytes_read.sum', scale)
    wr = val(r, 'dram__bytes_write.sum', scale)
    res[name] = {'kernel': k, 'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_launch': rd + wr,
                 'ncu_duration_ms': val(r, 'gpu__time_duration.sum', tscale),
                 'issue_active_pct': val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                 'dram_throughput_pct': val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                 'fma_pipe_active_pct': val(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
                 'source': os.path.basename(rep)}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, 'profiles', 'ncu_traffic.json'), 'w') as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
