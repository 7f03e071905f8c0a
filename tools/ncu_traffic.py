"""profiles/ncu_traffic.json from an `ncu --set full` capture of one frame's four cost-volume kernels
(bench.py reads it for roofline.traffic):  python tools/ncu_traffic.py gpurun_out/prof_line_r1f.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
iK = hdr.index('Kernel Name')
iR, iW, iT = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
names = ['ci_h1', 'v2', 'v3', 'h4_wta']
res = {}
for name, r in zip(names, rows[2:6]):
    rd = float(r[iR]) * scale[units[iR]]
    wr = float(r[iW]) * scale[units[iW]]
    res[name] = {'kernel': r[iK], 'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_launch': rd + wr,
                 'ncu_duration_ms': float(r[iT]) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'msecond': 1.0, 'nsecond': 1e-6}.get(units[iT], 1.0),
                 'source': os.path.basename(rep)}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, 'profiles', 'ncu_traffic.json'), 'w') as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
