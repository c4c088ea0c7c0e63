"""Collect the DRAM bytes per launch of the quantizer + InfoNCE path kernels from the ncu --set full summaries
(gpurun_out/r02_ncu_<kernel>.txt, written by tools/idx_profile.sh) into profiles/r02_path_traffic.json - the `traffic` entry
of bench.py's roofline_quantizer_infonce.   python tools/path_traffic.py [dir]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out')
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out = {}
for f in sorted(os.listdir(src)):
    m = re.match(r'r02_ncu_(\w+)\.txt$', f)
    if not m:
        continue
    txt = open(os.path.join(src, f)).read()
    ent = {}
    for key, name in (('dram__bytes_read.sum', 'dram_read_bytes'), ('dram__bytes_write.sum', 'dram_write_bytes'),
                      ('gpu__time_duration.sum', 'ncu_us')):
        mm = re.search(re.escape(key) + r' = ([0-9.,]+) (\w+)', txt)
        if mm:
            v = float(mm.group(1).replace(',', ''))
            ent[name] = v * UNIT.get(mm.group(2), 1) if name != 'ncu_us' else v
    if ent:
        out[m.group(1)] = ent
out['note'] = ('per launch at B = 4096 (557 056 vectors), ncu --set full --clock-control none; ncu durations are cold-cache and '
               'serialised (never a bench value)')
path = os.path.join(ROOT, 'profiles', 'r02_path_traffic.json')
json.dump(out, open(path, 'w'), indent=1)
print(path, json.dumps(out)[:600])
