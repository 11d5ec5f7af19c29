"""Condense the output of tools/probe_convs.sh into one line per (shape, probe)."""
import re
import sys

t = open(sys.argv[1]).read().split('== ')
for blk in t[1:]:
    lines = blk.strip().split('\n')
    prof = [l for l in lines if 'igemm prof' in l]
    tf = [l for l in lines if 'TF/s' in l and 'prof' not in l]
    m = re.search(r'mma total (\d+) \(wait full (\d+), wait tempty (\d+)\) \| producer wait empty (\d+) \| epi warp total (\d+) '
                  r'\(wait tfull (\d+).*-> (\d+) cyc/tile, (\d+) cyc/slice', prof[-1]) if prof else None
    us = re.search(r'([\d.]+) us\s+([\d.]+) TF/s\s+(.*)', tf[-1]) if tf else None
    if not (m and us):
        print(lines[0], '| no data')
        continue
    print(f"{lines[0]} | {us.group(1)} us {us.group(3)} | mma {m.group(1)} wfull {m.group(2)} wtempty {m.group(3)} "
          f"prod_wempty {m.group(4)} epi {m.group(5)} epi_wtfull {m.group(6)} cyc/slice {m.group(8)}")
