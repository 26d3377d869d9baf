"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, time, share.
Usage: python profiles/summarize_launches.py gpurun_out/launches.csv [--per-launch REGEX]"""
import collections
import csv
import re
import sys


def load(path):
  with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
  out = []
  for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
      continue
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    us = v / 1000.0 if unit in ('ns', 'nsecond') else (v * 1000.0 if unit in ('ms', 'msecond') else v)
    out.append((int(row['ID']), re.sub(r'\(.*', '', row['Kernel Name']), row['Grid Size'], row['Block Size'], us))
  return out


def main():
  rows = load(sys.argv[1])
  agg = collections.OrderedDict()
  tot = sum(r[4] for r in rows)
  for _, name, _, _, us in rows:
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
  print('launches %d, total %.1f us (cold-cache, serialised under ncu: compare SHARES)' % (len(rows), tot))
  for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-70s n=%4d %10.1f us %5.1f%%' % (k[:70], n, us, 100 * us / tot))
  if len(sys.argv) > 3 and sys.argv[2] == '--per-launch':
    pat = re.compile(sys.argv[3])
    for i, name, grid, blk, us in rows:
      if pat.search(name):
        print('%5d %-50s grid %-14s %9.1f us' % (i, name[:50], grid, us))


if __name__ == '__main__':
  main()
