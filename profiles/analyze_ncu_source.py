"""Per-instruction PC-sampling tables from an `ncu --set full --import-source on` report, read offline:

    ncu -i gpurun_out/tc_v12.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > src.csv
    python profiles/analyze_ncu_source.py src.csv [first_index last_index]

Prints the sample share of the instructions that matter for a warp-specialised tcgen05 kernel (barrier waits, TMA
issue, MMA issue, TMEM loads, global loads / stores) and, for an index range (one warp role), how its samples split
between waiting on mbarriers and issuing instructions.
"""
import csv
import sys

KEYS = ('TRYWAIT', 'UTCHMMA', 'UTCQMMA', 'UTMALDG', 'LDTM', 'UTCBAR', 'ARRIVE', 'STG', 'LDG', 'BAR.', 'RED.')


def load(path):
  rows = list(csv.reader(open(path)))
  hdr = rows[1]
  ix = {h: i for i, h in enumerate(hdr)}
  seen, out = set(), []
  for r in rows[2:]:
    if len(r) != len(hdr) or not r[ix['# Samples']].isdigit() or r[0] in seen:
      continue                                   # the page lists every instruction twice; keep the first copy
    seen.add(r[0])
    out.append(r)
  return hdr, ix, out


def main():
  hdr, ix, rows = load(sys.argv[1])
  S, IE = ix['# Samples'], ix['Instructions Executed']
  total = sum(int(r[S]) for r in rows)
  print('instructions %d, samples %d' % (len(rows), total))
  for i, r in enumerate(rows):
    if any(k in r[1] for k in KEYS):
      loop = sum(int(x[S]) for x in rows[i:i + 4])      # the wait instruction plus its spin branch
      print('%5d %6s %7d %9s  %s' % (i, r[S], loop, r[IE], r[1].strip()[:100]))
  if len(sys.argv) >= 4:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    region = rows[a:b]
    samples = sum(int(r[S]) for r in region)
    waits = sum(sum(int(x[S]) for x in region[i:i + 2]) for i, r in enumerate(region) if 'TRYWAIT' in r[1])
    print('region [%d, %d): %d samples (%.1f %% of all), %d of them waiting on mbarriers (%.0f %%), %d instructions'
          % (a, b, samples, 100.0 * samples / total, waits, 100.0 * waits / max(samples, 1), len(region)))


if __name__ == '__main__':
  main()
