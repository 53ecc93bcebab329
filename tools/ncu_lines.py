"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line."""
import collections
import csv
import sys


def num(x):
  try:
    return int(x)
  except ValueError:
    return 0


def main(path, top=45):
  rows = list(csv.reader(open(path)))
  hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
  half = len(hdr_idx) // 2 or len(hdr_idx)     # one set of sections per profiled launch
  tot, stall, ninstr, srcs = collections.Counter(), collections.Counter(), collections.Counter(), {}
  nsass = 0
  for hi in hdr_idx[:half]:
    hdr = rows[hi]
    fpath = rows[hi - 2][1].split('/')[-1]
    iE, iS = hdr.index('Instructions Executed'), hdr.index('# Samples')
    cur = None
    for r in rows[hi + 1:]:
      if not r or r[0] in ('File Path', 'Function Name', 'Line No'):
        break
      if r[0] != '':
        cur = (fpath, num(r[0]))
        srcs[cur] = r[1]
        tot[cur] += num(r[iE])
        stall[cur] += num(r[iS])
      elif r[2].startswith('0x'):
        ninstr[cur] += 1
        nsass += 1
  T, S = sum(tot.values()), sum(stall.values())
  print('total warp-inst', T, 'samples', S, 'static SASS rows shown', nsass)
  for k, v in tot.most_common(top):
    print('%-24s %5.1f%% inst %5.1f%% smp | %s' % (k[0][:18] + ':' + str(k[1]), 100 * v / T, 100 * stall[k] / max(S, 1), srcs[k].strip()[:90]))


if __name__ == '__main__':
  main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
