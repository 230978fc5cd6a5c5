"""opcode mix and hottest source lines of one kernel from an ncu report: python scratch/sass_mix.py rep kernel-regex warps"""
import collections, csv, subprocess, sys
rep, rx, nwarp = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [r for r in rows if len(r) > iSamp and r[iE].isdigit()]
by, bys = collections.Counter(), collections.Counter()
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    by[op] += int(r[iE]); bys[op] += int(r[iSamp])
tot = sum(by.values())
print('total', tot, 'per warp', tot / nwarp, 'samples', sum(bys.values()))
for op, c in by.most_common(28):
    print(f"{op:10s} {c / nwarp:8.1f}  samples {bys[op]}")
