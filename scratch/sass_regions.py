"""per-instruction samples of one kernel from an ncu report, printed as a compact listing with cumulative shares:
python scratch/sass_regions.py rep kernel-regex [min_samples]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [r for r in rows if len(r) > iSamp and r[iE].isdigit()]
# the page may list the kernel twice (SASS + PTX views): keep the first run of increasing addresses
iA = hdr.index('Address')
out = []
last = -1
for r in data:
    try: a = int(r[iA], 16)
    except ValueError: break
    if a <= last: break
    last = a; out.append(r)
tot = sum(int(r[iSamp]) for r in out)
print("instructions", len(out), "samples", tot)
cum = 0
for i, r in enumerate(out):
    cum += int(r[iSamp])
    print(f"{i:5d} {int(r[iE]):10d} {int(r[iSamp]):7d} {100*cum/tot:6.1f}%  {r[iS].strip()[:90]}")
