"""Hot instructions of one kernel launch from an ncu report (source page): python tools/ncu_hot.py rep.ncu-rep [launch] [min_frac]"""
import csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.008
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[which]:starts[which + 1]]
print(blk[0][:2])
H = blk[1]; data = [r for r in blk[2:] if len(r) >= len(H) - 2]
isrc = H.index('Source'); isamp = H.index('# Samples'); iex = H.index('Instructions Executed')
stalls = [i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
print('total samples', tot, 'total warp-inst', totex, 'sass lines', len(data))
for b in range(0, len(data), 250):
    c = data[b:b + 250]
    print(f"  sass[{b:5d}..] samples {sum(int(r[isamp]) for r in c):7d}  inst {sum(int(r[iex]) for r in c):10d}")
for n, r in enumerate(data):
    s = int(r[isamp])
    if s > tot * frac:
        st = sorted(((int(r[i]), H[i]) for i in stalls), reverse=True)[:2]
        print(n, r[isrc].strip()[:70], s, r[iex], st)
