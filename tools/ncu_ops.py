import csv, subprocess, sys, collections, io
rep, pat = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; data = []
for r in rows:
    if 'Source' in r and '# Samples' in r:
        if h is not None: break
        h = r; continue
    if h and len(r) == len(h): data.append(r)
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
ops = collections.Counter(); ex = collections.Counter()
for r in data:
    s = r[isrc].split()
    o = s[1] if s[0].startswith('@') else s[0]
    o = o.split('.')[0] + ('.' + o.split('.')[1] if o.startswith(('LD', 'ST', 'ATOM', 'RED', 'BAR')) and '.' in o else '')
    ops[o] += int(r[isamp]); ex[o] += int(r[iex])
T = sum(ops.values()); E = sum(ex.values())
for k, v in ops.most_common(18): print("%-14s samples %5.1f%%  instr %5.1f%%" % (k, 100 * v / T, 100 * ex[k] / E))
