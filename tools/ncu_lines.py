"""Aggregate an ncu --page source (cuda,sass) export per CUDA source line."""
import csv, collections, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kernel}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname = None; agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0] and r[0].isdigit():
        key = (fname, int(r[0]), r[1][:95])
        try: n = float(r[7]); s = float(r[6])
        except ValueError: n = s = 0
        a = agg.setdefault(key, [0, 0]); a[0] += n; a[1] += s
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-inst %.4g samples %d" % (tot, ts))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]/tot*100:5.1f}% inst {a[1]/ts*100:5.1f}% samp  {k[0]}:{k[1]:>4} | {k[2]}")
