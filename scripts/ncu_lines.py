"""aggregate warp-stall samples per CUDA source line from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
cur = None
agg = collections.Counter(); ex = collections.Counter(); text = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0].isdigit() and len(r) > 7 and r[4].isdigit():
        k = (cur, int(r[0])); agg[k] += int(r[4]); ex[k] += int(r[7]); text[k] = r[1].strip()[:100]
tot = sum(agg.values())
print("total samples", tot)
for k, v in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print("%6d %5.1f%% ex=%-9d %s:%d  %s" % (v, 100 * v / tot, ex[k], k[0], k[1], text[k]))
