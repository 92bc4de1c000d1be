"""Summarise `ncu --page source --csv` output: top SASS instructions by stall samples + stall-reason totals."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
recs = []
for r in data:
    try:
        n = int(r[ix["# Samples"]])
    except Exception:
        continue
    recs.append((n, r))
    for s in stall_cols:
        try: tot[s] += int(r[ix[s]])
        except Exception: pass
total = sum(n for n, _ in recs)
print("total samples", total)
print("stall totals:", ", ".join(f"{k[6:]}={v} ({100*v/max(total,1):.1f}%)" for k, v in tot.most_common(12)))
recs_sorted = sorted(recs, key=lambda x: -x[0])[:top]
for n, r in recs_sorted:
    reasons = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:3]
    print(f"{n:7d} {100*n/total:5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:90]:90s}  " + " ".join(f"{b}:{a}" for a, b in reasons if a))
