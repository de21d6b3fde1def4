"""Per-source-line instruction / stall-sample shares from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            data.append((cur, int(r[0]), r[1], int(r[4] or 0), int(r[7] or 0)))
        except ValueError:
            pass
ts = sum(d[3] for d in data) or 1
ti = sum(d[4] for d in data) or 1
print("total stall samples", ts, "instructions", ti)
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%5.1f%% stall %5.1f%% inst  %s:%d  %s" % (100 * d[3] / ts, 100 * d[4] / ti, d[0], d[1], d[2][:100]))
