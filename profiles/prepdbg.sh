#!/bin/bash
for d in 0 16 32 48; do
ART_B200_UDBG=$d ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file /tmp/l.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-streams 1 > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("/tmp/l.csv")) if len(r) > 5 and r[0].isdigit()]
print("UDBG=$d", [(r[4][:20], r[-1]) for r in rows if "prep" in r[4]][:3])
PY
done
