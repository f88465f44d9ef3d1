"""median SM clock / instantaneous power over the loaded samples (> 400 W) of a tools/power_trace.sh csv"""
import sys, csv, statistics
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) >= 5]
    busy = [(float(r[1]), float(r[2]), r[4].strip()) for r in rows if r[2].strip() not in ("N/A", "[N/A]") and float(r[2]) > 400]
    if not busy:
        print(f, "no loaded samples"); continue
    print(f"{f}: loaded samples {len(busy)}  sm clock median {statistics.median(b[0] for b in busy):.0f} MHz  "
          f"power median {statistics.median(b[1] for b in busy):.0f} W  max {max(b[1] for b in busy):.0f} W  "
          f"sw_power_cap active in {sum(b[2] == 'Active' for b in busy)} samples")
