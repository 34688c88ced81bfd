"""Aggregate an ncu report's warp-stall samples per CUDA source line:  python tools/ncu_lines.py report.ncu-rep [N]
(needs a capture taken with --import-source on and a build with -lineinfo)."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, agg = None, []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0].isdigit():
        try:
            agg.append((int(r[6]), cur, int(r[0]), r[1][:120]))
        except (ValueError, IndexError):
            pass
tot = sum(a[0] for a in agg) or 1
print("total samples", tot)
for n, f, l, src in sorted(agg, reverse=True)[:top]:
    print(f"{n:6d} {n / tot * 100:5.1f}%  {f}:{l}  {src}")
