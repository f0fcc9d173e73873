"""Mean duration / SM cycles / tensor-pipe share per kernel name from an ncu --csv log taken with
   --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
   grouped by consecutive runs of the same kernel (so A/B variants launched one after another stay apart)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r[0], {"k": r[4]})[r[12]] = float(r[14].replace(",", ""))
runs = []
for d in by.values():
    if runs and runs[-1][0] == d["k"]:
        runs[-1][1].append(d)
    else:
        runs.append((d["k"], [d]))
for k, ds in runs:
    if len(ds) < 3:
        continue
    ds = ds[len(ds) // 4:]                      # drop warm-up launches
    t = sum(d["gpu__time_duration.sum"] for d in ds) / len(ds)
    c = sum(d["sm__cycles_elapsed.max"] for d in ds) / len(ds)
    tp = sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) for d in ds) / len(ds)
    print(f"{k[:60]:<60} n={len(ds):3d} {t / 1e6:.3f} ms {c / 1e6:.3f} Mcyc {c / t:.3f} GHz tensor {tp:.1f}%")
