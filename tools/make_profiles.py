"""Turn one tools/capture_profiles.sh run (gpurun_out/TAG_*) into the tracked files under profiles/ (prefix OUT):
   python tools/make_profiles.py TAG OUT"""
import collections, csv, json, os, shutil, subprocess, sys
tag, out = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"

def keys(rep, dst, *hdr):
    txt = subprocess.run([sys.executable, "tools/ncu_keys.py", f"{G}/{tag}_{rep}.ncu-rep", *hdr], capture_output=True, text=True).stdout
    open(f"{P}/{out}_{dst}_ncu_key_metrics.txt", "w").write(txt)

def launches(name, what):
    src = f"{G}/{tag}_{name}_launches.csv"
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and (r[0].isdigit() or r[0] == "ID")]
    with open(f"{P}/{out}_{name}_launch_list.csv", "w", newline="") as f:
        csv.writer(f).writerows(rows)
    by = collections.OrderedDict()
    for r in rows[1:]:
        by.setdefault(r[0], {"k": r[4]})[r[12]] = float(r[14].replace(",", ""))
    agg = collections.OrderedDict()
    for d in by.values():
        a = agg.setdefault(d["k"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0) / 1e3
        a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values())
    with open(f"{P}/{out}_{name}_launch_summary.txt", "w") as f:
        f.write(f"# {what}\n# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':<72}{'launches':>9}{'total us':>12}{'share':>8}{'dram rd MB':>12}{'dram wr MB':>12}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:70]:<72}{a[0]:>9}{a[1]:>12.1f}{a[1] / tot:>8.3f}{a[2] / 1e6:>12.1f}{a[3] / 1e6:>12.1f}\n")
        f.write(f"{'total':<72}{sum(a[0] for a in agg.values()):>9}{tot:>12.1f}\n")

shutil.copy(f"{G}/{tag}_bench.json", f"{P}/{out}_bench_1gpu.json")
shutil.copy(f"{G}/{tag}_bench_reference.json", f"{P}/{out}_bench_reference_arm.json")
shutil.copy(f"{G}/{tag}_pytest.log", f"{P}/{out}_pytest_gpu.log")
keys("fine", "bench_fine_launch",
     "ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 7 -c 1 python bench.py --only-render --no-cpu-baseline --steps 1 --warmup 3",
     "= the fine-pass network query of one bench step (640,000 rays x 192 samples) with alpha compositing fused in: raw[R,S,4] is not written",
     "algorithmic HBM bytes of this launch: 4 B (z) per sample + 28 B of maps per ray = 509.5 MB")
keys("train", "train_kernels",
     "ncu --set full --clock-control none --import-source on -k regex:'k_mlp_tc|k_mlp_bwd_pipe' -s 8 -c 4 python tools/bench_train.py --steps 2 --warmup 2",
     "one training step of 4096 rays (64 + 128 samples): tape forward coarse, tape forward fine, pipelined backward fine, pipelined backward coarse")
keys("stages", "stage_kernels",
     "ncu --set full --clock-control none --import-source on -k regex:'k_composite|k_importance|k_stratified' python tools/stage_ncu.py",
     "R = 2^20 rays per launch; order: stratified 64, importance 64/128, importance 256/768, then per S in (64, 192, 1024): composite fwd, composite bwd")
launches("render", "ncu launch list of: python bench.py --only-render --no-cpu-baseline --steps 2 --warmup 3 (5 render steps of 640,000 rays + the e2e steps)")
launches("train", "ncu launch list of: python tools/bench_train.py --steps 3 --warmup 2 (5 training steps of 4096 rays)")
print(os.listdir(P))
