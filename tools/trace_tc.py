"""Timeline tracer for k_mlp_tc: builds libfnerf_trace.so with -DFNERF_TRACE, runs one launch and prints
the MMA issuer's and the worker groups' clock64() stamps for CTA 0's 3rd tile (cycles relative to the
tile's first event).  Usage (GPU box): python tools/trace_tc.py"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "fashion_nerf_b200", "csrc")
OUT = os.path.join(ROOT, "fashion_nerf_b200", "libfnerf_trace.so")


def build(extra=(), out=None):
    from fashion_nerf_b200._build import SOURCES
    srcs = [os.path.join(CSRC, f) for f in SOURCES]
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DFNERF_TRACE", *extra,
           "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared", "-o", out or OUT, *srcs, "-lcudart"]
    subprocess.run(cmd, check=True)


if __name__ == "__main__":
    if sys.argv[1:2] == ["build"]:
        build()
        for tag in sys.argv[2:]:            # experiment variants: python tools/trace_tc.py build EXP_NOFENCE ...
            build([f"-D{x}" for x in tag.split("+")], OUT.replace(".so", f"_{tag}.so"))
        sys.exit(0)
    if os.environ.get("FNERF_TRACE_VARIANT"):
        OUT = OUT.replace(".so", f"_{os.environ['FNERF_TRACE_VARIANT']}.so")
    BRIEF = bool(os.environ.get("FNERF_TRACE_BRIEF"))
    import torch
    import fashion_nerf_b200 as F
    from fashion_nerf_b200 import _lib
    _lib.LIB_PATH = OUT
    lib = F.load_library()
    dev = torch.device("cuda:0")
    buf = torch.zeros(1024, dtype=torch.int64, device=dev)
    assert lib.fnerf_debug_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
    net = F.NerfNetwork.random(1, dev)
    R, S = 16384, 192
    g = torch.Generator().manual_seed(1)
    o = (torch.rand(R, 3, generator=g) * 2 - 1).to(dev); d = torch.randn(R, 3, generator=g).to(dev)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0].to(dev)
    vd, _ = F.ops.ray_setup(d)
    for _ in range(3):
        F.ops.mlp_fwd(net.packed, o, d, vd, z, precision="bf16")
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    mma = [x for x in t[:256] if x]
    t0 = min(x for x in t if x)
    if BRIEF:
        w = t[256:320]
        units = [w[4 * st + 2] - w[4 * st + 1] for st in range(9)] + [w[4 * st + 3] - w[4 * st + 2] for st in range(9)]
        waits = [w[4 * st + 1] - w[4 * st] for st in range(9)]
        print(os.environ.get("FNERF_TRACE_VARIANT", "base"), "tile span", max(t[:117]) - min(x for x in t[:117] if x),
              "mean unit", sum(units) / len(units), "mean acc wait", sum(waits) / len(waits))
        sys.exit(0)
    print("MMA issuer: per chunk (before weights wait, weights ready, issued) relative cycles")
    names = ["L0pe", "L0b"] + [f"L{l}{k}" for l in range(1, 5) for k in ("k0", "k1", "k2", "k3", "b")] + ["L5pe"] + \
            [f"L5{k}" for k in ("k0", "k1", "k2", "k3", "b")] + [f"L{l}{k}" for l in (6, 7) for k in ("k0", "k1", "k2", "k3", "b")] + \
            [f"F{k}" for k in ("k0", "k1", "k2", "k3", "b")] + [f"V{k}" for k in ("k0", "k1", "k2", "k3", "d")]
    for i, n in enumerate(names):
        a, b, c = t[3 * i:3 * i + 3]
        if a:
            print(f"  {n:6s} arrive {a - t0:7d}  w_wait {b - a:5d}  issue {c - b:5d}   (gap to next {(t[3 * i + 3] - c) if 3 * i + 3 < 256 and t[3 * i + 3] else 0:6d})")
    for grp in range(4):
        w = t[256 + grp * 64:256 + grp * 64 + 64]
        print(f"worker group {grp} (row 0): per step (wait start, acc ready, unit0 done, unit1 done)")
        for st in range(9):
            a, b, c, dd = w[4 * st:4 * st + 4]
            if a:
                print(f"  step {st}: wait_start {a - t0:7d}  waited {b - a:5d}  unit0 {c - b:5d}  unit1 {dd - c:5d}")
