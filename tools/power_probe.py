"""Is k_mlp_tc power-capped?  Loop the kernel for ~3 s while sampling nvidia-smi (power, SM clock, throttle
reasons); repeat on a smaller problem that only fills half the SMs (FNERF-independent: fewer tiles)."""
import subprocess, sys, time, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
net = F.NerfNetwork.random(1, dev)


def run(R, S, secs=3.0):
    g = torch.Generator().manual_seed(1)
    o = (torch.rand(R, 3, generator=g) * 2 - 1).to(dev); d = torch.randn(R, 3, generator=g).to(dev)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0].to(dev)
    vd, _ = F.ops.ray_setup(d)
    for _ in range(5): F.ops.mlp_fwd(net.packed, o, d, vd, z, precision="bf16")
    torch.cuda.synchronize()
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): F.ops.mlp_fwd(net.packed, o, d, vd, z, precision="bf16")
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    p.terminate()
    lines = [l.strip() for l in p.stdout.read().splitlines() if l.strip()]
    ms = e0.elapsed_time(e1) / n
    tfl = R * S * 1186816 / ms / 1e9
    mid = lines[len(lines) // 2:] if lines else []
    print(f"R={R} S={S}: {ms:.3f} ms/launch, {tfl:.1f} TFLOP/s; nvidia-smi (2nd half): {mid[:3]} ... {mid[-2:]}")


run(16384, 192)          # 24576 tiles: all 148 SMs busy
run(74 * 8, 128)         # 592 tiles * ... = 74*8*128/128 = 592 tiles -> 4 per CTA on 148 CTAs
