"""Generates tests/golden/*.pt from the CPU oracle (the reference ships no fixtures -- SURVEY.md 8c --
so these pin the ORACLE's behaviour over time and across machines, and give the GPU tests inputs and
expected outputs that do not depend on the test host's CPU).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(1234)
    R, Nc, Nf = 48, 64, 128
    near, far = 1.5 + torch.rand(R, generator=g), 5.0 + torch.rand(R, generator=g)
    t = torch.linspace(0, 1, Nc)
    u_s, u_f = torch.rand(R, Nc, generator=g), torch.rand(R, Nf, generator=g)
    z = O.stratified(near, far, t, u_s)
    raw = torch.randn(R, Nc, 4, generator=g) * 2
    dn = 1 + torch.rand(R, generator=g)
    comp = O.raw2outputs(raw, z, dn)
    sp = O.sample_pdf(z, comp["weights"], u_f)
    g_rgb, g_d, g_a = torch.randn(R, 3, generator=g), torch.randn(R, generator=g), torch.randn(R, generator=g)
    g_raw = O.composite_bwd(raw.double(), z.double(), dn.double(), g_rgb.double(), g_d.double(), g_a.double()).float()
    torch.save({"near": near, "far": far, "t_vals": t, "u_strat": u_s, "u_fine": u_f, "z": z, "raw": raw, "dnorm": dn,
                "rgb": comp["rgb"], "depth": comp["depth"], "acc": comp["acc"], "disp": comp["disp"],
                "weights": comp["weights"], "inds": sp["inds"].int(), "z_samples": sp["z_samples"], "z_f": sp["z_f"],
                "g_rgb": g_rgb, "g_depth": g_d, "g_acc": g_a, "g_raw_fp64": g_raw},
               os.path.join(OUT, "sampling_compositing.pt"))

    # network query: 96 samples, seed-7 network, fp64 evaluation of A.3/A.4 as the expected value
    Rq, Sq = 12, 8
    o = torch.rand(Rq, 3, generator=g) * 2 - 1
    d = torch.randn(Rq, 3, generator=g)
    zq = torch.sort(torch.rand(Rq, Sq, generator=g) * 4 + 2, -1)[0]
    vd, _ = O.ray_setup(d)
    p = O.init_params(7)
    pts = o[:, None, :] + d[:, None, :] * zq[:, :, None]
    p64 = {k: v.double() for k, v in p.items()}
    raw64 = O.run_network(p64, pts.double(), vd.double())
    torch.save({"seed": 7, "rays_o": o, "rays_d": d, "viewdirs": vd, "z": zq, "raw_fp64": raw64.float(),
                "raw_fp32_oracle": O.run_network(p, pts, vd)}, os.path.join(OUT, "network_query.pt"))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
