"""Generate the golden vectors that pin the CPU oracle: outputs of the REFERENCE's own CUDA extensions
(oracle/_ref/_ref_*.so, built unmodified from /root/reference by oracle/build_ref.py) on seeded inputs.

Run on a B200 box:   python tests/golden/make_golden.py gpurun_out/golden     (then copy the .npz into tests/golden/)
Inputs are regenerated from seeds by tests/golden/cases.py on both sides, so only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from conftest import load_ref  # noqa: E402
import cases  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    rm, ge, fe, she = load_ref("raymarching"), load_ref("gridencoder"), load_ref("freqencoder"), load_ref("shencoder")
    assert rm and ge and fe and she, "build the reference extensions first (python oracle/build_ref.py)"

    # ---------------- raymarching ----------------
    c = cases.raymarch_case()
    N, H, cas = c["rays_o"].shape[0], c["H"], c["cascade"]
    out = {}
    ro, rd, aabb, bits = T(c["rays_o"]), T(c["rays_d"]), T(c["aabb"]), T(c["bitfield"])
    nears = torch.empty(N, device=dev); fars = torch.empty(N, device=dev)
    rm.near_far_from_aabb(ro, rd, aabb, N, c["min_near"], nears, fars)
    out["nears"], out["fars"] = nears.cpu().numpy(), fars.cpu().numpy()
    coords = T(c["coords"]); ind = torch.empty(coords.shape[0], dtype=torch.int32, device=dev)
    rm.morton3D(coords, coords.shape[0], ind)
    inv = torch.empty(coords.shape[0], 3, dtype=torch.int32, device=dev)
    rm.morton3D_invert(ind, coords.shape[0], inv)
    out["morton"], out["morton_inv"] = ind.cpu().numpy(), inv.cpu().numpy()
    grid = T(c["density"]); pk = torch.empty(grid.numel() // 8, dtype=torch.uint8, device=dev)
    rm.packbits(grid, grid.numel() // 8, c["thresh"], pk)
    out["packbits"] = pk.cpu().numpy()
    for tag, dt_gamma, noises in (("g0", 0.0, np.zeros(N, np.float32)), ("g128", 1.0 / 128, c["noises"])):
        M = N * 256
        xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
        rays = torch.zeros(N, 3, dtype=torch.int32, device=dev); counter = torch.zeros(2, dtype=torch.int32, device=dev)
        rm.march_rays_train(ro, rd, bits, c["bound"], dt_gamma, c["max_steps"], N, cas, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, T(noises))
        r = rays.cpu().numpy(); x = xyzs.cpu().numpy(); de = deltas.cpu().numpy()
        order = np.argsort(r[:, 0])
        r = r[order]
        cnt = r[:, 2]
        # ray-ordered repacking (the reference's own packing order is atomics dependent)
        xs = np.concatenate([x[o:o + k] for _, o, k in r]) if cnt.sum() else np.zeros((0, 3), np.float32)
        ds = np.concatenate([de[o:o + k] for _, o, k in r]) if cnt.sum() else np.zeros((0, 2), np.float32)
        out["march_%s_counts" % tag] = cnt.astype(np.int32)
        out["march_%s_xyzs" % tag] = xs
        out["march_%s_deltas" % tag] = ds
        out["march_%s_counter" % tag] = counter.cpu().numpy()
        if tag == "g0":
            # composite forward / backward on the ray-ordered packing with seeded sigmas / rgbs
            tot = int(cnt.sum())
            offs = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32)
            rays_sorted = np.stack([np.arange(N, dtype=np.int32), offs, cnt.astype(np.int32)], 1)
            sig, rgb, gws, gim = cases.composite_inputs(tot, N)
            ws = torch.empty(N, device=dev); dp = torch.empty(N, device=dev); im = torch.empty(N, 3, device=dev)
            rm.composite_rays_train_forward(T(sig), T(rgb), T(ds), T(rays_sorted), tot, N, 1e-4, ws, dp, im)
            gs = torch.zeros(tot, device=dev); gc = torch.zeros(tot, 3, device=dev)
            rm.composite_rays_train_backward(T(gws), T(gim), T(sig), T(rgb), T(ds), T(rays_sorted), ws, im, tot, N, 1e-4, gs, gc)
            out["comp_ws"], out["comp_depth"], out["comp_image"] = ws.cpu().numpy(), dp.cpu().numpy(), im.cpu().numpy()
            out["comp_gsig"], out["comp_grgb"] = gs.cpu().numpy(), gc.cpu().numpy()
    # one inference round: march_rays (n_step 4) + composite_rays
    n_step = 4
    alive = torch.arange(N, dtype=torch.int32, device=dev); rays_t = nears.clone()
    Mi = N * n_step
    xyzs = torch.zeros(Mi, 3, device=dev); dirs = torch.zeros(Mi, 3, device=dev); deltas = torch.zeros(Mi, 2, device=dev)
    rm.march_rays(N, n_step, alive, rays_t, ro, rd, c["bound"], 0.0, c["max_steps"], cas, H, bits, nears, fars, xyzs, dirs, deltas, torch.zeros(N, device=dev))
    sig, rgb, _, _ = cases.composite_inputs(Mi, N, seed=5)
    ws = torch.zeros(N, device=dev); dp = torch.zeros(N, device=dev); im = torch.zeros(N, 3, device=dev)
    rm.composite_rays(N, n_step, 1e-2, alive, rays_t, T(sig), T(rgb), deltas, ws, dp, im)
    out.update(inf_xyzs=xyzs.cpu().numpy(), inf_deltas=deltas.cpu().numpy(), inf_alive=alive.cpu().numpy(), inf_rays_t=rays_t.cpu().numpy(),
               inf_ws=ws.cpu().numpy(), inf_depth=dp.cpu().numpy(), inf_image=im.cpu().numpy())
    np.savez_compressed(os.path.join(out_dir, "raymarch.npz"), **out)

    # ---------------- grid encoder ----------------
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    out = {}
    for name, g in cases.grid_cases().items():
        x, table, offsets = T(g["x"]), T(g["table"]), T(g["offsets"])
        B, D = g["x"].shape
        L, C = g["L"], g["C"]
        o = torch.empty(L, B, C, device=dev); dy = torch.empty(B, L * D * C, device=dev)
        ge.grid_encode_forward(x, table, offsets, o, B, D, C, L, g["S"], g["H"], dy, g["gridtype"], g["align"], g["interp"])
        grad = T(g["grad"])
        gt = torch.zeros_like(table); gx = torch.zeros(B, D, device=dev)
        ge.grid_encode_backward(grad.view(B, L, C).permute(1, 0, 2).contiguous(), x, table, offsets, gt, B, D, C, L, g["S"], g["H"], dy, gx,
                                g["gridtype"], g["align"], g["interp"])
        out[name + "_out"] = o.permute(1, 0, 2).reshape(B, L * C).cpu().numpy()
        out[name + "_dydx"] = dy.view(B, L, D, C).cpu().numpy()
        out[name + "_gtable"] = gt.cpu().numpy()
        out[name + "_gx"] = gx.cpu().numpy()
        # per-level scales as the device evaluates exp2f (ex2.approx): needed by the oracle for bit-exact indices
        idx = torch.empty(1, L, 1 << D, dtype=torch.int32, device=dev); sc = torch.empty(L, device=dev); rs = torch.empty(L, dtype=torch.int32, device=dev)
        _lib.call("seald_grid_debug_indices", ptr(torch.zeros(1, D, device=dev)), ptr(offsets), ptr(idx), ptr(sc), ptr(rs), 1, D, L, g["S"], g["H"],
                  g["gridtype"], int(g["align"]), _lib.stream())
        out[name + "_scales"] = sc.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "grid.npz"), **out)

    # ---------------- small encoders ----------------
    out = {}
    e = cases.encoder_case()
    for D, deg, key in ((3, 10, "x"), (1, 6, "t")):
        xin = T(e[key]); Cc = D + 2 * D * deg
        o = torch.empty(xin.shape[0], Cc, device=dev)
        fe.freq_encode_forward(xin, xin.shape[0], D, deg, Cc, o)
        gi = torch.zeros_like(xin)
        fe.freq_encode_backward(T(e["g_" + key]), o, xin.shape[0], D, deg, Cc, gi)
        out["freq_%s" % key], out["freq_%s_gin" % key] = o.cpu().numpy(), gi.cpu().numpy()
    d = T(e["dirs"]); o = torch.empty(d.shape[0], 16, device=dev); dy = torch.empty(d.shape[0], 48, device=dev)
    she.sh_encode_forward(d, o, d.shape[0], 3, 4, dy)
    out["sh"], out["sh_dydx"] = o.cpu().numpy(), dy.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "encoders.npz"), **out)
    print("golden vectors written to", out_dir, {f: os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir)})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
