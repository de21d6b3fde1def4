"""Run the REFERENCE (its own Python over its own extensions + cuBLAS, oracle/ref_runtime.py) and the product side by side on a
B200, print how far apart they are per case, and write the reference's outputs as fixtures.

    python tests/golden/make_ref_golden.py [out_dir=gpurun_out/golden_ref] [case ...]      (on the GPU box)

then copy the *.npz into tests/golden/.  tests/test_gpu_ref_parity.py holds the tolerances; this script is where they were
calibrated (it prints max-abs / relative-L2 / 99.9th-percentile differences).  Cases: field train pretrain frame occupancy ffmlp teacher.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import ref_cases as rc  # noqa: E402


def diff(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.shape != b.shape:
        return {"shape_mismatch": [list(a.shape), list(b.shape)]}
    if a.size == 0:
        return {"empty": True}
    d = np.abs(a - b)
    ok = np.isfinite(d)
    den = float(np.sqrt((b[ok] ** 2).sum())) or 1.0
    return {"max_abs": float(d[ok].max()) if ok.any() else None, "p999": float(np.quantile(d[ok], 0.999)), "rel_l2": float(np.sqrt((d[ok] ** 2).sum()) / den),
            "ref_absmax": float(np.abs(b[ok]).max()), "nonfinite": int((~ok).sum())}


def report(name, ours, ref):
    rows = {}
    for k in ref:
        if k in ours:
            if ref[k].dtype == np.uint8:
                x = np.unpackbits(np.asarray(ours[k])) != np.unpackbits(np.asarray(ref[k]))
                rows[k] = {"bit_mismatch_frac": float(x.mean()), "bits": int(x.size)}
            else:
                rows[k] = diff(ours[k], ref[k])
    print(json.dumps({"case": name, "diff": rows}), flush=True)
    return rows


def small(res, case):
    """Trim a reference result to fixture size."""
    out = dict(res)
    if case == "frame":
        out = {k: v[::rc.FRAME_STRIDE].copy() for k, v in res.items()}
    return out


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_ref")
    which = sys.argv[2:] or ["field", "train", "pretrain", "frame", "occupancy", "teacher", "ffmlp"]
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")

    if "ffmlp_ref_child" in which:  # the reference's ffmlp (Sm70-tagged CUTLASS GEMMs) in its own process: a crash must not take the rest down
        from oracle import ref_runtime as rr
        rr.install()
        from ffmlp.ffmlp import FFMLP as RefFFMLP
        res = rc.run_ffmlp(RefFFMLP, dev)
        np.savez_compressed(os.path.join(out_dir, "ref_ffmlp.npz"), **res)
        print("ffmlp reference ok", flush=True)
        return

    t0 = time.time()
    ours = rc.ours_model(dev)
    ref = rc.ref_model(ours)
    print("models built %.1fs" % (time.time() - t0), flush=True)

    if "field" in which:
        r, o = rc.ref_field(ref, dev), rc.ours_field(ours, dev)
        report("field", o, r)
        np.savez_compressed(os.path.join(out_dir, "ref_field.npz"), **r)
    if "train" in which:
        r, o = rc.ref_train(ref, dev), rc.ours_train(ours, dev)
        report("train", o, r)
        print(json.dumps({"train_loss": [float(o["loss"]), float(r["loss"])], "samples": [int(o["samples"]), int(r["samples"])]}), flush=True)
        np.savez_compressed(os.path.join(out_dir, "ref_train.npz"), **r)
    if "pretrain" in which:
        so = rc.ours_model(dev, seald=True)
        sr = rc.ref_model(so, seald=True)
        r, o = rc.ref_pretrain(sr, dev), rc.ours_pretrain(so, dev)
        report("pretrain", o, r)
        solid = np.abs(r["grad_table_head"]) > 1e-2 * np.abs(r["grad_table_head"]).max()
        same = np.abs(o["table_after_head"] - r["table_after_head"])[solid] < 1e-3 * rc.PRETRAIN_LR
        print(json.dumps({"pretrain_loss": [float(o["loss"]), float(r["loss"])], "solid_entries": int(solid.sum()), "same_update_frac": float(same.mean())}), flush=True)
        np.savez_compressed(os.path.join(out_dir, "ref_pretrain.npz"), **r)
        del so, sr
    if "frame" in which:
        for tag, T in (("", None),):
            r, o = rc.ref_frame(ref, dev, T_thresh=T), rc.ours_frame(ours, dev, T_thresh=T)
            report("frame" + tag, o, r)
            od = rc.ours_frame_dropin(ours, dev)
            report("frame_dropin" + tag, od, r)
            np.savez_compressed(os.path.join(out_dir, "ref_frame%s.npz" % tag), **small(r, "frame"))
    if "occupancy" in which:
        t0 = time.time()
        r = rc.run_occupancy(ref, dev, fused=False)
        t1 = time.time()
        o = rc.run_occupancy(ours, dev, fused=True)
        t2 = time.time()
        report("occupancy", o, r)
        print(json.dumps({"occupancy_wall_s": {"reference": t1 - t0, "ours": t2 - t1},
                          "occupied_cells": {k: [int(o[k].sum()), int(r[k].sum())] for k in ("full_occupied_cells_per_frame", "partial_occupied_cells_per_frame")},
                          "mean_density": {k: [float(o[k]), float(r[k])] for k in ("full_mean_density", "partial_mean_density")}}), flush=True)
        np.savez_compressed(os.path.join(out_dir, "ref_occupancy.npz"), **r)
        # restore the analytic occupancy grid for the cases below
        ours2 = rc.ours_model(dev)
        ours.load_state_dict(ours2.state_dict())
        ours.mean_density = ours2.mean_density
        del ours2
    if "teacher" in which:
        so = rc.ours_model(dev, seald=True)
        sr = rc.ref_model(so, seald=True)
        for kind in ("bbox", "brush_dry", "brush_linear"):
            r = rc.ref_teacher(sr, dev, kind)
            o = rc.ours_teacher(so, dev, kind)
            report("teacher_" + kind, o, r)
            o1 = rc.ours_teacher(so, dev, kind, one_pass=True)
            report("teacher_one_pass_" + kind, o1, r)
            np.savez_compressed(os.path.join(out_dir, "ref_teacher_%s.npz" % kind), **r)
    if "ffmlp" in which:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), out_dir, "ffmlp_ref_child"], capture_output=True, text=True, timeout=600)
        print("ffmlp reference child rc", p.returncode, p.stdout[-300:], p.stderr[-1500:], flush=True)
        path = os.path.join(out_dir, "ref_ffmlp.npz")
        if p.returncode == 0 and os.path.exists(path):
            from seald_nerf_b200.ffmlp import FFMLP
            r = dict(np.load(path))
            o = rc.run_ffmlp(FFMLP, dev)
            report("ffmlp", o, r)


if __name__ == "__main__":
    main()
