"""Golden vectors for get_rays: runs the REFERENCE'S OWN source of `custom_meshgrid` and `get_rays`
(/root/reference/nerf/utils.py:34-137), extracted with `ast` and executed here on CPU (the module itself cannot be imported: it pulls
in lpips / tensorboardX / torchmetrics), and stores inputs + outputs in tests/golden/rays.npz.

    python tests/golden/make_rays_golden.py        (needs /root/reference; the committed .npz is what the tests read)
"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/nerf/utils.py"


def reference_get_rays():
    src = open(REF).read()
    tree = ast.parse(src)
    wanted = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("custom_meshgrid", "get_rays"):
            node.decorator_list = []  # @torch.cuda.amp.autocast(enabled=False): a no-op for fp32 CPU tensors
            wanted[node.name] = ast.get_source_segment(src, node)
            # get_source_segment keeps decorators out when we cut from `def`
            wanted[node.name] = wanted[node.name][wanted[node.name].index("def "):]
    from packaging import version as pver
    ns = {"torch": torch, "pver": pver, "np": np}
    exec(wanted["custom_meshgrid"], ns)
    exec(wanted["get_rays"], ns)
    return ns["get_rays"]


def main():
    from seald_nerf_b200 import synthetic as syn
    get_rays = reference_get_rays()
    out = {}
    for k, (H, W, N) in enumerate(((800, 800, 4096), (60, 80, 1000))):
        intr = syn.intrinsics(H, W)
        poses = syn.orbit_poses(3, "cpu", seed=k)
        torch.manual_seed(100 + k)
        r = get_rays(poses[1:2], intr, H, W, N)
        out["case%d_pose" % k] = poses[1].numpy()
        out["case%d_intr" % k] = np.asarray(intr, np.float64)
        out["case%d_HWN" % k] = np.array([H, W, N], np.int64)
        out["case%d_inds" % k] = r["inds"][0].numpy()
        out["case%d_rays_o" % k] = r["rays_o"][0].contiguous().numpy()
        out["case%d_rays_d" % k] = r["rays_d"][0].contiguous().numpy()
        # collate's ground-truth gather (dnerf/provider.py:340-343) on a synthetic RGBA image
        g = torch.Generator().manual_seed(7 + k)
        img = torch.rand(1, H, W, 4, generator=g)
        C = 4
        gt = torch.gather(img.view(1, -1, C), 1, torch.stack(C * [r["inds"]], -1))[0]
        out["case%d_image" % k] = img[0].reshape(-1, 4).numpy() if H * W <= 10000 else np.zeros((0, 4), np.float32)
        out["case%d_image_seed" % k] = np.array([7 + k], np.int64)
        out["case%d_gt_rgba" % k] = gt.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rays.npz"), **out)
    print("wrote tests/golden/rays.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
