"""optim.py-shaped mesh optimisation on synthetic views -- the drop-in demonstration for config C3
("full optim.py step").  Mirrors the reference loop (optim.py:145-219): per iteration
`vertices = init + parameter` -> `scene.update_verticex` -> ray loss + silhouette loss + smoothness ->
backward -> NaN-zero/clamp hook -> SGD Nesterov; only the imports differ from the reference
(`drt_b200.DiffRender` for `DiffRender`, SyntheticData for the .h5 loaders, no MeshLab remesh).

    python examples/optimize_synthetic.py --mesh mouse_vh --iters 50 --res 240 320
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drt_b200.DiffRender as Render  # noqa: E402
from drt_b200 import configs, losses, synthetic_data  # noqa: E402


def limit_hook(grad, cap=1.0):  # optim.py:155-162
    grad = torch.nan_to_num(grad, nan=0.0)
    return grad.clamp(-cap, cap)


def optimize(vertices0, faces, data, hp, iters, log_every=10, fused_loss=True, remesh_len=None):
    Render.intIOR, Render.resy, Render.resx = hp["IOR"], data.resy, data.resx   # optim.py:178-180
    scene = Render.Scene(vertices=vertices0, faces=faces)
    if remesh_len:  # optim.py:198 `meshlabserver.remesh(scene, remesh_len)`, served by the in-tree stand-in (no MeshLab here)
        from drt_b200 import remesh
        remesh.Remesher().remesh(scene, remesh_len)
    init_vertices = scene.vertices
    parameter = torch.zeros_like(init_vertices, requires_grad=True)
    parameter.register_hook(limit_hook)
    opt = torch.optim.SGD([parameter], lr=hp["start_lr"], momentum=hp["momentum"], nesterov=True)  # optim.py:169
    ray_view, silh_view = data.ray_view_generator(), data.silh_view_generator()
    history = []
    for it in range(iters):
        opt.zero_grad()
        vertices = init_vertices + parameter
        scene.update_verticex(vertices)                                           # optim.py:203
        # ray loss (optim.py:91-108)
        k = next(ray_view)
        if fused_loss:
            # the whole of optim.py:93-106 (+ its backward) in one library call, from the loader's compact view
            ray_loss = losses.ray_loss_view(scene, data.get_view_compact(k))
        else:
            screen, valid, _, origin, ray_dir, _ = data.get_view(k)
            out_ori, out_dir, mask = scene.render_transparent(origin, ray_dir)
            target = screen - out_ori.detach()
            target = target / target.norm(dim=1, keepdim=True)
            ray_loss = (out_dir - target)[valid * mask[:, 0]].pow(2).sum()
        # silhouette loss over 8 views (optim.py:67-80)
        vh_loss = torch.zeros((), dtype=torch.float64, device=init_vertices.device)
        if hp["vh_w"]:
            for _ in range(8):
                _, _, sil, origin, _, cam = data.get_view(next(silh_view))
                edges = scene.silhouette_edge(origin[0])
                index, output = scene.primary_visibility(edges, cam, origin[0], detach_depth=True)
                vh_loss = vh_loss + (sil.view(data.resy, data.resx)[index[:, 1], index[:, 0]] - output).abs().sum()
        # smoothness (optim.py:82-89)
        sm_loss = (-torch.log(1 + scene.dihedral_angle())).sum() if hp["sm_w"] else torch.zeros_like(vh_loss)
        loss = (hp["ray_w"] * 217.5 / data.resy / data.resy * ray_loss + hp["vh_w"] * 217.5 / data.resy * vh_loss
                + hp["sm_w"] * scene.mean_len / 10 * sm_loss)                      # optim.py:127-129
        loss.backward()
        history.append((ray_loss.item(), vh_loss.item(), sm_loss.item()))
        if log_every and it % log_every == 0:  # before the step, like optim.py:212-215 (foreach-SGD rewrites .grad in place)
            print(f"Iteration {it}: ray={history[-1][0]:g} vh={history[-1][1]:g} sm={history[-1][2]:g} "
                  f"maxgrad={parameter.grad.abs().max():g}", flush=True)
        opt.step()
    return scene, history


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="hand_vh")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--res", type=int, nargs=2, default=[240, 320])
    ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--momentum", type=float, default=0.95)
    ap.add_argument("--scale", type=float, default=0.6, help="amplitude (mm) of the target perturbation")
    args = ap.parse_args()
    hp = {"IOR": 1.4723, "ray_w": 40, "sm_w": 0.08, "vh_w": 2e-3, "momentum": args.momentum, "start_lr": args.lr}  # config.py:18-39
    v, f = configs.load_mesh(args.mesh)
    target = configs.perturbed_target_mesh(v, scale=args.scale)
    data = synthetic_data.SyntheticData(target, f, args.res[0], args.res[1], n_views=args.views, num_view=args.views, int_ior=hp["IOR"])
    t0 = time.time()
    scene, hist = optimize(v, f, data, hp, args.iters)
    torch.cuda.synchronize()
    err0 = np.abs(v - target).mean()
    err1 = np.abs(scene.vertices.detach().cpu().numpy() - target).mean()
    print(f"optimize time: {time.time() - t0:.2f} s for {args.iters} iterations; mean |vertex - target| {err0:.4f} -> {err1:.4f} mm; "
          f"ray loss {hist[0][0]:.4g} -> {hist[-1][0]:.4g}")


if __name__ == "__main__":
    main()
