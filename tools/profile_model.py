#!/usr/bin/env python
"""Where a whole-network training step spends its GPU time (torch.profiler, one GPU, eager launches).

    python tools/profile_model.py --task scene_segmentation --la pospool_xyz > profiles/model_step_seg.txt

Prints the kernels of ONE step (after warm-up) grouped by name, this library's kernels (cl3d::) first, then the
torch / cuDNN / cuBLAS kernels of the reference-shaped module tree around them (1x1 convs, BatchNorm, ReLU, residual
adds, loss, optimizer).  A diagnostic for SURVEY.md section 8 row (f)2 (what block-level fusion would buy).
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="scene_segmentation", choices=["classification", "scene_segmentation"])
    ap.add_argument("--la", default="pospool_xyz")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--points", type=int, default=None)
    args = ap.parse_args()
    import train_synth as ts
    from closerlook3d_b200 import backbone as bb
    device = torch.device("cuda", 0)
    la, over = ts.LA[args.la]
    cfg = bb.model_config(args.task, la, **over)
    B = args.batch or (16 if args.task == "classification" else 8)
    N = args.points or cfg.num_points
    build = bb.build_classification if args.task == "classification" else bb.build_scene_segmentation
    torch.manual_seed(0)
    model, criterion = build(cfg)
    model.init_weights()
    model = model.to(device).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.002, momentum=0.98, weight_decay=0.001)
    xyz, mask, feats = ts.synth_batch(B, N, cfg.input_features_dim, 1, device, 1.0 if args.task == "classification" else 2.0)
    tgt = torch.randint(0, cfg.num_classes, (B,) if args.task == "classification" else (B, N), device=device)

    def step():
        pred = model(xyz, mask, feats)
        loss = criterion(pred, tgt) if args.task == "classification" else criterion(pred, tgt, mask.float())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    by = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            by[ev.name][0] += 1
            by[ev.name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    total = sum(v[1] for v in by.values())
    mine = sum(v[1] for k, v in by.items() if "cl3d" in k)
    print(f"# {args.task} / {args.la}: {B} x {N} points, one training step, {len(prof.events())} profiler events")
    print(f"# GPU busy time {total / 1000:.2f} ms in {sum(v[0] for v in by.values())} kernels / copies; "
          f"libcl3d kernels {mine / 1000:.2f} ms ({100 * mine / max(total, 1e-9):.0f} %)")
    for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{t / 1000:9.3f} ms  {n:5d} x  {k[:110]}")


if __name__ == "__main__":
    main()
