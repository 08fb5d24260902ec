"""micro-benchmark: ball query brute vs grid on the BASELINE shapes (GPU)"""
import sys, torch
sys.path.insert(0, '.')
from closerlook3d_b200 import ops, synth
from closerlook3d_b200.config import baseline_config
dev = torch.device('cuda:0')
for ci, B in [(2, 32), (1, 2), (4, 4), (3, 8)]:
    t = baseline_config(ci)
    d = synth.make_cloud_batch(B, t['N'], 4, 1000 + ci)
    xyz, mask = d['xyz'].to(dev), d['mask'].to(dev)
    r = synth.ball_radius(t['N'], t['K'])
    for algo in (1, 2):
        if algo == 1 and t['N'] > 8192: continue
        for _ in range(3): ops.ball_query(xyz, xyz, mask, mask, r, t['K'], want_mask=False, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.ball_query(xyz, xyz, mask, mask, r, t['K'], want_mask=False, algo=algo)
        e1.record(); torch.cuda.synchronize()
        print(f"c{ci} B={B} N={t['N']} K={t['K']} algo={'brute' if algo==1 else 'grid'}: {e0.elapsed_time(e1)/20*1000:.1f} us")
