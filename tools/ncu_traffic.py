"""Turn an `ncu --set full` report of ONE bench step into a profiles/traffic.json entry:
   python tools/ncu_traffic.py gpurun_out/prof.ncu-rep c2 [steps] [clouds_per_gpu]
   -> prints {cfg: {entry_point: {dram_bytes_per_step, kernels, time_us_under_ncu}}} (merge into traffic.json)"""
import csv, json, os, subprocess, sys, collections
rep, cfg = sys.argv[1], sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
clouds = int(sys.argv[4]) if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ki, ri, wi, ti = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
MAP = [("sgemm", "cl3d_sgemm_algo"), ("gemm_tf32x3", "cl3d_sgemm_algo"), ("splitk_reduce", "cl3d_sgemm_algo"),
       # round 2: the search and the transposed lists are one entry point (cl3d_ball_query_csr)
       ("ball_query", "cl3d_ball_query_csr"), ("grid_params", "cl3d_ball_query_csr"), ("grid_build_fused", "cl3d_ball_query_csr"),
       ("cell_", "cl3d_ball_query_csr"), ("zero_cells", "cl3d_ball_query_csr"), ("csr_", "cl3d_ball_query_csr"),
       ("pg2_kernel<0>", "cl3d_agg_fwd"), ("pg2_kernel<(bool)0>", "cl3d_agg_fwd"), ("pg2_kernel<1>", "cl3d_agg_bwd"),
       ("pg2_kernel<(bool)1>", "cl3d_agg_bwd"),
       ("pwmlp_fwd_kernel", "cl3d_pwmlp_fwd_stats"), ("pwmlp_out", "cl3d_pwmlp_fwd_out"), ("pwmlp_bwd", "cl3d_pwmlp_bwd"),
       ("agg_fwd", "cl3d_agg_fwd"), ("sincos_fwd", "cl3d_agg_fwd"), ("agg_bwd", "cl3d_agg_bwd"), ("sincos_bwd", "cl3d_agg_bwd"),
       ("bn_relu_fwd", "cl3d_bn_relu_fwd"), ("bn_relu_bwd", "cl3d_bn_relu_bwd"), ("bn_reduce2", "cl3d_bn_relu_bwd"),
       ("bn_finalize", "cl3d_bn_finalize"), ("to_point_major_aug", "cl3d_to_point_major_aug"),
       ("to_point_major", "cl3d_to_point_major"), ("pwmlp_prep", "cl3d_pwmlp_prep_weights"),
       ("pwmlp_wgrad", "cl3d_pwmlp_weight_grad"), ("reduce_partials", "cl3d_reduce_partials"), ("to_channel_major", "cl3d_to_channel_major")]
acc = collections.defaultdict(lambda: [0.0, 0, 0.0])
for r in rows[2:]:
    name = r[ki]
    ent = next((e for k, e in MAP if k in name), None)
    if ent is None: continue
    b = float(r[ri]) * scale.get(units[ri], 1) + float(r[wi]) * scale.get(units[wi], 1)
    acc[ent][0] += b; acc[ent][1] += 1; acc[ent][2] += float(r[ti])
d = {e: {"dram_bytes_per_step": v[0] / steps, "kernels": v[1] / steps, "time_us_under_ncu": v[2] / steps}
     for e, v in acc.items()}
d["_report"] = os.path.basename(rep)
d["_total_dram_bytes_per_step"] = sum(v[0] for v in acc.values()) / steps
if clouds is not None:
    d["_clouds_per_gpu"] = clouds
print(json.dumps({cfg: d}, indent=1))
