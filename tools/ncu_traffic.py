"""profiles/traffic_r1.json from an ncu report: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum)
per VIEW for every kernel of the fused path, keyed by the C-ABI entry point that launches it (bench.py reads
the file for roofline.traffic).   usage: ncu_traffic.py report.ncu-rep views out.json"""
import csv, io, json, re, subprocess, sys
rep, views, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
SYM = [("trace_hits_kernel", "voge_trace_hits"), ("select_topk_kernel", "voge_select_topk"),
       ("blend_weights_kernel", "voge_blend_weights"), ("blend_pair_kernel", "voge_blend_weights"),
       ("render_bwd_fused_kernel", "voge_render_backward_fused"), ("render_bwd_pair_kernel", "voge_render_backward_fused"),
       ("merge_fwd", "voge_merge_final"), ("merge_bwd", "voge_merge_final_backward"),
       ("bin_count_kernel", "voge_bin_count"), ("bin_fill_kernel", "voge_bin_fill"),
       ("render_fwd_kernel", "voge_render_forward")]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res, n = {}, {}
for r in rows[2:]:
    for pat, sym in SYM:
        if pat in r[ik]:
            b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            res[sym] = res.get(sym, 0.0) + b
            n[sym] = n.get(sym, 0) + 1
            break
res = {k: v / n[k] / views for k, v in res.items()}
res["_source"] = "%s (ncu --set full --clock-control none, %d views per launch; bytes per view)" % (rep, views)
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(res, indent=1, sort_keys=True))
