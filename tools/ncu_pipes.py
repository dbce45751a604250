"""profiles/pipes_r2.json from an ncu report: per C-ABI entry point of the fused path, what the hardware issued (FMA / XU
pipe, issue slots, L1 LSU wavefronts, L2 tag requests, DRAM: % of peak) and the DRAM bytes per view -- bench.py puts
them next to the algorithmic roofline fraction.   usage: ncu_pipes.py report.ncu-rep views out.json"""
import csv, io, json, subprocess, sys
rep, views, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
SYM = [("trace_hits_kernel", "voge_trace_hits"), ("select_topk_kernel", "voge_select_topk"),
       ("blend_weights_kernel", "voge_blend_weights"), ("blend_pair_kernel", "voge_blend_weights"),
       ("render_bwd_pair_kernel<128, 9, 0, 1>", "voge_render_backward_image"), ("render_bwd_pair_kernel", "voge_render_backward_fused"),
       ("render_bwd_fused_kernel", "voge_render_backward_fused"),
       ("merge_fwd", "voge_merge_final"), ("merge_bwd", "voge_merge_final_backward"),
       ("bin_count_kernel", "voge_bin_count"), ("bin_fill_kernel", "voge_bin_fill"),
       ("pack_gaussians_kernel", "voge_pack_gaussians"), ("unpack_gradients_kernel", "voge_unpack_gradients")]
M = {"fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "l1_wavefront_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
     "l2_tag_request_pct": "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "duration_us": "gpu__time_duration.sum"}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
acc = {}
for r in rows[2:]:
    for pat, sym in SYM:
        if pat in r[ik]:
            a = acc.setdefault(sym, {"n": 0, "dram": 0.0, "kernel": r[ik].split("(")[0], **{k: 0.0 for k in M}})
            a["n"] += 1
            a["dram"] += float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            for k, col in M.items():
                if col in hdr:
                    v = float(r[hdr.index(col)])
                    if k == "duration_us":
                        v *= tscale.get(units[hdr.index(col)], 1.0)
                    a[k] += v
            break
res = {}
for sym, a in acc.items():
    res[sym] = {k: round(a[k] / a["n"], 3) for k in M}
    res[sym]["dram_bytes_per_view"] = a["dram"] / a["n"] / views
    res[sym]["kernel"] = a["kernel"]
    res[sym]["launches_captured"] = a["n"]
res["_source"] = "%s (ncu --set full --clock-control none, %d views per launch; percentages of peak, bytes per view)" % (rep, views)
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(res, indent=1, sort_keys=True))
