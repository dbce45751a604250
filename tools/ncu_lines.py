"""Per CUDA source line: stall samples and executed warp instructions of a kernel in an ncu report (needs -lineinfo and
--import-source on).  usage: ncu_lines.py report.ncu-rep [kernel-regex] [min-sample-pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; rx = sys.argv[2] if len(sys.argv) > 2 else None; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["-k", "regex:" + rx] if rx else [])
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None; out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; hdr = None
    elif r[0] == "Line No":
        hdr = r
    elif hdr is not None and len(r) == len(hdr) and r[2] == "-":        # a CUDA source line row (Address == "-")
        i_s, i_e, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        try:
            out.append((cur_file, int(r[0]), r[1].strip()[:100], int(r[i_s]), int(r[i_e]), int(r[i_t])))
        except ValueError:
            pass
ts = sum(o[3] for o in out) or 1; te = sum(o[4] for o in out) or 1
print("total samples %d, warp instructions %d" % (ts, te))
for f, ln, src, s, e, t in out:
    if 100.0 * s / ts >= thr or 100.0 * e / te >= 2 * thr:
        print("%-22s %4d  samples %5.1f%%  exec %5.1f%%  thr/inst %4.1f  | %s" % (f, ln, 100.0 * s / ts, 100.0 * e / te, t / max(e, 1), src))
