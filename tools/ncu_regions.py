"""Summarise an ncu report's source page per kernel: share of executed instructions / stall samples and
average active threads per block of SASS instructions.  usage: ncu_regions.py report.ncu-rep [kernel-regex] [block]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; rx = sys.argv[2] if len(sys.argv) > 2 else None; B = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", "regex:" + rx] if rx else [])
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
for blk in blocks:
    hdr = blk['rows'][0]; data = [r for r in blk['rows'][1:] if len(r) == len(hdr)]
    isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); ith = hdr.index('Thread Instructions Executed'); isamp = hdr.index('# Samples')
    tot = sum(int(r[iex]) for r in data); tots = sum(int(r[isamp]) for r in data)
    print(blk['name'][:70], '| warp-instr', tot, '| samples', tots, '| SASS instrs', len(data))
    for b in range(0, len(data), B):
        ch = data[b:b + B]
        ex = sum(int(c[iex]) for c in ch); th = sum(int(c[ith]) for c in ch); s = sum(int(c[isamp]) for c in ch)
        if ex / max(tot, 1) > 0.012 or s / max(tots, 1) > 0.012:
            ops = {}
            for c in ch:
                t = c[isrc].split(); op = t[0] if not t[0].startswith('@') else t[1]
                ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda x: -x[1])[:7]
            print('  instr %4d-%4d: exec %5.1f%%  samples %5.1f%%  avg thr %4.1f  %s' % (b, b + B, 100 * ex / tot, 100 * s / tots, th / max(ex, 1), top))
