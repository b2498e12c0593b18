#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + stall breakdown + hottest SASS.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'smsp__inst_executed.sum', 'smsp__inst_executed_op_shared_atom.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('== launch:', r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr:
            print(f"  {w:66s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
    cyc = float(r[hdr.index('sm__cycles_elapsed.avg')].replace(',', ''))
    wf = float(r[hdr.index('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')].replace(',', ''))
    print(f"  shared-memory wavefronts per SM-cycle (peak 1.0): {wf / 148 / cyc:.3f}")
    break
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
        continue
    if cur is not None:
        cur.append(r)
b = blocks[0]
hdr = b[0]
data = [r for r in b[1:] if len(r) > 10]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in data:
    for s in stalls:
        agg[s] += int(r[ix[s]])
tot = sum(agg.values())
print("== warp stall sampling (all samples):", tot)
for k, v in agg.most_common(10):
    print(f"  {k:26s} {v:8d} {100.0 * v / tot:5.1f}%")
print("== hottest SASS instructions")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    st = {s: int(r[ix[s]]) for s in stalls if int(r[ix[s]])}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"  {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip()[:64]:64s} {st}")
