#!/usr/bin/env python
"""Key numbers of an `ncu --set full` report (.ncu-rep): duration, throughputs, DRAM bytes, occupancy, top stall reasons,
and (with --source) the hottest source lines by warp-stall samples.  Usage: tools/ncu_summary.py report.ncu-rep [--source N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:100], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    keys = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_bytes.sum"]
    for k in keys:
        if k in d:
            print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    stalls = [(float(v.replace(",", "")), k) for k, v in d.items() if "smsp__average_warp_latency_issue_stalled" in k and k.endswith(".ratio") and v not in ("", "n/a")]
    if not stalls:
        stalls = [(float(v.replace(",", "")), k) for k, v in d.items() if "smsp__average_warps_issue_stalled" in k and "_per_issue_active" in k and v not in ("", "n/a")]
    for v, k in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {k.split('stalled_')[-1][:50]:50s} {v:10.2f}")
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, h, out = "?", None, []
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            h = r; continue
        if h is None or not r[0].isdigit():
            continue
        d = dict(zip(h[4:], r[4:]))
        try:
            smp = float(d.get("# Samples", "0").replace(",", "") or 0)
            ins = d.get("Instructions Executed", "")
        except ValueError:
            continue
        st = sorted(((float(v.replace(",", "") or 0), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "-")), reverse=True)[:2]
        out.append((smp, cur, int(r[0]), r[1].strip()[:110], ins, " ".join(f"{k[6:]}={int(v)}" for v, k in st if v > 0)))
    tot = sum(o[0] for o in out) or 1.0
    for smp, f, ln, text, ins, st in sorted(out, reverse=True)[:n]:
        print(f"  {100 * smp / tot:5.1f}%  {f}:{ln:<4d} inst={ins:>9s} [{st}]  {text}")
