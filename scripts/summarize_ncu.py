#!/usr/bin/env python
"""Turns the ncu reports brought back in gpurun_out/ into the tracked summaries under profiles/ (run on the CPU box)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]


def ncu_csv(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def summarize(rep, name, frames):
    rows = list(csv.reader(io.StringIO(ncu_csv(rep, "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full --clock-control none: {m.get('Kernel Name', ('?',))[0]}  ({frames} frames in this launch)"]
    for k in KEYS:
        if k in m:
            lines.append(f"{k:72s} {m[k][0]:>18s} {m[k][1]}")
    src = list(csv.reader(io.StringIO(ncu_csv(rep, "source"))))
    h = src[1]; data = src[2:]; ix = {c: i for i, c in enumerate(h)}
    stalls = collections.Counter(); ops = collections.Counter()
    for r in data:
        for c in h:
            if c.startswith("stall_") and "Not Issued" not in c:
                try: stalls[c] += int(r[ix[c]])
                except Exception: pass
        parts = r[ix["Source"]].split()
        if parts:
            op = (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]
            try: ops[op] += int(r[ix["Instructions Executed"]])
            except Exception: pass
    tot = sum(ops.values()) or 1
    lines.append("\nwarp-level stall samples: " + ", ".join(f"{k[6:]}={v}" for k, v in stalls.most_common(8)))
    lines.append(f"executed warp instructions: {tot}  ({tot / frames:.0f} per frame)")
    lines.append("instruction mix: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(16)))
    sass_flags = [k for k in ops if k in ("UBLKCP", "SYNCS", "VIMNMX", "VIMNMX3", "VIADD", "UTMALDG")]
    lines.append("Blackwell/Hopper-class SASS present: " + ", ".join(f"{k} x{ops[k]}" for k in sass_flags))
    open(os.path.join(OUT, name + "_summary.txt"), "w").write("\n".join(lines) + "\n")
    return m


def main():
    """usage: summarize_ncu.py <frames in the captured launch> [round prefix, default r02]"""
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(ROOT, "gpurun_out")
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    R = sys.argv[2] if len(sys.argv) > 2 else "r02"
    reports = [("prof_ofdm", "ofdm_demod_kernel"), ("prof_ofdm_fastnco", "ofdm_demod_kernel_oscillator_fast"), ("prof_ofdm_exactnco", "ofdm_demod_kernel_oscillator_exact"),
               ("prof_viterbi", "viterbi_kernel_msc"), ("prof_viterbi_fic", "viterbi_kernel_fic")]
    for rep, name in reports:
        path = os.path.join(g, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        m = summarize(path, f"{R}_{name}", frames)
        if rep == "prof_ofdm":
            def val(k):
                v, u = m[k]; f = float(v)
                return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            per = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / frames
            json.dump({"dram_bytes_per_frame": per, "frames_in_capture": frames, "source": f"profiles/{R}_ofdm_demod_kernel_summary.txt"}, open(os.path.join(OUT, f"{R}_ofdm_traffic.json"), "w"))
        print(open(os.path.join(OUT, f"{R}_{name}_summary.txt")).read())
    lc = os.path.join(g, "launches.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(open(lc)) if len(r) > 5 and r[0].isdigit()]
        agg = collections.OrderedDict()
        for r in rows:
            name = r[4].split("(")[0].replace("dabb::<unnamed>::", "").replace("void ", "")
            agg.setdefault(name, []).append(float(r[-1]))
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(OUT, f"{R}_launch_list.txt"), "w") as f:
            f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows)} launches of `python bench.py --batch {frames} ...` (cold-cache, serialised: compare shares)\n")
            f.write(f"{'kernel':60s} {'launches':>8s} {'avg_us':>10s} {'share':>7s}\n")
            for k, v in agg.items():
                f.write(f"{k[:60]:60s} {len(v):8d} {sum(v) / len(v) / 1e3:10.1f} {100 * sum(v) / tot:6.1f}%\n")
        print(open(os.path.join(OUT, f"{R}_launch_list.txt")).read())


if __name__ == "__main__":
    main()
