#!/usr/bin/env python
"""ncu metrics driver: turns an .ncu-rep (ncu --set full) into the markdown summary kept under
profiles/.  Replaces the reference's nvprof scripts (/root/reference/src/cuda/metrics/
{flop,stall,util}_metrics.sh: flop count/efficiency, stall reasons, unit utilisation) and the
PAPI summary (/root/reference/src/cpu/papi.c:459-489) with the counters the north star asks for:
achieved HBM / L2 GB/s for the gathers, MUFU / FMA pipe utilisation, atomic (RED) throughput.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--title "..."] [--intersections N] > profiles/x.md

Capture on the GPU box (one GPU, never a multi-rank command):
    ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 \
        -o gpurun_out/prof python bench.py --steps 1 --warmup 1 --segments 20000000 --no-cpu-baseline
"""
import argparse
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe cycles active"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe instructions"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe instructions"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe instructions"),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "TMA pipe instructions"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2 -> SM gather bandwidth (source rows)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum.per_second", "SM -> L2 bandwidth (tally REDs)"),
    ("lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "L2 tag sectors, % of peak"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 RED sectors (atomic tallies)"),
    ("lts__t_sectors_srcunit_tex_op_red.sum.per_second", "L2 RED sector rate"),
    ("lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed", "L2 RED throughput, % of peak"),
    ("l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_red.sum.per_second", "TMA bulk-reduce bandwidth"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second", "TMA bulk-load bandwidth"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__bytes_read.sum.per_second", "DRAM read bandwidth"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--title", default=None)
    ap.add_argument("--intersections", type=float, default=None, help="segment x group intersections in the launch")
    ap.add_argument("--top", type=int, default=12)
    a = ap.parse_args()

    rows = ncu_csv(a.report, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    name = d.get("Kernel Name", ("?", ""))[0]
    print(f"# {a.title or 'ncu --set full: ' + name}\n")
    print(f"kernel `{name}`, grid {d.get('Grid Size', ('?',))[0]}, block {d.get('Block Size', ('?',))[0]}; "
          f"report `{a.report}` (not committed)\n")
    print("| metric | value | unit |\n|---|---|---|")
    for key, label in METRICS:
        if key in d and d[key][0] != "":
            print(f"| {label} (`{key}`) | {d[key][0]} | {d[key][1]} |")
    if a.intersections:
        t = float(d["gpu__time_duration.sum"][0])
        unit = d["gpu__time_duration.sum"][1]
        sec = t * {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}.get(unit, 1e-3)
        inst = float(d["smsp__inst_executed.sum"][0])
        print(f"\n{a.intersections:.3e} intersections in this launch: {a.intersections / sec:.3e} intersections/s under the "
              f"profiler, {inst * 32 / a.intersections:.1f} thread-instructions per intersection, "
              f"algorithmic bandwidth {18.4 * a.intersections / sec / 1e9:.0f} GB/s (18.4 B/intersection)")

    print("\n## warp stall reasons (cycles per issued instruction)\n\n| reason | value |\n|---|---|")
    stalls = [(h, float(d[h][0] or 0)) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for h, v in sorted(stalls, key=lambda x: -x[1]):
        if v >= 0.05:
            print(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.2f} |")

    try:
        src = ncu_csv(a.report, "source")
        h2 = src[1]
        ci = {h: i for i, h in enumerate(h2)}
        reasons = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in src[2:] if len(r) >= len(h2)]
        total = sum(int(r[ci["# Samples"]] or 0) for r in data) or 1
        print(f"\n## hottest SASS instructions ({total} stall samples)\n\n| SASS | samples | executed | top stall reasons |\n|---|---|---|---|")
        for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[: a.top]:
            rs = sorted(((k[6:], int(r[ci[k]] or 0)) for k in reasons), key=lambda x: -x[1])[:2]
            print(f"| `{r[ci['Source']].strip()[:72]}` | {r[ci['# Samples']]} ({100 * int(r[ci['# Samples']]) / total:.1f}%) | "
                  f"{r[ci['Instructions Executed']]} | {', '.join(f'{k} {v}' for k, v in rs)} |")
    except Exception as e:  # source page needs -lineinfo / --import-source
        print(f"\n(source page unavailable: {e})", file=sys.stderr)


if __name__ == "__main__":
    main()
