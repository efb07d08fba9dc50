"""Summarise an .ncu-rep (captured with `ncu --set full`) into the text form kept under profiles/:
   python tools/ncu_summary.py <report.ncu-rep> [n_leapfrogs] > profiles/<name>.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
nlf = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"# kernel {d.get('Kernel Name')}  grid {d.get('Grid Size')}  block {d.get('Block Size')}")
    for k in WANT:
        if k in d:
            print(f"{k},{units[hdr.index(k)]},{d[k]}")
    st = sorted(((float(d[k]), k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k] not in ("", "n/a")), reverse=True)
    print("# stall reasons (warps stalled per issue-active cycle)")
    for v, k in st[:8]:
        print(f"{k},{v:.3f}")
    if nlf:
        dr = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"]))
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index("dram__bytes_read.sum")]]
        print(f"# per leapfrog ({nlf:.0f} in this launch): DRAM bytes {dr * mult / nlf:.0f}, warp-instructions {float(d['smsp__inst_executed.sum']) / nlf:.0f}, "
              f"SM cycles per CTA {float(d['sm__cycles_elapsed.max']) * float(d['Grid Size'].strip('()').split(',')[0]) / nlf:.0f}")
