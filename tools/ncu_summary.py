"""Summarise ncu reports for profiles/: `python tools/ncu_summary.py <rep.ncu-rep> [...] > profiles/x.md`.
Reads each report with `ncu -i … --page raw --csv` and prints one block per profiled launch with the
counters DESIGN.md / bench.py quote (duration, DRAM bytes, DRAM %, tensor-pipe %, occupancy, registers)."""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__cycles_elapsed.max", "sm cycles elapsed"),
    ("smsp__cycles_active.avg", "smsp cycles active"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe % of peak (of active cycles)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.avg.per_cycle_active", "warp instructions per cycle per SM"),
]

def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"## {rep.split('/')[-1]}\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print(f"**{name}**\n")
            print("| counter | value |\n|---|---|")
            for key, label in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
            print()

if __name__ == "__main__":
    main()
