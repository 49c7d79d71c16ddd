"""Summarise ncu --set full reports (one line block per captured launch): python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]
Reads the raw page through `ncu -i ... --page raw --csv` (works without a GPU) and prints the metrics the roofline section of
bench.py / DESIGN.md quotes: duration, DRAM bytes (traffic), DRAM and tensor-pipe utilisation, occupancy, registers."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (of elapsed)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (MUFU) pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * mult.get(unit, 1)


def main():
    for path in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            print(f"== {path}: no launches")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"== {path}")
        for r in rows[2:]:
            print(f"-- {r[col['Kernel Name']]}")
            rd = wr = None
            for key, label in WANT:
                if key in col and r[col[key]] != "":
                    print(f"   {label:42s} {r[col[key]]} {units[col[key]]}")
                    if key == "dram__bytes_read.sum":
                        rd = to_bytes(r[col[key]], units[col[key]])
                    if key == "dram__bytes_write.sum":
                        wr = to_bytes(r[col[key]], units[col[key]])
            if rd is not None and wr is not None:
                dur_i = col.get("gpu__time_duration.sum")
                dur = float(r[dur_i].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(units[dur_i], 1e-6)
                print(f"   {'traffic (dram read + write)':42s} {(rd + wr) / 1e6:.3f} Mbyte -> {(rd + wr) / dur / 1e9:.1f} GB/s under ncu")


if __name__ == "__main__":
    main()
