"""Print the roofline-relevant metrics of every launch in an .ncu-rep, optionally labelled.
usage: python tools/ncu_summary.py rep.ncu-rep [label-file]   (label file: one line per launch: name | algorithmic bytes | flops)"""
import csv, subprocess, sys
rep = sys.argv[1]
labels = [l.rstrip("\n").split("|") for l in open(sys.argv[2])] if len(sys.argv) > 2 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]; idx = {k: i for i, k in enumerate(h)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg"]
units = rows[1]
def val(r, k):
    if k not in idx: return None
    v = r[idx[k]].replace(",", "")
    try: return float(v)
    except ValueError: return v
def to_bytes(x, unit):
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
def to_us(x, unit):
    return x * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
for n, r in enumerate(rows[2:]):
    name = r[idx["Kernel Name"]].split("(")[0]
    us = to_us(val(r, "gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]])
    rd = to_bytes(val(r, "dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]]); wr = to_bytes(val(r, "dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]])
    line = (f"launch {n}: {name[:48]:48s} {us:9.1f} us  dram {rd / 1e6:8.1f} MB rd + {wr / 1e6:8.1f} MB wr = {(rd + wr) / 1e6:8.1f} MB "
            f"({(rd + wr) / us / 1e3:7.1f} GB/s)  tensor {val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')}%  "
            f"xu {val(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')}%  warps {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')}%  "
            f"regs {val(r, 'launch__registers_per_thread')}  L2 hit {val(r, 'lts__t_sector_hit_rate.pct')}%")
    if n < len(labels):
        lab, alg, fl = labels[n][0].strip(), float(labels[n][1]), float(labels[n][2])
        line = f"[{lab}] " + line + f"  | algorithmic {alg / 1e6:.1f} MB -> traffic ratio {(rd + wr) / alg:.2f}"
        if fl: line += f", {fl / us / 1e6:.0f} TFLOP/s"
    print(line)
