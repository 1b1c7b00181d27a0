"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into the few numbers the roofline uses."""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "lts__t_bytes.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(rows) - 2} kernel launches (ncu --set full --clock-control none --import-source on)")
    for r in rows[2:]:
        print("kernel:", r[idx["Kernel Name"]][:90], " grid", r[idx.get("launch__grid_size", 0)])
        for w in WANT:
            if w in idx:
                print(f"    {w:68s} {r[idx[w]]:>16s} {units[idx[w]]}")
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
            print(f"    {'traffic = dram read + write':68s} {(rd + wr) / 1e6:16.1f} Mbyte")
            t = float(r[idx["gpu__time_duration.sum"]]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(units[idx["gpu__time_duration.sum"]], 1e-6)
            print(f"    {'dram traffic / duration':68s} {(rd + wr) / t / 1e9:16.1f} GB/s")
        except Exception:
            pass


if __name__ == "__main__":
    main(sys.argv[1])
