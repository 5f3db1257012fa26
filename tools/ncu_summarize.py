"""Condense an `ncu --page raw --csv` export of one step's trunk launches into the two files bench.py / DESIGN.md cite:

    ncu -i gpurun_out/prof_trunk.ncu-rep --page raw --csv > gpurun_out/prof_trunk_raw.csv
    python tools/ncu_summarize.py gpurun_out/prof_trunk_raw.csv profiles/r01/ncu_full_trunk_v8.csv \
        profiles/r01/trunk_traffic_v8.json

  *.csv   one row per launch, selected columns (time, DRAM bytes, tensor-pipe activity, L2 hit rate, smem wavefronts ...)
  *.json  DRAM traffic summed over the launches (= roofline.traffic of bench.py) + time-weighted tensor-pipe activity
"""
import csv
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "implementation-phd-lab-vision_b200"))
from phdfx.synthetic import csrc_sha  # noqa: E402  (content hash of the kernel sources the capture was taken with)

COLS = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    raw, out_csv, out_json = sys.argv[1:4]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in idx]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for d in data:
            w.writerow([d[idx[c]] for c in cols])

    def val(d, name):
        return float(d[idx[name]].replace(",", "")) * SCALE.get(units[idx[name]], 1.0)

    rd = sum(val(d, "dram__bytes_read.sum") for d in data)
    wr = sum(val(d, "dram__bytes_write.sum") for d in data)
    us = [val(d, "gpu__time_duration.sum") for d in data]
    tc = [float(d[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) for d in data]
    per_kernel = {}
    for d, t in zip(data, us):
        name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        per_kernel[name] = per_kernel.get(name, 0.0) + t
    json.dump({"source": f"ncu --set full, one step at batch 256 ({len(data)} trunk launches), {out_csv}",
               "csrc_sha": csrc_sha(),
               "launches": len(data), "dram_bytes_read": rd, "dram_bytes_write": wr,
               "traffic_bytes_per_step": rd + wr, "sum_kernel_us": sum(us),
               "time_weighted_tensor_pipe_pct": sum(a * b for a, b in zip(us, tc)) / sum(us),
               "kernel_share_us": {k: round(v, 1) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])}},
              open(out_json, "w"), indent=1)
    print(open(out_json).read())


if __name__ == "__main__":
    main()
