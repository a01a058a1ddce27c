"""Condense `ncu --set full` reports (.ncu-rep) into a small JSON with the metrics the design argues from.
Usage: python tools/ncu_summary.py out.json rep1.ncu-rep [rep2.ncu-rep ...]   (needs the ncu CLI; CPU only)."""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}

out = []
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"report": rep.split("/")[-1], "kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k in KEYS:
            if k in hdr:
                v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
                d[k] = v * SCALE[u] if u in SCALE else v
                if u in SCALE:
                    d[k + ".unit"] = "byte" if "byte" in u else "s"
        if "dram__bytes_read.sum" in d:
            d["dram_bytes_total"] = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        out.append(d)
json.dump(out, open(sys.argv[1], "w"), indent=1)
for d in out:
    print(d["report"], d["kernel"], "%.3f ms" % (d["gpu__time_duration.sum"] * 1e3), "dram %.3f GB" % (d.get("dram_bytes_total", 0) / 1e9),
          "issue %.1f%%" % d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
          "tensor %.1f%%" % d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0))
