"""`ncu -i X.ncu-rep --page raw --csv` -> a small JSON with the metrics the docs cite.
usage: python tools/ncu_raw_summary.py X.ncu-rep out.json "<how it was captured>" ["<what the kernel is>"]"""
import csv, io, json, subprocess, sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum"]


def main():
    rep, out, how = sys.argv[1], sys.argv[2], sys.argv[3]
    what = sys.argv[4] if len(sys.argv) > 4 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    rec = {"source": how, "kernel": (vals[col["Kernel Name"]] if "Kernel Name" in col else "") + (" — " + what if what else ""), "metrics": {}}
    for m in WANT:
        if m in col:
            rec["metrics"][m] = {"value": vals[col[m]], "unit": units[col[m]]}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec["metrics"], indent=1)[:1200])


if __name__ == "__main__":
    main()
