"""Turn an .ncu-rep into the text summaries kept under profiles/ (run in the authoring container, where ncu can read reports).

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x            # -> r01_x_details.txt, r01_x_raw.txt

`_details.txt` = every metric of the `--set full` sections per profiled launch; `_raw.txt` = the raw counters the bench and DESIGN.md
quote (DRAM / L2 bytes, thread-instruction efficiency, issue slots, stall reasons)."""
import csv
import io
import subprocess
import sys

RAW_KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
    "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
)


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def details(rep):
    out = []
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "details", "--csv"]))))
    head = rows[0]
    ix = {k: i for i, k in enumerate(head)}
    last = None
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"] or not r[ix["Metric Name"]]:
            continue
        key = (r[ix["ID"]], r[ix["Kernel Name"]], r[ix["Grid Size"]], r[ix["Block Size"]])
        if key != last:
            out.append("\n=== launch %s: %s grid %s block %s" % key)
            last = key
        out.append("%s | %s | %s | %s" % (r[ix["Section Name"]], r[ix["Metric Name"]], r[ix["Metric Unit"]], r[ix["Metric Value"]]))
    return "\n".join(out) + "\n"


def raw(rep):
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    head, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        out.append("\n=== launch %s: %s grid %s block %s" % (d.get("ID"), d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for i, k in enumerate(head):
            if k in RAW_KEEP or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") \
                    or k.startswith("smsp__pcsamp_warps_issue_stalled"):
                if r[i] not in ("", "0"):
                    out.append("%s | %s | %s" % (k, units[i], r[i]))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    rep, stem = sys.argv[1], sys.argv[2]
    open(stem + "_details.txt", "w").write(details(rep))
    open(stem + "_raw.txt", "w").write(raw(rep))
    print("wrote", stem + "_details.txt", stem + "_raw.txt")
