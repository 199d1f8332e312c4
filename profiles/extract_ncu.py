"""Extract the key metrics of an `ncu --set full` report (read with `ncu -i rep --page raw --csv`) into markdown."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
SHORT = ["time", "dram rd", "dram wr", "dram %", "sm %", "tensor %", "xu %", "issue %", "warps %", "regs", "dyn smem", "st long_sb", "st lg_thr", "st math_thr", "st barrier"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {rep}\n\n| kernel | grid | " + " | ".join(SHORT) + " |\n|---|---|" + "---:|" * len(SHORT) + "\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].replace("<unnamed>::", "").split("(")[0].replace("void ", "")[:48]
            vals = []
            for w in WANT:
                i = idx.get(w)
                vals.append("-" if i is None or not r[i] else f"{float(r[i].replace(',', '')):.4g} {units[i]}".replace(" register/thread", "").replace(" inst", ""))
            f.write(f"| `{name}` | {r[idx['Grid Size']]} | " + " | ".join(vals) + " |\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
