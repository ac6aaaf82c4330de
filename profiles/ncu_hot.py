"""Prints headline metrics and the hottest SASS instructions (by stall samples) of one kernel in an .ncu-rep.
    python profiles/ncu_hot.py report.ncu-rep [min %% of samples] [regex:kernel]"""
import csv, subprocess, sys, io

KERNEL = sys.argv[3] if len(sys.argv) > 3 else None   # e.g. regex:stage2_thread


def page(rep, name):
    cmd = ["ncu", "-i", rep, "--page", name, "--csv"] + (["--kernel-name", KERNEL] if KERNEL else [])
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, v = raw[0], raw[1], raw[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
            "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_utcmma_cycles_active.avg.pct_of_peak_sustained_active" if False else "x"]
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:70s} {units[i]:12s} {v[i]}")
    st = [(float(v[i].replace(",", "")), h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
    tot = sum(x for x, _ in st)
    print("stall samples:", ", ".join(f"{h.split('stalled_')[1]} {100 * x / tot:.0f}%" for x, h in sorted(st, reverse=True)[:8]))
    src = [r for r in page(rep, "source")[2:] if len(r) > 5 and r[4].strip().isdigit()]
    tot = sum(int(r[4]) for r in src) or 1
    acc = 0
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    for i, r in enumerate(src):
        s = int(r[4]); acc += s
        if s >= tot * thr / 100:
            print(f"{i:5d} {r[1].strip()[:64]:64s} {100 * s / tot:5.1f}%  cum {100 * acc / tot:3.0f}%  exec {r[5]}")

if __name__ == "__main__":
    main()
