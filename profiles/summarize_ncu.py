"""Turns an `ncu --page raw --csv` dump into the per-kernel table committed under profiles/ and a small JSON
(dram traffic per launch etc.) that bench.py attaches to its roofline object.
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python profiles/summarize_ncu.py raw.csv r1"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
tag = sys.argv[2] if len(sys.argv) > 2 else "r1"
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_insts"),
        ("lts__t_bytes.sum", "l2_bytes"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("sm__cycles_elapsed.max", "cycles")]


def num(r, k):
    if k not in idx:
        return None
    v, u = r[idx[k]].replace(",", ""), units[idx[k]]
    try:
        x = float(v)
    except ValueError:
        return None
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
    return x * scale


title = sys.argv[3] if len(sys.argv) > 3 else ""
out, md = {}, [f"# ncu --set full --metrics lts__t_bytes.sum summary ({tag}) {title}\n", "| kernel | time ms | DRAM rd MB | DRAM wr MB | DRAM % | tensor % | issue % | warps % | regs | warp insts | L2 bytes MB | grid x block |",
               "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[idx["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("adn::", "")
    if "mlp_umma_kernel" in name:
        short = "mlp_umma_kernel" + name[name.index("<"):name.index(">") + 1].replace("(int)", "")
    v = {k2: num(r, k1) for k1, k2 in want}
    out.setdefault(short, v)
    f = lambda x, s=1.0, p=2: "-" if x is None else f"{x * s:.{p}f}"
    md.append(f"| `{short}` | {f(v['time'], 1e3, 3)} | {f(v['dram_rd'], 1e-6, 1)} | {f(v['dram_wr'], 1e-6, 1)} | {f(v['dram_%'])} | {f(v['tensor_%'])} | "
              f"{f(v['issue_%'])} | {f(v['warps_%'])} | {f(v['regs'], 1, 0)} | {f(v['warp_insts'], 1, 0)} | {f(v['l2_bytes'], 1e-6, 1)} | {f(v['grid'], 1, 0)} x {f(v['block'], 1, 0)} |")
open(f"profiles/ncu_{tag}_summary.md", "w").write("\n".join(md) + "\n")
json.dump(out, open(f"profiles/ncu_{tag}_summary.json", "w"), indent=1)
print("\n".join(md))
