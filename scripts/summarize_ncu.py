"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof_gemm_r1.ncu-rep profiles/r1_gemm_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    seq = []
    for row in r:
        if len(row) < len(hdr):
            continue
        v = float(row[ix["Metric Value"]].replace(",", ""))
        u = row[ix["Metric Unit"]]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        seq.append((re.sub(r"\(.*", "", row[ix["Kernel Name"]]).replace("void ", ""), v))
    starts = [i for i, (n, _) in enumerate(seq) if "timestep_features" in n]
    # the last COMPLETE evaluation in the capture (earlier ones include first-touch effects)
    a, b = (starts[-2], starts[-1]) if len(starts) > 1 else (starts[0], len(seq))
    agg = collections.OrderedDict()
    tot = 0.0
    for n, v in seq[a:b]:
        agg.setdefault(n, [0.0, 0])
        agg[n][0] += v
        agg[n][1] += 1
        tot += v
    out = io.StringIO()
    title = sys.argv[4] if len(sys.argv) > 4 else "DiT-L/2, batch 64 (M = 16384 token rows), `python bench.py --steps 1 --warmup 3`"
    out.write(f"# ncu launch list, one network evaluation (NFE): {title}\n\n")
    out.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n")
    out.write(f"{b - a} kernel launches per NFE; sum {tot / 1000:.3f} ms.\n\n")
    out.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.write(f"| `{n}` | {c} | {v:.1f} | {100 * v / tot:.1f}% |\n")
    out.write("\nFirst kernels in launch order (us):\n\n")
    for n, v in seq[a:a + 13]:
        out.write(f"- {v:8.1f}  {n}\n")
    open(dst, "w").write(out.getvalue())
    print(out.getvalue())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    out = io.StringIO()
    out.write(f"# ncu --set full summary of {src.split('/')[-1]} (selected metrics; the .ncu-rep itself is scratch)\n\n")
    for row in r[2:]:
        name = re.sub(r"\(.*", "", row[hdr.index("Kernel Name")])
        out.write(f"## {name}  grid {row[hdr.index('Grid Size')]} block {row[hdr.index('Block Size')]}\n\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.write(f"- {k} = {row[i]} {units[i]}\n")
        out.write("\n")
    open(dst, "w").write(out.getvalue())
    print(out.getvalue()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
