"""Turn gpurun_out ncu artefacts into the text summaries committed under profiles/.
usage: python tools/summarize_ncu.py <tag> <launches.csv> [<full.ncu-rep>]"""
import collections, csv, subprocess, sys, os

tag, launches = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 else None
out = []
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("ddpm3d::", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out.append(f"# ncu launch list ({tag}): gpu__time_duration.sum, --clock-control none, {len(data)} launches, "
           f"{tot / 1e3:.3f} ms total (cold-cache, serialised: compare shares)")
out.append(f"{'ms':>10} {'share':>7} {'n':>5}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{v[1] / 1e3:10.3f} {100 * v[1] / tot:6.1f}% {v[0]:5d}  {k}")
open(f"profiles/{tag}_launches_summary.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u, d = rr[0], rr[1], rr[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
            "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]
    o2 = [f"# ncu --set full --clock-control none ({tag}): selected metrics per captured launch"]
    for w in want:
        if w in h:
            i = h.index(w)
            o2.append(f"{w} [{u[i]}]: " + " | ".join(r[i][-70:] for r in d))
    open(f"profiles/{tag}_conv_tc_full.txt", "w").write("\n".join(o2) + "\n")
    print("\n".join(o2))
