"""Summaries committed under profiles/ from the ncu artefacts in gpurun_out/ (run in the build container)."""
import collections, csv, json, re, subprocess, sys

def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, seq, tot = collections.OrderedDict(), [], 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("rd::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "").replace("void ", "")
        t = float(row["Metric Value"].replace(",", "")) * (1000 if row["Metric Unit"] == "us" else 1)
        seq.append((name, t, row["Grid Size"], row["Block Size"])); tot += t
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    with open(out, "w") as f:
        f.write("# one eager training step (P19 B=128, TrainStep._enqueue), ncu --metrics gpu__time_duration.sum --clock-control none\n")
        f.write("# cold-cache, serialised per-launch times: compare SHARES, not absolutes\n")
        f.write("launches %d   sum %.1f us\n\n%-44s %5s %10s %7s\n" % (len(seq), tot / 1000, "kernel", "n", "us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-44s %5d %10.1f %6.1f%%\n" % (k[:44], v[0], v[1] / 1000, 100 * v[1] / tot))
        f.write("\n# launch order\n")
        for i, (n, t, g, b) in enumerate(seq):
            f.write("%3d %-44s %8.1f us  grid %-16s block %s\n" % (i, n[:44], t / 1000, g, b))

def full(rep, out, json_out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
            "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_tmem.sum",
            "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed"]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, %s\n" % rep)
        for r in rows[2:]:
            f.write("\nkernel: %s  (launch id %s)\n" % (r[hdr.index("Kernel Name")][:70], r[hdr.index("ID")]))
            for k in hdr:
                if k in keys:
                    f.write("  %-78s %s %s\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    if json_out:
        def val(r, k):
            v = float(r[hdr.index(k)].replace(",", "")); u = units[hdr.index(k)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        per = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in rows[2:]]
        json.dump({"dram_bytes_per_launch": sum(per) / len(per), "launches": len(per), "source": out,
                   "rows": 557056, "C": 240}, open(json_out, "w"), indent=1)

if __name__ == "__main__":
    # usage: ncu_summary.py launches <csv> <out.txt> | full <rep> <out.txt> [traffic.json]
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
