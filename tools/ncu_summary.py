"""Summarise ncu captures into small text files under profiles/ (the .ncu-rep binaries stay in gpurun_out/).

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_bench.txt
    python tools/ncu_summary.py full gpurun_out/prof_gemm.ncu-rep profiles/r01_ncu_gemm.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
        name = row["Kernel Name"]
        m = re.search(r"(\w+_kernel)", name)
        key = m.group(1) if m else re.sub(r"[<(].*", "", name)[:70]
        tot[key] += v
        cnt[key] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {src}; {sum(cnt.values())} launches, {total / 1e6:.3f} ms total\n")
        f.write(f"{'ms':>10} {'share':>7} {'count':>6} {'avg_us':>9}  kernel\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{v / 1e6:10.3f} {100 * v / total:6.1f}% {cnt[k]:6d} {v / cnt[k] / 1e3:9.1f}  {k}\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none --import-source on; source: {src}"]
    for row in rows[2:3]:
        d = dict(zip(hdr, zip(row, units)))
        out.append(f"kernel: {d['Kernel Name'][0]}")
        for m in METRICS:
            if m in d:
                out.append(f"  {m} = {d[m][0]} {d[m][1]}")
    src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src_csv)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "Address":
            if hdr is None:
                hdr = r
                continue
            break
        if hdr and len(r) == len(hdr):
            data.append(r)
    if hdr:
        i_src, i_samp = hdr.index("Source"), hdr.index("# Samples")
        tot = sum(int(r[i_samp] or 0) for r in data) or 1
        out.append(f"\ntop sampled instructions (of {tot} warp samples):")
        for r in sorted(data, key=lambda r: -int(r[i_samp] or 0))[:25]:
            out.append(f"  {100 * int(r[i_samp]) / tot:5.1f}%  {r[i_src][:110]}")
        scols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
        agg = sorted(((hdr[i], sum(float(r[i] or 0) for r in data)) for i in scols), key=lambda x: -x[1])[:10]
        out.append("\nstall reasons (samples): " + ", ".join(f"{k}={int(v)}" for k, v in agg))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:30]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
