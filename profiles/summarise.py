#!/usr/bin/env python
"""Turn ncu artefacts from gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarise.py launches gpurun_out/launches_r01.csv  profiles/r01_launches.md
    python profiles/summarise.py kernel   gpurun_out/prof_x.ncu-rep    profiles/r01_x.md [kernel-regex]
    python profiles/summarise.py traffic  gpurun_out/prof_x.ncu-rep    <key> <kernel-substring>   # -> profiles/traffic.json

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (count, total, share);
`kernel` extracts the roofline-relevant raw metrics of every captured launch plus the hottest SASS
instructions of the first one (needs `ncu` on PATH; runs on the CPU box).
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_fp32.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mn, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    order = []
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("dcd::<unnamed>::", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
        order.append(name)
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("source: `%s`, %d launches captured\n\n" % (src, len(order)))
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %.1f |\n" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / v[0] / 1e3))
        f.write("\nfirst 40 launches in order: " + ", ".join(order[:40]) + "\n")


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def kernel(rep, dst, regex=None):
    rows = ncu_csv(rep, "raw")
    hdr = rows[0]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary of `%s`\n\n" % rep)
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            name = r[hdr.index("Kernel Name")]
            f.write("## %s\n\n| metric | value |\n|---|---:|\n" % name.replace("|", "/")[:160])
            for k in RAW_KEYS:
                if k in hdr:
                    f.write("| %s | %s |\n" % (k, r[hdr.index(k)]))
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            f.write("\nwarp stall reasons (avg warps stalled per issue-active cycle, top 6): " +
                    ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]) + "\n\n")
        src = ncu_csv(rep, "source", ("--kernel-name", "regex:" + regex) if regex else ())
        if len(src) > 2:
            h = src[1]
            try:
                isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
            except ValueError:
                return
            data = []
            for r in src[2:]:
                if r and r[0] == "Kernel Name":
                    break
                if len(r) < len(h):
                    continue
                try:
                    data.append((r[isrc].strip(), int(r[isamp]), int(r[iex])))
                except ValueError:
                    pass
            tot = sum(d[1] for d in data) or 1
            ops = collections.Counter()
            for d in data:
                ops[d[0].split()[1] if d[0].startswith("@") and len(d[0].split()) > 1 else d[0].split()[0]] += d[2]
            f.write("## SASS of the first captured launch: %d instructions, %d stall samples\n\n" % (len(data), tot))
            f.write("executed warp-instructions by opcode (top 16): " +
                    ", ".join("%s %d" % kv for kv in ops.most_common(16)) + "\n\n")
            f.write("| # | stall samples | share | executed | SASS |\n|---:|---:|---:|---:|---|\n")
            for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][1])[:20]):
                f.write("| %d | %d | %.1f%% | %d | `%s` |\n" % (i, data[i][1], 100 * data[i][1] / tot, data[i][2], data[i][0][:90]))


def traffic(rep, key, needle):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first captured launch whose name contains `needle`, stored under
    `key` in profiles/traffic.json (what bench.py reports as roofline.traffic)."""
    import json
    import os
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot = None
    for r in rows[2:]:
        if len(r) < len(hdr) or needle not in r[hdr.index("Kernel Name")]:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * scale[units[i]]
        break
    if tot is None:
        raise SystemExit("no launch of %r in %s" % (needle, rep))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = tot
    data.setdefault("_source", {})[key] = os.path.basename(rep)
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(key, tot)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
