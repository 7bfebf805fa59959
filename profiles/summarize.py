#!/usr/bin/env python
"""Turn the scratch captures of profiles/capture.sh (gpurun_out/launches.csv, gpurun_out/prof_*.ncu-rep) into the tracked
markdown summaries profiles/rNN_launches.md and profiles/rNN_ncu_summary.md.   usage: python profiles/summarize.py r01"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = "gpurun_out"


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"lrp::", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    m = re.match(r"([A-Za-z0-9_]+(<[^(]*>)?)", name)
    return m.group(1) if m else name


def launches():
    rows = [l for l in open(f"{OUT}/launches.csv") if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    total = sum(tot.values())
    n = sum(cnt.values())
    lines = [f"# {tag} — launch list of one bench micro-batch (8 prompts x 2048 tokens, Llama-3-8B dims, 32 layers)\n",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` on `bench.py --per-gpu-batch 8 --micro-batch 8 --steps 1 --warmup 3`",
             "(`profiles/capture.sh`; cold-cache, serialised launches: compare SHARES, not absolutes)\n",
             f"total device time of the {n} launches: {total:.1f} ms\n", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:14]:
        lines.append(f"| `{k}` | {cnt[k]} | {v:.2f} | {100 * v / total:.1f} % |")
    gemm = sum(v for k, v in tot.items() if k.startswith("gemm_bf16"))
    lines.append(f"\nGEMM share under ncu: {100 * gemm / total:.1f} %  (bench.py's live CUDA-event share of the step: `roofline.share_of_step`).")
    open(f"profiles/{tag}_launches.md", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


METRICS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
           ("lts__t_bytes.sum", "L2 bytes"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
           ("sm__inst_executed_pipe_tensor_op_gen5.avg.pct_of_peak_sustained_active", "tcgen05 pipe %"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
           ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__cluster_dim_x", "cluster")]


def full(name):
    try:
        txt = subprocess.run(["ncu", "-i", f"{OUT}/prof_{name}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    except Exception as ex:  # noqa
        return [f"(no capture for {name}: {ex})"]
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    have = [(m, t) for m, t in METRICS if m in col]
    lines = ["| kernel | " + " | ".join(t for _, t in have) + " |", "|---|" + "---|" * len(have)]
    for r in data:
        lines.append(f"| `{short(r[col['Kernel Name']])}` | " + " | ".join(f"{r[col[m]]} {units[col[m]]}".strip() for m, _ in have) + " |")
    return lines


if __name__ == "__main__":
    launches()
    body = [f"# {tag} — ncu `--set full --clock-control none` captures (one B200, `bench.py --per-gpu-batch 8 --micro-batch 8 --steps 1 --warmup 3`)\n",
            "Command: `profiles/capture.sh` through gpurun; raw `.ncu-rep` files are scratch (`gpurun_out/`), this table is `profiles/summarize.py`.",
            "Kernel times under ncu are cold-cache and serialised (compare shares, not absolutes).\n"]
    for name in ("gemm", "attn_fwd", "attn_bwd"):
        body += [f"\n## {name}\n"] + full(name)
    try:   # hand-written interpretation, kept next to the tables
        body += ["", open(f"profiles/{tag}_ncu_reading.md").read()]
    except OSError:
        pass
    open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(body) + "\n")
    print("\n".join(body))
