"""Turn the raw ncu outputs of tools/run_gpu_round.sh (gpurun_out/) into the committed summaries under profiles/.

    python tools/summarize_profiles.py [round_tag]

Writes  profiles/<tag>_launches.csv        per-launch device time (ncu --metrics gpu__time_duration.sum), trimmed
        profiles/<tag>_launch_summary.txt  time share per kernel over the captured launches
        profiles/<tag>_gemm_metrics.txt    key `ncu --set full` metrics of the GEMM launches of one block
        profiles/<tag>_attn_metrics.txt    same for the attention / row_norm kernels
        profiles/traffic.json               dram bytes per launch of the FF1 GEMM (bench.py reads it as roofline.traffic)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(PR, exist_ok=True)


def short(name):
    name = re.sub(r"\(CUtensorMap.*", "", name)
    return name.replace("void f5::", "").replace("f5::", "")


def launches():
    src = os.path.join(GO, "launches.csv")
    if not os.path.exists(src):
        return
    lines = open(src).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    with open(os.path.join(PR, f"{tag}_launches.csv"), "w") as f:
        f.write("id,kernel,grid,block,duration_us\n")
        for d in rows:
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v * 1000 if u == "ms" else v / 1000 if u == "ns" else v
            k = short(d["Kernel Name"])
            f.write(f'{d["ID"]},"{k}","{d["Grid Size"]}","{d["Block Size"]}",{v:.3f}\n')
            a = agg.setdefault((k, d["Grid Size"]), [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PR, f"{tag}_launch_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none over {len(rows)} consecutive launches of\n"
                f"# `python bench.py --steps 1 --warmup 0 --no-cpu-baseline` (cfg2).  Cold-cache, serialised: compare SHARES.\n")
        f.write(f"{'total_us':>10} {'n':>5} {'avg_us':>8} {'share':>6}  kernel  grid\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:10.1f} {v[0]:5d} {v[1] / v[0]:8.2f} {100 * v[1] / tot:5.1f}%  {k[0]}  {k[1]}\n")
        f.write(f"total {tot:.1f} us\n")
    print(open(os.path.join(PR, f"{tag}_launch_summary.txt")).read())


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def metrics(rep, out, traffic_key=None):
    path = os.path.join(GO, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = None
    with open(os.path.join(PR, out), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({rep}); one column per captured launch\n")
        for r in rows[2:]:
            f.write(f"\n== {short(r[idx['Kernel Name']])}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"   {k} [{units[idx[k]]}] = {r[idx[k]]}\n")
            def tobytes(k):
                v, u = float(r[idx[k]].replace(",", "")), units[idx[k]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

            try:  # derived: achieved DRAM / L2 bandwidth of this launch (the HBM peak is MEASURED_PEAKS.json hbm_gbs)
                dur, du = float(r[idx["gpu__time_duration.sum"]].replace(",", "")), units[idx["gpu__time_duration.sum"]]
                dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(du, 1e-9)
                dram = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
                f.write(f"   derived: dram bytes {dram:.0f}  -> {dram / dur_s / 1e9:.1f} GB/s DRAM"
                        f"  | L2 traffic {tobytes('lts__t_bytes.sum') / dur_s / 1e9:.1f} GB/s\n")
            except Exception:  # noqa: BLE001
                pass
            if traffic_key and traffic_key in r[idx["Kernel Name"]] and traffic is None:
                traffic = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    print(open(os.path.join(PR, out)).read()[:3000])
    return traffic


launches()
KEYS += ["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
traffic = {}
t = metrics("prof_gemm.ncu-rep", f"{tag}_gemm_metrics.txt", traffic_key="gemm_tcgen05")
if t is not None:
    traffic["gemm_dram_bytes_per_launch"] = t
t = metrics("prof_attn.ncu-rep", f"{tag}_attn_metrics.txt", traffic_key="attn_fwd")
if t is not None:
    traffic["attention_dram_bytes_per_launch"] = t
metrics("prof_bw.ncu-rep", f"{tag}_bandwidth_kernels_metrics.txt")
if traffic:
    traffic["source"] = f"profiles/{tag}_gemm_metrics.txt / {tag}_attn_metrics.txt (first captured launch of each kernel)"
    json.dump(traffic, open(os.path.join(PR, "traffic.json"), "w"))
    print("traffic", traffic)
