"""profiles/kernel_traffic.json from an ncu CSV of ONE optimizer step (dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum per launch): measured DRAM bytes per kernel family and for the whole step -- `roofline.traffic` of bench.py.

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/r2_ncu_traffic_c2.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e --cuda-profiler
    python tools/traffic_from_ncu.py gpurun_out/r2_ncu_traffic_c2.csv c2_lstm
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILY = [("gemm_fwd_dgrad", r"gemm_tf32x3_(wtmem_)?kernel"), ("gemm_wgrad", r"gemm_wgrad_atmem_kernel|wgrad_reduce_kernel"),
          ("rnn", r"(fwd|bwd)_(resident|cluster|generic)_kernel|(fwd|bwd)_gate_kernel"), ("unit_dgrad_fused", r"unit_dgrad_fused_kernel"),
          ("encoder", r"unit_|target_unit|env_(fwd|bwd)"),
          ("ppo_loss", r"ppo_"), ("grad_finish", r"grad_sumsq|adam_kernel|finish_tail|grad_flags")]


def main(path, key):
    per = collections.defaultdict(lambda: [0.0, 0.0, 0])          # kernel -> [bytes, ns, launches]
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
    for r in csv.reader(open(path)):
        if len(r) < 15 or not r[0].isdigit():
            continue
        name, metric, u, val = r[4], r[12], r[13], float(r[14].replace(",", ""))
        k = re.sub(r"\(.*", "", name).replace("void ", "")
        if metric.startswith("dram__bytes"):
            per[k][0] += val * unit.get(u, 1.0)
        elif metric.startswith("gpu__time_duration"):
            per[k][1] += val * unit.get(u, 1.0)
            per[k][2] += 1
    fam = collections.defaultdict(float)
    for k, (b, ns, n) in per.items():
        for f, pat in FAMILY:
            if re.search(pat, k):
                fam[f] += b
                break
        else:
            fam["other"] += b
    fam["step_total"] = sum(v[0] for v in per.values())
    out_path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    doc = json.load(open(out_path)) if os.path.exists(out_path) else {}
    doc["_doc"] = ("measured DRAM bytes per optimizer step (dram__bytes_read.sum + dram__bytes_write.sum, ncu, ONE step of `bench.py "
                   "--config c2` replayed from its graph), summed over the launches of each kernel family (tools/traffic_from_ncu.py); "
                   "read by bench.py for roofline.traffic")
    doc[key] = dict(fam)
    doc[key + "_kernels"] = {k: {"dram_bytes": v[0], "us": v[1] / 1e3, "launches": v[2]} for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])}
    json.dump(doc, open(out_path, "w"), indent=1)
    for f, b in sorted(fam.items(), key=lambda kv: -kv[1]):
        print("%-16s %8.3f GB" % (f, b / 1e9))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
