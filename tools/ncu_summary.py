#!/usr/bin/env python
"""Summarise an ncu --csv metrics pass over one forward + write_results (tools/ncu_traffic.sh):

    python tools/ncu_summary.py <launches.csv> <out prefix>

writes <prefix>.txt (one line per launch: kernel, duration, DRAM bytes read / written, tensor-pipe activity, L2 hit
rate) and <prefix>.json (per kernel class: launches, total time, share of the step, DRAM bytes per launch; for the
tcgen05 convolution class also the algorithmic bytes per launch).  bench.py reads the committed copy
profiles/r2_conv_traffic.json for `roofline.traffic`."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def algorithmic_conv_bytes(batch=64, reso=416, cfg_name="yolov3"):
    """sum over the tcgen05 convolutions of (input + output + weights [+ shortcut operand]) at 2 bytes/element, and
    the number of those launches (the stem runs its own kernel)"""
    from realtimeobjectdetection_b200 import synth
    from realtimeobjectdetection_b200.cfg import builtin_cfg, parse_cfg
    blocks = parse_cfg(builtin_cfg(cfg_name))
    table = synth.layer_table(blocks)
    size, total, n = reso, 0, 0
    sizes = []
    for row in table:
        kind = row["type"]
        if kind == "convolutional":
            out = (size + 2 * row["pad"] - row["size"]) // row["stride"] + 1
            if row["index"] > 0:
                nxt = table[row["index"] + 1] if row["index"] + 1 < len(table) else None
                res = nxt is not None and nxt["type"] == "shortcut"
                head = nxt is not None and nxt["type"] == "yolo"
                act_in = batch * size * size * row["cin"] * 2
                act_out = batch * out * out * row["cout"] * (4 if head else 2)
                total += act_in + act_out * (2 if res else 1) + row["cout"] * row["cin"] * row["size"] ** 2 * 2
                n += 1
            size = out
        elif kind == "upsample":
            size *= 2
        elif kind == "maxpool":
            size = size // row["stride"] if row["stride"] != 1 else size
        elif kind == "route":
            size = sizes[row["sources"][0]]
        sizes.append(size)
    return total, n


def main():
    path, prefix = sys.argv[1], sys.argv[2]
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    launches = {}
    for r in csv.DictReader(lines):
        if "Kernel Name" not in r or "Metric Name" not in r:
            continue
        key = int(r["ID"])
        d = launches.setdefault(key, {"kernel": r["Kernel Name"]})
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0,
                 "nsecond": 1e-6, "ns": 1e-6, "second": 1e3, "s": 1e3}.get(unit, 1.0)
        d[r["Metric Name"]] = val * scale
    for key in sorted(launches):
        d = launches[key]
        rows.append({"kernel": d["kernel"], "ms": d.get("gpu__time_duration.sum", 0.0),
                     "rd": d.get("dram__bytes_read.sum", 0.0), "wr": d.get("dram__bytes_write.sum", 0.0),
                     "tensor": d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0.0),
                     "l2hit": d.get("lts__t_sector_hit_rate.pct", 0.0), "l2bytes": d.get("lts__t_bytes.sum", 0.0)})
    total_ms = sum(r["ms"] for r in rows) or 1.0
    with open(prefix + ".txt", "w") as fh:
        fh.write("# one forward + write_results, YOLOv3-416 batch 64 (tools/ncu_traffic.sh); times are cold-cache and "
                 "serialised by the profiler: compare SHARES, not absolutes\n")
        for r in rows:
            fh.write("%-60s %9.1f us  dram rd %8.2f MB wr %8.2f MB  L2 %8.2f MB hit %5.1f %%  tensor %5.1f %%\n"
                     % (r["kernel"][:60], r["ms"] * 1e3, r["rd"] / 1e6, r["wr"] / 1e6, r["l2bytes"] / 1e6, r["l2hit"], r["tensor"]))
    classes = {}
    for r in rows:
        name = r["kernel"]
        cls = "conv (tcgen05)" if ("conv_tc_kernel" in name or "conv_pair_kernel" in name) else name.split("(")[0].split("::")[-1][:40]
        c = classes.setdefault(cls, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0, "l2_bytes": 0.0})
        c["launches"] += 1
        c["ms"] += r["ms"]
        c["dram_bytes"] += r["rd"] + r["wr"]
        c["l2_bytes"] += r["l2bytes"]
    out = {"classes": {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share_of_step": round(v["ms"] / total_ms, 4),
                           "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
                           "l2_bytes_per_launch": v["l2_bytes"] / v["launches"]} for k, v in classes.items()},
           "total_ms_under_ncu": round(total_ms, 4)}
    conv = classes.get("conv (tcgen05)")
    if conv:
        alg, n = algorithmic_conv_bytes()
        out["yolov3-416-B64-fp16"] = {"dram_bytes_per_launch": conv["dram_bytes"] / conv["launches"],
                                      "algorithmic_bytes_per_launch": alg / n, "launches": conv["launches"],
                                      "share_of_step_under_ncu": round(conv["ms"] / total_ms, 4),
                                      "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d tcgen05 "
                                                "convolution launches of one forward (tools/ncu_traffic.sh)" % conv["launches"]}
    with open(prefix + ".json", "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps(out.get("yolov3-416-B64-fp16", out), indent=1))


if __name__ == "__main__":
    main()
