"""Summarise ncu captures into profiles/ncu_traffic.json, the file bench.py reads `roofline.traffic` from.

    python tools/ncu_traffic.py render <render.ncu-rep> [<kernel regex> ...]     # per-launch DRAM bytes of the MLP kernel(s)
    python tools/ncu_traffic.py train  <launches.csv> <samples per step>          # DRAM bytes per sample over one train step

`render`: an `ncu --set full` report of bench.py's render launches; the LAST captured launch of each kernel (the fine pass,
128 samples/ray) is recorded.  `train`: an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--csv` launch list of tools/train_profile.py covering exactly one step.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def _load():
    return json.load(open(OUT)) if os.path.exists(OUT) else {}


def _num(x):
    return float(str(x).replace(",", ""))


def _scale(unit):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def render(rep, patterns):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    out = _load()
    for pat in patterns or ["mlp_tc7_kernel", "mlp_tc32_kernel", "mlp_tc_kernel", "bwd_tc_kernel", "dw_grouped_kernel"]:
        hits = [r for r in body if re.search(pat, r[col["Kernel Name"]])]
        if not hits:
            continue
        r = hits[-1]
        g = lambda m: _num(r[col[m]]) * _scale(units[col[m]]) if m in col else None
        out[pat] = {"kernel": r[col["Kernel Name"]][:120], "dram_bytes_read": g("dram__bytes_read.sum"),
                    "dram_bytes_write": g("dram__bytes_write.sum"),
                    "duration_ns": _num(r[col["gpu__time_duration.sum"]]) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(units[col["gpu__time_duration.sum"]], 1),
                    "tensor_pipe_active_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
                    "source": "ncu --set full, " + os.path.basename(rep)}
        print(pat, out[pat])
    json.dump(out, open(OUT, "w"), indent=1)


def train(csv_path, samples):
    txt = open(csv_path).read()
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(txt)))
    tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    per = {}
    for r in rows:
        m = r["Metric Name"]
        if m not in tot:
            continue
        v = _num(r["Metric Value"]) * (_scale(r["Metric Unit"]) if "bytes" in m else {"ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1))
        tot[m] += v
        k = re.sub(r"\(.*", "", r["Kernel Name"])[:60]
        per.setdefault(k, {"dram_bytes": 0.0, "ns": 0.0, "launches": 0})
        if "bytes" in m:
            per[k]["dram_bytes"] += v
        else:
            per[k]["ns"] += v
            per[k]["launches"] += 1
    out = _load()
    out["train_step"] = {"dram_bytes_per_sample": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / samples,
                         "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
                         "kernel_time_ns": tot["gpu__time_duration.sum"], "samples_per_step": samples,
                         "per_kernel": dict(sorted(per.items(), key=lambda kv: -kv[1]["ns"])[:12]),
                         "source": "ncu launch list, " + os.path.basename(csv_path)}
    print(json.dumps(out["train_step"], indent=1)[:1500])
    json.dump(out, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "render":
        render(sys.argv[2], sys.argv[3:])
    else:
        train(sys.argv[2], int(sys.argv[3]))
