"""profiles/r2_traffic.json from an `ncu --set full` capture: per kernel, the per-launch average of
dram__bytes_read.sum / dram__bytes_write.sum, the duration, the tensor-pipe share and the SM clock (bench.py reads the
bytes for `roofline.traffic`).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_traffic.py raw.csv [--suffix=@arm] [note]
Entries are keyed by kernel name (+ suffix: the weights arm the capture was taken on)."""
import collections
import csv
import json
import os
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
        "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def short_name(name):
    m = re.search(r"(gemm_pair_kernel|gemm_tc_kernel)<(?:dlc::)?(\w+)(?:<([^>]*)>)?", name)
    if m:
        args = m.group(3)
        if args is not None:
            parts = [a.strip() for a in args.split(",")]
            if m.group(2) == "BiasActPolicy":
                parts = parts[:2] + (["conv"] if len(parts) > 2 and parts[2] in ("1", "true") else [])
            return "%s<%s<%s>>" % (m.group(1), m.group(2), ",".join(parts))
        return "%s<%s>" % (m.group(1), m.group(2))
    return re.sub(r"^(void )?(dlc::)?", "", re.sub(r"\(.*", "", name)).strip()


def main(path, note="", suffix=""):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        if key not in idx or r[idx[key]] in ("", "n/a"):
            return None
        return float(r[idx[key]].replace(",", "")) * UNIT.get(units[idx[key]], 1.0)

    agg = collections.OrderedDict()
    for r in rows[2:]:
        k = short_name(r[idx["Kernel Name"]]) + suffix
        a = agg.setdefault(k, collections.defaultdict(list))
        for key, out in (("dram__bytes_read.sum", "dram_bytes_read"), ("dram__bytes_write.sum", "dram_bytes_write"),
                         ("gpu__time_duration.sum", "gpu_time_ms"),
                         ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
                         ("sm__cycles_elapsed.avg.per_second", "sm_clock_ghz"),
                         ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                         ("lts__t_sector_hit_rate.pct", "l2_hit_rate_pct")):
            v = val(r, key)
            if v is not None:
                a[out].append(v)
    out = {}
    for k, a in agg.items():
        out[k] = {m: sum(v) / len(v) for m, v in a.items()}
        out[k]["launches_captured"] = len(a["gpu_time_ms"])
        out[k]["source"] = note or os.path.basename(path)
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json")
    prev = {}
    if os.path.exists(dst):
        prev = json.load(open(dst))
    prev.update(out)
    json.dump(prev, open(dst, "w"), indent=1)
    for k, v in out.items():
        print("%-60s %3d launches  %8.3f ms  read %8.1f MB  write %8.1f MB  tensor %5.1f %%" % (
            k[:60], v["launches_captured"], v.get("gpu_time_ms", 0), v.get("dram_bytes_read", 0) / 1e6,
            v.get("dram_bytes_write", 0) / 1e6, v.get("tensor_pipe_active_pct", 0)))


if __name__ == "__main__":
    args = sys.argv[2:]
    sfx = ""
    if args and args[0].startswith("--suffix="):
        sfx = args.pop(0).split("=", 1)[1]
    main(sys.argv[1], " ".join(args), sfx)
