"""ncu --csv launch list of ONE whole evaluation (tools/ncu_step.py) -> profiles/r02_step_<workload>.md and an entry of
profiles/r02_step_traffic.json (read by bench.py for roofline.traffic):

    python tools/summarize_step_ncu.py <workload> <csv> [weights_bytes]

Per kernel name: launches, total / mean time, DRAM bytes read + written.  Whole-step DRAM bytes vs the algorithmic minimum of the step
(pre-split weights streamed once + boundary activations) is the re-read ratio VERDICT r01 asked for."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").strip()


def main(workload, path, weight_bytes=None):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    iid, ik, im, iu, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    launches = OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[iid], dict(name=short(r[ik])))
        d[r[im]] = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1.0)
    per = OrderedDict()
    for d in launches.values():
        k = per.setdefault(d["name"], dict(n=0, us=0.0, rd=0.0, wr=0.0))
        k["n"] += 1
        k["us"] += d.get("gpu__time_duration.sum", 0.0)
        k["rd"] += d.get("dram__bytes_read.sum", 0.0)
        k["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot_us = sum(k["us"] for k in per.values())
    tot_b = sum(k["rd"] + k["wr"] for k in per.values())
    gemm = [k for n, k in per.items() if "gemm_tc_kernel" in n]
    g_n = sum(k["n"] for k in gemm)
    g_b = sum(k["rd"] + k["wr"] for k in gemm)
    lines = [f"# r02 — whole-step ncu launch list, workload {workload}", "",
             f"`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none` over ONE "
             f"eager evaluation of the launch plan (tools/ncu_step.py): {len(launches)} launches, {tot_us / 1e3:.3f} ms of serialised kernel time "
             f"(compare shares, not absolutes: kernels run back to back in the graph), {tot_b / 1e6:.1f} MB of DRAM traffic.", "",
             "| kernel | launches | total us | share | avg us | DRAM read MB | DRAM write MB |", "|---|---:|---:|---:|---:|---:|---:|"]
    for n, k in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        lines.append(f"| `{n}` | {k['n']} | {k['us']:.1f} | {100 * k['us'] / tot_us:.1f}% | {k['us'] / k['n']:.2f} | {k['rd'] / 1e6:.2f} | {k['wr'] / 1e6:.2f} |")
    entry = dict(launches=len(launches), kernel_time_ms_serialised=tot_us / 1e3, dram_bytes_per_step=tot_b, gemm_launches=g_n,
                 gemm_dram_bytes=g_b, gemm_bytes_per_launch=g_b / max(g_n, 1), source=os.path.basename(path))
    if weight_bytes:
        entry["algorithmic_bytes_per_step"] = float(weight_bytes)
        entry["traffic_over_algorithmic"] = tot_b / float(weight_bytes)
        lines += ["", f"Algorithmic bytes of the step (tensor-core weight operands streamed once): {float(weight_bytes) / 1e6:.1f} MB -> measured / algorithmic = "
                      f"**{tot_b / float(weight_bytes):.2f}**."]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"r02_step_{workload}.md"), "w").write("\n".join(lines) + "\n")
    jp = os.path.join(ROOT, "profiles", "r02_step_traffic.json")
    allj = json.load(open(jp)) if os.path.exists(jp) else {}
    allj[workload] = entry
    json.dump(allj, open(jp, "w"), indent=1)
    print("\n".join(lines[:12]))
    print(json.dumps(entry))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
