#!/usr/bin/env python3
"""Fold what tools/gpu_profile.sh left in gpurun_out/ into the committed evidence under profiles/<round>/ and
profiles/traffic.json (read by bench.py for roofline.traffic).   usage: tools/profile_summary.py r2"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
OUT = os.path.join(ROOT, "profiles", rnd)
os.makedirs(OUT, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max"]


def launches():
    path = os.path.join(GO, "prof_launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mbpe::", "")
        name = re.sub(r"<.*", "", name) if not name.startswith("k_par") else name
        t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values()) or 1.0
    cmd = next((l.split("running: ", 1)[-1].strip() for l in open(os.path.join(ROOT, "tools", "gpu_profile.sh")) if l.startswith("CMD=")), "")
    with open(os.path.join(OUT, "launches_bench.md"), "w") as f:
        f.write(f"# ncu launch list of `{cmd}`\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and serialised: "
                f"shares, not absolutes). {sum(v[0] for v in agg.values())} launches, {total / 1e3:.1f} ms of kernel time.\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| `{k[:90]}` | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / total:.1f} % |\n")
    print("launch list:", len(agg), "kernels")


def full(rep, kernel, note):
    path = os.path.join(GO, rep + ".ncu-rep")
    if not os.path.exists(path):
        return None
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    with open(os.path.join(OUT, f"ncu_full_{kernel}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch of {d.get('Kernel Name', ('?',))[0]}\n# {note}\n")
        for k in KEYS:
            if k in d:
                f.write(f"{k:72s} {d[k][0]:>22s} {d[k][1]}\n")
        f.write("\n# warp stall reasons, cycles per issued instruction\n")
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                f.write(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {float(d[h][0]):8.2f}\n")
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), path, "30"], capture_output=True, text=True).stdout
    open(os.path.join(OUT, f"{kernel}_lines.txt"), "w").write(f"# stall samples / executed instructions by CUDA source line ({rep}.ncu-rep)\n" + lines)

    def num(k, scale={"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}):
        v, u = d[k]
        return float(v.replace(",", "")) * scale.get(u, 1)
    return {"dram_bytes_per_launch": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
            "dram_bytes_read": int(num("dram__bytes_read.sum")), "dram_bytes_write": int(num("dram__bytes_write.sum")),
            "launch_seconds_under_ncu": num("gpu__time_duration.sum"), "warp_instructions": int(num("smsp__inst_executed.sum")),
            "source": f"profiles/{rnd}/ncu_full_{kernel}.txt", "launch": note}


launches()
traffic = {}
for rep, kernel, note in (("prof_persistent", "k_persistent", "launch 300 of `bench.py` (1 GiB corpus, vocab 32768, lexical): the resident merge loop between two grid-wide steps"),
                          ("prof_encode", "k_encode_tiles", "tools/enc_ab.py 512 0: first launch of a warm pass over 512 MiB of text = 2^26 chunks (about 306 MiB of text, 141 M ids)"),
                          ("prof_decode", "k_decode_lean", "tools/dec_ab.py 1024 (default configuration): one launch = 1 GiB of text, 442 M ids")):
    t = full(rep, kernel, note)
    if t:
        traffic[kernel] = t
        print(kernel, t)
if traffic:
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
