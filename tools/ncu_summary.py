"""Summarise `ncu --set full` reports (*.ncu-rep) into profiles/<name>.md + profiles/traffic.json.
usage: python tools/ncu_summary.py gpurun_out/prof_r1_conv3x3.ncu-rep [...]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__inst_executed_pipe_tensor.*", r"^sm__pipe_tensor.*cycles_active.*pct.*",
    r"^sm__pipe_tensor_subpipe.*", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__t_bytes\.sum$",
    r"^lts__t_sector_hit_rate\.pct$", r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^launch__registers_per_thread$", r"^launch__grid_size$",
    r"^launch__block_size$", r"^launch__shared_mem_per_block_dynamic$", r"^launch__occupancy_limit.*", r"^sm__cycles_active\.avg$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^smsp__average_warp.*issue_stalled.*ratio$", r"^smsp__average_warps_issue_stalled.*", r"^sm__inst_executed\.avg\.per_cycle_active$",
    r"^smsp__cycles_active\.avg$", r"^dram__cycles_active.*", r"^sm__ctas_launched\.sum$",
]


def main(paths):
    os.makedirs(os.path.join(REPO, "profiles"), exist_ok=True)
    traffic_path = os.path.join(REPO, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.PIPE).stdout.decode()
        lines = [l for l in out.splitlines() if not l.startswith("==")]
        rows = list(csv.reader(io.StringIO("\n".join(lines))))
        if len(rows) < 3:
            print("no data in", path)
            continue
        hdr, units = rows[0], rows[1]
        name = os.path.splitext(os.path.basename(path))[0]
        md = ["# ncu --set full summary: %s" % name, "", "source: `%s` (captured under gpurun with `--clock-control none`)" % os.path.basename(path), ""]
        for r in rows[2:]:
            rec = dict(zip(hdr, r))
            kname = rec.get("Kernel Name", "?")
            md.append("## %s" % kname[:160])
            md.append("")
            md.append("| metric | value | unit |")
            md.append("|---|---|---|")
            for h, u in zip(hdr, units):
                if any(re.search(k, h) for k in KEYS):
                    md.append("| %s | %s | %s |" % (h, rec.get(h, ""), u))
            try:
                rd = float(rec["dram__bytes_read.sum"].replace(",", "")); wr = float(rec["dram__bytes_write.sum"].replace(",", ""))
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                u_rd = units[hdr.index("dram__bytes_read.sum")]; u_wr = units[hdr.index("dram__bytes_write.sum")]
                total = rd * scale.get(u_rd, 1) + wr * scale.get(u_wr, 1)
                md.append("")
                md.append("DRAM traffic per launch: %.1f MB" % (total / 1e6))
                key = re.sub(r"^prof_r\d+_", "", name)
                traffic[key] = total
            except Exception as e:
                md.append("(no dram bytes: %s)" % e)
            md.append("")
        with open(os.path.join(REPO, "profiles", name + ".md"), "w") as f:
            f.write("\n".join(md) + "\n")
        print("wrote profiles/%s.md" % name)
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
