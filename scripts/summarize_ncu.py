"""Summarise an .ncu-rep (read here, on the CPU box) into profiles/<name>.md + traffic.json."""
import csv, io, json, os, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_bytes.sum", "L1 bytes"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe / issue"),
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary: {os.path.basename(rep)}", "", note, ""]
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]]
        lines.append(f"## {name[:100]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        dram = 0.0
        for key, label in KEYS:
            if key in col:
                v, u = r[col[key]], units[col[key]]
                lines.append(f"| {label} (`{key}`) | {v} | {u} |")
                if key.startswith("dram__bytes"):
                    dram += float(v.replace(",", "")) * UNIT_SCALE.get(u, 1.0)
        lines.append(f"| **DRAM traffic read+write** | {dram:.0f} | byte |")
        lines.append("")
        traffic.setdefault(name.split("(")[0], []).append(dram)
    with open(out_md, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out_md)
    return traffic


if __name__ == "__main__":
    t = main()
    print(json.dumps({k: v for k, v in t.items()}))
