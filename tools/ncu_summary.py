"""Summarise an .ncu-rep (raw page) into a small JSON + markdown table for profiles/."""
import csv, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__cycles_elapsed.max": "cycles",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}


def main(rep, out_json):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for k, name in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                v = float(r[i].replace(",", ""))
                d[name] = v * UNIT_SCALE.get(units[i], 1)
        d["dram_bytes"] = d.get("dram_read", 0) + d.get("dram_write", 0)
        out.append(d)
    json.dump(out, open(out_json, "w"), indent=1)
    for d in out:
        print(f"{d['kernel'][:60]:60s} {d['duration']*1e3:8.3f} ms  dram {d['dram_bytes']/1e9:6.3f} GB  l1tex {d.get('l1tex_pct',0):5.1f}%  "
              f"lts {d.get('lts_pct',0):5.1f}%  issue {d.get('issue_active_pct',0):5.1f}%  L2hit {d.get('l2_hit_pct',0):5.1f}%  regs {int(d.get('regs',0))}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
