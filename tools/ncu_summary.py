"""Summarise .ncu-rep files (read with `ncu -i ... --page raw --csv`) into one table."""
import csv, io, subprocess, sys, json, os
KEYS = {
 "gpu__time_duration.sum": "dur_us",
 "dram__bytes_read.sum": "dram_rd",
 "dram__bytes_write.sum": "dram_wr",
 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
 "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1_pct",
 "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
 "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
 "launch__registers_per_thread": "regs",
 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
 "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pct",
 "smsp__issue_active.avg.pct": "issue_pct",
 "lts__t_sector_hit_rate.pct": "l2_hit",
 "l1tex__t_sector_hit_rate.pct": "l1_hit",
 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
}
def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        e = {"kernel": d.get("Kernel Name", "?"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k, n in KEYS.items():
            if k in d and d[k] != "":
                try: v = float(d[k].replace(",", ""))
                except ValueError: continue
                un = u.get(k, "")
                if n == "dur_us":
                    v = v / 1e3 if un in ("ns", "nsecond") else (v * 1e3 if un in ("ms", "msecond") else v)
                if n in ("dram_rd", "dram_wr"):
                    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(un, 1)
                    v *= mult
                e[n] = v
        res.append(e)
    return res
if __name__ == "__main__":
    allr = []
    for p in sys.argv[1:]:
        for e in load(p):
            e["file"] = os.path.basename(p); allr.append(e)
            tr = e.get("dram_rd", 0) + e.get("dram_wr", 0)
            print(f"{e['kernel'][:34]:34s} grid {e['grid']:>16s} {e.get('dur_us',0):9.1f} us  dram {tr/1e6:9.1f} MB ({tr/max(e.get('dur_us',1),1e-9)/1e3:7.1f} GB/s, {e.get('dram_pct',0):5.1f}%)  "
                  f"sm {e.get('sm_pct',0):5.1f}% l1 {e.get('l1_pct',0):5.1f}% l2 {e.get('l2_pct',0):5.1f}% occ {e.get('occ_pct',0):5.1f}% tensor {e.get('tensor_pct',0):5.1f}% lsu {e.get('lsu_pct',0):5.1f}% fp64 {e.get('fp64_pct',0):5.1f}% issue {e.get('issue_pct',0):5.1f}% regs {int(e.get('regs',0))} l1hit {e.get('l1_hit',0):4.0f} l2hit {e.get('l2_hit',0):4.0f}")
    json.dump(allr, open("/tmp/ncu_summary.json", "w"), indent=1)
