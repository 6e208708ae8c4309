"""profiles/ncu_traffic.json from ONE ncu launch list of a step (tools/profile_step.py under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) and the library's own kernel table of the
same step (VQA_PROF_REPORT): per kernel, DRAM bytes summed over all launches of the step against the algorithmic bytes the
library declares for them (VQA_BYTES in csrc/) -> the ratio bench.py applies to `roofline.traffic`.
usage: traffic_from_csv.py launches.csv report.json [tag]"""
import csv, json, os, re, sys


def norm(name):
    n = name.split("(const")[0].split("(unsigned")[0].split("(float")[0].split("(int")[0].split("(double")[0].split("(CUtensorMap")[0]
    n = n.replace("void ", "").replace("vqa::", "").replace("(int)", "").replace("(bool)", "").replace(" ", "")
    n = n.split("(")[0]
    n = re.sub(r"k_dct_umma<256,1(,64)?>", "k_dct_umma<BN1,1>", n)
    n = re.sub(r"k_dct_umma<128,2(,64)?>", "k_dct_umma<BN2,2>", n)
    n = re.sub(r"k_fb_pyramid_dec<[0-9,]+>", "k_fb_pyramid_dec<S,R,TO>", n)
    n = re.sub(r"<1>$", "<true>", n) if n.startswith(("k_yuv420_gray_hist", "k_gray_hist", "k_orb64")) else n
    n = re.sub(r"<0>$", "<false>", n) if n.startswith(("k_yuv420_gray_hist", "k_gray_hist", "k_orb64")) else n
    return n


def main():
    launches, report = sys.argv[1], json.load(open(sys.argv[2]))
    tag = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(launches)
    rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    per = {}
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        k = norm(r[ik])
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        e = per.setdefault(k, {"dram": 0.0, "us": 0.0, "n": 0})
        if r[im].startswith("dram__bytes"):
            e["dram"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif r[im] == "gpu__time_duration.sum":
            e["us"] += v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
            e["n"] += 1
    out = {}
    for k, e in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        rep = report["kernels"].get(k)
        ab = rep["bytes"] if rep and rep["bytes"] else None
        out[k] = {"dram_bytes": e["dram"], "algorithmic_bytes": ab, "ratio": (e["dram"] / ab) if ab else None,
                  "duration_us": round(e["us"], 1), "launches": e["n"], "launches_declared": rep["launches"] if rep else None,
                  "note": None if ab else "no declared bytes: data-dependent work list (8 B per tile-local root) or a fixed few hundred bytes",
                  "report": tag, "step": "%d frames %dx%d, all launches of the step summed (ncu times are cold-cache and serialised)" % (
                      report["frames"], report["width"], report["height"])}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    for k, v in out.items():
        print("%-28s %3d launches %9.1f us  dram %9.1f MB  algorithmic %s  ratio %s" % (
            k, v["launches"], v["duration_us"], v["dram_bytes"] / 1e6,
            "%9.1f MB" % (v["algorithmic_bytes"] / 1e6) if v["algorithmic_bytes"] else "     none", "%.2f" % v["ratio"] if v["ratio"] else "-"))


if __name__ == "__main__":
    main()
