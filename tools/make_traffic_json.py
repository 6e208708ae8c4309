"""profiles/ncu_traffic.json from `ncu --set full` reports: per kernel, DRAM bytes of the captured launch,
its algorithmic bytes (same model as VQA_BYTES in csrc/) and the ratio bench.py applies."""
import json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_summary as S

HW = 1080 * 1920
def alg_bytes(kernel, grid):
    g = [int(v) for v in re.findall(r"\d+", grid)]
    z = g[2] if len(g) > 2 else 1
    if "k_fb_blur_solve" in kernel: return 28.0 * HW * z
    if "k_fb_matrices<0>" in kernel or "k_fb_matrices_v4<0>" in kernel or "k_fb_matrices_t4<0>" in kernel: return 68.0 * HW * z
    if "k_fb_matrices<1>" in kernel: return (60.0 * HW + 8.0 * HW / 4) * z
    if "k_fb_polyexp" in kernel: return 24.0 * HW * z
    if "k_canny_nms" in kernel: return 2.0 * HW * z
    if "k_gray_hist" in kernel: return 4.0 * HW * g[1]
    if "k_psnr_ssim" in kernel: return 2.0 * HW * z
    return None

def norm(name):
    n = name.split("(")[0].replace("void ", "").replace(" ", "")
    n = n.replace("<256,1>", "<BN1,1>").replace("<128,2>", "<BN2,2>").replace("vqa::", "")
    n = re.sub(r"k_gray_hist<1,0>", "k_gray_hist<true,false>", n)
    return n

path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
out = json.load(open(path)) if os.path.exists(path) and "--merge" in sys.argv else {}
for p in [a for a in sys.argv[1:] if not a.startswith("--")]:
    for e in S.load(p):
        tr = e.get("dram_rd", 0) + e.get("dram_wr", 0)
        ab = alg_bytes(e["kernel"], e["grid"])
        out[norm(e["kernel"])] = {"dram_bytes": tr, "algorithmic_bytes": ab, "ratio": (tr / ab) if ab else None,
                                  "duration_us": e.get("dur_us"), "grid": e["grid"], "report": os.path.basename(p),
                                  "dram_pct": e.get("dram_pct"), "tensor_pct": e.get("tensor_pct"), "l1_pct": e.get("l1_pct")}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
