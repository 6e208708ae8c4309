"""Per CUDA source line: warp instructions executed and stall samples of one kernel of an
`ncu --set full --import-source on` report (needs -lineinfo).  usage: ncu_lines.py report kernel-substring [top]"""
import csv, io, subprocess, sys


def main():
    path, want = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    blocks = out.split('"File Path"')
    for blk in blocks[1:]:
        rows = list(csv.reader(io.StringIO('"File Path"' + blk)))
        name = rows[1][1]
        if want not in name:
            continue
        hdr = rows[2]
        il, isrc, iex, ismp = hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        per = {}
        for r in rows[3:]:
            if len(r) != len(hdr) or not r[il].isdigit() or not r[iex].isdigit():
                continue
            per[int(r[il])] = [r[isrc], int(r[iex]), int(r[ismp]) if r[ismp].isdigit() else 0]
        tot = sum(v[1] for v in per.values()) or 1
        tots = sum(v[2] for v in per.values()) or 1
        print("==", name[:90], "warp instr", tot, "samples", tots)
        for ln, (src, n, s) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
            print("%5d %5.1f%% inst %5.1f%% smp | %s" % (ln, 100 * n / tot, 100 * s / tots, src.strip()[:110]))
        break


if __name__ == "__main__":
    main()
