"""Where the instructions of a kernel go: reads the SOURCE page of an `ncu --set full --import-source on`
report and prints, per kernel, thread instructions per launch, the share of warp instructions and stall
samples between consecutive barriers (the phases of a tiled kernel) and the opcode mix.  This view found the
two wins of round 1's last session (Canny's tile-wide CCL passes, UpdateMatrices' 16-sector gathers).

usage: python tools/ncu_phases.py report.ncu-rep [pixels-per-launch]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def pages(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name"')
    for blk in blocks[1:]:
        rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
        hdr = rows[1]
        yield rows[0][1], hdr, [r for r in rows[2:] if len(r) == len(hdr)]


def main():
    path = sys.argv[1]
    px = float(sys.argv[2]) if len(sys.argv) > 2 else None
    seen = set()
    for name, hdr, data in pages(path):
        if name in seen:
            continue
        seen.add(name)
        ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
        tot = sum(int(r[ia]) for r in data) or 1
        tots = sum(int(r[isamp]) for r in data) or 1
        print("== %s\n   warp instructions %d, SASS lines %d%s" % (
            name[:100], tot, len(data), ", thread instructions per pixel %.1f" % (tot * 32 / px) if px else ""))
        acc = accs = seg = start = 0
        for i, r in enumerate(data):
            acc += int(r[ia])
            accs += int(r[isamp])
            if "BAR.SYNC" in r[isrc] or i == len(data) - 1:
                print("   phase %d (SASS %4d-%4d): instructions %5.1f %%, stall samples %5.1f %%" % (
                    seg, start, i, 100 * acc / tot, 100 * accs / tots))
                acc = accs = 0
                seg += 1
                start = i + 1
        ops, ops_s = Counter(), Counter()
        for r in data:
            parts = r[isrc].split()
            if not parts:
                continue
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            op = ".".join(op.split(".")[:2]) if op.startswith(("LDG", "STG", "LDS", "STS", "F2F", "I2F", "F2I")) else op.split(".")[0]
            ops[op] += int(r[ia])
            ops_s[op] += int(r[isamp])
        print("   opcodes: " + ", ".join("%s %.1f%% (%.1f%% of samples)" % (o, 100 * n / tot, 100 * ops_s[o] / tots)
                                         for o, n in ops.most_common(10)))


if __name__ == "__main__":
    main()
