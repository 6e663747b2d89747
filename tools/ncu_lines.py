"""Join the SASS page of an .ncu-rep (per-instruction samples / executed counts) with nvdisasm's line table of the
shipped cubin and aggregate by source line. Usage:
  python tools/ncu_lines.py REP KERNEL_REGEX CUBIN MANGLED_SUBSTR [file.cu] [--top N]
Lines inside inlined helpers are attributed to the outermost call site in `file.cu` (default: the kernel's own file)."""
import collections
import csv
import re
import subprocess
import sys


def line_table(cubin, mangled, want_file):
    txt = subprocess.check_output(["nvdisasm", "-g", "-c", cubin], text=True)
    out, cur, on = [], None, False
    for l in txt.splitlines():
        if l.startswith("\t.section\t.text."):
            on = mangled in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            f, n, rest = m.group(1), int(m.group(2)), m.group(3)
            chain = [(f, n)] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', rest)]
            cur = chain
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            out.append(cur)
    return out


def pick(chain, want_file):
    if not chain:
        return ("?", 0)
    for f, n in reversed(chain):   # outermost first
        if f.endswith(want_file):
            return (want_file, n)
    return (chain[0][0].split("/")[-1], chain[0][1])


def main():
    rep, kre, cubin, mangled = sys.argv[1:5]
    want = sys.argv[5] if len(sys.argv) > 5 and not sys.argv[5].startswith("--") else "update_tc.cu"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 60
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                                   "--launch-count", "1"], text=True)
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lt = line_table(cubin, mangled, want)
    if len(lt) != len(data):
        print("warning: %d SASS rows vs %d disassembled instructions" % (len(data), len(lt)))
    agg = collections.defaultdict(lambda: [0, 0])
    tot_s = tot_i = 0
    for r, ch in zip(data, lt):
        k = pick(ch, want)
        s, i = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        agg[k][0] += s; agg[k][1] += i; tot_s += s; tot_i += i
    src = {}
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    print("%-22s %8s %6s %10s %6s" % ("line", "samples", "%", "instr", "%"))
    for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %8d %6.1f %10d %6.1f" % ("%s:%d" % k, s, 100.0 * s / max(tot_s, 1), i, 100.0 * i / max(tot_i, 1)))


if __name__ == "__main__":
    main()
