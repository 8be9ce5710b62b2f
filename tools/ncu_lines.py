"""Development aid: per-source-line instruction counts and stall samples of one kernel in an
.ncu-rep (captured with --import-source on), joined by instruction order with the line table of
the cubin inside libb200sa.so (nvdisasm -g).  Runs here (no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep msd_local_sort_kernel [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stralg_b200", "lib", "libb200sa.so")


def sass_lines(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        secs = re.split(r"\n(?=\.text\.)", out)
        for s in secs:
            head = s.split("\n", 1)[0]
            if head.startswith(".text.") and kernel in head:
                lines, cur = [], None
                for ln in s.split("\n"):
                    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                    if m:
                        cur = (os.path.basename(m.group(1)), int(m.group(2)))
                        continue
                    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
                        lines.append((cur, ln.split("*/", 1)[1].strip().rstrip(";")))
                return head, lines
    raise SystemExit(f"kernel {kernel} not found in {LIB}")


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    head, sl = sass_lines(kernel)
    # NCU_NAME: ncu-side (demangled) name regex when it differs from the cubin symbol substring;
    # NCU_INDEX: which of the matching launches in the report (default 0)
    ncu_name = os.environ.get("NCU_NAME", kernel)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{ncu_name}"],
                         capture_output=True, text=True).stdout
    blocks = txt.split('"Kernel Name"')
    blk = '"Kernel Name"' + blocks[1 + int(os.environ.get("NCU_INDEX", "0"))]
    rows = list(csv.reader(io.StringIO(blk)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    ci = hdr.index("Instructions Executed")
    si = hdr.index("Warp Stall Sampling (All Samples)")
    print(f"{head}: {len(sl)} SASS instructions in the cubin, {len(data)} in the report")
    n = min(len(sl), len(data))
    agg = {}
    tot_i = tot_s = 0
    for k in range(n):
        key = sl[k][0]
        a = agg.setdefault(key, [0, 0])
        a[0] += int(data[k][ci] or 0)
        a[1] += int(data[k][si] or 0)
        tot_i += int(data[k][ci] or 0)
        tot_s += int(data[k][si] or 0)
    src_cache = {}

    def src(key):
        if not key:
            return ""
        f, ln = key
        p = os.path.join(ROOT, "stralg_b200", "csrc", f)
        if p not in src_cache:
            src_cache[p] = open(p).read().split("\n") if os.path.exists(p) else []
        L = src_cache[p]
        return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""

    print(f"total warp instructions {tot_i:,}  stall samples {tot_s:,}")
    print("--- by instructions executed")
    for key, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0*i/tot_i:5.1f}% inst {100.0*s/max(tot_s,1):5.1f}% samples  {key}  {src(key)}")
    print("--- by stall samples")
    for key, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100.0*s/max(tot_s,1):5.1f}% samples {100.0*i/tot_i:5.1f}% inst  {key}  {src(key)}")


if __name__ == "__main__":
    main()
