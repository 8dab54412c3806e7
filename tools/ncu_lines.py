"""Per-source-line executed-instruction counts of one kernel: joins an .ncu-rep's SASS page with the line table of the SAME
cubin (extracted from libexl2b200.so).  python tools/ncu_lines.py rep.ncu-rep <kernel substring> [cubin substring] [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
cub = sys.argv[3] if len(sys.argv) > 3 else "gemv_i8"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
if os.environ.get("CUBIN"):          # a cubin saved from the build the report was taken with
    cubin_path = os.environ["CUBIN"]
else:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "exllamav2_b200", "libexl2b200.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin_path = os.path.join(tmp, [f for f in os.listdir(tmp) if cub in f and f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin_path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
seq, cur, on = [], None, False
for line in dis.split("\n"):
    if line.startswith(".text."): on = kern in line
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line): seq.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[1]; iE, iN = h.index("Instructions Executed"), h.index("# Samples")
data = r[2:]
print(f"instructions: cubin {len(seq)}, report {len(data)}" + ("" if len(seq) == len(data) else "   MISMATCH: report is from another build"))
agg, smp = collections.Counter(), collections.Counter()
for ln, row in zip(seq, data):
    agg[ln] += int(row[iE]); smp[ln] += int(row[iN])
tot = sum(agg.values()); ts = max(1, sum(smp.values()))
srcs = {}
def text(k):
    if not k: return ""
    p = os.environ.get("SRC") if (os.environ.get("SRC") and k[0] == os.path.basename(os.environ["SRC"]).split("_")[-2] + "_" + os.path.basename(os.environ["SRC"]).split("_")[-1]) else os.path.join(root, "exllamav2_b200", "csrc", k[0])
    if k[0] not in srcs: srcs[k[0]] = open(p).read().split("\n") if os.path.exists(p) else None
    return srcs[k[0]][k[1] - 1].strip()[:110] if srcs[k[0]] else ""
print(f"total warp instructions {tot}, stall samples {ts}")
for k, v in agg.most_common(top):
    print(f"{100 * v / tot:5.1f}% {100 * smp[k] / ts:5.1f}%s  {k[0] if k else '?'}:{k[1] if k else 0:<5d} {text(k)}")
