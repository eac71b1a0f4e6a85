"""profiles/traffic.json from an ncu capture of one assembly pass (bench.py --cells N --steps 1 under `ncu --set full`):
DRAM bytes per element and the kernels' shares of the pass, stamped with the digest of the CUDA sources so that bench.py
refuses a capture taken with other kernels.  usage: python tools/traffic_json.py X.ncu-rep CELLS [mode]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import build
rep, cells = sys.argv[1], int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "gather"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
def to_bytes(v, u): return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
def to_ms(v, u): return float(v) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}[u]
kern = {}
for r in rows[2:]:
    name = r[col("Kernel Name")].split("<")[0].replace("void ", "").replace("nsb::", "")
    b = to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")]) + to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
    t = to_ms(r[col("gpu__time_duration.sum")], units[col("gpu__time_duration.sum")])
    k = kern.setdefault(name, {"dram_bytes": 0.0, "ms": 0.0, "launches": 0})
    k["dram_bytes"] += b; k["ms"] += t; k["launches"] += 1
n_elem = cells ** 3
tot_b = sum(k["dram_bytes"] / k["launches"] for k in kern.values())
tot_t = sum(k["ms"] / k["launches"] for k in kern.values())
tp = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(tp)) if os.path.exists(tp) else {}
if tj.get("sources_digest") != build._sources_digest():
    tj = {"bytes_per_element": {}}
tj["sources_digest"] = build._sources_digest()
tj["bytes_per_element"][mode] = tot_b / n_elem
tj["kernel_share_ncu"] = {n: round(k["ms"] / k["launches"] / tot_t, 3) for n, k in kern.items()}
tj["kernels"] = {n: {"dram_GB_per_pass": k["dram_bytes"] / k["launches"] / 1e9, "ms_under_ncu": k["ms"] / k["launches"]} for n, k in kern.items()}
tj["note"] = "dram__bytes_read.sum + dram__bytes_write.sum of every kernel of one pass, ncu --set full at hex %d^3 (%s), scaled per element" % (cells, os.path.basename(rep))
json.dump(tj, open(tp, "w"), indent=1)
print(json.dumps(tj, indent=1))
