"""profiles/traffic.json is bound to the CUDA sources by build._sources_digest() (every file under csrc/), so that bench.py never
reports an ncu capture taken with other kernels.  When a change touches csrc/ WITHOUT changing the machine code of the measured
pass (a new element type in its own translation unit, host-side dispatch), the capture is still the capture of these kernels.
This tool proves that instead of assuming it: it disassembles the translation unit that holds both kernels of the measured pass
(build/fv1_e3.o: fv1_flux_kernel + fv1_rows_owner_kernel for hexahedra, and their launch code) with `cuobjdump -sass`, and re-stamps
traffic.json with the new sources digest ONLY if the SASS text and the launch-code sources are byte-identical to the ones recorded
with the capture.

  python tools/restamp_traffic.py --record    # after an ncu capture: store the SASS / launch-code hashes next to the capture
  python tools/restamp_traffic.py             # after a rebuild: verify, then re-stamp (exit 1 and no change if anything differs)
"""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import build

TP = os.path.join(ROOT, "profiles", "traffic.json")
UNIT_OBJ = os.path.join(build.OBJ, "fv1_e3.o")
LAUNCH_SOURCES = ["fv1_inst.cu", "ns_launch.h"]          # grid / block / shared-memory sizes of the two kernels


def sass_sha():
    build.build_cuda()
    out = subprocess.run(["cuobjdump", "-sass", UNIT_OBJ], capture_output=True, check=True).stdout
    return hashlib.sha256(out).hexdigest()


def launch_sha():
    h = hashlib.sha256()
    for f in LAUNCH_SOURCES:
        h.update(open(os.path.join(build.CSRC, f), "rb").read())
    h.update(" ".join(build.NVCC_FLAGS).encode())
    return h.hexdigest()


def main():
    tj = json.load(open(TP))
    now = {"unit": "build/fv1_e3.o (fv1_inst.cu -DNSB_ELEM=3)", "sass_sha256": sass_sha(), "launch_sources_sha256": launch_sha()}
    if "--record" in sys.argv:
        if tj.get("sources_digest") != build._sources_digest():
            sys.exit("traffic.json does not belong to the present sources: capture first (tools/traffic_json.py)")
        tj["machine_code"] = now
        json.dump(tj, open(TP, "w"), indent=1)
        print("recorded", now)
        return
    rec = tj.get("machine_code")
    if not rec:
        sys.exit("traffic.json carries no machine-code record (run with --record right after the capture)")
    if rec["sass_sha256"] != now["sass_sha256"] or rec["launch_sources_sha256"] != now["launch_sources_sha256"]:
        sys.exit("the measured kernels changed (SASS or launch code differ): re-capture, not re-stamp")
    new = build._sources_digest()
    if tj["sources_digest"] != new:
        tj.setdefault("restamped", []).append({"from": tj["sources_digest"], "to": new,
                                               "why": "csrc/ changed outside the measured pass; cuobjdump -sass of %s and the launch sources are byte-identical to the captured build" % rec["unit"]})
        tj["sources_digest"] = new
        json.dump(tj, open(TP, "w"), indent=1)
    print("traffic.json valid for sources digest", new)


if __name__ == "__main__":
    main()
