#!/usr/bin/env python
"""Per-phase instruction shares of render_kernel from an ncu report captured with --import-source on.

usage: python tools/ncu_phases.py gpurun_out/prof_k2.ncu-rep [miniworld_b200/libmwb.so]
Maps every SASS address to its source line through `nvdisasm -gi` of the library's cubin (which must be
the build that was profiled) and sums ncu's per-instruction counters by the phase the line belongs to.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else "miniworld_b200/libmwb.so"
kern = sys.argv[3] if len(sys.argv) > 3 else "_Z13render_kernelILi8ELi320ELi3ELb1E"
if os.environ.get("CSRC"):
    pass
root = os.environ.get("CSRC") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "miniworld_b200", "csrc", "")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout


def find(fn, pat):
    for i, l in enumerate(open(root + fn).read().split("\n"), 1):
        if pat in l:
            return i
    raise KeyError(pat)


rc, rk = "raster_core.cuh", "raster.cuh"
marks = {
    rc: sorted([(find(rc, "MWB_DEV Camera make_camera"), "A camera"), (find(rc, "struct HVert"), "B geometry/setup"),
                (find(rc, "MWB_DEV float edge_value"), "C2 exact (sample_key)"),
                (find(rc, "// ---- lazy visibility"), "C1 classify"),
                (find(rc, "MWB_DEV uint32_t sample_key"), "C2 exact (sample_key)"),
                (find(rc, "MWB_DEV float texel_f"), "D shading"), (find(rc, "MWB_DEV uint8_t to_unorm8"), "D unorm/depth"),
                (find(rc, "struct FrameMap"), "B geometry/setup"), (find(rc, "struct Segment"), "D seg lookup"),
                (find(rc, "MWB_DEV uint32_t key_id"), "D resolve bookkeeping")]),
    rk: sorted([(1, "A prologue/TMA"), (find(rk, "B. room + box"), "B compaction"),
                (find(rk, "visiting order of the block-resident"), "B sort"), (find(rk, "candidate lists: one THREAD"), "B2 tile lists"), (find(rk, "C/D. one warp"), "C0 tile loop/tri test"),
                (find(rk, "auto flush"), "C2 flush (sample-parallel exact)"),
                (find(rk, "auto enqueue"), "C2 enqueue"), (find(rk, "hot path: this half-tile"), "C0 listed path"), (find(rk, "generic path: the mesh lists"), "C0 generic path"),
                (find(rk, "Lazy pixels join"), "D resolve/store")]),
}


def phase(f, l):
    if f in marks:
        p = "?"
        for s, name in marks[f]:
            if l >= s:
                p = name
        return p
    if f == "hd.h":
        return "exact ops (f_mul/f_add/f_div)"
    if f == "libm_sincos.cuh":
        return "A camera"
    return f


start = dis.index("//--------------------- .text." + kern)
end = dis.find("//---------------------", start + 30)
sec = dis[start:end if end > 0 else len(dis)]
amap, pend, cur = {}, None, ("?", 0)
for ln in sec.split("\n"):
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm:
        if pend is None:
            pend = (mm.group(1).split("/")[-1], int(mm.group(2)))
        continue
    mi = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(\S.*?);", ln)
    if mi:
        if pend:
            cur = pend
        amap[int(mi.group(1), 16)] = cur
        pend = None
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if "Address" in r)
hdr = rows[h]
ia, ie, it, iss = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
base, tot, tots = None, 0, 0
agg = collections.defaultdict(lambda: [0, 0, 0])
lines = collections.defaultdict(lambda: [0, 0])
for r in rows[h + 1:]:
    try:
        a = int(r[ia], 16)
    except (ValueError, IndexError):
        continue
    base = a if base is None else base
    k = amap.get(a - base, ("?", 0))
    v = agg[phase(*k)]
    v[0] += int(r[ie]); v[1] += int(r[it]); v[2] += int(r[iss])
    lines[k][0] += int(r[ie]); lines[k][1] += int(r[iss])
    tot += int(r[ie]); tots += int(r[iss])
print("total warp-inst %.0f" % tot)
for p, v in sorted(agg.items()):
    if v[0] / tot > 0.003:
        print("%-40s inst %5.1f%%  samples %5.1f%%  lanes %4.1f" % (p, 100 * v[0] / tot, 100 * v[2] / tots, v[1] / max(1, v[0])))
if os.environ.get("TOP"):
    print("-- hottest source lines")
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:int(os.environ["TOP"])]:
        print("%-22s %5d  inst %5.2f%% samples %5.2f%%" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / tots))
