import csv, sys, re, collections, os
sass_all, fn, ncu_csv = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. offsets -> (file,line)
lines = open(sass_all).read().split("\n")
start = next(i for i,l in enumerate(lines) if l.startswith(".text."+fn+":"))
loc = {}; cur = ("?",0)
for l in lines[start+1:]:
    if l.startswith(".text.") or l.startswith("\t.section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', l)
    if m: loc[int(m.group(1),16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
data = [dict(zip(hdr,r)) for r in rows[2:] if len(r)==len(hdr)]
base = int(data[0]["Address"],16)
def num(x):
    try: return float(x)
    except: return 0.0
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for d in data:
    off = int(d["Address"],16)-base
    (fl, op) = loc.get(off, (("?",0),"?"))
    for k in ("Instructions Executed","# Samples","stall_long_sb","Thread Instructions Executed","stall_wait","stall_short_sb","stall_branch_resolving","stall_not_selected","stall_selected","stall_math","stall_mio","stall_lg","stall_membar","stall_no_inst","stall_dispatch"):
        v = num(d.get(k,0)); agg[fl][k]+=v; tot[k]+=v
print("kernel", fn[:60], "total warp inst %.3g samples %d"%(tot["Instructions Executed"], tot["# Samples"]))
print("stall mix:", {k:round(100*tot[k]/max(1,tot["# Samples"]),1) for k in tot if k.startswith("stall")})
srcs = {}
def src(fl):
    f,l = fl
    path = {"search_group.cuh":"mapad_b200/csrc/search_group.cuh","search_core.cuh":"mapad_b200/csrc/search_core.cuh","dev_index.cuh":"mapad_b200/csrc/dev_index.cuh","simt.cuh":"mapad_b200/csrc/simt.cuh","libm_emu.cuh":"mapad_b200/csrc/libm_emu.cuh"}.get(f)
    if not path: return ""
    if path not in srcs: srcs[path] = open(os.environ.get("SRC_ROOT","/root/repo/")+path).read().split("\n")
    return srcs[path][l-1].strip()[:100] if 0<l<=len(srcs[path]) else ""
key = sys.argv[5] if len(sys.argv) > 5 else "Instructions Executed"
for fl,c in sorted(agg.items(), key=lambda kv:-kv[1][key])[:topn]:
    print("%-17s %4d inst %5.2f%% (thr/inst %4.1f) samp %5.2f%% lsb %5.2f%% | %s"%(fl[0],fl[1],100*c["Instructions Executed"]/tot["Instructions Executed"], c["Thread Instructions Executed"]/max(1,c["Instructions Executed"]),100*c["# Samples"]/max(1,tot["# Samples"]),100*c["stall_long_sb"]/max(1,tot["# Samples"]),src(fl)))

# ---- buckets by function (line ranges of the current tree) ----
import bisect
def func_table(path):
    out=[]; 
    for i,l in enumerate(open(path).read().split("\n"),1):
        m=re.match(r'\s*(?:template <[^>]*>\s*)?(?:MAPAD_DEV|MAPAD_HD|MAPAD_DEV_NOINLINE|__device__ __forceinline__|static MAPAD_DEV|static __device__ __forceinline__)\s+[\w:<>\*&\s]+?\b(\w+)\s*\(', l)
        if m: out.append((i,m.group(1)))
    return out
tabs={f:func_table(os.environ.get("SRC_ROOT","/root/repo/")+"mapad_b200/csrc/"+f) for f in ("search_group.cuh","search_core.cuh","dev_index.cuh","simt.cuh","libm_emu.cuh")}
b=collections.Counter(); bs=collections.Counter()
for fl,c in agg.items():
    f,l=fl
    name="?"
    if f in tabs and tabs[f]:
        ls=[x[0] for x in tabs[f]]; k=bisect.bisect_right(ls,l)-1
        name=tabs[f][k][1] if k>=0 else "top"
    b[(f,name)]+=c["Instructions Executed"]; bs[(f,name)]+=c["# Samples"]
print("---- by function: inst% / samples%")
for k,v in sorted(b.items(), key=lambda kv:-kv[1])[:32]:
    print("  %-18s %-22s %5.1f%%  %5.1f%%"%(k[0],k[1],100*v/tot["Instructions Executed"],100*bs[k]/max(1,tot["# Samples"])))
