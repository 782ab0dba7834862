"""Turns the files tools/gpu_profile_round.sh brought back (gpurun_out/<tag>_*) into the tracked summaries under profiles/:
bench line, launch list + shares, ncu full summary (key metrics, stall sites, shared-memory wavefronts, instruction segments), DRAM
traffic, SASS mnemonic counts of the default instantiation.   usage: python tools/make_profile_summary.py <tag> "<one-line description>" """
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, desc = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
POINTS, GROUPS = 205542400, 3211600  # the 592-sample config-3 batch of tools/prof_fused.py (PROF_UNIQUE=8 PROF_REPS=74)

def tool(name, *args, head=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", name), *map(str, args)], capture_output=True, text=True).stdout
    return "\n".join(out.splitlines()[:head]) if head else out.rstrip("\n")

# bench line + launch list
open(os.path.join(P, "r2_bench_line.json"), "w").write(open(os.path.join(G, f"{tag}_bench.json")).read())
open(os.path.join(P, "r2_launches_bench.csv"), "w").write(open(os.path.join(G, f"{tag}_launches.csv")).read())
rows = list(csv.reader(l for l in open(os.path.join(G, f"{tag}_launches.csv")) if l.startswith('"')))
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    c, t = agg.get(r[ki], (0, 0.0)); agg[r[ki]] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
out = ["launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)",
       f"{'kernel':100s} {'count':>6s} {'total us':>11s} {'share':>7s}"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k[:100]:100s} {c:6d} {t:11.1f} {100 * t / tot:6.1f}%")
open(os.path.join(P, "r2_launch_share.txt"), "w").write("\n".join(out) + "\n")

# ncu full summary
raw, src = os.path.join(G, f"{tag}_full_raw.csv"), os.path.join(G, f"{tag}_full_src.csv")
txt = [f"ncu --set full --clock-control none, {desc}, 592-sample config-3 batch ({POINTS:,} points = {GROUPS:,} groups of 64 points), grid 148, "
       f"one launch (gpurun_out/{tag}_full.ncu-rep; tools/gpu_profile_round.sh, tools/make_profile_summary.py)", "",
       "== key metrics (ncu raw page)", tool("ncu_key.py", raw, POINTS), "",
       "== stall reasons (warp samples) and the top stall sites (ncu source page)", tool("ncu_stalls.py", src, 14, GROUPS), "",
       "== shared-memory wavefronts per 64 points by instruction", tool("ncu_wavefronts.py", src, POINTS, head=18), "",
       "== instruction segments per 64 points (>= 4 instructions)", tool("ncu_segments.py", src, GROUPS, 4)]
open(os.path.join(P, "r2_stream4_ncu_full.txt"), "w").write("\n".join(txt) + "\n")
r = list(csv.reader(open(raw))); h, v = r[0], r[2]
rd = float(v[h.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[r[1][h.index("dram__bytes_read.sum")]]
wr = float(v[h.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[r[1][h.index("dram__bytes_write.sum")]]
json.dump({"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "source": f"profiles/r2_stream4_ncu_full.txt (ncu --set full, 592-sample config-3 batch, {desc}: gpurun_out/{tag}_full.ncu-rep)",
           "algorithmic_bytes_per_launch": 4403509120}, open(os.path.join(P, "fused_traffic.json"), "w"), indent=1)

# SASS mnemonics of the default instantiation
obj = os.path.join(ROOT, "multimodal-scene-captioning_b200", "csrc", "stream4.o")
sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3msc14stream4_kernelILb1ELb1ELi2ELb1EEEvNS_9FusedArgsENS_11TableLayoutEPh", obj], capture_output=True, text=True).stdout
cnt = collections.Counter()
for line in sass.splitlines():
    parts = line.split()
    if len(parts) > 1 and parts[0].startswith("/*") and len(parts[0]) == 8:
        i = 1
        if parts[i].startswith("@"): i += 1
        cnt[parts[i].rstrip(";")] += 1
head = ["SASS mnemonics of stream4_kernel<FOV=true, FASTDIV=true, PPT=2, STD=true> (cuobjdump -sass -fun ... stream4.o; static instruction counts):",
        "bulk copy through the TMA unit (UBLKCP) completing on mbarriers (SYNCS), native shared-memory integer atomics (ATOMS), global reductions (REDG),",
        "the f64 transform chain (F2F + DFMA), packed f32 pairs (FADD2 / FMUL2 / FFMA2); no tensor-core instruction."]
open(os.path.join(P, "r2_sass_evidence.txt"), "w").write("\n".join(head + [f"{n} {m}" for m, n in cnt.most_common()]) + "\n")
print(open(os.path.join(P, "r2_launch_share.txt")).read())
