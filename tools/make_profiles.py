"""Turn the scratch outputs of tools/gpu_round.sh (gpurun_out/) into the tracked evidence under profiles/.
usage: python tools/make_profiles.py <tag>   (e.g. r1)"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

# ---- launch list
with open(os.path.join(G, "launches.csv")) as f:
    r = list(csv.reader([l for l in f if l.startswith('"')]))
h = r[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg, order = collections.OrderedDict(), []
for row in r[1:]:
    name = row[ki]
    short = re.sub(r"\(.*", "", name.replace("void ", ""))[:70]
    if "at::" in name: short = "torch:" + short[:50]
    t = float(row[vi]) / 1e6
    order.append((short, t)); a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
out = ["# ncu launch list of `python bench.py --steps 2 --warmup 1 --no_cpu_baseline --train_steps 0` (B200)",
       "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv (tools/gpu_round.sh); per-launch times are cold-cache and serialised",
       "# %d launches captured, %.2f ms total device time" % (len(order), tot), "", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, n, t, 100 * t / tot))
out += ["", "## featurisation launches in order (first 24)", ""]
for k, t in [o for o in order if o[0].startswith("idl::")][:24]:
    out.append("- %s: %.3f ms" % (k, t))
open(os.path.join(P, "launches_%s.md" % tag), "w").write("\n".join(out) + "\n")

# ---- full capture of the dominant kernel
rep = os.path.join(G, "prof_pc_%s.ncu-rep" % tag)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); h, u, row = r[0], r[1], r[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
keys += [k for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")]
lines = ["# ncu --set full --clock-control none: profiles_pc_kernel<STD_F32>, 4000 sequences x 10 kb, 51 variants (tools/gpu_round.sh)",
         "# algorithmic bytes of this launch: 4000 x (51 x 16384 + 2500) = 3.3523 GB", "", "| metric | unit | value |", "|---|---|---:|"]
val = {}
for k in keys:
    if k in h:
        i = h.index(k); lines.append("| %s | %s | %s |" % (k, u[i], row[i])); val[k] = (float(row[i]), u[i])
open(os.path.join(P, "ncu_pc_%s_summary.md" % tag), "w").write("\n".join(lines) + "\n")
st = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), rep, "profiles_pc.cuh", "1", "2000", "25"], capture_output=True, text=True).stdout
open(os.path.join(P, "ncu_pc_%s_stalls.txt" % tag), "w").write(st)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = val["dram__bytes_read.sum"][0] * scale[val["dram__bytes_read.sum"][1]]
wr = val["dram__bytes_write.sum"][0] * scale[val["dram__bytes_write.sum"][1]]
json.dump({"kernel": "profiles_pc_kernel<STD_F32>", "dram_bytes_per_sequence": (rd + wr) / 4000,
           "note": "ncu --set full --clock-control none, 4000 sequences x 10 kb x 51 variants (profiles/ncu_pc_%s_summary.md): dram__bytes_read.sum %.1f MB + "
                   "dram__bytes_write.sum %.4f GB per launch = %.1f kB per sequence (algorithmic 838.1 kB; the tail of the output is still dirty in L2 when the "
                   "kernel ends); scaled linearly to the bench's sequences per launch" % (tag, rd / 1e6, wr / 1e9, (rd + wr) / 4000 / 1e3)},
          open(os.path.join(P, "dominant_kernel_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "launches_%s.md" % tag)).read()[:1500])
