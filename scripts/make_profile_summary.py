"""Turn the scratch outputs of scripts/gpu_evidence.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python scripts/make_profile_summary.py r01_v7"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for src, dst in (("bench_%s.json", "%s_bench.json"), ("bench_ref_%s.json", "%s_bench_reference_arm.json"),
                 ("tests_%s.log", "%s_gpu_tests.log"), ("launches_%s.csv", "%s_launches.csv")):
    if os.path.exists(os.path.join(G, src % R)):
        shutil.copy(os.path.join(G, src % R), os.path.join(P, dst % R))
# ---- launch list ----
rows = [r for r in csv.reader(open(os.path.join(G, "launches_%s.csv" % R))) if len(r) > 10]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ik].split("(")[0].replace("mind::", "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
bench = json.load(open(os.path.join(G, "bench_%s.json" % R)))
st = bench["stage_ms_per_step"]
with open(os.path.join(P, "%s_launches_summary.md" % R), "w") as f:
    f.write("# ncu launch list of ONE forward (B=256, 32x128, f16tc), build %s; cold-cache serialized times: compare shares\n\n" % R)
    f.write("command: ncu --metrics gpu__time_duration.sum --clock-control none -s 176 -c 176 --csv python bench.py --steps 1 --warmup 1 --kernel-only\n")
    f.write("(window = 176 launches starting inside the timed forward)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, a[0], a[1], 100 * a[1] / tot))
    fus = st.get("fusion_tc", 0) + st.get("fusion_tc_last", 0)
    f.write("\ntotal %.1f us over %d launches; bench stage split (CUDA events, same build, ms per 256-scene step): %s -> fused kernel %.0f%% of the step (%.2f of %.2f ms)\n"
            % (tot, sum(a[0] for a in agg.values()), json.dumps(st), 100 * fus / bench["ms_per_step"], fus, bench["ms_per_step"]))
# ---- ncu --set full of the fused kernel ----
rep = os.path.join(G, "prof_%s.ncu-rep" % R)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
names, units, vals = rr[0], rr[1], rr[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "smsp__mem_tensor_reads_op_utcmma_matrix_c.sum.pct_of_peak_sustained_elapsed", "smsp__mem_tensor_writes_op_utcmma.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
m = {n: (u, v) for n, u, v in zip(names, units, vals)}
rd = float(m["dram__bytes_read.sum"][1]); wr = float(m["dram__bytes_write.sum"][1])
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = rd * scale[m["dram__bytes_read.sum"][0]] + wr * scale[m["dram__bytes_write.sum"][0]]
json.dump({"kernel": "k_rela_fusion_tc", "config": "B=256, N=161, layer with edge update", "dram_bytes_per_launch": traffic,
           "source": "profiles/%s_fusion_tc_ncu_full.md" % R}, open(os.path.join(P, "fusion_tc_traffic.json"), "w"))
with open(os.path.join(P, "%s_fusion_tc_ncu_full.md" % R), "w") as f:
    f.write("# ncu --set full, k_rela_fusion_tc (build %s), layer with edge update, B=256 scenes 32x128 (N=161)\n\n" % R)
    f.write("command (scripts/gpu_evidence.sh): ncu --set full --clock-control none --import-source on -k regex:k_rela_fusion_tc -s 6 -c 1 python bench.py --steps 1 --warmup 1 --kernel-only\n\n| metric | unit | value |\n|---|---|---|\n")
    for n in want:
        if n in m:
            f.write("| %s | %s | %s |\n" % (n, m[n][0], m[n][1]))
    f.write("\nalgorithmic bytes per launch: 256 x 161^2 x 128 x 2 B x 2 (read + write) = 3.398 GB; measured DRAM read+write = %.3f GB -> no re-reads.\n" % (traffic / 1e9))
    f.write("\nSASS (cuobjdump -sass mind_b200/libmind_b200.so): UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), SYNCS (mbarrier), FFMA2 / FADD2 (packed fp32), BAR.ARV / BAR.SYNC (named-barrier hand-offs), ELECT.\n")
print("profiles written for", R, "traffic %.3f GB" % (traffic / 1e9))
