"""Turn the scratch outputs of scripts/r2_evidence.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python scripts/make_profile_summary_r2.py r02_v1"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for src, dst in (("bench_%s.json", "%s_bench.json"), ("bench_ref_%s.json", "%s_bench_reference_arm.json"),
                 ("tests_%s.log", "%s_gpu_tests.log"), ("launches_%s.csv", "%s_launches.csv")):
    if os.path.exists(os.path.join(G, src % R)):
        shutil.copy(os.path.join(G, src % R), os.path.join(P, dst % R))
bench = json.load(open(os.path.join(G, "bench_%s.json" % R)))
st = bench["stage_ms_per_step"]
# ---- launch list ----
rows = [r for r in csv.reader(open(os.path.join(G, "launches_%s.csv" % R))) if len(r) > 10]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
# window = the last forward of the run: from its k_actor_prep launch to the end of the list
body = rows[1:]
starts = [i for i, r in enumerate(body) if "k_actor_prep" in r[ik]]
body = body[starts[-1]:] if starts else body
agg = collections.OrderedDict()
for r in body:
    k = r[ik].split("(")[0].replace("mind::", "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, "%s_launches_summary.md" % R), "w") as f:
    f.write("# ncu launch list of ONE forward (B=256, 32x128, f16tc), build %s; cold-cache serialized times: compare shares\n\n" % R)
    f.write("command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 1 --warmup 1 --kernel-only\n")
    f.write("(window = the %d launches of the timed forward: from its k_actor_prep to k_bezier)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % len(body))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, a[0], a[1], 100 * a[1] / tot))
    fus = st.get("fusion_tc", 0) + st.get("fusion_tc_last", 0)
    f.write("\ntotal %.1f us over %d launches; bench stage split (CUDA events, same build, ms per 256-scene step): %s -> fused kernel %.0f%% of the step (%.2f of %.2f ms); "
            "launch-list share of k_rela_fusion_tc: %.0f%%\n"
            % (tot, sum(a[0] for a in agg.values()), json.dumps(st), 100 * fus / bench["ms_per_step"], fus, bench["ms_per_step"],
               100 * agg.get("tc::k_rela_fusion_tc", [0, 0.0])[1] / tot))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TSCALE = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def read(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    names, units, vals = rr[0], rr[1], rr[-1]
    return {n: (u, v) for n, u, v in zip(names, units, vals)}


def table(f, m):
    f.write("| metric | unit | value |\n|---|---|---|\n")
    for n in WANT:
        if n in m:
            f.write("| %s | %s | %s |\n" % (n, m[n][0], m[n][1]))


def traffic_time(m):
    rd = float(m["dram__bytes_read.sum"][1].replace(",", "")) * SCALE[m["dram__bytes_read.sum"][0]]
    wr = float(m["dram__bytes_write.sum"][1].replace(",", "")) * SCALE[m["dram__bytes_write.sum"][0]]
    t = float(m["gpu__time_duration.sum"][1].replace(",", "")) * TSCALE[m["gpu__time_duration.sum"][0]]
    return rd, wr, t


# ---- fused kernel ----
m = read(os.path.join(G, "prof_%s.ncu-rep" % R))
rd, wr, t = traffic_time(m)
json.dump({"kernel": "k_rela_fusion_tc", "config": "B=256, N=161, layer with edge update", "dram_bytes_per_launch": rd + wr,
           "source": "profiles/%s_fusion_tc_ncu_full.md" % R}, open(os.path.join(P, "fusion_tc_traffic.json"), "w"))
with open(os.path.join(P, "%s_fusion_tc_ncu_full.md" % R), "w") as f:
    f.write("# ncu --set full, k_rela_fusion_tc (build %s), layer with edge update, B=256 scenes 32x128 (N=161)\n\n" % R)
    f.write("command (scripts/r2_evidence.sh): ncu --set full --clock-control none --import-source on -k regex:k_rela_fusion_tc -s 6 -c 1 python bench.py --steps 1 --warmup 1 --kernel-only\n\n")
    table(f, m)
    f.write("\nalgorithmic bytes per launch: 256 x 161^2 x 128 x 2 B x 2 (read + write) = 3.398 GB; measured DRAM read + write = %.3f GB (%.2fx).\n" % ((rd + wr) / 1e9, (rd + wr) / 3.3975e9))
    lp = os.path.join(G, "prof_last_%s.ncu-rep" % R)
    if os.path.exists(lp):
        ml = read(lp)
        f.write("\n## last layer (no edge update; G1 of the next tile issued ahead of K|V, D1 alternating between two TMEM regions)\n\n")
        table(f, ml)
    f.write("\nSASS (cuobjdump -sass mind_b200/libmind_b200.so): UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), SYNCS (mbarrier), FFMA2 / FADD2 (packed fp32), BAR.ARV / BAR.SYNC (named-barrier hand-offs), ELECT.\n")
# ---- the other kernels of the step ----
others = [("gemm_actor", "tcg::k_tc_gemm, an ActorNet conv GEMM (3-term fp16 split) with the GroupNorm epilogue (group 3, N = 256)"), ("gemm_lane", "tcg::k_tc_gemm, a LaneNet 128x128 linear over 330k rows (3-term)"),
          ("lane_chain", "lane::k_lane_net_tc (the whole LaneNet on chip, 33k polylines)"),
          ("node_chain", "node::k_node_chain_tc (out-proj + LN2 + FFN + LN3 + next layer's S|T|q, 41k token rows)"),
          ("edge_init", "k_edge_init_ch: edge init (5 -> 128 + closed-form LN + ReLU, writes the fp16 edge stream: 1.70 GB)"), ("gn_apply", "tcg::k_gn_apply (lateral GroupNorm + FPN top-down step -> fp16 hi/lo, finest level)"),
          ("node_fields", "k_node_fields (cost fields of the trajectory-tree optimiser, fp64)")]
with open(os.path.join(P, "%s_other_kernels_ncu_full.md" % R), "w") as f:
    f.write("# ncu --set full of the kernels next to the fused layer (build %s, B=256 benchmark step; cost fields: demo_2 trees)\n\n" % R)
    f.write("commands: scripts/r2_evidence.sh (one launch each, `-k regex:<kernel> -s <skip> -c 1`).  HBM peak measured on this pool: see MEASURED_PEAKS.json.\n")
    for tag, title in others:
        rp = os.path.join(G, "prof_%s_%s.ncu-rep" % (tag, R))
        if not os.path.exists(rp):
            continue
        mm = read(rp)
        rd, wr, t = traffic_time(mm)
        f.write("\n## %s\n\n" % title)
        table(f, mm)
        f.write("\nDRAM read %.1f MB + write %.1f MB in %.1f us = %.2f TB/s\n" % (rd / 1e6, wr / 1e6, t * 1e6, (rd + wr) / t / 1e12))
print("profiles written for", R)
