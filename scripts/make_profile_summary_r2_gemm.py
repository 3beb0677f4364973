"""profiles/<tag>_gemm_epilogues_ncu_full.md from the three captures of scripts/r2_ncu_gemm.sh, plus the launch-list summary of
scripts/r2_final.sh.   usage: python scripts/make_profile_summary_r2_gemm.py r02_v4"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def read(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    return {n: (u, v) for n, u, v in zip(rr[0], rr[1], rr[-1])}


caps = [("lat0", "lateral 0 conv (K = 96 -> 128, N = 128, 8192 actors x 48 steps, r_in = 64: 2 actors per tile) with GroupNorm + FPN top-down step in the epilogue, output = padded fp16 (hi, lo) operand of the output block"),
        ("g0c2", "group-0 block-0 conv2 (two steps per GEMM row: K = 128, N = 64, r_in = 32: 4 actors per tile) with bn2 + GroupNorm of the down-sampled shortcut + ReLU in the epilogue"),
        ("reg0", "decoder reg.0 linear (49,152 actor x mode rows, K = N = 128, 3-term), plain epilogue")]
with open(os.path.join(P, "%s_gemm_epilogues_ncu_full.md" % R), "w") as f:
    f.write("# ncu --set full of three k_tc_gemm launches of build %s (scripts/r2_ncu_gemm.sh; B = 256 benchmark step)\n" % R)
    for tag, title in caps:
        rp = os.path.join(G, "prof_gemm_%s_%s.ncu-rep" % (tag, R))
        if not os.path.exists(rp):
            continue
        m = read(rp)
        f.write("\n## %s\n\n| metric | unit | value |\n|---|---|---|\n" % title)
        for n in WANT:
            if n in m:
                f.write("| %s | %s | %s |\n" % (n, m[n][0], m[n][1]))
# launch list
lp = os.path.join(G, "launches_%s.csv" % R)
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
    h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    body = rows[1:]
    starts = [i for i, r in enumerate(body) if "k_actor_prep" in r[ik]]
    body = body[starts[-1]:] if starts else body
    agg = collections.OrderedDict()
    for r in body:
        k = r[ik].split("(")[0].replace("mind::", "")
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", "")) / 1e3
    tot = sum(a[1] for a in agg.values())
    bench = json.load(open(os.path.join(G, "bench_%s.json" % R)))
    with open(os.path.join(P, "%s_launches_summary.md" % R), "w") as f:
        f.write("# ncu launch list of ONE forward (B=256, 32x128, f16tc), build %s; cold-cache serialized times: compare shares\n\n" % R)
        f.write("command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 1 --warmup 1 --kernel-only\n")
        f.write("(window = the %d launches of the timed forward)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % len(body))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, a[0], a[1], 100 * a[1] / tot))
        f.write("\ntotal %.1f us; bench stage split of the same build (CUDA events, ms per 256-scene step): %s\n" % (tot, json.dumps(bench["stage_ms_per_step"])))
print("ok")
