"""Builds the tracked round-2 profile artefacts from what the gpurun calls of the round brought back (gpurun_out/ is scratch):

    profiles/r02_launches_hg38.csv     ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tb_ -c 3000 --csv
                                       python bench.py --steps 1 --warmup 1 --no-cpu-baseline                  (gpurun_out/r02_launches_hg38.csv)
    profiles/r02_ncu_chr50mb.json      counters of one `ncu --set full --clock-control none --import-source on` launch per kernel,
                                       python bench.py --workload chr50mb --steps 1 --warmup 0 --no-cpu-baseline (gpurun_out/r2_prof{1,2,3,4}.ncu-rep)
    profiles/r02_roofs.json            binding unit per stage (read by bench.py for the `bound` labels of its `stages`)
    profiles/r02_sass_grep.txt         cuobjdump -sass of the shipped library: tcgen05 / TMA / spill mnemonics per kernel
    profiles/r02_bench_*.json          the bench lines quoted in DESIGN.md / profiles/r02_summary.md (not under a profiler)

usage: python profiles/make_summary_r02.py            (needs ncu and cuobjdump on PATH)"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = {"gpu__time_duration.sum": "time_ms", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_smem_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts", "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
        "launch__grid_size": "grid", "launch__block_size": "block", "smsp__inst_executed.sum": "warp_inst", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct"}
SCALE = {"Gbyte": 1000.0, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "second": 1000.0, "msecond": 1.0, "ms": 1.0, "usecond": 1e-3, "us": 1e-3, "nsecond": 1e-6, "ns": 1e-6}


def counters():
    out = collections.OrderedDict()
    for rep in ("r2_prof4", "r2_prof2", "r2_prof3", "r2_prof1"):      # prof4: the kernels that changed last (K3c, K4, signal), captured on the final code
        path = os.path.join(OUT, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
        hdr, units = rows[0], rows[1]
        stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h]
        seen = collections.Counter()
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            seen[name] += 1
            key = name if seen[name] == 1 else "%s [launch %d]" % (name, seen[name])
            if key in out:
                continue
            d = {"capture": rep}
            for h, u, v in zip(hdr, units, vals):
                if h in WANT and v not in ("", "n/a"):
                    d[WANT[h]] = round(float(v.replace(",", "")) * SCALE.get(u, 1.0), 3)
            tot = sum(float(vals[hdr.index(h)] or 0) for h in stall) or 1.0
            top = sorted(stall, key=lambda h: -float(vals[hdr.index(h)] or 0))[:5]
            d["top_stalls_pct"] = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): round(100 * float(vals[hdr.index(h)] or 0) / tot, 1) for h in top}
            out[key] = d
    out.pop("tb_footprint_kernel<0>", None)          # first form of K4: no longer on the default path
    # prof4 also saw the chunked launches of the streaming (e2e) pass: keep the whole-workload launch of every kernel
    best = {}
    for k, d in out.items():
        if d["capture"] == "r2_prof4":
            base = k.split(" [launch")[0]
            if base not in best or d["time_ms"] > out[best[base]]["time_ms"]:
                best[base] = k
    for k in [k for k, d in out.items() if d["capture"] == "r2_prof4"]:
        base = k.split(" [launch")[0]
        d = out.pop(k)
        if best[base] == k:
            out[base] = d
    return out


def launches():
    src = os.path.join(OUT, "r02_launches_hg38.csv")
    shutil.copy(src, os.path.join(PROF, "r02_launches_hg38.csv"))
    tot = collections.OrderedDict()
    with open(src) as fh:
        rows = [r for r in csv.reader(l for l in fh if l.startswith('"'))]
    hdr = rows[0]
    for r in rows[1:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        ms = float(r[hdr.index("Metric Value")].replace(",", "")) * SCALE.get(r[hdr.index("Metric Unit")], 1.0)
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ms
    return tot


def sass_grep():
    lib = os.path.join(ROOT, "tobias_b200", "csrc", "libtobias_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    pats = ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "LDS", "ATOMS", "LDL", "STL", "DFMA", "DMUL", "DADD", "BAR.SYNC")
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            per[cur] = collections.Counter()
        elif cur:
            for p in pats:
                if " " + p in line:
                    per[cur][p] += 1
    with open(os.path.join(PROF, "r02_sass_grep.txt"), "w") as fh:
        fh.write("cuobjdump -sass tobias_b200/csrc/libtobias_b200.so | per kernel: count of lines holding the mnemonic (sm_100a build of this commit)\n")
        fh.write("LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory); UTMALDG / UTMASTG = TMA; UTC*MMA = tcgen05.mma (none: no dense contraction on this path)\n\n")
        for k, c in per.items():
            if k.startswith("_Z") and c:
                fh.write("%-110s %s\n" % (k[:110], "  ".join("%s=%d" % (p, c[p]) for p in pats if c[p])))
    return per


def main():
    cnt = counters()
    json.dump(cnt, open(os.path.join(PROF, "r02_ncu_chr50mb.json"), "w"), indent=1)
    lt = launches()
    sass_grep()
    for f, dst in (("r2_bench_j.json", "r02_bench_hg38_n1.json"), ("r2_bench_n2.json", "r02_bench_hg38_n2.json"), ("r2_bench_n8.json", "r02_bench_hg38_n8.json"),
                   ("r2_fp_genome.json", "r02_bench_fp_genome.json"), ("r2_batch24_n1.json", "r02_bench_batch24_n1.json"),
                   ("r2_batch24_n8.json", "r02_bench_batch24_n8.json"), ("r2_cli_chr50mb.json", "r02_cli_chr50mb.json"),
                   ("r2_ref.json", "r02_bench_hg38_reference_arm.json")):
        p = os.path.join(OUT, f)
        if os.path.exists(p):
            lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
            if lines:
                open(os.path.join(PROF, dst), "w").write(lines[-1] + "\n")
    g = lambda k, f: cnt.get(k, {}).get(f)
    roofs = {
        "_source": "profiles/r02_ncu_chr50mb.json (ncu --set full, one launch per kernel, chr50mb workload, B200); traffic of K3a at hg38 size from the round-1 capture (kernel unchanged)",
        "K1_pileup": {"bound": "L2 atomics / integer issue (ALU pipe)", "bound_util_pct": g("tb_pileup_kernel<0>", "alu_pipe_pct")},
        "K2_counts": {"bound": "sites kernel: barriers + load latency over ~1400 small regions per block; histogram kernel: shared-memory atomics (L1 pipe)",
                      "bound_util_pct": g("tb_k2_hist_smem_kernel<13>", "l1_smem_throughput_pct")},
        "K3a_score": {"bound": "shared-memory data pipe (table rows + shuffles) together with instruction issue; half of the rows come from tensor memory",
                      "bound_util_pct": g("tb_score_dwm2_kernel<1>", "l1_smem_throughput_pct"), "dram_bytes_per_launch_hg38": 4058331807.9999995},
        "K3b_fit": {"bound": "latency of the window-data loads (long-scoreboard stalls 39-62 %); issue 57 -> 32 % as the work lists thin out",
                    "bound_util_pct": g("tb_fitg_round_kernel<0, 1>", "issue_active_pct")},
        "K3c_correct": {"bound": "instruction issue (10-row mean, verification counts, block scans)", "bound_util_pct": g("tb_correct_kernel<2>", "issue_active_pct")},
        "K4_footprint": {"bound": "ALU pipe (three-input maxima of the max-plus register tile) + shared-memory pipe at 14 warps per SM; the load / scan / staging phases of a tile are latency bound",
                         "bound_util_pct": g("tb_footprint_tile_kernel<0, 16>", "alu_pipe_pct")},
    }
    json.dump(roofs, open(os.path.join(PROF, "r02_roofs.json"), "w"), indent=1)
    tot_ms = sum(v[1] for v in lt.values())
    with open(os.path.join(PROF, "r02_launch_shares.md"), "w") as fh:
        fh.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, ms) in sorted(lt.items(), key=lambda kv: -kv[1][1]):
            fh.write("| %s | %d | %.3f | %.1f%% |\n" % (k, n, ms, 100 * ms / tot_ms))
    print("kernels with counters:", len(cnt), "| launch list kernels:", len(lt), "| total ms under ncu: %.1f" % tot_ms)


if __name__ == "__main__":
    main()
