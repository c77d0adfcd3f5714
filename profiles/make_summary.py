"""Builds profiles/r01_summary.md, r01_ncu_full_chr50mb.json and r01_roofs.json from what a gpurun profiling call brought back:

    gpurun_out/bench_final.json          python bench.py --steps 3 --warmup 3 --no-cpu-baseline          (not under a profiler)
    gpurun_out/r01_launches_hg38.csv     ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tb_ -c 3000 --csv
                                         python bench.py --steps 1 --warmup 1 --no-cpu-baseline
    gpurun_out/r01_full_{a,b,c}.ncu-rep  ncu --set full --clock-control none --import-source on, one launch per kernel, chr50mb workload
                                         (c: the fit rounds with --lm-exact)

usage: python profiles/make_summary.py            (needs ncu on PATH to read the .ncu-rep files)
The hand-written part (section 4 and the history paragraph) lives in this script."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

WANT = {"gpu__time_duration.sum": "time_ms", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_pipe_pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
        "launch__grid_size": "grid", "launch__block_size": "block", "smsp__inst_executed.sum": "warp_inst"}
SCALE = {"Gbyte": 1000.0, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "second": 1000.0, "msecond": 1.0, "ms": 1.0, "usecond": 1e-3, "us": 1e-3,
         "nsecond": 1e-6, "ns": 1e-6}


def full_counters():
    out = collections.OrderedDict()
    for tag in "abc":
        rep = os.path.join(OUT, "r01_full_%s.ncu-rep" % tag)
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            if tag == "c":
                name += " [--lm-exact]"
            if name in out:
                continue
            d = {}
            for h, u, v in zip(hdr, units, vals):
                if h in WANT and v not in ("", "n/a"):
                    d[WANT[h]] = round(float(v.replace(",", "")) * SCALE.get(u, 1.0), 3)
            out[name] = d
    return out


def stage_of(kernel):
    for pre, st in (("tb_fitg", "K3b_fit"), ("tb_score", "K3a_score"), ("tb_correct_kernel", "K3c_correct"), ("tb_k2", "K2_counts"),
                    ("tb_footprint", "K4_footprint"), ("tb_corrected_to_signal", "K4_footprint"), ("tb_pileup_kernel<0>", "K1_pileup")):
        if kernel.startswith(pre):
            return st
    return None


HISTORY = """History of the same line this round (ms per step, hg38, 1 GPU): 1377 (warp-per-window fits with tree sums, sum-form DWM scoring)
-> 1082 (thread-per-window exact fits) -> 883 (product-form scoring) -> 754 (K2 site list) -> 614 (K3c without fp64 shared atomics, K3a trinucleotide
group + butterfly reduction) -> 604 (K4) -> 478 (one-pass outer iterations for DWM fits) -> 444 (K3a rows from tensor memory) -> 427 (K3a compile-time
grouping) -> 367 (K3a ratio form, lane = column) -> 338 (plane-major traversal in the one-pass fits) -> 317 (K3c one tile per region, K2 joint histograms
in L2) -> 309 (K3c rounding in the fp64 pipe, fit records in shared memory) -> 291 (K2 joint histograms of 2-base chunks in shared memory) -> 288 (K4
pre-scaled flank sums, unconditional width loop) -> 283 (K3c verification weights through per-base position lists) -> 276 (K3b one pass per
accepted step) -> 261 (K3b window statistics taken in the init pass, no divergent finish pass) -> 259 (K3b loads two points ahead) -> 254 (K3a
epilogue: branch-free column selection, half the exponent splits, no gather shuffles) -> 252.7 (K3b one-pass rounds at five blocks per SM) -> 247.9 (K3c: quotients of the 10-row mean through per-window reciprocals, correctly rounded) -> 243.8 (K3b: first outer-iteration head inside the init pass).  e2e: 1496 -> ~430 -> 396 -> 355 (genome upload on the copy
stream, overlapped with the read counts and the training counts) -> 321 ms; with 3.94 GB of H2D (76 ms at 52 GB/s, only ~20 ms of kernels can
overlap it because the model needs the counts of the whole genome) and 2.10 GB of D2H (40 ms, hidden behind the footprint kernels except for the
last 10 ms) the e2e line sits ~5 ms above what the PCIe link allows for these kernel times.
"""

READINGS = """## 4. Readings: which unit binds each kernel

DRAM traffic of every kernel is within ~2x of its algorithmic bytes and none of them is HBM bound; the binding unit (from the table above) is:

| kernel | binding unit | what was done about it this round | what is left |
|---|---|---|---|
| K1 `tb_pileup_kernel` | L2 atomics / integer issue (scatter), 1.1 ms for 49M records + 124M positions | -- (0.3 % of the step) | -- |
| K2 `tb_k2_scatter` + `tb_k2_weights` + `tb_k2_hist_smem` + `tb_k2_fold` | scatter / weights: DRAM latency of sector-sized read-modify-writes on the 24 GB multiplicity slab (12.6 of the 16 ms); histogram: shared-memory atomics | site list instead of a slab scan (174 -> 40 ms); joint histograms of chunk pairs instead of one atomic per position pair: 21 L2 atomics per k-mer (32 ms, bound by the L2 atomic rate), then 78 shared-memory atomics per k-mer on 16 x 16 tables (histogram kernel 25 -> 3.2 ms, K2 16 ms) | the slab passes: a sort-based dedup of (region, strand, cut site) would avoid the 24 GB slab |
| K3a `tb_score_dwm2_kernel` | **shared-memory data pipe (76 %: table rows 30 + shuffles 10 wavefronts per window) and instruction issue (70 %)** together | product form (no exp2), pair / trinucleotide row groups (25 -> 10 rows), ratio form (3 values per column, lane = column), 5 groups read from **tensor memory** through tcgen05.ld, butterfly reduction, compile-time grouping (446 -> 98 ms); branch-free column selection, raw level-16 products split afterwards, idle lanes as exact constants, no gather shuffles: 631 -> 527 instructions per four windows (-> 91 ms) | both pipes are near their limit with fp64 at 33 %: fewer shared-memory rows (a fourth table store does not exist on the SM), k-mers from uniform loads instead of 2 shuffles per window |
| K3b `tb_fitg_round_kernel` | **instruction issue + L1 miss latency** (thread-sequential LM arithmetic; fp64 and integer instructions issue at half rate) | one thread per window, rounds over compacted lists, exact MINPACK order without m-vector storage (2854 -> 227 ms exact); one-pass outer iterations, plane-major traversal, tail finished in one launch for DWM (-> ~80 ms); one pass per accepted step, window statistics in the init pass, loads two points ahead (-> 59.8); five resident blocks per SM (-> 58.2); first outer-iteration head inside the init pass (-> 54.0) | a quarter of the stall samples sit on the first use of the look-ahead loads; software prefetch measured slower both ways tried (64.5 / 61.2 ms) |
| K3c `tb_correct_kernel` | instruction issue, spread over four phases (was: conversion unit, two conversions per term of the float-accumulator rolling sums) | no fp64 shared atomics, integer prefix / sparse sums where exact, zero addends skipped (103 -> 68 ms); one tile per 824-bp region (-> 57); float rounding in the fp64 pipe by adding and subtracting 1.5 * 2^(E+29) -- exact, no conversions (-> 56); fit records staged in shared memory, no per-row modulo, sign-split verification adds (-> 52.5); verification weights through per-base position lists (-> 47.4); quotients of the 10-row mean through per-window reciprocals + two fused corrections, correctly rounded (-> 43.3) | three dense sequential rolling sums of 100 terms per position -- the semantics are the reference's |
| K4 `tb_footprint_kernel` | issue (fp32 + shared loads), 651 width combinations per position | fitted block size, half the barriers, suffix-maximum spreading (49 -> 30 ms); flank sums divided by 2 x width once per element and an unconditional, padded width loop: LDS + FADD + FMNMX per combination instead of five instructions (-> 26 ms) | two footprint starts per thread with 64-bit shared loads (tried: 128-bit loads of the 32 values from four shifted copies of the flank sums -- 33.7 ms instead of 26.1: the 32 result registers in flight next to the 32 accumulators leave no room at 64 registers per thread) |

Bytes-based roofline fractions (algorithmic bytes of SURVEY 8d / measured time / 6554.9 GB/s measured copy bandwidth) are therefore small for
K2-K4 (0.1 - 0.6 %) and are reported in every bench line (`stages[*].alg_GBps`, `roofline`) because the contract asks for them; K1 reaches 16 %.
"""


def main():
    os.makedirs(PROF, exist_ok=True)
    full = full_counters()
    json.dump(full, open(os.path.join(PROF, "r01_ncu_full_chr50mb.json"), "w"), indent=1)
    shutil.copyfile(os.path.join(OUT, "r01_launches_hg38.csv"), os.path.join(PROF, "r01_launches_hg38.csv"))
    rows = [r for r in csv.reader(open(os.path.join(PROF, "r01_launches_hg38.csv"))) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        k = r[ik].split("(")[0].replace("void ", "")
        tot[k] = tot.get(k, 0.0) + float(r[iv].replace(",", "")) / 1e6
        cnt[k] += 1
    total = sum(tot.values())
    bench = json.loads(open(os.path.join(OUT, "bench_final.json")).read().strip().splitlines()[-1])
    st, ms = bench["stages"], bench["ms_per_step"]
    ssum = sum(v["ms"] for v in st.values())
    L = ["# Round 1 profile summary (B200, sm_100a)\n",
         "All numbers in this file were measured on the GPU box through `gpurun`; raw artefacts: `profiles/r01_launches_hg38.csv` (ncu launch list),\n"
         "`profiles/r01_ncu_full_chr50mb.json` (counters of one `ncu --set full` launch per kernel); `profiles/make_summary.py` builds this file from them.\n",
         "## 1. Bench line the profiles belong to (not under a profiler)\n",
         "`python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (hg38 workload: 3.1 Gb, 150k peaks, 49.3M records, 105.1 Mb of output regions), 1 GPU:\n",
         "| quantity | value |\n|---|---|",
         "| `value` (inputs resident) | %.4g bp/s, %.1f ms per step (the six stages %.1f ms; the rest: read-count kernels 3.3 ms, footprint input 3.5 ms, host work between launches; one run in five shows 30 ms more of it) |"
         % (bench["value"], ms, ssum),
         "| `e2e` (host buffers, H2D 3.94 GB + D2H 2.10 GB per step inside the timed region; 76 ms of H2D at 52 GB/s is the floor under it) | %.4g bp/s, %.1f ms per step |"
         % (bench["e2e"]["value"], bench["e2e"]["ms_per_step"]),
         "| `gpu_launches` in the timed region (3 steps) | %d |" % bench["gpu_launches"],
         "| clocks during the timed region | %s MHz of %s MHz, throttle reasons %s |"
         % (bench["clocks"]["sm_mhz"], bench["clocks"]["sm_max_mhz"], bench["clocks"]["reasons"]),
         "| reference on the box's 16 host cores (`--impl reference`; `cpu_baseline` of the default bench line) | 1.0e5 - 1.1e5 bp/s |",
         "\nStage times (CUDA events on the library stream, mean over the timed steps; DWM default = one-pass outer iterations in the window fits):\n",
         "| stage | ms | share of the kernels | algorithmic GB/s (SURVEY 8d bytes / time) |\n|---|---|---|---|"]
    for k, v in st.items():
        L.append("| %s | %.2f | %.1f%% | %s |" % (k, v["ms"], 100 * v["ms"] / ssum, v["alg_GBps"]))
    L.append("\nWith `--lm-exact` (MINPACK-exact window fits, the PWM default): K3b ~215 ms.  Whole-contig footprint scan of BASELINE configs[3] (fp 20-50, flank 100, one 120 Mb region, `tests/test_properties_gpu.py`): 14.1 ms = 8.5e9 bp/s.\n")
    L.append(HISTORY)
    L.append("## 2. ncu launch list of the bench command\n")
    L.append("`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tb_ -c 3000 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline`\n"
             "(5 passes of the hot path: 1 first e2e pass, 1 + 1 resident, 1 + 1 e2e; durations are cold-cache and serialised -- compare SHARES):\n")
    L.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        L.append("| %s | %d | %.3f | %.1f%% |" % (k, cnt[k], v, 100 * v / total))
    g = collections.Counter()
    for k, v in tot.items():
        if stage_of(k):
            g[stage_of(k)] += v
    gt = sum(g.values())
    L.append("\nShares per stage (of the six stages), launch list vs bench (CUDA events): "
             + "; ".join("%s %.1f%% vs %.1f%%" % (k, 100 * g[k] / gt, 100 * st[k]["ms"] / ssum) for k in st) + ".\n")
    L.append("## 3. `ncu --set full --clock-control none --import-source on` (one launch per kernel, chr50mb workload: 16.4M extended positions, 4.9M records)\n")
    L.append("(Captured one commit before the last K3b change: `tb_fitg_round_kernel<1, 1>` is the first-round kernel of that commit; since the init pass "
             "runs the first outer-iteration head, the DWM default launches only `tb_fitg_init_kernel<1>` and `tb_fitg_round_kernel<0, 1>` -- the launch list "
             "of section 2 and the bench line of section 1 are of the final commit.)\n")
    L.append("| kernel | ms | grid x block | regs | DRAM read / write MB | SM thr % | issue % | fp64 pipe % | ALU pipe % | XU pipe % | shared pipe % | lanes / inst | warps active % |\n"
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for k, d in full.items():
        L.append("| %s | %.3f | %d x %d | %d | %.0f / %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
            k, d["time_ms"], d["grid"], d["block"], d["registers"], d["dram_read_MB"], d["dram_write_MB"], d["sm_throughput_pct"],
            d["issue_active_pct"], d["fp64_pipe_pct"], d["alu_pipe_pct"], d["xu_pipe_pct"], d["smem_pipe_pct"], d["lanes_per_inst"],
            d["warps_active_pct"]))
    L.append("")
    L.append(READINGS)
    open(os.path.join(PROF, "r01_summary.md"), "w").write("\n".join(L))

    def pick(prefix):
        for k in full:
            if k.startswith(prefix):
                return full[k]
        return {}

    def roof(prefix, bound, key):
        d = pick(prefix)
        return {"bound": bound, "bound_util_pct": d.get(key), "dram_MB_per_launch_chr50mb": round(d.get("dram_read_MB", 0) + d.get("dram_write_MB", 0), 1)}
    roofs = {"_source": "profiles/r01_summary.md section 3-4 (ncu --set full, one launch per kernel, chr50mb workload, B200)",
             "K1_pileup": roof("tb_pileup_kernel<0>", "L2 atomics / integer issue", "alu_pipe_pct"),
             "K2_counts": roof("tb_k2_scatter", "DRAM latency of the slab read-modify-writes (scatter, weights); shared-memory atomics (histogram)", "sm_throughput_pct"),
             "K3a_score": roof("tb_score_dwm2", "shared-memory data pipe (table rows + shuffles) together with instruction issue (~70 %); half of the rows come from tensor memory", "smem_pipe_pct"),
             "K3b_fit": roof("tb_fitg_round_kernel<0, 1>", "instruction issue (thread-sequential LM arithmetic) + L1 miss latency", "issue_active_pct"),
             "K3c_correct": roof("tb_correct_kernel", "instruction issue (sequential float-accumulator rolling sums, 10-row mean, verification counts)", "issue_active_pct"),
             "K4_footprint": roof("tb_footprint_kernel", "instruction issue (fp32 + shared loads)", "issue_active_pct")}
    # DRAM traffic of the dominant kernel at the bench's own size: one --set full launch of tb_score_dwm2_kernel in the hg38 bench
    # (gpurun_out/r01_full_hg38.ncu-rep, step 4 of profiles/collect.sh)
    rep = os.path.join(OUT, "r01_full_hg38.ncu-rep")
    if os.path.exists(rep):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rws = list(csv.reader(io.StringIO(txt)))
        h, u = rws[0], rws[1]
        for v in rws[2:]:
            if "tb_score_dwm2" in v[h.index("Kernel Name")]:
                rd = float(v[h.index("dram__bytes_read.sum")]) * SCALE[u[h.index("dram__bytes_read.sum")]] * 1e6
                wr = float(v[h.index("dram__bytes_write.sum")]) * SCALE[u[h.index("dram__bytes_write.sum")]] * 1e6
                roofs["K3a_score"]["dram_bytes_per_launch_hg38"] = rd + wr
                roofs["K3a_score"]["dram_note"] = ("read %.3g + write %.3g bytes; the write side is twice the 16 B per position of the final scores because the "
                                                   "four (strand, model) block classes write partial scores that a 2 ms kernel combines" % (rd, wr))
    json.dump(roofs, open(os.path.join(PROF, "r01_roofs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
