#!/usr/bin/env python
"""
bench.py -- headline benchmark of the TOBIAS per-base hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload hg38|chr50mb|small]

Metric (BASELINE.json): corrected + footprint-scored bp/s -- output-region base pairs (trimmed, strands combined) that went
through ATACorrect (read counts -> bias estimation -> bias model -> correction) AND ScoreBigwig footprint scoring, per second.
Workload at N=1: BASELINE.json configs[1], the synthetic hg38-sized genome (24 contigs with hg38 lengths + chrM, 3.1 Gb,
150 k peaks, 50 M alignment records).  At N>1 the same workload is sharded by contig (strong scaling, config[2]); the only
collective is the all-reduce of the bias-training counts / normalisation counters (NCCL) and of the verification PWMs.

One "step" = one pass of the whole path over the rank's shard:
   value : inputs (2-bit genome, read columns) resident in HBM; timed with CUDA events on the library's stream.
   e2e   : same pass through the public host API with HOST buffers: H2D of the read columns + ASCII genome from pinned
           memory, all kernels, D2H of the four float32 tracks + the footprint track into pinned memory.
Host-side region algebra (BED parsing, merge / subtract / split) happens once outside the timed region on both arms.

--impl reference times the reference's own CPU implementation (compiled, unmodified: oracle/_ref) with all host cores on
a bounded sample of the same workload (two small contigs of the synthetic genome at the same read density).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "corrected+footprint-scored bp/s"
SEED = 20261019
K_FLANK, WINDOW, BG_SHIFT, EXTEND, READ_SHIFT = 12, 100, 100, 100, (4, -5)
FP = dict(flank_min=10, flank_max=30, fp_min=20, fp_max=50)
LM_EXACT = 0          # fit kernel: 0 = one-pass outer iterations (DWM default), 1 = MINPACK-exact arithmetic (--lm-exact); --lm-mode 5 = moment form (opt-in)


def workload(name):
    from tobias_b200 import synth
    if name == "hg38":
        return dict(lengths=dict(synth.HG38_LENGTHS), n_peaks=150000, n_records=50_000_000,
                    desc="ATACorrect+ScoreBigwig, synthetic hg38-sized genome (3.1 Gb, 150k peaks, 50M reads)")
    if name == "batch24":                # one sample of the batch: the hg38-sized genome and peaks with 5 M records
        return dict(lengths=dict(synth.HG38_LENGTHS), n_peaks=150000, n_records=5_000_000,
                    desc="one sample of the batch: synthetic hg38-sized genome (3.1 Gb, 150k peaks), 5M records")
    if name == "chr50mb":
        return dict(lengths={"chr1": 50_000_000}, n_peaks=20000, n_records=5_000_000,
                    desc="ATACorrect+ScoreBigwig, synthetic 50 Mb chromosome, 20k 500bp peaks, 5M reads")
    if name == "small":
        return dict(lengths={"chr1": 3_000_000, "chr2": 2_000_000, "chrM": 16569}, n_peaks=600, n_records=400000,
                    desc="small smoke workload (not a benchmark configuration)")
    raise SystemExit("unknown workload " + name)


def shard_contigs(lengths, world):
    """Greedy longest-first bin packing of contigs onto ranks (deterministic)."""
    loads = [0] * world
    owner = {}
    for name, ln in sorted(lengths.items(), key=lambda kv: (-kv[1], kv[0])):
        r = min(range(world), key=lambda i: (loads[i], i))
        owner[name] = r
        loads[r] += ln
    return owner


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.samples, self.reasons, self.proc = [], set(), None
        self.idx = gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((float(f[0]), float(f[1])))
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except (ValueError, IndexError):
                pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(s for s, _ in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- B200 arm
def prepare_shard(args, rank, world):
    """Synthetic data of this rank's contigs + region tables as arrays (host driver work, untimed)."""
    import torch
    from tobias_b200 import synth_torch
    from tobias_b200.pipeline import prepare_regions
    from tobias_b200.regions import OneRegion, RegionList, to_arrays
    wl = workload(args.workload)
    owner = shard_contigs({n: ln for n, ln in wl["lengths"].items()}, world)
    mine = {n: ln for n, ln in wl["lengths"].items() if owner[n] == rank}
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    # every contig is seeded by its index in the whole workload: the same bases, peaks and records whichever rank builds it, so
    # the N-rank job processes exactly the data of the 1-rank job (`checksum` in the bench line shows it)
    ds = synth_torch.TorchDataset(SEED, mine, wl["n_peaks"], wl["n_records"], device=dev, pin=True, universe=wl["lengths"])
    names = ds.names
    chrom_info = {n: ln for n, ln in zip(names, ds.lengths) if n != "chrM"}          # --drop-chroms default
    peaks = RegionList([OneRegion(list(p)) for p in ds.peaks])
    regs = prepare_regions(peaks, chrom_info, [n for n in names if n in chrom_info], extend=EXTEND, k_flank=K_FLANK, window=WINDOW)
    index = {n: i for i, n in enumerate(names)}
    arrays = {k: to_arrays(sorted(regs[k], key=lambda r: (index[r.chrom], r.start, r.end)), index)
              for k in ("peak_regions", "nonpeak_regions", "input_regions", "output_regions")}
    return ds, arrays, wl


def allreduce_np(buf, device):
    """Sum a host int64/float64 array over ranks with NCCL (torch.distributed plumbing); no-op at world size 1."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return buf
    t = torch.from_numpy(buf).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    buf[...] = t.cpu().numpy()
    return buf


def run_step(eng, arrays, device, resident, host=None, out=None, genome_resident=False, collective=True):
    """One pass of the hot path over this rank's shard.  Returns (AtacBias, stage_ms list).
    genome_resident (batch of samples over one genome): only the records of the sample travel in an e2e step."""
    from tobias_b200.pipeline import AtacBias, STRANDS
    ctx = eng.ctx
    trace = os.environ.get("TB_BENCH_TRACE") and (not resident or os.environ.get("TB_BENCH_TRACE") == "2")   # diagnostic: host wall clock per phase (stderr)
    marks = [("start", time.perf_counter())]

    def mark(name):                       # resident passes: every library call below returns synchronised, so the host clock is enough
        if trace:
            import torch
            torch.cuda.synchronize()
        marks.append((name, time.perf_counter()))
    upload_reads = ctx.reads_upload_compact if host is not None and host.get("reads_compact") else ctx.reads_upload
    if not resident and genome_resident:
        upload_reads(*host["reads"])
        mark("reads_h2d")
    elif not resident:                    # e2e: inputs come from (pinned) host buffers every step
        ctx.genome_create(host["lengths"])
        upload_reads(*host["reads"])
        mark("reads_h2d")
        # the genome streams in on the copy stream while the read counts and the first contigs' training counts run
        if host["packed"] is not None and host.get("nruns") is not None:
            for i, ((sq, nm), (rs, re_)) in enumerate(zip(host["packed"], host["nruns"])):
                ctx.genome_upload_packed_runs(i, sq, rs, re_, host["lengths"][i], wait=False)
        elif host["packed"] is not None:
            for i, (sq, nm) in enumerate(host["packed"]):
                ctx.genome_upload_packed(i, sq, nm, host["lengths"][i], wait=False)
        else:
            for i, g in enumerate(host["genome"]):
                ctx.genome_upload_async(i, g)
        if trace:
            marks.append(("genome_h2d enqueued (no sync)", time.perf_counter()))
    # --- normalisation counters + bias-training counts, then the one exchange step of the path
    reads_peaks = ctx.count_reads(*arrays["peak_regions"], READ_SHIFT)
    reads_nonpeaks = ctx.count_reads(*arrays["nonpeak_regions"], READ_SHIFT)
    mark("count_reads")
    counts, scalars = ctx.bias_counts(*arrays["input_regions"], K_FLANK, BG_SHIFT, READ_SHIFT, 10, True)
    k2_ms = ctx.last_kernel_ms()
    if not resident and not genome_resident:
        ctx.genome_wait()
    mark("bias_counts")
    packed = np.concatenate([counts.ravel(), scalars, np.array([reads_peaks, reads_nonpeaks], dtype=np.int64)])
    if collective:
        allreduce_np(packed, device)
    counts = packed[:counts.size].reshape(counts.shape)
    scalars = packed[counts.size:counts.size + 5]
    reads_peaks, reads_nonpeaks = int(packed[-2]), int(packed[-1])
    reads_total = reads_peaks + reads_nonpeaks
    bias = AtacBias(2 * K_FLANK + 1, "DWM")
    for si, st in enumerate(STRANDS):
        m = bias.bias[st]
        m.counts = counts[si, 0].astype(np.float64)
        m.bg_counts = counts[si, 1].astype(np.float64)
        m.prepare_mat()                   # host numpy, identical on every rank
    bias.no_reads = int(scalars[0])
    bias.correction_factor = (10000000 / reads_total) * (reads_total / reads_peaks)      # atacorrect.py:298-300
    mark("prepare_mat")
    eng.set_bias(bias)
    mark("counts+model")
    # --- correction + footprints of this rank's output regions
    c, s, e = arrays["output_regions"]
    if not resident:                      # the four tracks travel to the host chunk by chunk while the correction's last stage and the footprint kernels run
        n_out = int((e - s).sum())
        ctx.correct_set_sink(out["tracks"], n_out, K_FLANK, as_f64=False, prepost=True)
    ctx.correct_run(c, s, e, K_FLANK, WINDOW, bias.correction_factor, READ_SHIFT, LM_EXACT)
    stage = ctx.stage_ms(4)
    mark("correct_run")
    ctx.signal_from_corrected(FP["flank_max"], e - s)
    if not resident:
        ctx.score_set_sink(out["footprint"])
    ctx.score_run(ctx.score_params("footprint", flank=FP["flank_max"], **FP))
    k4_ms = ctx.last_kernel_ms()
    mark("footprints")
    if not resident:
        ctx.correct_download_wait()       # waits for the copy stream: tracks and footprint scores
        ctx.score_sink_wait()
        mark("tracks_d2h_wait")
    if trace:
        sys.stderr.write(("e2e" if not resident else "resident") + " trace (ms): " + ", ".join("%s %.1f" % (n, (t - marks[i][1]) * 1e3) for i, (n, t) in enumerate(marks[1:])) + "\n")
    phases = {n: (t - marks[i][1]) * 1e3 for i, (n, t) in enumerate(marks[1:])}
    return bias, {"K1_pileup": stage[0], "K3a_score": stage[1], "K3b_fit": stage[2], "K3c_correct": stage[3],
                  "K2_counts": k2_ms, "K4_footprint": k4_ms}, phases


def _mix(x):
    """64-bit mixer (splitmix64 finaliser) on a uint64 array."""
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def checksum_block(eng, ds, arrays, bias, out, world, device):
    """Order-independent checksums of what the last e2e step produced: sum over output positions of mix(global contig, position) *
    float32 bits, per track (mod 2^64, summed over ranks), and the same over the all-reduced count tensors.  The 1-rank job and
    the N-rank job process the same data (contig-keyed seeds), so equal checksums show that sharding leaves the results unchanged."""
    import torch
    import torch.distributed as dist
    wl_names = list(workload_names())
    c, s, e = arrays["output_regions"]
    gid = np.array([wl_names.index(n) for n in ds.names], dtype=np.uint64)
    lens = (e - s).astype(np.int64)
    pos = np.concatenate([np.arange(a, b, dtype=np.uint64) for a, b in zip(s, e)]) if len(s) else np.zeros(0, np.uint64)
    key = _mix(np.repeat(gid[c] << np.uint64(40), lens) + pos + np.uint64(0x9E3779B97F4A7C15))
    sums = {}
    with np.errstate(over="ignore"):
        for t, a in list(out["tracks"].items()) + [("footprint", out["footprint"])]:
            bits = a[:key.shape[0]].view(np.uint32).astype(np.uint64)
            sums[t] = int((key * (bits + np.uint64(1))).sum(dtype=np.uint64))
    names = sorted(sums)
    v = torch.tensor([sums[n] & 0x7FFFFFFF for n in names] + [sums[n] >> 31 for n in names], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(v)                # pieces of 31 / 33 bits: the sums over <= 8 ranks cannot overflow int64
    k = len(names)
    total = {n: (int(v[i]) + (int(v[k + i]) << 31)) & 0xFFFFFFFFFFFFFFFF for i, n in enumerate(names)}
    cnt = 0
    with np.errstate(over="ignore"):
        for st in ("forward", "reverse"):
            for m in (bias.bias[st].counts, bias.bias[st].bg_counts):
                a = np.asarray(m).astype(np.uint64).ravel()
                cnt = (cnt * 1000003 + int((a * _mix(np.arange(a.shape[0], dtype=np.uint64))).sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF
    return {"tracks": {n: "%016x" % total[n] for n in names}, "count_tensors": "%016x" % cnt, "no_reads": int(bias.no_reads),
            "note": "order-independent; identical at every --gpus N when sharding leaves the results unchanged"}


_WL_NAMES = []


def workload_names():
    return _WL_NAMES


def bench_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1 and not os.environ.get("TB_NO_NUMA_BIND"):      # before any pinned buffer exists (first touch decides the node)
        from tobias_b200.dist import bind_to_device_numa
        _NUMA["cpus"] = len(bind_to_device_numa(local) or ())
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from tobias_b200.pipeline import Engine
    t_setup = time.time()
    ds, arrays, wl = prepare_shard(args, rank, world)
    _WL_NAMES[:] = list(wl["lengths"])
    eng = Engine(local)
    ctx = eng.ctx
    for kv in args.opt:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    eng.names, eng.index, eng.lengths = ds.names, {n: i for i, n in enumerate(ds.names)}, ds.lengths
    rc = ds.reads
    from tobias_b200 import lib
    flags = (rc.is_reverse.astype(np.uint8) * lib.TB_READ_REVERSE | rc.five_prime_is_match().astype(np.uint8) * lib.TB_READ_5P_MATCH
             | (~rc.usable()).astype(np.uint8) * lib.TB_READ_SKIP)
    # the record columns wait on the host in the compact form a sorted BAM decodes to (9 bytes per record: per-contig offsets, start int32,
    # reference length and 5' clip uint16, flags; --wide-inputs: the five 4 / 1-byte columns of the first interface)
    comp = None if args.wide_inputs else lib.compact_read_columns(rc.chrom_id, rc.ref_start, rc.ref_end, rc.clip5(), flags, len(ds.lengths))
    cols = []
    for a in (comp if comp is not None else (rc.chrom_id, rc.ref_start, rc.ref_end, rc.clip5(), flags)):          # pinned copies
        p = ctx.pinned_empty(a.shape, a.dtype)
        p[...] = a
        cols.append(p)
    # the genome waits on the host in the device layout (2 bit per base; the N mask as the runs of N it consists of -- what
    # tobias_b200.io.fasta.PackedGenome caches next to a FASTA): 0.25 bytes per base cross the link every step instead of 1
    # (--wide-inputs: 2 bit + mask words, 0.375; --ascii-genome: the ASCII contigs, packed on the device)
    packed, nruns = None, None
    if not args.ascii_genome:
        from tobias_b200.io.fasta import pack_ascii
        packed = []
        for g in ds.genome:
            n = int(g.shape[0])
            packed.append(pack_ascii(g, ctx.pinned_empty(((n + 15) // 16,), np.uint32), ctx.pinned_empty(((n + 31) // 32,), np.uint32)))
        if not args.wide_inputs:
            nruns = [lib.mask_runs(nm, int(g.shape[0])) for (sq, nm), g in zip(packed, ds.genome)]
    host = {"lengths": ds.lengths, "genome": ds.genome, "reads": cols, "reads_compact": comp is not None, "packed": packed, "nruns": nruns}
    c, s, e = arrays["output_regions"]
    out_bp = int((e - s).sum())
    out = {"tracks": {t: ctx.pinned_empty((out_bp,), np.float32) for t in lib.TRACKS}, "footprint": ctx.pinned_empty((out_bp,), np.float32)}
    if packed is None:
        genome_bytes = sum(int(g.nbytes) for g in ds.genome)
    elif nruns is not None:
        genome_bytes = sum(int(a.nbytes) for a, _ in packed) + sum(int(rs.nbytes) + int(re_.nbytes) for rs, re_ in nruns)
    else:
        genome_bytes = sum(int(a.nbytes) + int(b.nbytes) for a, b in packed)
    h2d = genome_bytes + sum(int(a.nbytes) for a in cols) + sum(int(x.nbytes) for v in arrays.values() for x in v)
    d2h = 5 * out_bp * 4 + 2 * 3 * 5 * 25 * 8
    setup_s = time.time() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(resident, n_warm, n_steps):
        stages, last = None, None
        for _ in range(n_warm):
            last, stages, _ = run_step(eng, arrays, device, resident, host, out)
        barrier()
        launches0 = ctx.launch_count()
        ctx.timer_begin()
        t0 = time.perf_counter()
        acc, pacc = {}, {}
        for _ in range(n_steps):
            last, stages, phases = run_step(eng, arrays, device, resident, host, out)
            for k, v in stages.items():
                acc[k] = acc.get(k, 0.0) + v
            for k, v in phases.items():
                pacc.setdefault(k, []).append(round(v, 2))
        ms = ctx.timer_end()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([ms, wall], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), {k: v / n_steps for k, v in acc.items()}, ctx.launch_count() - launches0, last, pacc

    # resident inputs for `value`
    run_step(eng, arrays, device, False, host, out)
    sampler = ClockSampler(local)
    sampler.start()
    ms_res, wall_res, stages, launches, bias, phases_res = timed(True, args.warmup, args.steps)
    clocks = sampler.stop()
    ms_e2e, wall_e2e, _, _, _, phases_e2e = timed(False, max(1, min(args.warmup, 2)), args.steps)

    tot_bp = torch.tensor([out_bp, int((arrays["output_regions"][2] - arrays["output_regions"][1] + 124).sum())], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tot_bp)
    total_out_bp, total_ext_bp = float(tot_bp[0]), float(tot_bp[1])
    value = total_out_bp * args.steps / (ms_res * 1e-3)
    e2e_value = total_out_bp * args.steps / (ms_e2e * 1e-3)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    # algorithmic bytes per unit; per-rank units for per-rank launch durations
    ext_bp = float(tot_bp[1] / world)
    n_reads = float(len(rc))
    alg = {"K1_pileup": 5.0 * n_reads + 8.0 * ext_bp, "K2_counts": 5.0 * n_reads, "K3a_score": 16.375 * ext_bp,
           "K3b_fit": 24.0 * ext_bp, "K3c_correct": 25.9 * out_bp, "K4_footprint": 8.0 * out_bp}
    dom = max(stages, key=lambda k: stages[k])
    achieved = alg[dom] / (stages[dom] * 1e-3) / 1e9
    per_stage = {k: {"ms": round(v, 4), "alg_GBps": round(alg[k] / (v * 1e-3) / 1e9, 2) if v > 0 else None} for k, v in stages.items()}
    try:                                   # binding unit of every kernel, from the committed ncu summary of this round (not a live measurement)
        roofs = json.load(open(os.path.join(ROOT, "profiles", "r02_roofs.json")))
        for k in per_stage:
            if k in roofs:
                per_stage[k]["bound"] = roofs[k]["bound"]
                per_stage[k]["bound_util_pct_ncu"] = roofs[k]["bound_util_pct"]
    except Exception:
        roofs = {}
    # DRAM traffic of the dominant kernel from the committed --set full capture of this workload on one GPU (profiles/collect.sh, step 4)
    traffic = roofs.get(dom, {}).get("dram_bytes_per_launch_hg38") if (args.workload == "hg38" and world == 1) else None
    result = {
        "metric": METRIC, "value": value, "unit": "bp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(bench_config(wl, world), out_bp_total=int(total_out_bp), extended_bp_total=int(total_ext_bp), reads_this_rank=int(len(rc)),
                       genome_host_format="ascii" if args.ascii_genome else "2 bit + N mask (packed once on the host, as the .tb2bit cache of a FASTA)",
                       **({"options": args.opt} if args.opt else {})),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "bp/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                     "algorithmic_bytes_per_launch": alg[dom], "peak_source": peak_src,
                     "binding_unit": roofs.get(dom, {}).get("bound"), "binding_unit_util_pct_ncu": roofs.get(dom, {}).get("bound_util_pct"),
                     "note": "achieved = algorithmic bytes of the stage (SURVEY 8d) / its CUDA-event time (all launches of the stage); none of K2-K4 is "
                             "HBM bound -- the unit that binds each kernel and its ncu utilisation are in `stages` (profiles/r02_summary.md); "
                             "traffic: the ncu --set full capture of this kernel in this workload committed under profiles/ (null for other workloads / ranks)"},
        "stages": per_stage, "wall_ms_per_step": wall_res / args.steps, "setup_s": round(setup_s, 1),
        # host clock around every library call of the timed steps (one entry per step; resident: each call returns synchronised, so these
        # add up to ms_per_step and show where host work sits between the kernels; e2e: calls overlap with the copy stream)
        "host_phases_ms": {"resident": phases_res, "e2e": phases_e2e},
        "correction_factor": bias.correction_factor,
    }
    result["checksum"] = checksum_block(eng, ds, arrays, bias, out, world, device)
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        try:
            # the reference on a bounded sample of the workload (CPU baseline) -- and the SAME sample through the B200 path, compared
            ref = cpu_reference(args, steps=1, warmup=0, n_out=max(2048, (os.cpu_count() or 1) * 128), collect=True)
            result["cpu_baseline"] = ref["cpu_baseline"]
            try:
                result["parity"] = compare_with_reference(eng, ref["sample"], ref["kept"], LM_EXACT)
            except Exception as ex:
                result["parity"] = {"green": False, "error": repr(ex)}
        except Exception as ex:          # the baseline is a report, never a reason to lose the GPU number
            result["cpu_baseline"] = {"error": repr(ex)}
    if args.fit_stats and rank == 0:      # diagnostic: distribution of the window fits of the last step (stderr)
        tr = eng.ctx.correct_fit_trace()
        mode, nf, info = tr[:, :, 0].ravel().astype(int), tr[:, :, 5].ravel(), tr[:, :, 4].ravel().astype(int)
        f = nf[mode < 2]
        sys.stderr.write("fit stats: windows %d modes %s info %s nfev mean %.2f p50 %.0f p90 %.0f p99 %.0f p999 %.0f max %.0f\n" % (
            nf.size, np.bincount(mode, minlength=3).tolist(), np.bincount(info[mode < 2]).tolist(), f.mean(),
            *np.percentile(f, [50, 90, 99, 99.9]), f.max()))
    if rank == 0:
        print(json.dumps(result))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------- reference arm
class _ListQ:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


def _ref_worker(task):
    """Runs inside a pool worker: the compiled, unmodified reference on one chunk of regions.  With params.collect the arrays the
    reference produced travel back (tracks as the float32 the bigWig stores, footprints, one record per curve_fit call)."""
    kind, regions, params = task
    from oracle import ref_loader
    ref = ref_loader.load()
    R = ref.regions
    regs = R.RegionList([R.OneRegion(list(r)) for r in regions])
    if kind == "count":
        return ref.atacorrect_functions.count_reads(regs, params)
    if kind == "estimate":
        return ref.atacorrect_functions.bias_estimation(regs, params)
    if kind == "correct":
        collect = getattr(params, "collect", False)
        af = ref.atacorrect_functions
        calls = []
        orig = af.curve_fit
        if collect:                       # instrumentation around scipy's curve_fit as the reference calls it (:292): same call, same exceptions
            def traced(f, x, y):
                try:
                    popt, pcov, info, _msg, ier = orig(f, x, y, full_output=True)
                except Exception:
                    calls.append((1.0, 0.0, 0.0, 0.0, 0.0, 0.0))
                    raise
                pmax = float(np.max(f(x, *popt)))
                calls.append((0.0, float(popt[0]), float(popt[1]), float(ier), float(info["nfev"]), pmax))
                return popt, pcov
            af.curve_fit = traced
        tracks = ("uncorrected", "bias", "expected", "corrected")
        qs = {t + ":both": _ListQ() for t in (tracks if collect else ("corrected",))}
        marks = []                        # number of curve_fit calls made when a region's last track is queued (:381-400: "corrected" last)
        put_corrected = qs["corrected:both"].put
        qs["corrected:both"].put = lambda x: (marks.append(len(calls)), put_corrected(x))
        params.qs = qs
        out = []
        fp_bp = 0
        try:
            af.bias_correction(regs, params, params.bias_obj)
            n0 = 0
            for i, (_k, key, sig) in enumerate(qs["corrected:both"].items):       # ScoreBigwig footprint scoring of the corrected track
                ext = np.zeros(sig.shape[0] + 2 * FP["flank_max"])          # the flank outside the bigWig's data reads as 0 (score_bigwig.py:52-53)
                ext[FP["flank_max"]:-FP["flank_max"]] = sig.astype(np.float32)
                fp = ref.signals.tobias_footprint_array(ext, FP["flank_min"], FP["flank_max"], FP["fp_min"], FP["fp_max"])
                fp_bp += sig.shape[0]
                if collect:
                    out.append((key, {t: qs[t + ":both"].items[i][2].astype(np.float32) for t in tracks},
                                fp[FP["flank_max"]:-FP["flank_max"]].astype(np.float32), np.array(calls[n0:marks[i]], dtype=np.float64).reshape(-1, 6)))
                    n0 = marks[i]
        finally:
            af.curve_fit = orig
        return fp_bp, out
    raise ValueError(kind)


def reference_sample(workload_name, seed=None):
    """Bounded sample of a workload for the CPU side: the two smallest autosome-like contigs (capped) at the workload's
    read / peak density, written where the reference's (shimmed) pysam reads them."""
    import types
    from tobias_b200 import synth_torch
    from tobias_b200.pipeline import prepare_regions
    from tobias_b200.regions import OneRegion, RegionList
    from tobias_b200.io.fasta import write_fasta
    wl = workload(workload_name)
    names = [n for n in wl["lengths"] if n != "chrM"]
    sample_names = sorted(names, key=lambda n: wl["lengths"][n])[:2] if len(names) > 2 else names
    tot = float(sum(wl["lengths"].values()))
    cap = {"hg38": 24_000_000, "batch24": 24_000_000, "chr50mb": 50_000_000}.get(workload_name, 6_000_000)
    lengths = {n: min(wl["lengths"][n], cap) for n in sample_names}
    frac = sum(lengths.values()) / tot
    ds = synth_torch.TorchDataset((SEED + 999) if seed is None else seed, lengths, max(4, int(wl["n_peaks"] * frac)),
                                  int(wl["n_records"] * frac), device="cpu")
    tmp = tempfile.mkdtemp(prefix="tb_ref_")
    fa = os.path.join(tmp, "sample.fa")
    write_fasta(fa, zip(ds.names, ds.genome))
    npz = os.path.join(tmp, "sample_reads.npz")
    ds.reads.save(npz)
    chrom_info = dict(zip(ds.names, ds.lengths))
    regs = prepare_regions(RegionList([OneRegion(list(p)) for p in ds.peaks]), chrom_info, ds.names, extend=EXTEND, k_flank=K_FLANK, window=WINDOW)
    params = types.SimpleNamespace(bam=npz, genome=fa, k_flank=K_FLANK, bg_shift=BG_SHIFT, read_shift=list(READ_SHIFT), score_mat="DWM",
                                   verbosity=0, log_q=None, window=WINDOW, split_strands=False, qs={}, bias_obj=None, collect=False)
    return {"ds": ds, "regs": regs, "params": params, "wl": wl, "tmp": tmp}


def cpu_reference(args, steps, warmup, n_out=None, collect=False, sample=None):
    """The reference's own CPU implementation (oracle/_ref: compiled, unmodified) over multiprocessing on all host cores.
    collect: keep what the last step computed (read counts, AtacBias, per-region tracks / footprints / fit records) for the parity block."""
    import multiprocessing as mp
    from oracle import ref_loader
    ref_loader.load()
    cores = os.cpu_count() or 1
    sample = sample or reference_sample(args.workload)
    ds, regs, params, wl = sample["ds"], sample["regs"], sample["params"], sample["wl"]
    params.collect = bool(collect)
    out_all = [tuple(r[:3]) for r in regs["output_regions"]]
    n_out = min(len(out_all), n_out or max(cores * 96, 192))          # ~10 s of correction + footprints on the box's cores per step
    rng = np.random.default_rng(1)
    out_sample = [out_all[i] for i in sorted(rng.choice(len(out_all), n_out, replace=False))]
    inp = [tuple(r[:3]) for r in regs["input_regions"]]
    pk = [tuple(r[:3]) for r in regs["peak_regions"]]
    npk = [tuple(r[:3]) for r in regs["nonpeak_regions"]]

    def chunks(lst, n):
        per = max(1, -(-len(lst) // n))
        return [lst[i:i + per] for i in range(0, len(lst), per)]

    results = []
    kept = None
    ctxmp = mp.get_context("fork")
    with ctxmp.Pool(cores) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            n_split = cores * 4
            reads_peaks = sum(pool.map(_ref_worker, [("count", ch, params) for ch in chunks(pk, n_split)]))
            reads_nonpeaks = sum(pool.map(_ref_worker, [("count", ch, params) for ch in chunks(npk, n_split)]))
            t_count = time.perf_counter() - t0
            t0 = time.perf_counter()
            parts = pool.map(_ref_worker, [("estimate", ch, params) for ch in chunks(inp, n_split)])
            bias = parts[0]
            for p in parts[1:]:
                bias.join(p)
            for st in ("forward", "reverse"):
                bias.bias[st].prepare_mat()
            total = reads_peaks + reads_nonpeaks
            bias.correction_factor = (10000000 / total) * (total / max(reads_peaks, 1))
            t_est = time.perf_counter() - t0
            params.bias_obj = bias
            t0 = time.perf_counter()
            done = pool.map(_ref_worker, [("correct", ch, params) for ch in chunks(out_sample, n_split)])
            t_corr = time.perf_counter() - t0
            params.bias_obj = None
            sample_bp = int(sum(d[0] for d in done))
            all_bp = sum(e - s for _, s, e in out_all)
            # counts + estimation cover the whole sample genome; correction + footprints cover `sample_bp` of its `all_bp`
            t_full = t_count + t_est + t_corr * (all_bp / float(sample_bp))
            if it >= warmup:
                results.append((all_bp / t_full, t_count, t_est, t_corr, sample_bp, all_bp))
            if collect:
                kept = {"reads_peaks": int(reads_peaks), "reads_nonpeaks": int(reads_nonpeaks), "bias": bias,
                        "regions": [x for d in done for x in d[1]]}
    value = float(np.mean([r[0] for r in results]))
    r = results[-1]
    desc = ("contigs %s (%.1f Mb) of the synthetic workload at its read/peak density: count_reads + bias_estimation over all %d input regions "
            "(%d records), bias_correction + footprint scoring over %d of %d output regions (%d of %d bp), scaled by bp; stage seconds "
            "count=%.1f estimate=%.1f correct+footprint=%.1f" % ("+".join(ds.names), sum(ds.lengths) / 1e6, len(inp), len(ds.reads), n_out,
                                                                len(out_all), r[4], r[5], r[1], r[2], r[3]))
    return {"value": value, "cpu_baseline": {"value": value, "unit": "bp/s", "cores": cores, "kind": "reference", "sample": desc},
            "ms_per_step": 1e3 * (r[1] + r[2] + r[3]), "wl": wl, "kept": kept, "sample": sample}


# ---------------------------------------------------------------------------------------------------- parity on the benchmarked path
# UNSTABLE WINDOWS.  The reference divides every window prediction max(0, a * bias + b) by its own maximum (atacorrect_functions.py:
# 301-302).  Where the least-squares problem has a flat valley -- typically a window with a few reads whose best fit is "no bias signal":
# every (a, b) that keeps a * bias + b at or below zero is equally good -- scipy's answer is decided by the last bits of the bias input: a
# prediction of 1e-10 and a prediction of exactly 0 become O(1) and 0 after the division.  The reference does not keep its OWN answer on such
# windows when its bias input changes by 1e-13 .. 1e-12 relative, so no implementation whose bias track differs in the last bits (DWM mode: table
# products instead of glibc pow / log2, ~1e-14 absolute) can follow it there.  The parity block therefore (1) lists the windows whose fitted
# max(prediction) or fit-vs-fallback decision differs from the reference's beyond the tolerance, (2) re-runs scipy's curve_fit exactly as the
# reference calls it on each of them with the bias changed by 1e-13 .. 1e-12 (_reference_is_unstable), and (3) calls a window UNSTABLE when scipy's own result moves
# by more than a tenth of the tolerance.  Disagreements on unstable windows are counted (`unstable_flips`: the flip budget); a disagreement
# on a STABLE window (`stable_mismatches`) fails the block.
# Distance of two fits = largest difference of their NORMALISED predictions max(0, a x + b) / max over the window's bias values x; a window
# enters the 10-row mean with weight 1 / 10, so 1e-4 here is the 1e-5 of the track tolerance.  _fit_distance is the cheap screen on the
# parameters alone (an upper bound up to the magnitude of x, which needs no window data); _pred_distance the measure itself.
FIT_TOL = 1e-4


def _norm_pred(x, a, b):
    p = np.maximum(0.0, a * x + b)
    m = p.max()
    return p / m if m > 0 else p


def _pred_distance(x, a0, b0, a1, b1):
    return float(np.abs(_norm_pred(x, a0, b0) - _norm_pred(x, a1, b1)).max())


def _fit_distance(a0, b0, m0, a1, b1, m1):
    z0, z1 = m0 <= 0.0, m1 <= 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        d = 4.0 * np.abs(a0 / m0 - a1 / m1) + np.abs(b0 / m0 - b1 / m1)
    d = np.where(z0 & z1, 0.0, np.where(z0 | z1, np.inf, d))          # an all-zero prediction stays all zero (no normalisation, :301)
    return d
TOL = {"rtol": 1e-5, "atol": {"uncorrected": 0.0, "bias": 1e-6, "expected": 2e-6, "corrected": 2e-6, "footprint": 2e-6},
       "min_fraction_inside": 0.999, "min_fraction_inside_x10": 0.9999, "far_factor": 100.0,
       "atol_unit": "x max(1, correction_factor): expected / corrected / footprint are in units of one read times the correction factor "
                    "(atacorrect.py:292-300), so the absolute floor scales with it (hg38 workload: factor ~0.9 -> the values above)",
       "note": "tracks compared as the float32 the bigWig stores, strands summed: with lim = rtol * |ref| + atol, |gpu - ref| <= lim on at least "
               "min_fraction_inside of the positions, <= 10 lim on min_fraction_inside_x10 of them and <= far_factor * lim everywhere (the rescaling "
               "factor of `corrected` is a quotient of float32-accumulator sums that cancel, atacorrect_functions.py:331-350, so a few positions "
               "carry float32 noise well above 1e-5 in the reference itself); uncorrected: bit-exact; read counts, bias-training count tensors and "
               "the log tables derived from them: bit-exact"}


def compare_with_reference(eng, sample, kept, lm_exact):
    """Runs the B200 path on the reference arm's sample (same contigs, records and regions) and compares it with the arrays the
    reference computed (`kept` of cpu_reference(collect=True)).  Returns the `parity` block of the bench line."""
    from tobias_b200.pipeline import AtacBias, STRANDS
    from tobias_b200.regions import to_arrays
    from tobias_b200 import lib
    ds, regs = sample["ds"], sample["regs"]
    ctx = eng.ctx
    eng.load_genome(ds.names, ds.genome)
    eng.load_reads(ds.reads)
    index = eng.index
    arr = {k: to_arrays(sorted(regs[k], key=lambda r: (index[r.chrom], r.start, r.end)), index)
           for k in ("peak_regions", "nonpeak_regions", "input_regions", "output_regions")}
    res = {"tolerance": TOL, "lm_exact_order": lm_exact}
    rp, rn = ctx.count_reads(*arr["peak_regions"], READ_SHIFT), ctx.count_reads(*arr["nonpeak_regions"], READ_SHIFT)
    res["read_counts_exact"] = bool(rp == kept["reads_peaks"] and rn == kept["reads_nonpeaks"])
    counts, scalars = ctx.bias_counts(*arr["input_regions"], K_FLANK, BG_SHIFT, READ_SHIFT, 10, True)
    rb = kept["bias"]
    exact = int(scalars[0]) == int(rb.no_reads)
    bias = AtacBias(2 * K_FLANK + 1, "DWM")
    tables_equal = True
    for si, st in enumerate(STRANDS):
        exact = exact and np.array_equal(counts[si, 0], np.asarray(rb.bias[st].counts)) and np.array_equal(counts[si, 1], np.asarray(rb.bias[st].bg_counts))
        m = bias.bias[st]
        m.counts, m.bg_counts = counts[si, 0].astype(np.float64), counts[si, 1].astype(np.float64)
        m.prepare_mat()
        for k in ("bias_pwm_log", "bg_pwm_log", "bias_dwm_log", "bg_dwm_log"):
            tables_equal = tables_equal and np.array_equal(np.asarray(getattr(m, k)), np.asarray(getattr(rb.bias[st], k)), equal_nan=True)
    res["count_tensors_exact"] = bool(exact)
    res["log_tables_exact"] = bool(tables_equal)
    total = rp + rn
    bias.correction_factor = (10000000 / total) * (total / max(rp, 1))
    eng.set_bias(bias)
    c, s, e = arr["output_regions"]
    ctx.correct_run(c, s, e, K_FLANK, WINDOW, bias.correction_factor, READ_SHIFT, lm_exact)
    got = ctx.correct_download(split_strands=False, as_f64=False, prepost=False)
    trace = ctx.correct_fit_trace()
    ctx.signal_from_corrected(FP["flank_max"], e - s)
    ctx.score_run(ctx.score_params("footprint", flank=FP["flank_max"], **FP))
    got["footprint"] = ctx.score_download()
    out_off = np.concatenate([[0], np.cumsum(e - s)])
    nwin = np.maximum(0, -(-((e - s) + 2 * (K_FLANK + WINDOW // 2) - 2 * K_FLANK - WINDOW) // 10))
    win_off = np.concatenate([[0], np.cumsum(nwin)])
    where = {(ds.names[ci], int(a), int(b)): j for j, (ci, a, b) in enumerate(zip(c, s, e))}
    acc = {t: {"n": 0, "outside": 0, "far_outside": 0, "outside_unexplained": 0, "outside_x10_unexplained": 0, "far_outside_unexplained": 0,
               "max_abs": 0.0, "max_rel": 0.0} for t in list(lib.TRACKS) + ["footprint"]}
    fit = {"windows": 0, "fitted_ref": 0, "fallback_ref": 0, "empty": 0, "decision_flips": 0, "ier_differs": 0, "nfev_differs": 0,
           "regions_with_call_count_mismatch": 0, "mismatching_windows": 0, "parameter_only_differences": 0, "unstable_flips": 0, "stable_mismatches": 0,
           "stable_decision_flips": 0, "regions_with_unstable_flip": 0}
    examples = []
    fe = K_FLANK + WINDOW // 2
    reach = WINDOW + FP["flank_max"] + FP["fp_max"]              # expected <- +-50, rescale <- +-50 more, footprint windows on top
    for key, tr, fp, calls in kept["regions"]:
        j = where[(key[0], int(key[1]), int(key[2]))]
        lo, hi = int(out_off[j]), int(out_off[j + 1])
        w0, w1 = int(win_off[j]), int(win_off[j + 1])
        g = trace[:, w0:w1, :]                                   # [strand][window][mode, a, b, max(prediction), info, nfev]
        live = [np.nonzero(g[si, :, 0] < 2)[0] for si in range(2)]
        fit["windows"] += 2 * (w1 - w0)
        fit["empty"] += 2 * (w1 - w0) - len(live[0]) - len(live[1])
        explained = np.zeros(hi - lo, dtype=bool)                # output positions an unstable window can reach
        if len(live[0]) + len(live[1]) != calls.shape[0]:
            fit["regions_with_call_count_mismatch"] += 1
            fit["decision_flips"] += max(len(live[0]) + len(live[1]), calls.shape[0])
        else:
            gg = np.concatenate([g[0, live[0]], g[1, live[1]]])  # the reference fits the forward windows first, then the reverse ones (:275-309)
            widx = np.concatenate([live[0], live[1]])
            wstrand = np.concatenate([np.zeros(len(live[0]), dtype=int), np.ones(len(live[1]), dtype=int)])
            fit["fitted_ref"] += int((calls[:, 0] == 0).sum())
            fit["fallback_ref"] += int((calls[:, 0] == 1).sum())
            dflip = gg[:, 0] != calls[:, 0]
            fit["decision_flips"] += int(dflip.sum())
            both = (gg[:, 0] == 0) & (calls[:, 0] == 0)
            fit["ier_differs"] += int((gg[both, 4] != calls[both, 3]).sum())
            fit["nfev_differs"] += int((gg[both, 5] != calls[both, 4]).sum())
            pm_g, pm_r = gg[:, 3], calls[:, 5]
            mism = dflip | (both & (_fit_distance(gg[:, 1], gg[:, 2], pm_g, calls[:, 1], calls[:, 2], pm_r) > FIT_TOL))
            fit["mismatching_windows"] += int(mism.sum())
            if mism.any():
                ci, a_, b_ = int(c[j]), int(s[j]) - fe, int(e[j]) + fe
                pile = ctx.pileup([ci], [a_], [b_], READ_SHIFT)
                xs = [np.nan_to_num(ctx.score_sequence(si, ci, a_, b_)) for si in range(2)]
                any_unstable = False
                for kk in np.nonzero(mism)[0]:
                    p0 = K_FLANK + 10 * int(widx[kk])
                    x, y = xs[wstrand[kk]][p0:p0 + WINDOW], pile[wstrand[kk]][p0:p0 + WINDOW].astype(np.float64)
                    if not dflip[kk] and _pred_distance(x, gg[kk, 1], gg[kk, 2], calls[kk, 1], calls[kk, 2]) <= FIT_TOL:
                        fit["mismatching_windows"] -= 1          # the parameters differ along the fit's flat direction, the predictions do not
                        fit["parameter_only_differences"] += 1
                        continue
                    if _reference_is_unstable(x, y):
                        fit["unstable_flips"] += 1
                        any_unstable = True
                        a0 = p0 - fe - reach
                        explained[max(0, a0):max(0, a0 + WINDOW + 2 * reach)] = True
                    else:
                        fit["stable_mismatches"] += 1
                        fit["stable_decision_flips"] += int(dflip[kk])
                        if len(examples) < 24:
                            examples.append({"region": [key[0], int(key[1]), int(key[2])], "strand": int(wstrand[kk]), "window": int(widx[kk]),
                                             "x": x.tolist(), "y": y.tolist(), "gpu": gg[kk].tolist(), "ref": calls[kk].tolist()})
                fit["regions_with_unstable_flip"] += int(any_unstable)
        for t in acc:
            ref = (fp if t == "footprint" else tr[t]).astype(np.float64)
            x = got[t][lo:hi].astype(np.float64)
            d = np.abs(x - ref)
            a = acc[t]
            a["n"] += d.size
            lim = TOL["rtol"] * np.abs(ref) + TOL["atol"][t] * (1.0 if t in ("uncorrected", "bias") else max(1.0, bias.correction_factor))
            out_, far_ = d > lim, d > TOL["far_factor"] * lim
            a["outside"] += int(out_.sum())
            a["far_outside"] += int(far_.sum())
            a["outside_unexplained"] += int((out_ & ~explained).sum())
            a["outside_x10_unexplained"] += int(((d > 10.0 * lim) & ~explained).sum())
            a["far_outside_unexplained"] += int((far_ & ~explained).sum())
            a["max_abs"] = max(a["max_abs"], float(d.max()) if d.size else 0.0)
            nz = np.abs(ref) > 1e-3
            if nz.any():
                a["max_rel"] = max(a["max_rel"], float((d[nz] / np.abs(ref[nz])).max()))
    green = res["read_counts_exact"] and res["count_tensors_exact"] and res["log_tables_exact"]
    for t, a in acc.items():
        a["pct_outside"] = 100.0 * a["outside"] / max(1, a["n"])
        a["pct_outside_unexplained"] = 100.0 * a["outside_unexplained"] / max(1, a["n"])
        ok = (a["outside_unexplained"] <= (1.0 - TOL["min_fraction_inside"]) * a["n"] and
              a["outside_x10_unexplained"] <= (1.0 - TOL["min_fraction_inside_x10"]) * a["n"] and a["far_outside_unexplained"] == 0)
        if t == "uncorrected":
            ok = a["max_abs"] == 0.0
        a["ok"] = bool(ok)
        green = green and ok
    if examples and os.path.isdir(os.path.join(ROOT, "gpurun_out")):      # diagnostic: the windows themselves, for offline study
        with open(os.path.join(ROOT, "gpurun_out", "parity_stable_mismatches_lm%d.json" % lm_exact), "w") as fh:
            json.dump(examples, fh)
    fit["pct_unstable_flips"] = 100.0 * fit["unstable_flips"] / max(1, fit["fitted_ref"] + fit["fallback_ref"])
    # budget of the one-pass fits on windows the reference fits stably: no fit-vs-fallback decision may differ; at most 1 window in 10 000 may
    # differ by more than FIT_TOL in its normalised prediction (a tenth of that, weight 1 / 10, reaches the track -- already counted above)
    fit["stable_mismatch_budget"] = int(np.ceil(1e-4 * max(1, fit["fitted_ref"] + fit["fallback_ref"])))
    green = (green and fit["stable_decision_flips"] == 0 and fit["stable_mismatches"] <= fit["stable_mismatch_budget"]
             and fit["regions_with_call_count_mismatch"] == 0)
    res["tracks"] = acc
    res["lm_windows"] = fit
    res["correction_factor"] = bias.correction_factor
    res["regions_compared"] = len(kept["regions"])
    res["bp_compared"] = acc["corrected"]["n"]
    res["green"] = bool(green)
    res["note"] = ("green = counts / tensors / tables bit-exact; every window whose fit differs from the reference's beyond the tolerance is one on "
                   "which the reference does not keep its own answer under a 1e-13 .. 1e-12 change of its bias input (`unstable_flips`; none on a stable "
                   "window beyond `stable_mismatch_budget` = 1 in 10 000 windows, and no fit-vs-fallback decision differs on a stable window); and outside the reach of the unstable windows (`*_unexplained`) every track is inside the tolerance.  `outside` / "
                   "`far_outside` count ALL positions, unstable windows included")
    return res


def _reference_is_unstable(x, y):
    """scipy's curve_fit as the reference calls it (atacorrect_functions.py:185-188, :292-299) on the window's bias x and on x changed by
    1e-13 .. 1e-12 (scaled up / down by 1e-13 and 1e-12, shifted by +-2e-13 -- the size of the difference between this library's DWM bias
    track and the reference's: max 1.9e-13 absolute on the golden data, tests/test_kernels.py):
    True when the fit-vs-fallback decision changes or the normalised prediction moves by more than FIT_TOL / 10 between the seven fits (the
    one-pass fits differ from scipy by a different summation order on top of the 1e-13 of the input, hence the margin: a window whose
    answer moves by 1e-5 under the smallest representable change of its input is not one whose fourth digit means anything)."""
    import warnings
    from scipy.optimize import curve_fit, OptimizeWarning

    def relu(x, a, b):
        return np.maximum(0.0, a * x + b)
    seen = []
    for xx in (x, x * (1.0 + 1e-13), x * (1.0 - 1e-13), x * (1.0 + 1e-12), x * (1.0 - 1e-12), x + 2e-13, x - 2e-13):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("error", OptimizeWarning)
                popt, _ = curve_fit(relu, xx, y)
            seen.append((0, float(popt[0]), float(popt[1])))
        except (OptimizeWarning, RuntimeError):
            seen.append((1, 0.0, 0.0))
    if len({m[0] for m in seen}) > 1:
        return True
    if seen[0][0] == 1:
        return False
    return any(_pred_distance(x, seen[0][1], seen[0][2], o[1], o[2]) > FIT_TOL / 10.0 for o in seen[1:])


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        res = cpu_reference(args, steps=args.steps, warmup=args.warmup)
    except Exception as ex:
        print(json.dumps({"impl": "reference", "unavailable": repr(ex)}))
        return
    out = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": bench_config(res["wl"], args.gpus),
           "cpu_baseline": res["cpu_baseline"],
           "e2e": {"value": res["value"], "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


_NUMA = {"cpus": 0}


def bench_config(wl, world):
    """`config` of the bench line: the same keys on both arms."""
    return {"host_affinity": ("rank 0 bound to the %d CPUs next to its GPU (NVML)" % _NUMA["cpus"]) if _NUMA["cpus"] else "unbound",
            "workload": wl["desc"], "score_mat": "DWM", "k_flank": K_FLANK, "window": WINDOW, "footprint": FP, "lm_exact_order": LM_EXACT,
            "sharding": "contigs over %d rank(s), count all-reduce only" % world,
            "l2": "inputs (2-bit genome + read columns + per-position arrays) far exceed the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------- configs[3]: whole-genome footprint scan
FPG_VARIANTS = {"cli_default_flank10-30_fp20-50_651": dict(flank_min=10, flank_max=30, fp_min=20, fp_max=50),
                "flank100_fp20-50_31": dict(flank_min=100, flank_max=100, fp_min=20, fp_max=50)}
FPG_HEADLINE = "cli_default_flank10-30_fp20-50_651"


def fp_track(key, length, device):
    """Synthetic corrected-like track of one contig: N(0, 0.5) rounded to 5 decimals, ~70 % exact zeros, float32."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(np.random.SeedSequence([SEED, 77, key]).generate_state(1)[0]))
    v = torch.randn(length, generator=g, device=device, dtype=torch.float32) * 0.5
    v = torch.round(v * 1e5) / 1e5
    v[torch.rand(length, generator=g, device=device) < 0.7] = 0.0
    return v


def _fp_ref_worker(task):
    arr, p = task
    from oracle import ref_loader
    ref = ref_loader.load()
    return ref.signals.tobias_footprint_array(arr.astype(np.float64), p["flank_min"], p["flank_max"], p["fp_min"], p["fp_max"]).astype(np.float32)


def fp_reference(track, variant, seconds_per_core=10.0):
    """The reference's tobias_footprint_array (signals.pyx:184-230) over multiprocessing on all host cores, on chunks of the track."""
    import multiprocessing as mp
    from oracle import ref_loader
    ref_loader.load()
    p = FPG_VARIANTS[variant]
    cores = os.cpu_count() or 1
    rate = 7000.0 if p["flank_max"] - p["flank_min"] > 0 else 50000.0            # bp/s/core measured at survey time: sizes the sample only
    chunk = 20000
    n_chunks = cores * max(1, int(round(rate * seconds_per_core / chunk)))
    rng = np.random.default_rng(7)
    starts = np.sort(rng.choice(max(1, (track.shape[0] - chunk) // chunk), size=min(n_chunks, max(1, (track.shape[0] - chunk) // chunk)), replace=False)) * chunk
    arrs = [np.ascontiguousarray(track[a:a + chunk]) for a in starts]
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        outs = pool.map(_fp_ref_worker, [(a, p) for a in arrs], chunksize=max(1, len(arrs) // (cores * 4)))
        dt = time.perf_counter() - t0
    bp = sum(a.shape[0] for a in arrs)
    return {"value": bp / dt, "seconds": dt, "bp": bp, "cores": cores, "starts": starts, "chunk": chunk, "in": arrs, "out": outs}


def bench_fp_genome(args):
    """BASELINE.json configs[3]: ScoreBigwig --score footprint over a whole-genome corrected track (regions = whole contigs)."""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from tobias_b200 import lib, synth
    lengths = {n: ln for n, ln in synth.HG38_LENGTHS.items() if n != "chrM"}
    if args.fp_scale < 1.0:
        lengths = {n: max(200000, int(ln * args.fp_scale)) for n, ln in lengths.items()}
    owner = shard_contigs(lengths, world)
    names = [n for n in lengths if owner[n] == rank]
    keys = {n: i for i, n in enumerate(lengths)}
    ctx = lib.Context(local)
    for kv in args.opt:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    lens = np.array([lengths[n] for n in names], dtype=np.int64)
    host = ctx.pinned_empty((int(lens.sum()),), np.float32)
    off = 0
    for n in names:
        host[off:off + lengths[n]] = fp_track(keys[n], lengths[n], "cuda").cpu().numpy()
        off += lengths[n]
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    results = {}
    sampler = ClockSampler(local)
    sampler.start()
    launches = 0
    out_all = ctx.pinned_empty((int(lens.sum()),), np.float32)
    for name, p in FPG_VARIANTS.items():
        params = ctx.score_params("footprint", flank=p["flank_max"], **p)
        out_bp = int((lens - 2 * p["flank_max"]).sum())
        out = out_all[:out_bp]
        ctx.signal_upload(host, lens)
        for _ in range(args.warmup):
            ctx.score_run(params)
        barrier()
        l0 = ctx.launch_count()
        ctx.timer_begin()
        k4 = 0.0
        for _ in range(args.steps):
            ctx.score_run(params)
            k4 += ctx.last_kernel_ms()
        ms = ctx.timer_end()
        barrier()
        launches += ctx.launch_count() - l0
        for _ in range(max(1, min(args.warmup, 2))):
            ctx.signal_upload(host, lens)
            ctx.score_run(params)
            ctx.score_download(out=out)
        barrier()
        ctx.timer_begin()
        for _ in range(args.steps):
            ctx.signal_upload(host, lens)
            ctx.score_run(params)
            ctx.score_download(out=out)
        ms_e2e = ctx.timer_end()
        barrier()
        t = torch.tensor([ms, ms_e2e, k4], dtype=torch.float64, device=device)
        b = torch.tensor([out_bp], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(b)
        results[name] = {"value": float(b[0]) * args.steps / (float(t[0]) * 1e-3), "ms_per_step": float(t[0]) / args.steps,
                         "k4_kernel_ms": float(t[2]) / args.steps, "e2e_value": float(b[0]) * args.steps / (float(t[1]) * 1e-3),
                         "e2e_ms_per_step": float(t[1]) / args.steps, "scored_bp": int(b[0]), "combinations": (p["flank_max"] - p["flank_min"] + 1) * (p["fp_max"] - p["fp_min"] + 1),
                         "h2d_bytes_per_step": int(host.nbytes), "d2h_bytes_per_step": int(out.nbytes),
                         "out_first": np.array(out[:int(lens[0]) - 2 * p["flank_max"]]) if (world == 1 and not args.no_cpu_baseline) else None}
    clocks = sampler.stop()
    head = results[FPG_HEADLINE]
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    achieved = 8.0 * (head["scored_bp"] / world) / (head["k4_kernel_ms"] * 1e-3) / 1e9
    result = {"metric": METRIC, "value": head["value"], "unit": "bp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": "ScoreBigwig multi-width footprint scan over a whole-genome corrected track (BASELINE.json configs[3]): %d contigs, %.2f Gb, "
                                     "regions = whole contigs; headline = CLI defaults (flank 10-30 x fp 20-50: 651 combinations); `variants` also holds "
                                     "flank 100 x fp 20-50 (31 combinations)" % (len(lengths), sum(lengths.values()) / 1e9),
                         "footprint": FPG_VARIANTS[FPG_HEADLINE], "sharding": "contigs over %d rank(s), no collective" % world,
                         "l2": "the track (12 GB) far exceeds the 126 MB L2; no flush needed", **({"options": args.opt} if args.opt else {})},
              "clocks": clocks, "gpu_launches": int(launches),
              "e2e": {"value": head["e2e_value"], "unit": "bp/s", "ms_per_step": head["e2e_ms_per_step"], "h2d_bytes_per_step": head["h2d_bytes_per_step"],
                      "d2h_bytes_per_step": head["d2h_bytes_per_step"]},
              "roofline": {"bound": "hbm", "kernel": "K4_footprint", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                           "algorithmic_bytes_per_launch": 8.0 * head["scored_bp"] / world,
                           "note": "K4 is bound by fp32 issue / shared-memory bandwidth (651 add-max pairs per footprint start), not by HBM: 8 B per scored bp"},
              "variants": {k: {kk: vv for kk, vv in v.items() if not kk.startswith("out")} for k, v in results.items()}}
    if world == 1 and not args.no_cpu_baseline:
        try:
            # reference on chunks of the first contig; the same chunks through the GPU as regions of their own, and the whole-contig GPU
            # scan at the chunks' interior positions (tile borders of a 248-Mb region), both against the reference's arrays
            n0 = int(lens[0])
            track0 = np.asarray(host[:n0])
            parity = {"tolerance": "identical zero pattern; |gpu - ref| <= 1e-5 |ref| on every non-zero position (float32 outputs)"}
            base = {}
            for name, p in FPG_VARIANTS.items():
                rf = fp_reference(track0, name, seconds_per_core=8.0 if name == FPG_HEADLINE else 4.0)
                base[name] = {"value": rf["value"], "unit": "bp/s", "cores": rf["cores"], "kind": "reference",
                              "sample": "%d chunks of %d bp of %s (%d bp) through tobias_footprint_array on %d cores in %.1f s" % (
                                  len(rf["in"]), rf["chunk"], names[0], rf["bp"], rf["cores"], rf["seconds"])}
                params = ctx.score_params("footprint", flank=0, **p)
                got = ctx.score_regions(np.concatenate(rf["in"]), [a.shape[0] for a in rf["in"]], params)
                ref = np.concatenate(rf["out"])
                nz = ref != 0
                ok_chunks = bool(np.array_equal(got == 0, ref == 0) and np.all(np.abs(got[nz] - ref[nz]) <= 1e-5 * np.abs(ref[nz])))
                whole = results[name]["out_first"]                      # scores of [flank_max, len - flank_max) of the first contig
                margin = 2 * (p["flank_max"] + p["fp_max"])
                ok_whole, n_cmp = True, 0
                for a, o in zip(rf["starts"], rf["out"]):
                    lo, hi = int(a) + margin, int(a) + rf["chunk"] - margin
                    w = whole[lo - p["flank_max"]:hi - p["flank_max"]]
                    r = o[margin:rf["chunk"] - margin]
                    nzr = r != 0
                    ok_whole = ok_whole and bool(np.array_equal(w == 0, r == 0) and np.all(np.abs(w[nzr] - r[nzr]) <= 1e-5 * np.abs(r[nzr])))
                    n_cmp += r.shape[0]
                parity[name] = {"chunks_as_regions_ok": ok_chunks, "whole_contig_interior_ok": ok_whole, "bp_compared": int(ref.shape[0]), "interior_bp_compared": n_cmp}
            parity["green"] = all(v["chunks_as_regions_ok"] and v["whole_contig_interior_ok"] for k, v in parity.items() if isinstance(v, dict))
            result["cpu_baseline"] = base[FPG_HEADLINE]
            result["cpu_baseline_variants"] = base
            result["parity"] = parity
        except Exception as ex:
            result["cpu_baseline"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(result))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def bench_fp_genome_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from tobias_b200 import synth
    n = min(synth.HG38_LENGTHS["chr21"], 8_000_000)
    track = fp_track(list(synth.HG38_LENGTHS).index("chr21"), n, "cpu").numpy()
    vals = []
    for it in range(args.warmup + args.steps):
        rf = fp_reference(track, FPG_HEADLINE, seconds_per_core=6.0)
        if it >= args.warmup:
            vals.append(rf)
    v = float(np.mean([r["value"] for r in vals]))
    r = vals[-1]
    base = {"value": v, "unit": "bp/s", "cores": r["cores"], "kind": "reference",
            "sample": "%d chunks of %d bp (%d bp) through tobias_footprint_array (651 combinations) on %d cores in %.1f s per step" % (
                len(r["in"]), r["chunk"], r["bp"], r["cores"], r["seconds"])}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * r["seconds"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "ScoreBigwig multi-width footprint scan (BASELINE.json configs[3]), CLI defaults: 651 combinations",
                                 "footprint": FPG_VARIANTS[FPG_HEADLINE]},
                      "cpu_baseline": base, "e2e": {"value": v, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------- configs[4]: batch of samples
def bench_batch24(args):
    """BASELINE.json configs[4]: 24 pseudobulk samples (5 M records each) over one genome and one peak set, 3 samples per GPU, each through
    its own bias training + correction + footprint scoring (the reference runs one ATACorrect per BAM, atacorrect.py:53-487).  Replicas mode:
    the genome is resident once per GPU, every sample has its own bias model, no collective.  Weak scaling: 3 samples per GPU at every N."""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from tobias_b200 import lib, synth_torch
    from tobias_b200.io.fasta import pack_ascii
    from tobias_b200.pipeline import Engine, prepare_regions
    from tobias_b200.regions import OneRegion, RegionList, to_arrays
    per_gpu, n_rec = 3, 5_000_000
    wl = workload("hg38")
    t_setup = time.time()
    ds = synth_torch.TorchDataset(SEED, wl["lengths"], wl["n_peaks"], n_rec, device="cuda", pin=False, keep_codes=True)
    sample_ids = [rank * per_gpu + j for j in range(per_gpu)]
    read_sets = [ds.reads_for(SEED + 1000 + sid, n_rec) for sid in sample_ids]
    ds._state = None
    torch.cuda.empty_cache()
    names = ds.names
    chrom_info = {n: ln for n, ln in zip(names, ds.lengths) if n != "chrM"}
    regs = prepare_regions(RegionList([OneRegion(list(p)) for p in ds.peaks]), chrom_info, [n for n in names if n in chrom_info], extend=EXTEND,
                           k_flank=K_FLANK, window=WINDOW)
    index = {n: i for i, n in enumerate(names)}
    arrays = {k: to_arrays(sorted(regs[k], key=lambda r: (index[r.chrom], r.start, r.end)), index)
              for k in ("peak_regions", "nonpeak_regions", "input_regions", "output_regions")}
    eng = Engine(local)
    ctx = eng.ctx
    for kv in args.opt:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    eng.names, eng.index, eng.lengths = names, index, ds.lengths
    ctx.genome_create(ds.lengths)
    for i, g in enumerate(ds.genome):                        # the genome: once per GPU, untimed (shared by every sample of the batch)
        sq, nm = pack_ascii(g)
        ctx.genome_upload_packed(i, sq, nm, int(g.shape[0]))
    hosts = []
    for rc in read_sets:
        flags = (rc.is_reverse.astype(np.uint8) * lib.TB_READ_REVERSE | rc.five_prime_is_match().astype(np.uint8) * lib.TB_READ_5P_MATCH
                 | (~rc.usable()).astype(np.uint8) * lib.TB_READ_SKIP)
        comp = None if args.wide_inputs else lib.compact_read_columns(rc.chrom_id, rc.ref_start, rc.ref_end, rc.clip5(), flags, len(ds.lengths))
        cols = []
        for a in (comp if comp is not None else (rc.chrom_id, rc.ref_start, rc.ref_end, rc.clip5(), flags)):
            p = ctx.pinned_empty(a.shape, a.dtype)
            p[...] = a
            cols.append(p)
        hosts.append({"lengths": ds.lengths, "reads": cols, "reads_compact": comp is not None, "packed": None, "genome": None})
    c, s, e = arrays["output_regions"]
    out_bp = int((e - s).sum())
    out = {"tracks": {t: ctx.pinned_empty((out_bp,), np.float32) for t in lib.TRACKS}, "footprint": ctx.pinned_empty((out_bp,), np.float32)}
    setup_s = time.time() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_batch(download):
        acc = {}
        for h in hosts:                                      # a sample: its records in, its own model, its tracks (and footprints) out
            if download:
                _b, st, _p = run_step(eng, arrays, device, False, h, out, genome_resident=True, collective=False)
            else:
                (ctx.reads_upload_compact if h.get("reads_compact") else ctx.reads_upload)(*h["reads"])
                _b, st, _p = run_step(eng, arrays, device, True, h, out, collective=False)
            for k, v in st.items():
                acc[k] = acc.get(k, 0.0) + v
        return acc

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        one_batch(False)
    barrier()
    l0 = ctx.launch_count()
    ctx.timer_begin()
    stages = {}
    for _ in range(args.steps):
        for k, v in one_batch(False).items():
            stages[k] = stages.get(k, 0.0) + v / args.steps
    ms = ctx.timer_end()
    barrier()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    for _ in range(max(1, min(args.warmup, 2))):
        one_batch(True)
    barrier()
    ctx.timer_begin()
    for _ in range(args.steps):
        one_batch(True)
    ms_e2e = ctx.timer_end()
    barrier()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_samples = per_gpu * world
    total_bp = float(out_bp) * n_samples
    value, e2e_value = total_bp * args.steps / (float(t[0]) * 1e-3), total_bp * args.steps / (float(t[1]) * 1e-3)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ext_bp = float((e - s + 124).sum())
    dom = max(stages, key=lambda k: stages[k])
    alg = {"K1_pileup": 5.0 * n_rec + 8.0 * ext_bp, "K2_counts": 5.0 * n_rec, "K3a_score": 16.375 * ext_bp, "K3b_fit": 24.0 * ext_bp,
           "K3c_correct": 25.9 * out_bp, "K4_footprint": 8.0 * out_bp}
    achieved = per_gpu * alg[dom] / (stages[dom] * 1e-3) / 1e9
    reads_bytes = sum(int(a.nbytes) for h in hosts for a in h["reads"])
    result = {"metric": METRIC, "value": value, "unit": "bp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": float(t[0]) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": dict(bench_config(wl, world), workload="batch of %d pseudobulk samples (5M records each) over one synthetic hg38-sized genome and 150k peaks, "
                             "%d samples per GPU, bias training + correction + footprints per sample (BASELINE.json configs[4])" % (n_samples, per_gpu),
                             sharding="%d samples per GPU on %d GPU(s): replicas, genome resident once per GPU, one bias model per sample, no collective" % (per_gpu, world),
                             samples=n_samples, out_bp_per_sample=out_bp, **({"options": args.opt} if args.opt else {})),
              "samples_per_s": n_samples * args.steps / (float(t[0]) * 1e-3), "clocks": clocks, "gpu_launches": int(launches),
              "e2e": {"value": e2e_value, "unit": "bp/s", "ms_per_step": float(t[1]) / args.steps, "samples_per_s": n_samples * args.steps / (float(t[1]) * 1e-3),
                      "h2d_bytes_per_step": reads_bytes + per_gpu * sum(int(x.nbytes) for v in arrays.values() for x in v), "d2h_bytes_per_step": per_gpu * (5 * out_bp * 4 + 1500),
                      "note": "a step = the GPU's %d samples; every sample: its record columns H2D from pinned memory, four float32 tracks + footprints D2H; the genome "
                              "(shared by the batch) is uploaded once per GPU outside the timed region" % per_gpu},
              "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                           "algorithmic_bytes_per_launch": alg[dom]},
              "stages": {k: {"ms_per_batch": round(v, 3)} for k, v in stages.items()}, "setup_s": round(setup_s, 1)}
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        try:
            a2 = argparse.Namespace(**vars(args))
            a2.workload = "batch24"
            result["cpu_baseline"] = cpu_reference(a2, steps=1, warmup=0)["cpu_baseline"]
        except Exception as ex:
            result["cpu_baseline"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(result))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------- the CLI from files (host I/O included)
def bench_cli(args):
    """`TOBIAS ATACorrect` + `TOBIAS ScoreBigwig` end to end FROM FILES on one GPU: synthetic FASTA / BAM / BED written to a scratch directory,
    then the two drivers as a user runs them (BGZF inflate + record decode, genome packing, kernels, bigWig writing with zoom levels).
    Prints wall seconds with a per-phase split -- the measure of whether the on-disk formats cap the GPU."""
    import shutil
    import torch
    from tobias_b200 import atacorrect, parsers, score_bigwig, synth, synth_torch
    from tobias_b200.io.bam import write_bam_fast
    from tobias_b200.io.fasta import write_fasta
    wl = workload(args.workload)
    d = tempfile.mkdtemp(prefix="tb_cli_")
    t0 = time.perf_counter()
    ds = synth_torch.TorchDataset(SEED, wl["lengths"], wl["n_peaks"], wl["n_records"], device="cuda" if torch.cuda.is_available() else "cpu")
    fa, bam_f, bed = os.path.join(d, "genome.fa"), os.path.join(d, "reads.bam"), os.path.join(d, "peaks.bed")
    write_fasta(fa, zip(ds.names, ds.genome))
    write_bam_fast(bam_f, ds.reads)
    synth.write_bed(bed, ds.peaks)
    t_files = time.perf_counter() - t0
    sizes = {"fasta": os.path.getsize(fa), "bam": os.path.getsize(bam_f), "records": len(ds.reads)}
    del ds
    runs = []
    for it in range(2):                                        # first run: packs the genome (.tb2bit) and compiles nothing; second: caches warm
        out = os.path.join(d, "out%d" % it)
        a = parsers.add_atacorrect_arguments(argparse.ArgumentParser()).parse_args(
            ["--bam", bam_f, "--genome", fa, "--peaks", bed, "--outdir", out, "--prefix", "x", "--verbosity", "0"])
        t0 = time.perf_counter()
        res = atacorrect.run_atacorrect(a)
        t_ata = time.perf_counter() - t0
        b = parsers.add_scorebigwig_arguments(argparse.ArgumentParser()).parse_args(
            ["--signal", os.path.join(out, "x_corrected.bw"), "--regions", bed, "--output", os.path.join(out, "x_footprints.bw"), "--verbosity", "0"])
        t0 = time.perf_counter()
        res_sb = score_bigwig.run_scorebigwig(b)
        t_sb = time.perf_counter() - t0
        runs.append({"atacorrect_s": round(t_ata, 2), "scorebigwig_s": round(t_sb, 2), "atacorrect_phases_s": res["phases_s"],
                     "scorebigwig_phases_s": res_sb["phases_s"],
                     "output_bytes": {f: os.path.getsize(os.path.join(out, f)) for f in sorted(os.listdir(out))}})
    shutil.rmtree(d, ignore_errors=True)
    print(json.dumps({"metric": "cli_wall_s", "workload": wl["desc"], "cli_wall_s": runs[-1]["atacorrect_s"] + runs[-1]["scorebigwig_s"],
                      "runs": runs, "input_files": sizes, "synthetic_files_written_in_s": round(t_files, 1), "host_cores": os.cpu_count(),
                      "note": "run 0 builds the .tb2bit genome cache, run 1 reuses it; both include BGZF inflate + record decode, bigWig compression and zoom levels"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hg38", choices=["hg38", "chr50mb", "small", "fp_genome", "batch24"])
    ap.add_argument("--fp-scale", type=float, default=1.0, help="fp_genome: scale every contig length (quick runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wide-inputs", action="store_true", help="e2e: five-column records and N-mask words instead of the compact host forms (comparison runs)")
    ap.add_argument("--lm-exact", action="store_true")
    ap.add_argument("--fit-stats", action="store_true")
    ap.add_argument("--lm-mode", type=int, default=None, help="5: moment form of the one-pass fits (comparison runs); 2 / 3 need a -DTB_FIRST_GEN_FITS build")
    ap.add_argument("--cli", action="store_true", help="time `TOBIAS ATACorrect` + `TOBIAS ScoreBigwig` from files (host I/O included); prints its own JSON line")
    ap.add_argument("--ascii-genome", action="store_true", help="e2e: upload the genome as ASCII (packed on the device) instead of the host-packed 2-bit form")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="tb_set_option switch for comparison runs (recorded in config)")
    args = ap.parse_args()
    global LM_EXACT
    LM_EXACT = args.lm_mode if args.lm_mode is not None else int(bool(args.lm_exact))
    if args.cli:
        bench_cli(args)
    elif args.workload == "fp_genome":
        (bench_fp_genome_reference if args.impl == "reference" else bench_fp_genome)(args)
    elif args.workload == "batch24":
        (bench_reference if args.impl == "reference" else bench_batch24)(args)
    elif args.impl == "reference":
        bench_reference(args)
    else:
        bench_b200(args)


if __name__ == "__main__":
    main()
