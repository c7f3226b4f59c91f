#!/usr/bin/env python
"""bench.py -- candidate sites/sec of the s1+s2 hot path (BASELINE.json metric) on synthetic ONT-like reads.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (N>1: launched under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's CPU path on a bounded sample

One "step" = one pass of pileup -> candidate select -> window gather -> PileupModel forward over one synthetic
contig (default: BASELINE.json configs[1], 100 Mb at 30x, per GPU; weak scaling: every rank owns one contig).
  value  device-resident inputs, CUDA-event timed, max over ranks
  e2e    same work through the host-buffer API: pinned host arrays -> H2D -> kernels -> D2H of the call list
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "candidate sites/sec (s1+s2)"
FLOP_PER_SITE = 6.22e6           # exact-output minimal PileupModel FLOPs, SURVEY 8(d)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--contig-mb", type=float, default=100.0)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--region-mb", type=float, default=12.5)
    ap.add_argument("--cpu-sample-kb", type=float, default=500.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "f16x3"])
    return ap.parse_args()


def synth_cfg(args, rank, contig_len):
    from nanosnp_b200.synth import SynthConfig
    return SynthConfig(contig_len=int(contig_len), coverage=args.coverage, contig=f"ctg{rank + 1}",
                       seed_ref=1000 + rank, seed_var=2000 + rank, seed_reads=3000 + rank)


def load_weights():
    from nanosnp_b200.utils import load_weights_npz     # the committed checkpoint fixture (flat .npz of the shipped .chkpt)
    return load_weights_npz(ROOT / "tests" / "golden" / "ont_pileup_weights.npz")


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(args, steps, warmup):
    """Times the reference's CPU s1+s2 path (oracle/_ref binaries + CPU predict logic) on a bounded sample."""
    from nanosnp_b200.synth import generate_host
    from oracle.cpu_path import run_cpu_path
    cores = os.cpu_count() or 1
    cfg = synth_cfg(args, 0, args.cpu_sample_kb * 1e3)
    ref, reads = generate_host(cfg)
    weights = load_weights()
    times, res = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res = run_cpu_path(reads, ref, cfg.contig, weights, cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return {"value": res["n_sites"] / mean, "seconds_per_step": mean, "cores": cores, "kind": res["kind"],
            "n_sites": res["n_sites"], "stage_s": res["stage_s"],
            "sample": f"{args.cpu_sample_kb:.0f} kb synthetic contig at {args.coverage:g}x (same generator/config as the GPU workload), "
                      f"{res['n_sites']} sites; stages: mpileup restatement (ours, not samtools) + reference DNA_CreateCanSnpTensor/"
                      f"DNA_CreatePredictData ({res['kind']}) + text->int32 + predict.py logic on torch CPU fp32"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args, args.steps, min(args.warmup, 1) if args.warmup else 0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "sites/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 counts + fp32 model", "data": "synthetic",
            "config": {"workload": f"synthetic contig at {args.coverage:g}x, s1+s2 on host CPU; bounded sample", "sample_kb": args.cpu_sample_kb},
            "cpu_baseline": {"value": r["value"], "unit": "sites/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                             "stage_s": r["stage_s"]},
            "e2e": {"value": r["value"], "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Samples taken inside [t_begin, t_end] (the timed region); nvidia-smi needs a few hundred ms to start, so the
        sampler is started before the warm-up and, if the timed region is shorter than one sampling period, the
        samples of the load immediately before it (same kernels, same clocks) are used and the window says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = "timed region"
        lines = [l for (t, l) in self.lines if t_begin is None or (t_begin <= t <= t_end + 0.12)]
        if not lines and t_begin is not None:
            lines = [l for (t, l) in self.lines if t <= t_end + 0.12][-5:]
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        self.window = window
        for l in lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from nanosnp_b200 import _lib
    from nanosnp_b200.pipeline import PileupEngine, PileupModelForward, PileupModelWeights
    from nanosnp_b200.reads import FIELDS, PackedReads
    from nanosnp_b200.runner import RegionRunner, StageTimer
    from nanosnp_b200.shard import plan_regions, read_range_for_region
    from nanosnp_b200.synth import generate_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep stdout to the one JSON line (NCCL prints its version at INFO/VERSION)
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic workload: one contig per rank, generated on the GPU, cut into regions with 16-bp halos ----
    L = int(args.contig_mb * 1e6)
    cfg = synth_cfg(args, rank, L)
    ref, reads = generate_device(cfg, dev)
    regions = plan_regions([(cfg.contig, L)], int(args.region_mb * 1e6))
    pos_host = reads.pos.cpu().numpy()
    max_span = int(cfg.len_max * 1.3) + 1000
    n_total = reads.n_reads
    total_bases = int(reads.seq2.numel()) * 4
    dev_regions, host_regions, alg = [], [], {"pileup": 0.0, "n_bases": 0, "n_cigar": 0, "n_reads": 0}
    for rg in regions:
        lo, hi = read_range_for_region(pos_host, max_span, rg)
        c0, c1 = int(reads.cigar_off[lo].item()), int(reads.cigar_off[hi].item())
        b0 = int(reads.seq_off[lo].item()) if lo < n_total else total_bases
        b1 = int(reads.seq_off[hi].item()) if hi < n_total else total_bases - 64
        pad = torch.zeros(16, dtype=torch.uint8, device=dev)
        sl = PackedReads(reads.pos[lo:hi].clone(), reads.flag[lo:hi].clone(), reads.mapq[lo:hi].clone(),
                         (reads.cigar_off[lo:hi + 1] - c0), reads.cigar[c0:c1].clone(), (reads.seq_off[lo:hi] - b0),
                         torch.cat([reads.seq2[b0 // 4:(b1 + 3) // 4], pad]),
                         None if reads.nmask is None else torch.cat([reads.nmask[b0 // 8:(b1 + 7) // 8], pad]))
        dev_regions.append(sl)
        alg["n_bases"] += b1 - b0; alg["n_cigar"] += c1 - c0; alg["n_reads"] += hi - lo
        alg["pileup"] += 0.25 * (b1 - b0) + 4.0 * (c1 - c0) + 23.0 * (hi - lo) + 74.0 * rg.length     # SURVEY 8(d) B_A (+ ref byte)
    del reads
    torch.cuda.empty_cache()

    eng = PileupEngine(dev)
    enc, fwd = load_weights()
    prec = _lib.PREC_FP32 if args.precision == "fp32" else _lib.PREC_F16X3
    model = PileupModelForward(PileupModelWeights(enc, fwd, device=dev), precision=prec)
    runner = RegionRunner(eng, model)
    # fewer than 4 host cores per rank: host waits sleep instead of spinning, the core goes to the VCF text assembly
    runner_e2e = RegionRunner(eng, model, records=True, blocking_sync=(os.cpu_count() or 1) < 4 * world)      # e2e: numeric record logic on the GPU, 32 B/site D2H

    def step_device(timer=None):
        n = 0
        for rg, rd in zip(regions, dev_regions):
            n += runner.run_device(rd, ref, rg, timer).n
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident timing ----
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        n_sites = step_device()
    barrier()
    timer = StageTimer(True)
    lib = _lib.load()
    lib.nsnp_profile_enable(1)
    launches0 = runner.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        n_sites = step_device(timer)
    e1.record()
    barrier()
    t_end = time.time()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end)
    stage_ms = timer.totals_ms()
    stage_calls = timer.counts()
    import ctypes as C
    kms = (C.c_double * len(_lib.PROF_SLOTS))(); kln = (C.c_int64 * len(_lib.PROF_SLOTS))()
    _lib.check(lib.nsnp_profile_read(kms, kln))
    lib.nsnp_profile_enable(0)
    kernel_ms = {k: kms[i] for i, k in enumerate(_lib.PROF_SLOTS)}
    kernel_launches = {k: int(kln[i]) for i, k in enumerate(_lib.PROF_SLOTS)}
    # kernels of this library launched inside the timed region, counted by the library itself (the select slot is three kernels)
    launches = sum(kernel_launches.values()) + 2 * kernel_launches.get("select_kernels", 0)

    # ---- end to end through the host-buffer API ----
    e2e = None
    if not args.no_e2e:
        for rd in dev_regions:
            host_regions.append(PackedReads(*[None if getattr(rd, f) is None else getattr(rd, f).cpu().pin_memory() for f in FIELDS]))
        h2d_bytes = sum(h.nbytes() for h in host_regions)
        capn = max(1024, max(rg.emit_end - rg.emit_start for rg in regions) // 3)

        def pinned_out():
            return {"rec": torch.empty((capn, 32), dtype=torch.uint8).pin_memory()}
        host_outs = (pinned_out(), pinned_out())

        from concurrent.futures import ThreadPoolExecutor
        from nanosnp_b200.predict import ContigVcfAssembler
        pool = ThreadPoolExecutor(1)
        vcf_bytes = [0]

        def step_host(with_vcf=True):
            # H2D of region k+1 and D2H of region k-1 overlap the kernels of region k (copy streams + double buffers);
            # the VCF text of region k-1 is formatted on host threads (nsnp_vcf_format_contig) meanwhile
            asm = ContigVcfAssembler(cfg.contig, 1000, max(1, (os.cpu_count() or 1) // world), None)     # host cores are shared by the ranks

            def consume(k, res):
                if not with_vcf:
                    return None
                return pool.submit(asm.add_records, res["rec"].numpy())
            n = runner_e2e.run_host_many(host_regions, regions, ref, host_outs, consume)
            vcf_bytes[0] = pool.submit(asm.close).result()
            return n
        step_host()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record()
        for _ in range(args.steps):
            n_e2e = step_host()
        f1.record()
        barrier()
        wall = time.perf_counter() - t0
        e2e_ms = max(f0.elapsed_time(f1), wall * 1e3)
        d2h_bytes = n_e2e * 32
        e2e = {"ms": e2e_ms, "sites": n_e2e, "h2d": h2d_bytes, "d2h": d2h_bytes, "vcf_bytes": vcf_bytes[0]}

    # ---- reduce over ranks: max time, total sites ----
    vals = torch.tensor([ms_total, float(n_sites), e2e["ms"] if e2e else 0.0, float(e2e["sites"]) if e2e else 0.0],
                        dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, sites_all = float(mx[0]), float(sm[1])
        e2e_ms, e2e_sites = float(mx[2]), float(sm[3])
    else:
        ms_total, sites_all = float(vals[0]), float(vals[1])
        e2e_ms, e2e_sites = float(vals[2]), float(vals[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk_path = ROOT / "MEASURED_PEAKS.json"
    if pk_path.exists():
        peaks = json.loads(pk_path.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    ms_step = ms_total / args.steps
    K = args.steps

    def hbm(stage, bytes_per_step):
        ms = stage_ms[stage] / K
        a = bytes_per_step / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak, "ms_per_step": ms,
                "launches_per_step": stage_calls[stage] // K, "algorithmic_bytes_per_step": bytes_per_step, "traffic": None}
    Lr = sum(rg.length for rg in regions)
    stages = {
        "pileup": hbm("pileup", alg["pileup"]),
        "select": hbm("select", 1.0 * Lr + 4.0 * n_sites),
        "gather": hbm("gather", 4753.0 * n_sites),
    }
    # ---- roofline of the dominant kernel (largest share of the timed region), timed live with CUDA events around
    #      each of its launches on the launching stream (nsnp_profile_*) ----
    traffic = {}
    tr_path = ROOT / "profiles" / "traffic.json"
    if tr_path.exists():
        traffic = json.loads(tr_path.read_text())
    flop_per_site = {"lstm_layer0": 2 * 1385472.0, "lstm_layer1": 2 * 1671168.0, "tail_kernel": 2 * 55296.0}
    kernels = {}
    for name in _lib.PROF_SLOTS:
        ms = kernel_ms[name] / K
        ent = {"ms_per_step": ms, "launches_per_step": kernel_launches[name] // K, "share_of_step": ms / ms_step}
        if name in flop_per_site and ms > 0:
            tf = flop_per_site[name] * n_sites / (ms * 1e-3) / 1e12
            ent.update({"bound": "tensor" if name != "tail_kernel" else "fp32", "achieved": tf, "unit": "TFLOP/s", "peak": tf_peak, "frac": tf / tf_peak})
            if name in traffic:
                ent["traffic"] = traffic[name]["dram_bytes_per_site"] * n_sites / max(1, kernel_launches[name] // K)
        kernels[name] = ent
    dom = max(("lstm_layer0", "lstm_layer1", "pileup_tile_kernel"), key=lambda k: kernel_ms[k])
    if dom == "pileup_tile_kernel":
        d = kernels[dom]; a = alg["pileup"] / (d["ms_per_step"] * 1e-3) / 1e9
        roofline = {"kernel": "pileup_tile_kernel", "bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                    "traffic": traffic.get(dom, {}).get("dram_bytes_per_position", 0) * Lr / len(regions) or None}
    else:
        d = kernels[dom]
        roofline = {"kernel": {"lstm_layer0": "lstm0_pair2_kernel (layer-0 BiLSTM, tcgen05 cta_group::2, two alternating site groups per CTA)",
                               "lstm_layer1": "lstm_tc_kernel<1,2,4> (layer-1 BiLSTM, tcgen05 cta_group::2)"}[dom] if args.precision != "fp32" else dom,
                    "bound": "tensor", "achieved": d["achieved"], "peak": tf_peak, "unit": "TFLOP/s", "frac": d["frac"],
                    "traffic": d.get("traffic"), "traffic_note": "DRAM bytes per launch, scaled from the ncu capture in profiles/traffic.json",
                    "ms_per_step": d["ms_per_step"], "launches_per_step": d["launches_per_step"], "share_of_step": d["share_of_step"],
                    "peak_source": peak_src,
                    "note": ("fp32 FFMA path (no tensor cores): algorithmic FLOPs over the bf16 sustained peak" if args.precision == "fp32" else
                             "algorithmic (exact-minimal) FLOPs of this layer over the measured sustained bf16 peak; every algorithmic MAC "
                             "costs three fp16 MMAs (hi/lo split), so frac <= 1/3 by construction.  ncu: tensor pipe 51 % active, MUFU "
                             "(cell update: 7 ex2/rcp per unit and step) 79 % -- the layer-0 kernel is bound by the MUFU rate, layer 1 by "
                             "the sustained tensor rate; both run under sw_power_cap (DESIGN.md 4.4)")}
    model_ms = stage_ms["model"] / K
    tfs = FLOP_PER_SITE * n_sites / (model_ms * 1e-3) / 1e12 if model_ms > 0 else 0.0
    stages["model"] = {"bound": "tensor", "achieved": tfs, "peak": tf_peak, "unit": "TFLOP/s", "frac": tfs / tf_peak, "ms_per_step": model_ms,
                       "algorithmic_flop_per_site": FLOP_PER_SITE}
    stages["pileup"]["traffic"] = traffic.get("pileup_tile_kernel", {}).get("dram_bytes_per_position", 0) * Lr / len(regions) or None
    line = {
        "metric": METRIC, "value": sites_all / (ms_total * 1e-3) * K, "unit": "sites/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 counts, fp64 AF gate, " + ("fp32 model" if args.precision == "fp32" else "fp16 hi/lo split x3 tensor-core model (fp32 accumulate)"),
        "data": "synthetic",
        "config": {"workload": f"synthetic {args.contig_mb:g} Mb contig at {args.coverage:g}x per GPU: pileup tensor build + candidate filter + PileupModel inference",
                   "contig_mb": args.contig_mb, "coverage": args.coverage, "region_mb": args.region_mb, "regions_per_gpu": len(regions),
                   "sites_per_gpu_step": n_sites, "reads": alg["n_reads"], "aligned_bases": alg["n_bases"], "cigar_ops": alg["n_cigar"],
                   "weights": "shipped ont_pileup.chkpt (fp32)", "l2": "inputs (>3 GB per step) exceed the 126 MB L2; no explicit flush",
                   "seeds": [cfg.seed_ref, cfg.seed_var, cfg.seed_reads], "parallelism": f"contig-per-GPU x{world}, no data-path collective"},
        "clocks": clocks, "gpu_launches": launches,
        "roofline": roofline, "roofline_stages": stages, "kernels": kernels,
        "stage_ms_per_step": {k: v / K for k, v in stage_ms.items()},
    }
    if e2e:
        line["e2e"] = {"value": e2e_sites / (e2e_ms * 1e-3) * K, "unit": "sites/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "ms_per_step": e2e_ms / K, "vcf_bytes_per_step": e2e["vcf_bytes"],
                       "note": "pinned host read arrays -> H2D -> kernels (incl. site_record_kernel: argmax/QUAL/DP/AF per site) -> D2H of 32-byte site "
                               "records -> VCF record text (native multi-threaded text assembly, 1000-site batches per contig as predict.py) through "
                               "RegionRunner.run_host_many; copies and formatting overlap the kernels of the neighbouring regions; reference FASTA and "
                               "weights resident"}
    else:
        line["e2e"] = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args, 1, 0)
        line["cpu_baseline"] = {"value": r["value"], "unit": "sites/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                "stage_s": r["stage_s"]}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
