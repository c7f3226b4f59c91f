#!/usr/bin/env python
"""BASELINE.json configs[3]: whole-genome-shaped synthetic input (chr1-22, X, Y, M with GRCh38 lengths, 3.1 Gb) at 30x,
region-sharded across the GPUs of one box with 16-bp halos: STRONG scaling, no data-path collective.

    python tools/wgs_bench.py                                            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/wgs_bench.py

Every rank plans the same regions (plan_regions), takes its LPT share (assign_lpt), generates the reads of the contigs it
needs on its own GPU (counter-based generator: every rank sees the same genome), keeps only the slices of its regions and
then times K passes over them (device-resident, CUDA events, barrier + max over ranks).  Rank 0 prints one JSON line.
--scale shrinks every contig (e.g. 0.01 for a smoke run)."""
import argparse, json, os, sys, time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

GRCH38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
          ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
          ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
          ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
          ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415), ("chrM", 16569)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--region-mb", type=float, default=12.5)
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    from nanosnp_b200 import _lib
    from nanosnp_b200.pipeline import PileupEngine, PileupModelForward, PileupModelWeights
    from nanosnp_b200.reads import PackedReads
    from nanosnp_b200.runner import RegionRunner
    from nanosnp_b200.shard import assign_lpt, plan_regions, read_range_for_region
    from nanosnp_b200.synth import SynthConfig, generate_device
    from bench import load_weights

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)

    contigs = [(n, max(2000, int(L * args.scale))) for n, L in GRCH38]
    regions = plan_regions(contigs, int(args.region_mb * 1e6))
    mine = assign_lpt(regions, world)[rank]
    by_contig = {}
    for i in mine:
        by_contig.setdefault(regions[i].contig_index, []).append(regions[i])

    t_gen = time.perf_counter()
    work = []                                   # (region, reads slice, contig reference)
    n_bases = 0
    for ci, rgs in sorted(by_contig.items()):
        name, L = contigs[ci]
        cfg = SynthConfig(contig_len=L, coverage=args.coverage, contig=name, seed_ref=100 + ci, seed_var=200 + ci, seed_reads=300 + ci)
        ref, reads = generate_device(cfg, dev)
        pos_host = reads.pos.cpu().numpy()
        max_span = int(cfg.len_max * 1.3) + 1000
        n_total = reads.n_reads
        total_bases = int(reads.seq2.numel()) * 4
        pad = torch.zeros(16, dtype=torch.uint8, device=dev)
        for rg in rgs:
            lo, hi = read_range_for_region(pos_host, max_span, rg)
            c0, c1 = int(reads.cigar_off[lo].item()), int(reads.cigar_off[hi].item())
            b0 = int(reads.seq_off[lo].item()) if lo < n_total else total_bases
            b1 = int(reads.seq_off[hi].item()) if hi < n_total else total_bases - 64
            sl = PackedReads(reads.pos[lo:hi].clone(), reads.flag[lo:hi].clone(), reads.mapq[lo:hi].clone(),
                             (reads.cigar_off[lo:hi + 1] - c0), reads.cigar[c0:c1].clone(), (reads.seq_off[lo:hi] - b0),
                             torch.cat([reads.seq2[b0 // 4:(b1 + 3) // 4], pad]), None)
            work.append((rg, sl, ref))
            n_bases += b1 - b0
        del reads
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    eng = PileupEngine(dev)
    enc, fwd = load_weights()
    runner = RegionRunner(eng, PileupModelForward(PileupModelWeights(enc, fwd, device=dev), precision=_lib.PREC_F16X3))

    def step():
        n = 0
        for rg, sl, ref in work:
            n += runner.run_device(sl, ref, rg, None).n
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        n_sites = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        n_sites = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    vals = torch.tensor([ms, float(n_sites), float(sum(r.length for r, _, _ in work)), float(n_bases)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = vals.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = mn = sm = vals
    if rank == 0:
        total_len = sum(L for _, L in contigs)
        print(json.dumps({
            "metric": "candidate sites/sec (s1+s2)", "value": float(sm[1]) / (float(mx[0]) * 1e-3), "unit": "sites/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": float(mx[0]), "ms_per_step_fastest_rank": float(mn[0]),
            "higher_is_better": True, "scaling": "strong", "data": "synthetic",
            "config": {"workload": f"whole-genome-shaped synthetic {total_len / 1e9:.2f} Gb (chr1-22,X,Y,M) at {args.coverage:g}x, "
                                   f"{len(regions)} regions of <= {args.region_mb:g} Mb + 16-bp halo, LPT-assigned to {world} GPU(s); device-resident",
                       "regions": len(regions), "regions_rank0": len(work), "positions_per_rank_min_max": [float(mn[2]), float(mx[2])],
                       "sites_total": float(sm[1]), "aligned_bases_total": float(sm[3]), "setup_generate_s_rank0": t_gen},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
