#!/usr/bin/env python
"""what a single-pass fp16 LSTM (1 tensor-core MMA per product instead of the 3 hi/lo passes) would buy
and cost.  Needs an experimental library whose MMA loops read NSNP_TC_PASSES (see profiles/r02_precision_experiment.md); run as

    NSNP_LIB=nanosnp_b200/build/libx1.so NSNP_TC_PASSES=1 python tools/prec_experiment.py [region_mb]

Prints one JSON line: model time per region, |dp| / argmax flips / QUAL drift of the tensor-core path against the fp32 path on
one 12.5 Mb region of the bench workload, and how many sites a margin-based fp32 re-evaluation would have to touch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from nanosnp_b200 import _lib
from nanosnp_b200.pipeline import PileupEngine, PileupModelForward, PileupModelWeights
from nanosnp_b200.runner import RegionRunner
from nanosnp_b200.shard import Region
from nanosnp_b200.synth import SynthConfig, generate_device
from nanosnp_b200.utils import load_weights_npz

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 12.5
dev = torch.device("cuda:0")
cfg = SynthConfig(contig_len=int(mb * 1e6), coverage=30.0, contig="ctg1", seed_ref=1000, seed_var=2000, seed_reads=3000)
ref, reads = generate_device(cfg, dev)
eng = PileupEngine(dev)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = PileupModelWeights(*load_weights_npz(os.path.join(root, "tests", "golden", "ont_pileup_weights.npz")), device=dev)
tc = PileupModelForward(w, _lib.PREC_F16X3); f32 = PileupModelForward(w, _lib.PREC_FP32)
runner = RegionRunner(eng, tc, keep_windows=True)
rg = Region("ctg1", 0, cfg.contig_len, 0, cfg.contig_len)
out = runner.run_device(reads, ref, rg)
n = out.n
x = out.x[:n]
ts = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g, z = tc(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
g32, z32 = f32(x)
torch.cuda.synchronize()


def qual(p):                                    # predict.py calculate_score without the rounding
    p = p.double().clamp(max=1 - 1e-12)
    return torch.clamp(-10 * torch.log10((1 - p) / p) + 10, min=0)


res = {"passes": int(os.environ.get("NSNP_TC_PASSES", "3")), "sites": int(n), "model_ms": min(ts)}
for name, a, b in (("gt", g, g32), ("zy", z, z32)):
    d = (a - b).abs().max(1).values
    top = b.topk(2, dim=1).values
    margin = (top[:, 0] - top[:, 1])
    flips = a.argmax(1) != b.argmax(1)
    dq = (qual(a.max(1).values) - qual(b.max(1).values)).abs()
    res[name] = {"max_abs_dp": float(d.max()), "p99_abs_dp": float(d.quantile(0.99)), "argmax_flips": int(flips.sum()),
                 "max_margin_of_a_flip": float(margin[flips].max()) if flips.any() else 0.0,
                 "qual_drift_max": float(dq.max()), "qual_drift_p99": float(dq.quantile(0.99)),
                 "qual_changed_2dp": int((torch.round(qual(a.max(1).values) * 100) != torch.round(qual(b.max(1).values) * 100)).sum()),
                 "sites_with_margin_below": {str(t): int((margin < t).sum()) for t in (0.01, 0.03, 0.1, 0.3)}}
print(json.dumps(res))
