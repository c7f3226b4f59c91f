#!/usr/bin/env python
"""Stage-by-stage validation of the tcgen05 LSTM path on a GPU (run each stage under `timeout`).
usage: tests/tc_check.py {gates0_cg1|gates0_cg2|gates1_cg2|full}"""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from nanosnp_b200 import _lib
from nanosnp_b200.pipeline import PileupModelWeights, PileupModelForward
from oracle.s2_restate import load_weights_npz, PileupModelOracle

ROOT = Path(__file__).resolve().parent.parent
enc, fwd = load_weights_npz(ROOT / "tests/golden/ont_pileup_weights.npz")
z = np.load(ROOT / "tests/golden/s1_small.npz")
x = z["windows"].astype(np.int32)
lib = _lib.load()
W = PileupModelWeights(enc, fwd, device="cuda:0")
stream = torch.cuda.current_stream().cuda_stream


def _tile_layout(hi, lo):
    m = hi.shape[0]; tiles = (m + 127) // 128
    out = np.zeros((tiles, 33, 2, 16, 128, 8), np.float16)
    for part, a in enumerate((hi, lo)):
        pad = np.zeros((tiles * 128, 33, 128), np.float16); pad[:m] = a
        out[:, :, part] = pad.reshape(tiles, 128, 33, 16, 8).transpose(0, 2, 3, 1, 4)
    return out


def col_order():
    n = np.arange(256)
    return ((n >> 2) & 3) * 64 + (n >> 5) * 8 + ((n >> 4) & 1) * 4 + (n & 3)


def gates_ref(layer, d, inp):
    sfx = "_reverse" if d else ""
    wih = enc[f"lstm.weight_ih_l{layer}{sfx}"].astype(np.float64); b = (enc[f"lstm.bias_ih_l{layer}{sfx}"] + enc[f"lstm.bias_hh_l{layer}{sfx}"]).astype(np.float64)
    n = np.arange(256)
    scale = np.where(((n >> 2) & 3) == 2, -2.0, -1.0) * np.log2(np.e)      # activation scale folded into the packed weights
    return (inp.astype(np.float64) @ wih.T + b)[:, col_order()] * scale


def run_gates(layer, d, cg, xin, h0=None, m=300):
    out = torch.full((m, 256), float("nan"), device="cuda")
    xi = torch.from_numpy(xin[:m]).cuda() if xin is not None else None
    rc = lib.nsnp_debug_lstm_tc_gates(W.blob.data_ptr(), 0 if xi is None else xi.data_ptr(), layer, d, cg, 0 if h0 is None else h0.data_ptr(),
                                      out.data_ptr(), m, stream)
    _lib.check(rc)
    torch.cuda.synchronize()
    return out.cpu().numpy()


what = sys.argv[1]
if what.startswith("gates0"):
    cg = int(what[-1])
    for d in (0, 1):
        t = 0 if d == 0 else 32
        got = run_gates(0, d, cg, x)
        ref = gates_ref(0, d, x[:300, t, :])
        err = np.abs(got - ref)
        print(what, "dir", d, "max abs err", err.max(), "ref max", np.abs(ref).max(), "nan", np.isnan(got).sum(), flush=True)
        if err.max() > 1e-3:
            bad = np.argwhere(err > 1e-3)[:5]; print(" bad idx", bad.tolist(), got[tuple(bad[0])], ref[tuple(bad[0])])
elif what == "gates1_cg2":
    rng = np.random.default_rng(0)
    m = 300
    h = rng.uniform(-1, 1, size=(m, 33, 128)).astype(np.float32)
    hi = h.astype(np.float16); lo = (h - hi.astype(np.float32)).astype(np.float16)
    h0 = torch.from_numpy(_tile_layout(hi, lo)).cuda()
    for d in (0, 1):
        t = 0 if d == 0 else 32
        got = run_gates(1, d, 2, None, h0, m)
        ref = gates_ref(1, d, (hi.astype(np.float64) + lo.astype(np.float64))[:, t, :])
        err = np.abs(got - ref)
        print(what, "dir", d, "max abs err", err.max(), "ref max", np.abs(ref).max(), "nan", np.isnan(got).sum(), flush=True)
elif what == "full":
    m = PileupModelOracle(enc, fwd)
    g0, z0 = m.predict(x)
    f32 = PileupModelForward(W, _lib.PREC_FP32); tc = PileupModelForward(W, _lib.PREC_F16X3)
    xt = torch.from_numpy(x).cuda()
    g1, z1 = f32(xt); g1, z1 = g1.clone(), z1.clone()
    g2, z2 = tc(xt)
    torch.cuda.synchronize()
    for name, (g, zz) in {"fp32": (g1, z1), "f16x3": (g2, z2)}.items():
        eg = (g.cpu() - g0).abs().max().item(); ez = (zz.cpu() - z0).abs().max().item()
        flips = int((g.cpu().argmax(1) != g0.argmax(1)).sum()) + int((zz.cpu().argmax(1) != z0.argmax(1)).sum())
        print(name, "max abs err gt", eg, "zy", ez, "argmax flips", flips, "nan", int(torch.isnan(g).sum()), flush=True)
    import time
    xb = xt.repeat(40, 1, 1)[:200000].contiguous()
    for name, f in (("fp32", f32), ("f16x3", tc)):
        f(xb); torch.cuda.synchronize(); t0 = time.perf_counter(); f(xb); torch.cuda.synchronize()
        dt = time.perf_counter() - t0; print(name, f"{xb.shape[0] / dt / 1e6:.2f} M sites/s", flush=True)
