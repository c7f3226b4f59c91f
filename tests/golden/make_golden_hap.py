"""Golden vectors of the HaplotypeModel s5 path (BASELINE configs[4]) from the REAL reference Python
(only possible in the build container):

    python tests/golden/make_golden_hap.py

Imports /root/reference/HaplotypeModel/{dataset_dev,model_dev,predict_dev}.py (stub modules for the absent `tables` /
`ranger21` packages), feeds seeded read x position matrices through the reference's own get_frequency_feature, its
LSTMNetwork with seeded random-init weights (both shipped checkpoints are missing: .MISSING_LARGE_BLOBS) and its predict loop,
and stores inputs + outputs in tests/golden/hap_small.npz / hap_small.csv.  The weights are NOT stored (33 MB): the oracle
(oracle/hap_restate.py) rebuilds them from the same seed, which the stored probabilities verify.
"""
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/HaplotypeModel")
SEED = 20240
N, DEPTH = 48, 36


def make_inputs(seed=7):
    """Read x position matrices shaped like write_to_bins.py:15-63 writes them: bases 1..4, deletion -1, absent 0, pad rows -2."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, L in (("pileup", 33), ("haplotype", 11)):
        seq = rng.integers(1, 5, size=(N, DEPTH, L)).astype(np.int32)
        major = rng.integers(1, 5, size=(N, 1, L))
        seq = np.where(rng.random((N, DEPTH, L)) < 0.8, major, seq).astype(np.int32)
        seq[rng.random((N, DEPTH, L)) < 0.05] = -1
        absent = rng.random((N, DEPTH, L)) < 0.1
        seq[absent] = 0
        bq = rng.integers(1, 45, size=(N, DEPTH, L)).astype(np.int32)
        bq[seq <= 0] = 0
        mq = np.repeat(rng.integers(0, 61, size=(N, DEPTH, 1)), L, axis=2).astype(np.int32)
        mq[seq == 0] = 0
        hp = np.repeat(rng.integers(1, 4, size=(N, DEPTH, 1)), L, axis=2).astype(np.int32)
        hp[seq == 0] = 0
        depth = rng.integers(3, DEPTH + 1, size=N)
        depth[0] = DEPTH; depth[1] = 1
        for i in range(N):
            for a in (seq, bq, mq, hp):
                a[i, depth[i]:] = -2
        hp[2][hp[2] == 1] = 3                      # a site without maternal reads, one without unphased reads
        hp[3][hp[3] == 3] = 2
        out[name] = (seq, bq, mq, hp)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 5000))
    ref = ref[:700] + "N" + ref[701:900] + "acgtn" + ref[905:]
    pos = np.sort(rng.choice(np.arange(100, 4800), N, replace=False))
    pos[5] = 701; pos[6] = 903                        # reference N and soft-masked bases: code 0 (dataset_dev.py:111-115)
    pos.sort()
    hap_pos = np.stack([np.sort(np.concatenate([[p], rng.choice(np.setdiff1d(np.arange(max(1, p - 400), min(5000, p + 400)), [p]), 10, replace=False)])) for p in pos])
    return out, ref, pos, hap_pos


def main():
    import torch
    for name in ("tables", "ranger21"):
        m = types.ModuleType(name)
        m.Ranger21 = type("Ranger21", (), {})
        sys.modules[name] = m
    sys.path.insert(0, str(REF))
    import dataset_dev as D
    import model_dev as M
    import predict_dev as P
    from utils import AttrDict
    import yaml
    inp, ref, pos, hap_pos = make_inputs()
    refs = {"ctgH": ref}
    feats = {}
    for name, L in (("pileup", 33), ("haplotype", 11)):
        seq, bq, mq, hp = inp[name]
        f = []
        for i in range(N):
            a = D.get_frequency_feature(seq[i], bq[i], mq[i], hp[i])
            if name == "pileup":
                rs = [D.BASE2INT.get(refs["ctgH"][j - 1], 0) if 0 <= j - 1 < len(ref) else 0 for j in range(pos[i] - 16, pos[i] + 17)]
            else:
                rs = [D.BASE2INT.get(refs["ctgH"][j - 1], 0) for j in hap_pos[i]]
            f.append(np.concatenate((a, np.asarray(rs).reshape(1, -1)), axis=0))
        feats[name] = np.stack(f)                      # float64 [N, 105, L]
    cfg = AttrDict(yaml.load(open(REF / "config/ont_haplotype.yaml"), Loader=yaml.FullLoader))
    torch.manual_seed(SEED)
    net = M.LSTMNetwork(cfg)
    net.eval()
    torch.set_num_threads(1)
    with torch.no_grad():
        gt, zy = net.predict(torch.from_numpy(feats["pileup"]).type(torch.FloatTensor), torch.from_numpy(feats["haplotype"]).type(torch.FloatTensor))

    class Ds(torch.utils.data.Dataset):                # stands in for TestDataset (PyTables is absent)
        def __init__(self, **kw):
            pass
        def __len__(self):
            return N
        def __getitem__(self, i):
            return "ctgH:%d" % pos[i], feats["pileup"][i], feats["haplotype"][i]

    P.TestDataset = Ds
    P.load_reference_file = lambda p: refs
    import os
    real_listdir = os.listdir
    P.os.listdir = lambda d: ["x.bin"]
    loader = torch.utils.data.DataLoader
    P.torch.utils.data.DataLoader = lambda ds, batch_size, shuffle, num_workers: loader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=0)
    try:
        P.predict(net, "unused", "unused", 20, 33, 11, str(HERE / "hap_small.csv"), torch.device("cpu"))
    finally:
        P.os.listdir = real_listdir
        P.torch.utils.data.DataLoader = loader
    np.savez_compressed(HERE / "hap_small.npz", seed=np.int64(SEED), ref=np.frombuffer(ref.encode(), np.uint8), pos=pos.astype(np.int64),
                        hap_pos=hap_pos.astype(np.int64),
                        **{f"{n}_{k}": a for n in inp for k, a in zip(("seq", "bq", "mq", "hp"), inp[n])},
                        pileup_feat=feats["pileup"], haplotype_feat=feats["haplotype"], gt=gt.numpy(), zy=zy.numpy(),
                        n_params=np.int64(sum(p.numel() for p in net.parameters())))
    print("params", sum(p.numel() for p in net.parameters()), "gt", gt.shape, "csv lines", sum(1 for _ in open(HERE / "hap_small.csv")))


if __name__ == "__main__":
    main()
