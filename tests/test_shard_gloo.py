"""Host logic of the multi-GPU path on CPU: region planning, LPT assignment, and the world_size-2 gloo gather that
merges per-rank call lists into (contig, position) order before the record logic runs."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from nanosnp_b200.shard import Region, assign_lpt, merge_site_lists, plan_regions, read_range_for_region

GRCH38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
          ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
          ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
          ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
          ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415), ("chrM", 16569)]


def test_regions_tile_every_contig_exactly():
    regs = plan_regions(GRCH38, 16_000_000)
    for ci, (name, L) in enumerate(GRCH38):
        mine = [r for r in regs if r.contig_index == ci]
        assert mine[0].emit_start == 0 and mine[-1].emit_end == L
        assert all(a.emit_end == b.emit_start for a, b in zip(mine, mine[1:]))
        assert all(r.start == max(0, r.emit_start - 16) and r.end == min(L, r.emit_end + 16) for r in mine)
        assert all(r.emit_end - r.emit_start <= 16_000_000 for r in mine)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_lpt_balances_whole_genome(world):
    regs = plan_regions(GRCH38, 16_000_000)
    mine = assign_lpt(regs, world)
    assert sorted(i for m in mine for i in m) == list(range(len(regs)))
    loads = [sum(regs[i].length for i in m) for m in mine]
    assert max(loads) <= 1.06 * (sum(loads) / world)


def test_read_range_covers_all_overlapping_reads():
    rng = np.random.default_rng(0)
    pos = np.sort(rng.integers(0, 1_000_000, 5000)).astype(np.int32)
    span = rng.integers(500, 30_000, 5000)
    reg = Region("c", 0, 1_000_000, 400_000, 500_000)
    lo, hi = read_range_for_region(pos, 30_000, reg)
    ov = np.nonzero((pos < reg.end) & (pos + span > reg.start))[0]
    assert lo <= ov.min() and hi > ov.max()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from nanosnp_b200.shard import gather_to_rank0
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    contigs = [("ctgA", 100_000), ("ctgB", 37_000), ("ctgC", 64_000)]
    regs = plan_regions(contigs, 20_000)
    mine = assign_lpt(regs, world)[rank]
    parts = []
    for i in mine:                                     # what this rank's GPU would emit for its regions
        r = regs[i]
        rng = np.random.default_rng(1000 + r.contig_index)
        allpos = np.sort(rng.choice(r.contig_len, size=r.contig_len // 9, replace=False))
        sel = allpos[(allpos >= r.emit_start) & (allpos < r.emit_end)]
        parts.append({"contig_index": np.full(len(sel), r.contig_index, np.int32), "pos": sel.astype(np.int32),
                      "gt": np.stack([sel % 21, sel % 3], 1).astype(np.float32)})
    merged = gather_to_rank0(parts)
    if rank == 0:
        q.put({k: v.copy() for k, v in merged.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process expectation
    exp_c, exp_p = [], []
    for ci, (_, L) in enumerate([("ctgA", 100_000), ("ctgB", 37_000), ("ctgC", 64_000)]):
        rng = np.random.default_rng(1000 + ci)
        allpos = np.sort(rng.choice(L, size=L // 9, replace=False))
        exp_c.append(np.full(len(allpos), ci, np.int32)); exp_p.append(allpos.astype(np.int32))
    assert np.array_equal(merged["contig_index"], np.concatenate(exp_c))
    assert np.array_equal(merged["pos"], np.concatenate(exp_p))
    assert np.array_equal(merged["gt"][:, 0], merged["pos"] % 21)


def test_merge_orders_by_contig_then_position():
    a = {"contig_index": np.array([1, 1], np.int32), "pos": np.array([5, 9], np.int32)}
    b = {"contig_index": np.array([0, 1], np.int32), "pos": np.array([7, 2], np.int32)}
    m = merge_site_lists([a, None, b])
    assert m["contig_index"].tolist() == [0, 1, 1, 1] and m["pos"].tolist() == [7, 2, 5, 9]


# ---- the sharded predict driver (caller.call_contigs_sharded) with a stand-in for the GPU: VCF bytes must not depend on
#      the number of ranks, although predict.py's record logic depends on the 1000-site batch composition ----
_SH_CONTIGS = [("ctgA", 260_000), ("ctgB", 91_000), ("ctgC", 150_000)]


def _fake_records(ci, L):
    """Deterministic compact site records of a whole contig (what the GPU would emit), ascending positions."""
    from nanosnp_b200.predict_io import RECORD_DTYPE
    rng = np.random.default_rng(77 + ci)
    pos = np.sort(rng.choice(L, size=L // 23, replace=False)).astype(np.int32)
    n = len(pos)
    rec = np.zeros(n, RECORD_DTYPE)
    rec["gt"] = rng.choice([0, 1, 2, 3, 4, 5, 7, 9, 12], n); rec["zy"] = rng.integers(0, 3, n)
    rec["ref"] = rng.choice(np.frombuffer(b"ACGT", np.uint8), n); rec["pos1"] = pos + 1
    rec["q100_gt"] = rng.integers(0, 5000, n); rec["q100_zy"] = rng.integers(0, 5000, n)
    rec["depth"] = rng.integers(6, 80, n); rec["af_q"] = rng.integers(0, 1000001, n)
    rec["p_gt"] = 0.5; rec["p_zy"] = 0.5
    return rec


def _fake_produce(ci, rgs):
    rec = _fake_records(ci, _SH_CONTIGS[ci][1])
    out = []
    for r in rgs:
        p0 = rec["pos1"] - 1
        out.append(rec[(p0 >= r.emit_start) & (p0 < r.emit_end)])
    return out


def _sharded_worker(rank, world, port, q):
    import io
    import torch.distributed as dist
    from nanosnp_b200.caller import call_contigs_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sink = io.BytesIO() if rank == 0 else None
    res = call_contigs_sharded(_SH_CONTIGS, _fake_produce, sink, batch_size=1000, region_len=40_000, n_threads=2)
    if rank == 0:
        q.put((sink.getvalue(), res))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_predict_driver_two_ranks_equals_one():
    import io
    from nanosnp_b200.caller import call_contigs_sharded
    sink = io.BytesIO()
    res1 = call_contigs_sharded(_SH_CONTIGS, _fake_produce, sink, batch_size=1000, region_len=40_000, n_threads=2)
    assert res1["world"] == 1 and res1["sites"] == sum(L // 23 for _, L in _SH_CONTIGS)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    text2, res2 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res2["world"] == 2 and res2["sites"] == res1["sites"]
    assert text2 == sink.getvalue() and len(text2) > 100_000


# ---- write_sharded_vcf: records never leave their rank; counts, batch heads and text lengths are all-reduced and every rank
#      pwrite()s its own segments into one ordered file ----
def _sharded_writer_worker(rank, world, port, path):
    import torch.distributed as dist
    from nanosnp_b200.caller import write_sharded_vcf
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    regs = plan_regions(_SH_CONTIGS, 40_000)
    mine = assign_lpt(regs, world)[rank]
    recs = {}
    for ci in sorted({regs[i].contig_index for i in mine}):
        idx = [i for i in mine if regs[i].contig_index == ci]
        for i, r in zip(idx, _fake_produce(ci, [regs[i] for i in idx])):
            recs[i] = r
    res = write_sharded_vcf(path, b"##header\n", _SH_CONTIGS, regs, recs, batch_size=1000)
    assert res["world"] == world
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_writer_equals_single_process_formatter(tmp_path, world):
    import ctypes as C
    from nanosnp_b200 import _lib
    from nanosnp_b200.caller import write_sharded_vcf
    lib = _lib.load()
    # expectation: every contig formatted in one piece by the host contig formatter
    want = b"##header\n"
    for ci, (name, L) in enumerate(_SH_CONTIGS):
        rec = _fake_records(ci, L)
        cap = len(rec) * 160 + 1024
        buf = C.create_string_buffer(cap)
        n = lib.nsnp_vcf_format_contig_records(name.encode(), len(rec), rec.ctypes.data, 1000, 2, C.addressof(buf), cap)
        want += buf.raw[:n]
    # one process
    regs = plan_regions(_SH_CONTIGS, 40_000)
    recs = {}
    for ci in range(len(_SH_CONTIGS)):
        idx = [i for i, r in enumerate(regs) if r.contig_index == ci]
        for i, r in zip(idx, _fake_produce(ci, [regs[i] for i in idx])):
            recs[i] = r
    p1 = str(tmp_path / "one.vcf")
    write_sharded_vcf(p1, b"##header\n", _SH_CONTIGS, regs, recs, batch_size=1000)
    assert open(p1, "rb").read() == want
    # `world` gloo ranks
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    pn = str(tmp_path / "n.vcf")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_sharded_writer_worker, args=(r, world, port, pn)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert open(pn, "rb").read() == want
