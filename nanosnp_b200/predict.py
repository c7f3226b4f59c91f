"""Drop-in for the reference entry point PileupModel/predict.py (same CLI, same VCF text).

    python -m nanosnp_b200.predict -config CFG -model_path CKPT -data DIR -reference REF.fa -output OUT.vcf [-batch_size 1000]

`-data` is a directory of per-contig files: the reference's `.pd` text, or packed reads `.reads.npz` (then s1 runs on
the GPU too).  Records are produced per file in consecutive batches of `-batch_size` sites, because the reference's
record logic depends on batch composition (predict.py:106,119); the per-batch text comes from the
native formatter nsnp_vcf_format_batch.  `--no_cuda` is refused: there is no CPU path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .dataset import PredictDataset
from .model import LSTMNetwork
from .runner import COV_CHANNELS


def write_head(reference_index_file, fwriter):                     # predict.py:13-27
    fwriter.write("##fileformat=VCFv4.3\n")
    fwriter.write('##FILTER=<ID=PASS,Description="All filters passed">\n')
    fwriter.write('##FILTER=<ID=RefCall,Description="Reference call">\n')
    with open(reference_index_file) as f:
        for line in f:
            fields = line.strip().split()
            fwriter.write("##contig=<ID={},length={}>\n".format(fields[0], fields[1]))
    fwriter.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
    fwriter.write('##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype Quality">\n')
    fwriter.write('##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read Depth">\n')
    fwriter.write('##FORMAT=<ID=AF,Number=A,Type=Float,Description="Allele Frequency">\n')
    fwriter.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSample\n")


from .predict_io import ContigVcfAssembler, format_records, format_records_into, vcf_buffer_bytes  # noqa: E402,F401


def predict(model: LSTMNetwork, testing_paths, reference_index_file, batch_size, output_file, device, reference=None, region_len=12_500_000):
    """predict.py:37-195.  `.pd` text files: windows -> model -> records.  Read inputs (`.bam`, `.reads.npz`): the whole
    s1 + s2 path runs on the GPU region by region (caller.call_contig)."""
    from .caller import call_contig
    from .dataset import load_fasta, load_reads_npz
    from .pipeline import PileupEngine
    from .runner import RegionRunner
    runner = None
    fasta = None

    def get_runner():
        nonlocal runner
        if runner is None:
            runner = RegionRunner(PileupEngine(device), model._forward(), records=True)
        return runner

    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    if world > 1:
        # one process per GPU (torchrun): every rank decodes the reads, computes its LPT share of the regions, rank 0
        # merges the records in (contig, position) order and writes the file (caller.call_contigs_sharded)
        from .caller import call_contigs_sharded, records_of_regions
        if any(f.endswith(".pd") for f in testing_paths):
            raise NotImplementedError("multi-GPU runs take read inputs (.bam / .reads.npz); .pd files hold ready-made windows")
        if reference is None:
            raise ValueError("read inputs need the reference FASTA")
        fasta = load_fasta(reference)
        todo = []
        for testing_file in testing_paths:
            if testing_file.endswith(".bam"):
                from .bam import read_bam
                refs, by_contig = read_bam(testing_file)
                todo += [(name, by_contig[name]) for name, _ in refs if name in by_contig]
            else:
                reads, contig, _ = load_reads_npz(testing_file)
                todo.append((contig, reads))
        for contig, _ in todo:
            if contig not in fasta:
                raise KeyError(f"contig {contig} is not in the reference")
        contigs = [(contig, len(fasta[contig])) for contig, _ in todo]
        model.eval()

        def produce(ci, rgs):
            return records_of_regions(get_runner(), todo[ci][1], fasta[todo[ci][0]], rgs)
        import contextlib, io
        with (open(output_file, "wb") if rank == 0 else contextlib.nullcontext()) as fwriter:
            if rank == 0:
                head = io.StringIO(); write_head(reference_index_file, head)
                fwriter.write(head.getvalue().encode())
            call_contigs_sharded(contigs, produce, fwriter, batch_size, region_len)
        return

    with open(output_file, "wb") as fwriter:
        import io
        head = io.StringIO(); write_head(reference_index_file, head)
        fwriter.write(head.getvalue().encode())
        model.eval()
        for testing_file in testing_paths:
            if testing_file.endswith(".pd"):
                dataset = PredictDataset(datapath=testing_file, reference=reference, device=device)
                if len(dataset) == 0:
                    continue
                names = dataset.contig_names
                if any(nm != names[0] for nm in names):
                    raise NotImplementedError("one predict-data file must hold one contig (as make_predict_data.sh writes them)")
                x = torch.from_numpy(dataset.position_matrix).to(device)
                gt, zy = model.predict(x)
                cov8 = x[:, 16, COV_CHANNELS].to(torch.float32)
                fwriter.write(format_records(names[0], dataset.positions, dataset.reference_bases, gt.cpu().numpy(), zy.cpu().numpy(),
                                             cov8.cpu().numpy(), batch_size))
                continue
            if fasta is None:
                if reference is None:
                    raise ValueError("read inputs need the reference FASTA")
                fasta = load_fasta(reference)
            if testing_file.endswith(".bam"):
                from .bam import read_bam
                refs, by_contig = read_bam(testing_file)
                todo = [(name, by_contig[name]) for name, _ in refs if name in by_contig]
            else:
                reads, contig, _ = load_reads_npz(testing_file)
                todo = [(contig, reads)]
            for contig, reads in todo:
                if contig not in fasta:
                    raise KeyError(f"contig {contig} of {testing_file} is not in the reference")
                call_contig(get_runner(), reads, fasta[contig], contig, fwriter, batch_size, region_len)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-config", type=str, required=True, help="path to config file")
    parser.add_argument("-model_path", required=True, help="path to trained model")
    parser.add_argument("-data", required=True, help="directory of per-contig files (.pd text, .reads.npz, .bam) or one such file")
    parser.add_argument("-reference", required=True, help="path to reference file")
    parser.add_argument("-output", required=True, help="output vcf file")
    parser.add_argument("-batch_size", type=int, default=1000, help="batch size")
    parser.add_argument("--no_cuda", action="store_true", help="refused: the B200 path has no CPU fallback")
    parser.add_argument("--precision", default="f16x3", choices=["fp32", "f16x3"])
    parser.add_argument("--region_len", type=int, default=12_500_000, help="positions per GPU work unit (read inputs)")
    opt = parser.parse_args(argv)
    if opt.no_cuda:
        raise SystemExit("nanosnp_b200.predict: --no_cuda is not supported (no CPU fallback); use the reference's predict.py on CPU")
    import yaml
    from .utils import AttrDict
    device = torch.device("cuda")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:              # torchrun: one process per GPU, NCCL only gathers the call lists
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=device)
    config = AttrDict(yaml.load(open(opt.config), Loader=yaml.FullLoader))
    pred_model = LSTMNetwork(config.model, precision=opt.precision).to(device)
    if opt.model_path.endswith(".npz"):
        z = np.load(opt.model_path)
        checkpoint = {"encoder": {k[8:]: z[k] for k in z.files if k.startswith("encoder.")},
                      "forward_layer": {k[14:]: z[k] for k in z.files if k.startswith("forward_layer.")}}
    else:
        checkpoint = torch.load(opt.model_path, map_location="cpu")
    pred_model.encoder.load_state_dict(checkpoint["encoder"])
    pred_model.forward_layer.load_state_dict(checkpoint["forward_layer"])
    if os.path.isfile(opt.data):
        testing_paths = [opt.data]                      # a single BAM / .reads.npz / .pd
    else:
        testing_paths = sorted(opt.data + "/" + f for f in os.listdir(opt.data) if f.endswith((".pd", ".reads.npz", ".bam")))
    assert os.path.exists(opt.reference + ".fai"), "reference index file does not exist."
    predict(pred_model, testing_paths, opt.reference + ".fai", opt.batch_size, opt.output, device, reference=opt.reference,
            region_len=opt.region_len)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
