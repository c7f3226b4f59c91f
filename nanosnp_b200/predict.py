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


def format_records_into(buf: np.ndarray, contig: str, positions, reference_bases, gt, zy, cov8, batch_size: int, n_threads: int = 0) -> int:
    """Formats all records of consecutive batch_size-site batches into the uint8 array `buf`; returns the byte count
    (raises if buf is too small).  No copies of the inputs when they are already contiguous and typed."""
    lib = _lib.load()
    n = len(positions)
    if n == 0:
        return 0
    pos = np.ascontiguousarray(positions, np.int32); refb = np.ascontiguousarray(reference_bases, np.uint8)
    gt = np.ascontiguousarray(gt, np.float32); zy = np.ascontiguousarray(zy, np.float32); cov8 = np.ascontiguousarray(cov8, np.float32)
    w = lib.nsnp_vcf_format_contig(contig.encode(), n, pos.ctypes.data, refb.ctypes.data, gt.ctypes.data, zy.ctypes.data,
                                   cov8.ctypes.data, batch_size, n_threads or (os.cpu_count() or 1), buf.ctypes.data, buf.shape[0])
    if w < 0:
        raise _lib.NsnpError(_lib.E_WORKSPACE, f"VCF buffer too small: need {-w} bytes")
    return int(w)


def vcf_buffer_bytes(n: int, contig: str) -> int:
    return n * (80 + len(contig)) + 64


def format_records(contig: str, positions, reference_bases, gt: np.ndarray, zy: np.ndarray, cov8: np.ndarray, batch_size: int,
                   n_threads: int = 0) -> bytes:
    """All records of one contig file, consecutive batches of batch_size sites (host arrays) -> VCF text bytes."""
    buf = np.empty(vcf_buffer_bytes(len(positions), contig), np.uint8)
    w = format_records_into(buf, contig, positions, reference_bases, gt, zy, cov8, batch_size, n_threads)
    return buf[:w].tobytes()


class ContigVcfAssembler:
    """Streams one contig's sites region by region into VCF text while keeping the reference's batch composition:
    records are formatted in consecutive batches of `batch_size` sites counted from the contig's first site
    (predict.py:43), so a region boundary in the middle of a batch carries the partial batch over to the next region."""

    def __init__(self, contig: str, batch_size: int = 1000, n_threads: int = 0, sink=None):
        self.contig, self.batch, self.threads, self.sink = contig, batch_size, n_threads, sink
        self.carry = None
        self.n_bytes = 0
        self.n_sites = 0

    def _emit(self, pos1, refb, gt, zy, cov8):
        need = vcf_buffer_bytes(len(pos1), self.contig)
        if getattr(self, "_buf", None) is None or self._buf.shape[0] < need:
            self._buf = np.empty(int(need * 1.1), np.uint8)          # reused across regions
        w = format_records_into(self._buf, self.contig, pos1, refb, gt, zy, cov8, self.batch, self.threads)
        self.n_bytes += w
        if self.sink is not None:
            self.sink.write(self._buf[:w].tobytes())

    def add(self, pos0, refbase, gt, zy, cov8):
        """Host arrays of one region, ascending positions (0-based)."""
        n = len(pos0)
        self.n_sites += n
        pos1 = np.asarray(pos0, np.int32) + 1
        arrs = [pos1, np.asarray(refbase), np.asarray(gt), np.asarray(zy), np.asarray(cov8)]
        start = 0
        if self.carry is not None:
            need = self.batch - len(self.carry[0])
            take = min(need, n)
            merged = [np.concatenate([c, a[:take]]) for c, a in zip(self.carry, arrs)]
            start = take
            if len(merged[0]) == self.batch:
                self._emit(*merged)
                self.carry = None
            else:
                self.carry = merged
                return
        full = (n - start) // self.batch * self.batch
        if full:
            self._emit(*[a[start:start + full] for a in arrs])
        if start + full < n:
            self.carry = [np.array(a[start + full:]) for a in arrs]        # copy: the caller reuses its buffers

    def close(self):
        if self.carry is not None:
            self._emit(*self.carry)
            self.carry = None
        return self.n_bytes


def predict(model: LSTMNetwork, testing_paths, reference_index_file, batch_size, output_file, device, reference=None):
    with open(output_file, "w") as fwriter:
        write_head(reference_index_file, fwriter)
        fwriter.flush()
        model.eval()
        for testing_file in testing_paths:
            dataset = PredictDataset(datapath=testing_file, reference=reference, device=device)
            if len(dataset) == 0:
                continue
            x = dataset.x_device if dataset.x_device is not None else torch.from_numpy(dataset.position_matrix).to(device)
            gt, zy = model.predict(x)
            cov8 = x[:, 16, COV_CHANNELS].to(torch.float32)
            # the reference groups by file; a file holds one contig (make_predict_data.sh:231-239)
            names = dataset.contig_names
            start = 0
            gt_h, zy_h, cov_h = gt.cpu().numpy(), zy.cpu().numpy(), cov8.cpu().numpy()
            while start < len(names):                             # contiguous runs of one contig name
                end = start
                while end < len(names) and names[end] == names[start]:
                    end += 1
                if start == 0 and end == len(names):
                    fwriter.write(format_records(names[0], dataset.positions, dataset.reference_bases, gt_h, zy_h, cov_h, batch_size).decode())
                else:
                    raise NotImplementedError("one predict-data file must hold one contig (as make_predict_data.sh writes them)")
                start = end


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-config", type=str, required=True, help="path to config file")
    parser.add_argument("-model_path", required=True, help="path to trained model")
    parser.add_argument("-data", required=True, help="directory of predict-data files (.pd text or .reads.npz)")
    parser.add_argument("-reference", required=True, help="path to reference file")
    parser.add_argument("-output", required=True, help="output vcf file")
    parser.add_argument("-batch_size", type=int, default=1000, help="batch size")
    parser.add_argument("--no_cuda", action="store_true", help="refused: the B200 path has no CPU fallback")
    parser.add_argument("--precision", default="f16x3", choices=["fp32", "f16x3"])
    opt = parser.parse_args(argv)
    if opt.no_cuda:
        raise SystemExit("nanosnp_b200.predict: --no_cuda is not supported (no CPU fallback); use the reference's predict.py on CPU")
    import yaml
    from .utils import AttrDict
    device = torch.device("cuda")
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
    testing_paths = sorted(opt.data + "/" + f for f in os.listdir(opt.data) if f.endswith(".pd") or f.endswith(".reads.npz"))
    assert os.path.exists(opt.reference + ".fai"), "reference index file does not exist."
    predict(pred_model, testing_paths, opt.reference + ".fai", opt.batch_size, opt.output, device, reference=opt.reference)


if __name__ == "__main__":
    main()
