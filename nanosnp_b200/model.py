"""Drop-in for the reference's model seam: PileupModel/model.py `LSTMNetwork`.

    net = LSTMNetwork(config.model).to(device)
    net.encoder.load_state_dict(ck['encoder']); net.forward_layer.load_state_dict(ck['forward_layer'])   # predict.py:211-214
    net.eval()
    gt, zy = net.predict(x)      # x: FloatTensor[N,33,18] on the device -> softmaxed Float[N,21], Float[N,3]   (model.py:114-119)

The forward pass is the hand-written CUDA path (csrc/model.cu, csrc/model_tc.cu); only inference is provided and
it refuses to run anywhere but on a CUDA device.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .pipeline import PileupModelForward, PileupModelWeights, require_cuda

_ENC_KEYS = [f"lstm.{k}_l{l}{s}" for l in (0, 1) for s in ("", "_reverse") for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")] + \
            ["output_proj.weight", "output_proj.bias"]
_FWD_KEYS = ["dense.weight", "dense.bias", "genotype_layer.weight", "genotype_layer.bias", "zygosity_layer.weight", "zygosity_layer.bias"]
_FWD_OPTIONAL = ["indel1_layer.weight", "indel1_layer.bias", "indel2_layer.weight", "indel2_layer.bias"]   # never evaluated by predict()


class _StateHolder:
    """Holds one of the two checkpoint sub-dicts with torch's load_state_dict error behaviour (missing/unexpected keys raise)."""

    def __init__(self, owner, required, optional=()):
        self._owner, self._required, self._optional = owner, list(required), list(optional)
        self._state = None

    def load_state_dict(self, state, strict: bool = True):
        missing = [k for k in self._required if k not in state]
        unexpected = [k for k in state if k not in self._required and k not in self._optional]
        if missing or (strict and unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing}, unexpected keys {unexpected}")
        self._state = {k: torch.as_tensor(v).detach().to(torch.float32).cpu() for k, v in state.items()}
        self._owner._invalidate()

    def state_dict(self):
        if self._state is None:
            raise RuntimeError("state_dict requested before load_state_dict")
        return dict(self._state)


class LSTMNetwork:
    def __init__(self, config=None, precision: str = "f16x3"):
        # config.model of ont_pileup.yaml is fixed by the kernels: 18 -> 2 x BiLSTM(64) -> 128 -> 256 -> 21 / 3
        if config is not None:
            enc = getattr(config, "enc", None)
            dims = (getattr(config, "feature_dim", 18), getattr(enc, "hidden_size", 64), getattr(enc, "n_layers", 2),
                    getattr(enc, "output_size", 128), getattr(getattr(config, "joint", None), "inner_size", 256),
                    getattr(config, "gt_num_class", 21), getattr(config, "zy_num_class", 3))
            if dims != (18, 64, 2, 128, 256, 21, 3):
                raise NotImplementedError(f"the CUDA path is built for ont_pileup.yaml's architecture, got {dims}")
        self.config = config
        self.encoder = _StateHolder(self, _ENC_KEYS)
        self.forward_layer = _StateHolder(self, _FWD_KEYS, _FWD_OPTIONAL)
        self.precision = {"fp32": _lib.PREC_FP32, "f16x3": _lib.PREC_F16X3, "f16x1": _lib.PREC_F16X1}[precision]
        self.device: Optional[torch.device] = None
        self._fwd: Optional[PileupModelForward] = None

    def _invalidate(self):
        self._fwd = None

    def to(self, device):
        self.device = require_cuda(device)
        self._invalidate()
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def eval(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("nanosnp_b200 provides inference only")
        return self

    def _forward(self) -> PileupModelForward:
        if self._fwd is None:
            if self.device is None:
                raise _lib.NsnpError(_lib.E_NO_DEVICE, "call .to('cuda') first: there is no CPU fallback")
            w = PileupModelWeights(self.encoder.state_dict(), self.forward_layer.state_dict(), device=self.device)
            self._fwd = PileupModelForward(w, self.precision)
        return self._fwd

    @torch.no_grad()
    def predict(self, inputs: torch.Tensor):
        assert inputs.dim() == 3                                    # model.py:32
        if not inputs.is_cuda:
            raise _lib.NsnpError(_lib.E_NO_DEVICE, "inputs must live on the CUDA device (predict.py:49 moves them there)")
        if inputs.dtype not in (torch.float32, torch.int32):
            inputs = inputs.to(torch.float32)
        return self._forward()(inputs.contiguous())
