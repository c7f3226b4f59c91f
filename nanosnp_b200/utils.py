"""AttrDict: YAML config as attributes, missing key -> None (contract of PileupModel/utils.py:5-21)."""


class AttrDict(dict):
    def __getattr__(self, item):
        if item not in self:
            return None
        if type(self[item]) is dict:
            self[item] = AttrDict(self[item])
        return self[item]

    def __setattr__(self, item, value):
        self.__dict__[item] = value


def load_weights_npz(path):
    """Checkpoint stored as a flat .npz ("encoder.<key>", "forward_layer.<key>" arrays): returns the two state dicts of
    utils.py:67-77 (ck['encoder'], ck['forward_layer']) as {name: ndarray}."""
    import numpy as np
    z = np.load(path)
    enc = {k[len("encoder."):]: z[k] for k in z.files if k.startswith("encoder.")}
    fwd = {k[len("forward_layer."):]: z[k] for k in z.files if k.startswith("forward_layer.")}
    return enc, fwd
