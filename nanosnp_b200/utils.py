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
