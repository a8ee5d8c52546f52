"""Base output type of the plugin surface (reference models/model_output.py:11-17):
``loss`` is a SUM over examples, ``n_examples`` its denominator."""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class ModelOutput:
    loss: Optional[torch.FloatTensor] = None
    n_examples: Optional[torch.LongTensor] = None

    def to_dict(self):
        return {k: getattr(self, k) for k in self.__dataclass_fields__.keys()}
