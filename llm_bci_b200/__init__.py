"""B200-native implementation of the NDT1 hot path of colehurwitz/llm_bci.

Public surface (mirrors the reference's plugin API for this path):
  NAME2MODEL["NDT1"] / ["BCI"] / ["iTransformer"], NDT1, NDT1Output, BCI, iTransformer, ModelOutput, Masker,
  pad_collate_fn / padded_array (host semantics) and DevicePadCollate,
  DictConfig / update_config, format_ctc / greedy_ctc_decode / phoneme_error_rate.
"""
from .config import DictConfig, update_config, config_from_kwargs, default_model_config, default_trainer_config  # noqa: F401
from .model_output import ModelOutput  # noqa: F401
from .masker import Masker  # noqa: F401
from .ndt1 import NDT1, NDT1Output, create_context_mask  # noqa: F401
from .collate import pad_collate_fn, padded_array, DevicePadCollate  # noqa: F401
from .decode import format_ctc, greedy_ctc_decode, ctc_error_counts, phoneme_error_rate  # noqa: F401
from .bci import BCI, BCIOutput  # noqa: F401
from .itransformer import iTransformer, iTransformerOutput  # noqa: F401
from .trainer import NAME2MODEL, DataParallelTrainer  # noqa: F401
