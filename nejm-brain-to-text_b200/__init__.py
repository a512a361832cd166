"""b2t_b200: B200-native (sm_100a) GRU -> CTC -> n-gram decode hot path behind the reference's API.

Import name: ``nejm_brain_to_text_b200`` (the directory is ``nejm-brain-to-text_b200``; use
``b2t_pkg.load()`` at the repository root, which registers it under the importable name).

Modules mirror the reference's files for this path:
  rnn_model.GRUDecoder                  model_training/rnn_model.py
  rnn_trainer.BrainToTextDecoder_Trainer model_training/rnn_trainer.py
  data_augmentations.gauss_smooth       model_training/data_augmentations.py
  evaluate_model_helpers                model_training/evaluate_model_helpers.py
  lm_decoder                            language_model/runtime/server/x86/python/lm_decoder.cc

Importing the package loads libb2t_b200.so and raises if it is missing: there is no fallback path.
"""
from . import _native  # noqa: F401  (fails loudly when the native library is absent)

__all__ = ["_native"]
__version__ = "0.1.0"
