"""Drop-in for the inference call of ``model_training/evaluate_model_helpers.py``.

Reference: evaluate_model_helpers.py:9-20 (LOGIT_TO_PHONEME), :79-83 (rearrange_speech_logits_pt),
:87-115 (runSingleDecodingStep).  The Redis / hdf5 helpers of that file are host plumbing outside the
hot path and are not re-implemented.
"""
import numpy as np
import torch

LOGIT_TO_PHONEME = [
    'BLANK',
    'AA', 'AE', 'AH', 'AO', 'AW',
    'AY', 'B', 'CH', 'D', 'DH',
    'EH', 'ER', 'EY', 'F', 'G',
    'HH', 'IH', 'IY', 'JH', 'K',
    'L', 'M', 'N', 'NG', 'OW',
    'OY', 'P', 'R', 'S', 'SH',
    'T', 'TH', 'UH', 'UW', 'V',
    'W', 'Y', 'Z', 'ZH',
    ' | ',
]


def rearrange_speech_logits_pt(logits):
    # original order is [BLANK, phonemes..., SIL]; rearrange so the order is [BLANK, SIL, phonemes...]
    return np.concatenate((logits[:, :, 0:1], logits[:, :, -1:], logits[:, :, 1:-1]), axis=-1)


def runSingleDecodingStep(x, input_layer, model, model_args, device):
    """Smooth ('valid' padding) and run one trial through the model; returns float32 numpy logits [1, T', C].

    With a b2t_b200 GRUDecoder the smoothing is fused into the engine's input kernel (same taps, same
    'valid' semantics) instead of being a separate convolution."""
    tr = model_args['dataset']['data_transforms']
    with torch.no_grad():
        eng = model.engine(x.shape[0], x.shape[1], training=False)
        logits, _ = eng.forward(x, torch.tensor([input_layer] * x.shape[0], dtype=torch.int32), training=False, smooth_mode=2,
                                smooth_std=float(tr['smooth_kernel_std']), smooth_size=int(tr['smooth_kernel_size']))
    return logits.float().cpu().numpy()
