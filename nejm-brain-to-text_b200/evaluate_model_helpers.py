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


# ------------------------------------------------------------------------------------------------ LM client side (Redis protocol)
# evaluate_model_helpers.py:136-297.  `r` is a redis.Redis connection or, without a redis-server process, the in-process
# language_model.LoopbackRedis; stream names and payloads are the reference's.
import time as _time


def get_current_redis_time_ms(redis_conn):
    t = redis_conn.time()
    return int(t[0] * 1000 + t[1] / 1000)


def _wait_for(r, stream, last_seen, what):
    out = []
    while len(out) == 0:
        out = r.xread({stream: last_seen}, count=1, block=10000)
        if len(out) == 0:
            print(f'Still waiting for {what} from ts {last_seen}...')
    return out


def reset_remote_language_model(r, remote_lm_done_resetting_lastEntrySeen):
    r.xadd('remote_lm_reset', {'done': 0})
    _time.sleep(0.001)
    out = _wait_for(r, 'remote_lm_done_resetting', remote_lm_done_resetting_lastEntrySeen, 'remote lm reset')
    for entry_id, _ in out[0][1]:
        remote_lm_done_resetting_lastEntrySeen = entry_id
    return remote_lm_done_resetting_lastEntrySeen


def update_remote_lm_params(r, remote_lm_done_updating_lastEntrySeen, acoustic_scale=0.35, blank_penalty=90.0, alpha=0.55):
    r.xadd('remote_lm_update_params', {'acoustic_scale': acoustic_scale, 'blank_penalty': blank_penalty, 'alpha': alpha})
    _time.sleep(0.001)
    out = _wait_for(r, 'remote_lm_done_updating_params', remote_lm_done_updating_lastEntrySeen, 'remote lm to update parameters')
    for entry_id, _ in out[0][1]:
        remote_lm_done_updating_lastEntrySeen = entry_id
    return remote_lm_done_updating_lastEntrySeen


def send_logits_to_remote_lm(r, remote_lm_input_stream, remote_lm_output_partial_stream, remote_lm_output_partial_lastEntrySeen, logits):
    r.xadd(remote_lm_input_stream, {'logits': np.float32(logits).tobytes()})
    out = _wait_for(r, remote_lm_output_partial_stream, remote_lm_output_partial_lastEntrySeen, 'remote lm partial output')
    decoded = ''
    for entry_id, entry_data in out[0][1]:
        remote_lm_output_partial_lastEntrySeen = entry_id
        decoded = entry_data[b'lm_response_partial'].decode()
    return remote_lm_output_partial_lastEntrySeen, decoded


def finalize_remote_lm(r, remote_lm_output_final_stream, remote_lm_output_final_lastEntrySeen):
    r.xadd('remote_lm_finalize', {'done': 0})
    _time.sleep(0.005)
    out = _wait_for(r, remote_lm_output_final_stream, remote_lm_output_final_lastEntrySeen, 'remote lm final output')
    sent, ac, ng, llm, tot = [], [], [], [], []
    for entry_id, entry_data in out[0][1]:
        remote_lm_output_final_lastEntrySeen = entry_id
        f = entry_data[b'scoring'].decode().split(';') if b'scoring' in entry_data else []
        if len(f) >= 5:
            sent = [str(c) for c in f[::5]]
            ac, ng, llm, tot = ([float(c) for c in f[k::5]] for k in (1, 2, 3, 4))
    if len(sent) == 0 or len(tot) == 0:
        print('No candidate sentences were received from the language model.')
        sent, ac, ng, llm, tot = [''], [0], [0], [0], [0]
    else:
        order = np.argsort(tot)[::-1]                      # higher is better
        sent = [sent[i] for i in order]
        ac, ng, llm, tot = ([v[i] for i in order] for v in (ac, ng, llm, tot))
    for i in range(len(sent) - 1, 0, -1):                  # drop duplicates, keeping the best-scoring copy
        if sent[i] in sent[:i]:
            for v in (sent, ac, ng, llm, tot):
                v.pop(i)
    return remote_lm_output_final_lastEntrySeen, {'candidate_sentences': sent, 'candidate_acoustic_scores': ac, 'candidate_ngram_scores': ng,
                                                  'candidate_llm_scores': llm, 'candidate_total_scores': tot}
