/* b2t_b200 -- C ABI of the B200-native GRU -> CTC -> n-gram decode hot path.
 *
 * Every entry point takes plain pointers and sizes (device pointers where noted); no torch or
 * C++ types cross the boundary.  `stream` is a cudaStream_t passed as void*.
 * Functions return 0 (or a non-negative value) on success and a negative code on failure;
 * b2t_last_error() returns a human-readable message for the calling thread's last failure.
 *
 * Reference interfaces replaced (file:line under the upstream repository):
 *   b2t_forward ............ GRUDecoder.forward            model_training/rnn_model.py:88-134
 *                            + transform_data/gauss_smooth  model_training/rnn_trainer.py:436-484,
 *                                                           model_training/data_augmentations.py:6-37
 *                            + runSingleDecodingStep        model_training/evaluate_model_helpers.py:87-115
 *   b2t_ctc_loss ........... log_softmax + torch.nn.CTCLoss model_training/rnn_trainer.py:242,538-545
 *   b2t_backward ........... loss.backward()                model_training/rnn_trainer.py:547
 *   b2t_optimizer_step ..... clip_grad_norm_ + AdamW.step   model_training/rnn_trainer.py:550-558,259-292
 *   b2t_greedy_edit ........ greedy decode + edit distance  model_training/rnn_trainer.py:724-736
 *   b2t_decoder_* .......... lm_decoder pybind module       language_model/runtime/server/x86/python/lm_decoder.cc:14-75
 *                            BrainSpeechDecoder             language_model/runtime/core/decoder/brain_speech_decoder.h:22-148
 */
#ifndef B2T_B200_H_
#define B2T_B200_H_

#if defined(B2T_EXPORTS)
#define B2T_API __attribute__((visibility("default")))
#else
#define B2T_API
#endif

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2T_OK 0
#define B2T_ERR_ARG (-1)
#define B2T_ERR_CUDA (-2)
#define B2T_ERR_WORKSPACE (-3)
#define B2T_ERR_UNSUPPORTED (-4)
#define B2T_ERR_STATE (-5)

const char* b2t_last_error(void);
int b2t_version(void);

/* ------------------------------------------------------------------ model / engine */

typedef struct b2t_config {
  int neural_dim;    /* input channels (512) */
  int n_units;       /* GRU hidden size (768); multiple of 64, <= 768 */
  int n_layers;      /* stacked GRU layers (5) */
  int n_days;        /* number of day-specific input layers (45) */
  int n_classes;     /* CTC classes incl. blank (41); <= 64 */
  int patch_size;    /* 14 (0 disables patching) */
  int patch_stride;  /* 4 */
  float rnn_dropout;   /* between GRU layers, training only (0.4) */
  float input_dropout; /* after the day layer, training only (0.2) */
} b2t_config;

/* Flat parameter layout.  Parameters live in ONE fp32 buffer owned by the caller; segment i has the
 * reference state_dict name (day_weights.{d}, day_biases.{d}, gru.weight_ih_l{k}, gru.weight_hh_l{k},
 * gru.bias_ih_l{k}, gru.bias_hh_l{k}, out.weight, out.bias, h0), an element offset (multiple of 64)
 * and a [rows, cols] shape.  Returns the number of segments; with index >= 0 fills the outputs. */
int b2t_param_segments(const b2t_config* cfg);
int b2t_param_segment(const b2t_config* cfg, int index, char* name, int name_cap, long long* offset, long long* rows,
                      long long* cols);
/* Elements of the flat parameter buffer, and of the gradient buffer (= parameters + n_days "touched"
 * flags at the tail, so that ONE all-reduce carries both). */
long long b2t_param_elems(const b2t_config* cfg);
long long b2t_grad_elems(const b2t_config* cfg);
long long b2t_workspace_bytes(const b2t_config* cfg, int max_batch, int max_T, int max_label_len, int training);

typedef struct b2t_engine b2t_engine;

/* All buffers are device memory owned by the caller and must outlive the engine.
 * grads / exp_avg / exp_avg_sq may be NULL for an inference-only engine (training == 0). */
b2t_engine* b2t_engine_create(const b2t_config* cfg, int max_batch, int max_T, int max_label_len, int training,
                              float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* workspace,
                              long long workspace_bytes);
void b2t_engine_destroy(b2t_engine* e);

/* Re-derive the bf16 operand copies (and W_hh^T) from the fp32 parameters, e.g. after load_state_dict. */
int b2t_refresh_weights(b2t_engine* e, void* stream);

typedef struct b2t_forward_args {
  const float* x;           /* device [B][T][neural_dim] fp32 */
  int B, T;
  const int* day_idx;       /* device int32 [B] */
  int training;             /* 1: dropout active and activations saved for b2t_backward */
  int smooth_mode;          /* 0 none, 1 'same' (trainer), 2 'valid' (evaluate_model) */
  float smooth_std;         /* 2.0 */
  int smooth_size;          /* 100 */
  int cut;                  /* random_cut draw: frames dropped from the front */
  float white_noise_std;    /* 0 disables */
  float offset_noise_std;   /* 0 disables */
  const float* white_noise; /* optional explicit N(0,1) draws, device [B][T][neural_dim] */
  const float* offset_noise;/* optional explicit N(0,1) draws, device [B][neural_dim] */
  unsigned long long seed;  /* device RNG stream for noise (when no explicit draws) and dropout */
  const float* states;      /* optional device [n_layers][B][n_units] initial state; NULL => h0 */
  float* logits_out;        /* optional device [B][T'][n_classes] fp32 */
  float* hidden_out;        /* optional device [n_layers][B][n_units] fp32 */
} b2t_forward_args;

/* Returns T' (number of output frames) or a negative error code. */
int b2t_forward(b2t_engine* e, const b2t_forward_args* a, void* stream);
/* T' for an input of T frames under the given smoothing mode and cut. */
int b2t_output_frames(const b2t_config* cfg, int T, int smooth_mode, int smooth_std_taps, int cut);

/* Stand-alone Gaussian smoothing along time (data_augmentations.py:6-37): x, out device fp32 [B][T][D] ->
 * [B][T_out][D]; mode 1 = 'same', 2 = 'valid'.  Returns T_out. */
int b2t_gauss_smooth(const float* x, int B, int T, int D, float std, int size, int mode, float* out, void* stream);

/* CTC on the logits of the last b2t_forward.  labels: device int32 [B][Smax]; in_len/tgt_len: device
 * int32 [B]; loss_out: device fp32 [B] (per-trial, reduction 'none').  With want_grad, the gradient of
 * grad_scale * sum_b loss_b (grad_scale = 1/global_batch reproduces torch.mean) is kept inside the
 * engine for b2t_backward. */
int b2t_ctc_loss(b2t_engine* e, const int* labels, int Smax, const int* in_len, const int* tgt_len, float grad_scale,
                 float* loss_out, int want_grad, void* stream);
/* Stand-alone CTC on caller logits (device [T][B][C] fp32, the nn.CTCLoss layout).  dlogits may be NULL. */
int b2t_ctc_loss_tbc(const float* logits_tbc, int T, int B, int C, const int* labels, int Smax, const int* in_len,
                     const int* tgt_len, float grad_scale, float* loss_out, float* dlogits_tbc, void* workspace,
                     long long workspace_bytes, void* stream);
long long b2t_ctc_workspace_bytes(int T, int B, int Smax);

/* Provide d(loss)/d(logits) computed elsewhere (device [B][T'][n_classes]) instead of b2t_ctc_loss. */
int b2t_set_dlogits(b2t_engine* e, const float* dlogits, void* stream);
/* Back-propagate through head, GRU stack, patching and day layers into the flat gradient buffer.
 * Gradients of day layers absent from the batch are left untouched (their "touched" flag stays 0). */
int b2t_backward(b2t_engine* e, void* stream);

/* Data parallel: leave n_sms SMs free during the tail of backward (weight-gradient and layer-0 data-gradient GEMMs otherwise hold
 * every SM with one persistent CTA) so that the collective of the finished gradient buckets runs beside it.  0 = none (default). */
int b2t_set_comm_sms(b2t_engine* e, int n_sms);
/* Gradient buckets for data parallelism: contiguous ranges of the gradient buffer together with the point of b2t_backward after
 * which each is final, so that the caller's all-reduce of one bucket overlaps the rest of backward (the reference is single-GPU;
 * one all-reduce per step is the north-star's collective, here issued bucket by bucket).  After b2t_backward: bucket i in
 * completion order has the range [*offset, *offset + *count); b2t_grad_bucket_wait makes a stream wait for it. */
int b2t_grad_buckets(b2t_engine* e);
int b2t_grad_bucket(b2t_engine* e, int i, long long* offset, long long* count);
int b2t_grad_bucket_wait(b2t_engine* e, int i, void* stream);

typedef struct b2t_adamw_args {
  float lr[3];            /* per group: 0 biases, 1 day layers, 2 everything else */
  float weight_decay[3];
  float beta1, beta2, eps;
  float max_grad_norm;    /* <= 0 disables clipping */
} b2t_adamw_args;
/* stats_out: optional device fp32 [2] = {total gradient norm, clip coefficient}. */
int b2t_optimizer_step(b2t_engine* e, const b2t_adamw_args* a, float* stats_out, void* stream);
/* Per-segment AdamW step counters (device int32 [n_segments]) for checkpointing. */
int* b2t_step_counters(b2t_engine* e);

/* Greedy CTC decode + edit distance on the logits of the last forward (integer outputs). */
int b2t_greedy_edit(b2t_engine* e, const int* labels, int Smax, const int* in_len, const int* tgt_len, int* decoded,
                    int* dec_len, int* edit, void* stream);

/* Profiling aid: device buffer of 2*T'*8 int64; CTA 0 of the layer-0 forward and top-layer backward recurrence
 * kernels records clock64() at eight points of every time step.  NULL disables. */
int b2t_debug_set_trace(b2t_engine* e, long long* device_buf);

/* Profiling aid: with enable != 0 every task of forward/backward/optimizer is bracketed by CUDA events (lane = side stream,
 * 9 = caller's stream); b2t_debug_dump_timeline synchronises and writes "lane name start_us end_us" lines. */
int b2t_debug_timeline(int enable);
int b2t_debug_dump_timeline(char* buf, int cap);

/* Test hook: device pointer (and element count) of an internal activation buffer of the last forward/backward: "xs", "xd",
 * "dpre" (bf16 [Bpad][T][neural_dim], Bpad = batch rounded up to 16), per layer "hseq" (bf16 [(T'+1)][Bpad][H], slot 0 = initial
 * state), "hdrop" (bf16 [T'][Bpad][H], the dropped copy read by the next layer), "gx" (fp32 [T'][Bpad][3H]), "dGx", "dGh"
 * (bf16 [T'][Bpad][3H]), "dY" (fp32 [T'][Bpad][H]), "logits" (fp32 [T'][Bpad][64]).  Used by the dropout-mask parity tests. */
int b2t_debug_buffer(b2t_engine* e, const char* name, int layer, void** ptr, long long* elems);

/* Number of kernels this library has launched on behalf of the calling process (bench accounting). */
long long b2t_launch_count(void);

/* ------------------------------------------------------------------ n-gram CTC decoder (WFST token passing on the GPU)
 * Mirrors the lm_decoder pybind surface (lm_decoder.cc:51-75) with integer status codes instead of glog aborts.
 * A decoder owns `max_slots` independent utterance slots; slot 0 is what the single-utterance pybind API maps to,
 * b2t_decoder_decode_batch drives all of them concurrently (one CTA per utterance). */
typedef struct b2t_decode_options {   /* DecodeOptions(max_active, min_active, beam, lattice_beam, acoustic_scale, */
  int max_active;                     /*               blank_skip_threshold, length_penalty, nbest)               */
  int min_active;                     /* brain_speech_decoder.h:22-43                                              */
  float beam;
  float lattice_beam;
  float acoustic_scale;
  float blank_skip_threshold;
  float length_penalty;
  int nbest;
} b2t_decode_options;

typedef struct b2t_decoder b2t_decoder;

const char* b2t_decoder_last_error(void);
/* DecodeResource(fst_path, ..., dict_path, ...) + BrainSpeechDecoder(resource, opts) (brain_speech_decoder.h:45-98,110-148).
 * fst_path: OpenFST binary TLG.fst (vector/standard); words_path: words.txt symbol table.  max_frames bounds the frames
 * of one utterance; pools are sized from the options and overflow is reported as B2T_ERR_WORKSPACE. */
b2t_decoder* b2t_decoder_create(const char* fst_path, const char* words_path, const b2t_decode_options* opt, int max_frames,
                                int max_slots);
void b2t_decoder_destroy(b2t_decoder* d);
int b2t_decoder_set_options(b2t_decoder* d, const b2t_decode_options* opt);          /* SetOpt */
/* Search order.  0 (default): the two-pass parallel search, identical to the reference whenever max_active does not bind.
 * 1: strict serial order -- Kaldi's token-list order (HashList, hash-list-inl.h:124-171) and the online tightening of
 * next_cutoff (lattice-faster-decoder.cc:785-822) are reproduced, so token counts, 1-best and n-best equal the reference's
 * also when max_active binds; slower.  Switch between utterances only.  Env B2T_DECODER_STRICT=1 sets the default. */
int b2t_decoder_set_strict_order(b2t_decoder* d, int on);
int b2t_decoder_reset(b2t_decoder* d, int slot);                                      /* Reset */
/* Test hook: states and costs of the tokens of frame_plus_one fp1, in list order (strict mode) -- before finish only. */
int b2t_decoder_debug_frame_tokens(b2t_decoder* d, int slot, int fp1, int* states, float* costs, int cap);
/* DecodeNumpy: host logits [T][C] (+ optional log_priors [T][C]); log_softmax, minus priors, blank column minus
 * blank_penalty, then Decode().  After the call the slot's result list holds the partial 1-best. */
int b2t_decoder_decode_logits(b2t_decoder* d, int slot, const float* logits, const float* log_priors, int T, int C,
                              float blank_penalty);
int b2t_decoder_decode_logprobs(b2t_decoder* d, int slot, const float* logp, int T, int C);   /* DecodeNumpyLogProbs */
/* FinishDecoding: final costs, lattice pruning within lattice_beam (on the device), n-best (nbest > 1: top-n distinct word
 * sequences; nbest == 1: back-pointer best path on the device).  Idempotent: a second call keeps the first call's results. */
int b2t_decoder_finish(b2t_decoder* d, int slot);
/* DecodeResource.lm_fst_path / rescore_lm_fst_path (brain_speech_decoder.h:57-79): the LM the graph was built from and the
 * rescoring LM as OpenFST acceptors, loaded once; empty paths leave rescoring unavailable. */
int b2t_decoder_set_rescore_lms(b2t_decoder* d, const char* lm_fst_path, const char* rescore_lm_fst_path);
/* Rescore (brain_speech_decoder.cc:61-101), after finish: every distinct word sequence of the pruned lattice is re-scored as
 * graph' = graph - c_old + c_new (see b2t_lm_rescore_sequences); the result list keeps its first-pass length, best first. */
int b2t_decoder_rescore(b2t_decoder* d, int slot);
/* Host core of Rescore() (brain_speech_decoder.cc:47-101), no GPU involved: n word sequences (ids concatenated in `words`,
 * lengths in `lens`) with first-pass (graph, acoustic) costs are re-scored as graph' = graph - c_old + c_new, where c is the
 * cheapest path through the LM acceptor (OpenFST file; back-off arcs taken as epsilons, final cost included), and ordered by
 * graph' + acoustic; at most `keep` entries are written: order_out[i] = index of the i-th best input, graph_out[i] = graph'. */
int b2t_lm_rescore_sequences(const char* lm_fst_path, const char* rescore_lm_fst_path, int n, const int* words, const int* lens,
                             const float* graph, const float* acoustic, int keep, int* order_out, float* graph_out);
int b2t_decoder_num_results(b2t_decoder* d, int slot);                                /* len(result()) */
int b2t_decoder_get_result(b2t_decoder* d, int slot, int i, float* ac_score, float* lm_score, char* sentence, int cap);
/* Batched extension: reset + decode (+ finish) N <= max_slots utterances concurrently (one CTA per utterance for the search and
 * for the lattice pruning; the host part of finish runs on internal worker threads); logits host [N][T][C]. */
int b2t_decoder_decode_batch(b2t_decoder* d, const float* logits, const int* lens, int N, int T, int C, float blank_penalty,
                             int finish);
int b2t_decoder_stats(b2t_decoder* d, int slot, int* frames, long long* tokens, long long* links, double* kernel_ms);
int b2t_decoder_tokens_per_frame(b2t_decoder* d, int slot, int* out, int cap);

/* LM-free CTC prefix beam search (ctc_prefix_beam_search.cc:44-136), one warp per utterance; first_beam <= 64, second_beam <= 512.  The
 * per-frame candidates (second_beam * (first_beam + 1) * 32 B) live in shared memory when they fit its 227 KB (512 x 10 does), else in global memory (slower).
 * Hypotheses are walked in the iteration order of the reference's unordered_map (libstdc++), on which its Viterbi times depend.
 * logp: host [N][T][C] log-probabilities; lens: host [N].  Outputs (host, best first): ids [N][second_beam][max_len],
 * len / score / viterbi [N][second_beam], times [N][second_beam][max_len], n_hyp [N]. */
const char* b2t_prefix_last_error(void);
int b2t_prefix_beam_search(const float* logp, const int* lens, int N, int T, int C, int blank, int first_beam, int second_beam,
                           int max_len, int* out_ids, int* out_len, float* out_score, float* out_viterbi, int* out_times,
                           int* out_n);

/* ------------------------------------------------------------------ test hooks (also used by tests/) */
/* C[M,N] = A[M,K] * B^T with B given as [N,K] (b_mn == 0) or as [K,N] (b_mn == 1), A as [M,K]
 * (a_mn == 0) or [K,M] (a_mn == 1); bf16 inputs, fp32 or bf16 output, optional fp32 bias[N]. */
int b2t_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int a_mn, int b_mn, int out_bf16,
                  const float* bias, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2T_B200_H_ */
