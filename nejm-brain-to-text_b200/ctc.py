"""CTC loss on the sm_100a kernel with the call signature the trainer uses for ``self.ctc_loss``.

Reference: torch.nn.CTCLoss(blank=0, reduction='none', zero_infinity=False) at rnn_trainer.py:242,
called as ``ctc_loss(log_probs[T,N,C], targets[N,S], input_lengths[N], target_lengths[N]) -> [N]``
(rnn_trainer.py:538-543).  The kernel normalises its input with a log-softmax, which is the identity
on log-probabilities, so it accepts either logits or log-probs.
"""
import torch

from . import _native as N


class _CTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths):
        if not log_probs.is_cuda:
            raise N.B2TError("ctc_loss (b2t_b200) needs CUDA tensors; there is no CPU path")
        lp = log_probs.detach().contiguous().float()
        T, B, C = lp.shape
        dev = lp.device
        tg = targets.to(device=dev, dtype=torch.int32).contiguous()
        if tg.dim() != 2:
            raise ValueError("targets must be [N, S] (padded), as the trainer passes them")
        il = input_lengths.to(device=dev, dtype=torch.int32).contiguous()
        tl = target_lengths.to(device=dev, dtype=torch.int32).contiguous()
        if tg.shape[1] > 64 and tl.numel() > 0:
            # the reference dataset pads seq_class_ids to a fixed width (500): the kernels pick their variant from the padded
            # width, so cut it down to the longest target (one small D2H read; only when the padding is large enough to matter)
            tg = tg[:, :max(1, min(int(target_lengths.max()), tg.shape[1]))].contiguous()
        S = max(int(tg.shape[1]), 1)
        ws = torch.empty(N.lib.b2t_ctc_workspace_bytes(T, B, S), dtype=torch.uint8, device=dev)
        loss = torch.empty(B, device=dev)
        need_grad = log_probs.requires_grad
        grad = torch.empty_like(lp) if need_grad else None
        N.check(N.lib.b2t_ctc_loss_tbc(lp.data_ptr(), T, B, C, tg.data_ptr(), S, il.data_ptr(), tl.data_ptr(), 1.0, loss.data_ptr(),
                                       grad.data_ptr() if need_grad else None, ws.data_ptr(), ws.numel(),
                                       torch.cuda.current_stream().cuda_stream), "b2t_ctc_loss_tbc")
        ctx.save_for_backward(grad)
        ctx.in_dtype = log_probs.dtype
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return None, None, None, None
        return (grad * gloss.view(1, -1, 1)).to(ctx.in_dtype), None, None, None


def ctc_loss(log_probs, targets, input_lengths, target_lengths):
    return _CTCFn.apply(log_probs, targets, input_lengths, target_lengths)
