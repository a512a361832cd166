"""Drop-in for ``model_training/rnn_trainer.py``: same class, constructor argument schema, method
names and returned dictionaries, with the training step executed by the sm_100a engine.

Reference: rnn_trainer.py:27-770 (BrainToTextDecoder_Trainer).  Differences that are deliberate:
  * the step (augmentation, smoothing, forward, CTC, backward, clip, AdamW) runs as native kernels
    through ``Engine`` instead of PyTorch library calls; ``torch.compile`` is not used;
  * data parallelism: when launched with torchrun (WORLD_SIZE > 1) every rank builds the trainer,
    processes its own batches and the flat gradient buffer is all-reduced ONCE per step over NCCL
    (the reference is single-GPU, rnn_trainer.py:85-107); only rank 0 writes files;
  * ``args`` may be a plain nested dict (OmegaConf is optional);
  * there is no CPU fallback (the reference falls back to CPU at rnn_trainer.py:98-107).
"""
from __future__ import annotations

import json
import logging
import math
import os
import pathlib
import pickle
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import LambdaLR
from torch.utils.data import DataLoader

from . import _native as N
from .datasets import SyntheticBrainToTextDataset
from .rnn_model import GRUDecoder


def _get(args, key, default=None):
    try:
        return args[key] if key in args else default
    except TypeError:
        return getattr(args, key, default)


class FusedClipAdamW(torch.optim.Optimizer):
    """torch.optim-compatible front for the fused gradient-norm / clip / AdamW kernel.

    param_groups mirror rnn_trainer.py:270-281 ('bias', 'day_layer', 'other'); LR schedulers act on
    them as usual.  ``step()`` launches one native kernel over the flat parameter buffer; parameters of
    day layers that no rank touched in this step are skipped exactly like ``grad is None`` in torch."""

    def __init__(self, model: GRUDecoder, param_groups, lr, betas, eps, weight_decay, max_grad_norm):
        self.model = model
        self.max_grad_norm = float(max_grad_norm)
        self.last_stats = None
        super().__init__(param_groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _group(self, kind):
        for g in self.param_groups:
            if g.get("group_type") == kind:
                return g
        return None

    @torch.no_grad()
    def step(self, closure=None):
        eng = self.model._engine
        if eng is None or not eng.training_capable:
            raise N.B2TError("FusedClipAdamW.step() needs a preceding training forward/backward")
        self.apply_pending_state(eng)
        gb, gd, go = self._group("bias"), self._group("day_layer"), self._group("other")
        gd = gd or go
        lr = [gb["lr"], gd["lr"], go["lr"]]
        wd = [gb["weight_decay"], gd["weight_decay"], go["weight_decay"]]
        b1, b2 = go["betas"]
        self.last_stats = eng.optimizer_step(lr, wd, b1, b2, go["eps"], self.max_grad_norm)
        self.model.weights_synced = True
        return None

    def zero_grad(self, set_to_none=True):
        # gradients live in the engine's flat buffer and are re-initialised by every backward pass
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    def state_dict(self):
        sd = super().state_dict()
        eng = self.model._engine
        if eng is not None and eng.training_capable:
            steps = self._steps(eng)
            state, idx = {}, 0
            for g in self.param_groups:
                for p in g["params"]:
                    name = self._name_of(p)
                    off, n = self.model._slots[name]
                    state[idx] = {"step": torch.tensor(float(steps[name])), "exp_avg": eng.exp_avg[off:off + n].view(p.shape).clone(),
                                  "exp_avg_sq": eng.exp_avg_sq[off:off + n].view(p.shape).clone()}
                    idx += 1
            sd["state"] = state
        return sd

    def _name_of(self, p):
        for n, q in self.model._named_flat:
            if q is p:
                return n
        raise KeyError("parameter not owned by the model")

    def _steps_tensor(self, eng):
        """The engine's per-segment AdamW step counters as a device int32 tensor (aliasing the engine's memory)."""
        from .engine import param_layout
        names = [n for n, _, _, _ in param_layout(self.model._cfg)]
        return names, eng.steps_tensor()

    def _steps(self, eng):
        names, t = self._steps_tensor(eng)
        host = t.cpu()
        return {n: int(host[i]) for i, n in enumerate(names)}

    def load_state_dict(self, state_dict):
        """Accepts the reference's optimizer checkpoints (per-parameter step / exp_avg / exp_avg_sq)."""
        self._pending_state = state_dict.get("state", {})
        groups = state_dict.get("param_groups", [])
        for g, saved in zip(self.param_groups, groups):
            for k in ("lr", "weight_decay", "betas", "eps", "initial_lr"):
                if k in saved:
                    g[k] = saved[k]

    def apply_pending_state(self, eng):
        pending = getattr(self, "_pending_state", None)
        if not pending:
            return
        names, steps = self._steps_tensor(eng)
        idx = 0
        for g in self.param_groups:
            for p in g["params"]:
                st = pending.get(idx)
                idx += 1
                if st is None:
                    continue
                name = self._name_of(p)
                off, n = self.model._slots[name]
                eng.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1).to(eng.device))
                eng.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1).to(eng.device))
                steps[names.index(name)] = int(float(st["step"]))
        self._pending_state = None


class _DevicePrefetcher:
    """Iterates a (pinned-memory) batch loader one batch ahead: the host -> device copy of batch i+1 runs on a side stream
    while step i computes (the reference copies inside the step, rnn_trainer.py:513-519, and its loss.item() every step keeps
    that copy from overlapping with anything).  Yields the reference's batch dict with the tensors already on the device."""

    def __init__(self, loader, device):
        self.it, self.device = iter(loader), device
        self.stream = torch.cuda.Stream(device=device)
        self.batch = self.event = None
        self._load()

    def _load(self):
        try:
            b = next(self.it)
        except StopIteration:
            self.batch = None
            return
        max_len = int(b['phone_seq_lens'].max()) if torch.is_tensor(b.get('phone_seq_lens')) and b['phone_seq_lens'].numel() else None
        with torch.cuda.stream(self.stream):
            self.batch = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) and k != 'transcriptions' else v) for k, v in b.items()}
            self.batch['max_phone_seq_len'] = max_len      # the reference pads seq_class_ids to a fixed width; the CTC kernels want the real one
            self.event = torch.cuda.Event()
            self.event.record(self.stream)

    def __iter__(self):
        return self

    def __next__(self):
        if self.batch is None:
            raise StopIteration
        cur = self.batch
        torch.cuda.current_stream().wait_event(self.event)
        for v in cur.values():
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(torch.cuda.current_stream())          # allocated on the copy stream, consumed on this one
        self._load()
        return cur


class BrainToTextDecoder_Trainer:
    """
    This class will initialize and train a brain-to-text phoneme decoder
    """

    def __init__(self, args):
        self.args = args
        self.logger = None
        self.device = None
        self.model = None
        self.optimizer = None
        self.learning_rate_scheduler = None
        self.ctc_loss = None
        self.best_val_PER = torch.inf
        self.best_val_loss = torch.inf
        self.train_dataset = self.val_dataset = self.train_loader = self.val_loader = None
        self.transform_args = self.args['dataset']['data_transforms']

        # --- data-parallel context (torchrun); single process otherwise
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.is_main = self.rank == 0

        if args['mode'] == 'train' and self.is_main:
            os.makedirs(self.args['output_dir'], exist_ok=False)
        if self.is_main and (args['save_best_checkpoint'] or args['save_all_val_steps'] or args['save_final_model']):
            os.makedirs(self.args['checkpoint_dir'], exist_ok=False)

        self.logger = logging.getLogger(__name__ + f".r{self.rank}")
        for handler in self.logger.handlers[:]:
            self.logger.removeHandler(handler)
        self.logger.setLevel(logging.INFO if self.is_main else logging.WARNING)
        formatter = logging.Formatter(fmt='%(asctime)s: %(message)s')
        if args['mode'] == 'train' and self.is_main:
            fh = logging.FileHandler(str(pathlib.Path(self.args['output_dir'], 'training_log')))
            fh.setFormatter(formatter)
            self.logger.addHandler(fh)
        sh = logging.StreamHandler(sys.stdout)
        sh.setFormatter(formatter)
        self.logger.addHandler(sh)

        if not torch.cuda.is_available():
            raise N.B2TError("BrainToTextDecoder_Trainer (b2t_b200) needs a CUDA sm_100a device; there is no CPU fallback")
        if self.world_size > 1:
            gpu_num = self.local_rank
        else:
            gpu_num = _get(self.args, 'gpu_number', 0)
            try:
                gpu_num = int(gpu_num)
            except ValueError:
                self.logger.warning(f"Invalid gpu_number value: {gpu_num}. Using 0 instead.")
                gpu_num = 0
            if gpu_num > torch.cuda.device_count() - 1:
                self.logger.warning(f"Requested GPU {gpu_num} not available. Using GPU 0 instead.")
                gpu_num = 0
        self.device = torch.device(f"cuda:{gpu_num}")
        torch.cuda.set_device(self.device)
        if self.world_size > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.device)
        self.logger.info(f'Using device: {self.device}')

        if self.args['seed'] != -1:
            np.random.seed(self.args['seed'])
            random.seed(self.args['seed'])
            torch.manual_seed(self.args['seed'])
        # augmentation draws (random_cut, device RNG seed of noise and dropout): one host stream per rank, so that data-parallel
        # ranks apply different noise/dropout patterns while the weight init above stays identical (and is broadcast anyway)
        self._aug_rng = np.random.RandomState(None if self.args['seed'] == -1 else (int(self.args['seed']) * 9973 + 17 + self.rank) % (2 ** 31))

        self.model = GRUDecoder(
            neural_dim=self.args['model']['n_input_features'],
            n_units=self.args['model']['n_units'],
            n_days=len(self.args['dataset']['sessions']),
            n_classes=self.args['dataset']['n_classes'],
            rnn_dropout=self.args['model']['rnn_dropout'],
            input_dropout=self.args['model']['input_network']['input_layer_dropout'],
            n_layers=self.args['model']['n_layers'],
            patch_size=self.args['model']['patch_size'],
            patch_stride=self.args['model']['patch_stride'],
        )
        self.logger.info("Initialized RNN decoding model (b2t_b200 native engine)")
        total_params = sum(p.numel() for p in self.model.parameters())
        self.logger.info(f"Model has {total_params:,} parameters")
        day_params = sum(p.numel() for name, p in self.model.named_parameters() if 'day' in name)
        self.logger.info(f"Model has {day_params:,} day-specific parameters | {((day_params / total_params) * 100):.2f}% of total parameters")

        self._build_datasets()

        self.optimizer = self.create_optimizer()
        if self.args['lr_scheduler_type'] == 'linear':
            self.learning_rate_scheduler = torch.optim.lr_scheduler.LinearLR(
                optimizer=self.optimizer, start_factor=1.0, end_factor=self.args['lr_min'] / self.args['lr_max'],
                total_iters=self.args['lr_decay_steps'])
        elif self.args['lr_scheduler_type'] == 'cosine':
            self.learning_rate_scheduler = self.create_cosine_lr_scheduler(self.optimizer)
        else:
            raise ValueError(f"Invalid learning rate scheduler type: {self.args['lr_scheduler_type']}")

        self.ctc_loss = self._ctc_loss_callable

        if self.args['init_from_checkpoint']:
            self.load_model_checkpoint(self.args['init_checkpoint_path'])

        for name, param in self.model.named_parameters():
            if not self.args['model']['rnn_trainable'] and 'gru' in name:
                param.requires_grad = False
            elif not self.args['model']['input_network']['input_trainable'] and 'day' in name:
                param.requires_grad = False
        if not self.args['model']['rnn_trainable'] or not self.args['model']['input_network']['input_trainable']:
            raise N.B2TError("frozen parameter groups are not supported by the fused optimizer yet")

        self.model.to(self.device)
        if self.world_size > 1:   # identical initial weights on every rank
            dist.broadcast(self.model.flat_parameters, src=0)

    # ------------------------------------------------------------------ data
    def _build_datasets(self):
        ds = self.args['dataset']
        synth = _get(ds, 'synthetic', None)
        n_days = len(ds['sessions'])
        if synth is not None:
            common = dict(batch_size=ds['batch_size'], n_days=n_days, neural_dim=self.args['model']['n_input_features'],
                          n_classes=ds['n_classes'], T=_get(synth, 'T', 400), min_len=_get(synth, 'min_len', 6),
                          max_len=_get(synth, 'max_len', 14), noise=_get(synth, 'noise', 0.6))
            self.train_dataset = SyntheticBrainToTextDataset(n_batches=self.args['num_training_batches'], days_per_batch=ds['days_per_batch'],
                                                             seed=ds['seed'] * 1000 + self.rank, split="train", **common)
            self.val_dataset = SyntheticBrainToTextDataset(n_batches=_get(synth, 'val_batches', 8), days_per_batch=1, seed=ds['seed'],
                                                           split="test", **common)
        else:
            from .dataset import BrainToTextDataset, train_test_split_indicies      # hdf5 through h5py when importable, .npz shards otherwise
            train_paths = [os.path.join(ds["dataset_dir"], s, 'data_train.hdf5') for s in ds['sessions']]
            val_paths = [os.path.join(ds["dataset_dir"], s, 'data_val.hdf5') for s in ds['sessions']]
            if len(set(train_paths)) != len(train_paths):
                raise ValueError("There are duplicate sessions listed in the train dataset")
            train_trials, _ = train_test_split_indicies(file_paths=train_paths, test_percentage=0, seed=ds['seed'], bad_trials_dict=None)
            _, val_trials = train_test_split_indicies(file_paths=val_paths, test_percentage=1, seed=ds['seed'], bad_trials_dict=None)
            if self.is_main:
                with open(os.path.join(self.args['output_dir'], 'train_val_trials.json'), 'w') as f:
                    json.dump({'train': train_trials, 'val': val_trials}, f)
            fs = _get(ds, 'feature_subset', None)
            self.train_dataset = BrainToTextDataset(trial_indicies=train_trials, split='train', days_per_batch=ds['days_per_batch'],
                                                    n_batches=self.args['num_training_batches'], batch_size=ds['batch_size'],
                                                    must_include_days=None, random_seed=ds['seed'] + self.rank, feature_subset=fs)
            self.val_dataset = BrainToTextDataset(trial_indicies=val_trials, split='test', days_per_batch=None, n_batches=None,
                                                  batch_size=ds['batch_size'], must_include_days=None, random_seed=ds['seed'], feature_subset=fs)
        if synth is None and _get(ds, 'pinned_prefetch', True):
            # batches assembled inside pinned host buffers on a background thread (dataset.py: PinnedBatchLoader)
            from .dataset import PinnedBatchLoader
            self.train_loader = PinnedBatchLoader(self.train_dataset, max_T=int(_get(ds, 'max_time_steps', 2048)),
                                                  neural_dim=self.args['model']['n_input_features'] if not _get(ds, 'feature_subset', None) else len(ds['feature_subset']),
                                                  shuffle=ds['loader_shuffle'])
        else:
            self.train_loader = DataLoader(self.train_dataset, batch_size=None, shuffle=ds['loader_shuffle'],
                                           num_workers=ds['num_dataloader_workers'], pin_memory=True)
        self.val_loader = DataLoader(self.val_dataset, batch_size=None, shuffle=False, num_workers=0, pin_memory=True)
        self.logger.info("Successfully initialized datasets")

    # ------------------------------------------------------------------ optimizer / schedule
    def create_optimizer(self):
        '''
        Create the optimizer with special param groups (rnn_trainer.py:259-292): biases and day weights are not
        decayed; day weights have a separate learning rate.
        '''
        bias_params = [p for name, p in self.model.named_parameters() if 'gru.bias' in name or 'out.bias' in name]
        day_params = [p for name, p in self.model.named_parameters() if 'day_' in name]
        other_params = [p for name, p in self.model.named_parameters() if 'day_' not in name and 'gru.bias' not in name and 'out.bias' not in name]
        if len(day_params) != 0:
            param_groups = [
                {'params': bias_params, 'weight_decay': 0, 'group_type': 'bias'},
                {'params': day_params, 'lr': self.args['lr_max_day'], 'weight_decay': self.args['weight_decay_day'], 'group_type': 'day_layer'},
                {'params': other_params, 'group_type': 'other'},
            ]
        else:
            param_groups = [{'params': bias_params, 'weight_decay': 0, 'group_type': 'bias'}, {'params': other_params, 'group_type': 'other'}]
        return FusedClipAdamW(self.model, param_groups, lr=self.args['lr_max'], betas=(self.args['beta0'], self.args['beta1']),
                              eps=self.args['epsilon'], weight_decay=self.args['weight_decay'],
                              max_grad_norm=self.args['grad_norm_clip_value'])

    def create_cosine_lr_scheduler(self, optim):
        lr_max, lr_min, lr_decay_steps = self.args['lr_max'], self.args['lr_min'], self.args['lr_decay_steps']
        lr_max_day, lr_min_day, lr_decay_steps_day = self.args['lr_max_day'], self.args['lr_min_day'], self.args['lr_decay_steps_day']
        lr_warmup_steps, lr_warmup_steps_day = self.args['lr_warmup_steps'], self.args['lr_warmup_steps_day']

        def lr_lambda(current_step, min_lr_ratio, decay_steps, warmup_steps):
            if current_step < warmup_steps:
                return float(current_step) / float(max(1, warmup_steps))
            if current_step < decay_steps:
                progress = float(current_step - warmup_steps) / float(max(1, decay_steps - warmup_steps))
                cosine_decay = 0.5 * (1 + math.cos(math.pi * progress))
                return max(min_lr_ratio, min_lr_ratio + (1 - min_lr_ratio) * cosine_decay)
            return min_lr_ratio

        main = lambda step: lr_lambda(step, lr_min / lr_max, lr_decay_steps, lr_warmup_steps)                       # noqa: E731
        day = lambda step: lr_lambda(step, lr_min_day / lr_max_day, lr_decay_steps_day, lr_warmup_steps_day)         # noqa: E731
        if len(optim.param_groups) == 3:
            lambdas = [main, day, main]
        elif len(optim.param_groups) == 2:
            lambdas = [main, main]
        else:
            raise ValueError(f"Invalid number of param groups in optimizer: {len(optim.param_groups)}")
        return LambdaLR(optim, lambdas, -1)

    # ------------------------------------------------------------------ checkpoints
    def load_model_checkpoint(self, load_path):
        checkpoint = torch.load(load_path, weights_only=False, map_location="cpu")
        sd = {k.replace("_orig_mod.", "").replace("module.", ""): v for k, v in checkpoint['model_state_dict'].items()}
        self.model.load_state_dict(sd)
        self.learning_rate_scheduler.load_state_dict(checkpoint['scheduler_state_dict'])
        if checkpoint.get('optimizer_state_dict') is not None:
            self.optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
        self.best_val_PER = checkpoint['val_PER']
        self.best_val_loss = checkpoint['val_loss'] if 'val_loss' in checkpoint.keys() else torch.inf
        self.model.to(self.device)
        self.logger.info("Loaded model from checkpoint: " + load_path)

    def save_model_checkpoint(self, save_path, PER, loss=None):
        if not self.is_main:
            return
        checkpoint = {
            'model_state_dict': self.model.state_dict(),
            'optimizer_state_dict': self.optimizer.state_dict(),
            'scheduler_state_dict': self.learning_rate_scheduler.state_dict(),
            'val_PER': PER,
            'val_loss': loss,
        }
        torch.save(checkpoint, save_path)
        self.logger.info("Saved model to checkpoint: " + save_path)
        with open(os.path.join(self.args['checkpoint_dir'], 'args.yaml'), 'w') as f:
            try:
                from omegaconf import OmegaConf
                OmegaConf.save(config=self.args, f=f)
            except Exception:  # noqa: BLE001
                import yaml
                yaml.safe_dump(json.loads(json.dumps(self.args, default=lambda o: dict(o))), f)

    # ------------------------------------------------------------------ loss / transform (API parity)
    def _ctc_loss_callable(self, log_probs, targets, input_lengths, target_lengths):
        """self.ctc_loss(log_probs[T,N,C], targets[N,S], input_lengths[N], target_lengths[N]) -> loss[N]
        (torch.nn.CTCLoss(blank=0, reduction='none', zero_infinity=False), rnn_trainer.py:242)."""
        from .ctc import ctc_loss
        return ctc_loss(log_probs, targets, input_lengths, target_lengths)

    def transform_data(self, features, n_time_steps, mode='train'):
        '''
        Apply augmentations and smoothing (rnn_trainer.py:436-484) as a stand-alone call.  The fused training
        step does the same inside the engine's input kernel; this method exists for API parity.
        '''
        from .data_augmentations import gauss_smooth
        ta = self.transform_args
        data_shape = features.shape
        if mode == 'train':
            if ta['static_gain_std'] > 0:
                warp = torch.eye(data_shape[-1], device=self.device).unsqueeze(0).repeat(data_shape[0], 1, 1)
                warp += torch.randn_like(warp) * ta['static_gain_std']
                features = torch.matmul(features, warp)
            if ta['white_noise_std'] > 0:
                features = features + torch.randn(data_shape, device=self.device) * ta['white_noise_std']
            if ta['constant_offset_std'] > 0:
                features = features + torch.randn((data_shape[0], 1, data_shape[-1]), device=self.device) * ta['constant_offset_std']
            if ta['random_walk_std'] > 0:
                features = features + torch.cumsum(torch.randn(data_shape, device=self.device) * ta['random_walk_std'], dim=ta['random_walk_axis'])
            if ta['random_cut'] > 0:
                cut = np.random.randint(0, ta['random_cut'])
                features = features[:, cut:, :]
                n_time_steps = n_time_steps - cut
        if ta['smooth_data']:
            features = gauss_smooth(inputs=features, device=self.device, smooth_kernel_std=ta['smooth_kernel_std'],
                                    smooth_kernel_size=ta['smooth_kernel_size'])
        return features, n_time_steps

    # ------------------------------------------------------------------ fused step
    def _train_step(self, batch):
        ta = self.transform_args
        features = batch['input_features'].to(self.device, non_blocking=True)
        labels = batch['seq_class_ids'].to(self.device, non_blocking=True)
        n_time_steps = batch['n_time_steps'].to(self.device, non_blocking=True)
        phone_seq_lens = batch['phone_seq_lens'].to(self.device, non_blocking=True)
        day_indicies = batch['day_indicies'].to(self.device, non_blocking=True)
        B, T, _ = features.shape
        if ta['static_gain_std'] > 0 or ta['random_walk_std'] > 0:
            # rarely used augmentations (off in rnn_args.yaml): applied with torch ops before the fused kernel
            if ta['static_gain_std'] > 0:
                warp = torch.eye(features.shape[-1], device=self.device).unsqueeze(0).repeat(B, 1, 1)
                warp += torch.randn_like(warp) * ta['static_gain_std']
                features = torch.matmul(features, warp)
            if ta['random_walk_std'] > 0:
                features = features + torch.cumsum(torch.randn_like(features) * ta['random_walk_std'], dim=ta['random_walk_axis'])
        cut = int(self._aug_rng.randint(0, ta['random_cut'])) if ta['random_cut'] > 0 else 0
        eng = self.model.engine(B, T, training=True)
        self.model.fused_updates = True
        seed = int(self._aug_rng.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(self._aug_rng.randint(0, 2 ** 31 - 1))
        if self.world_size > 1:
            eng.reserve_comm_sms(int(os.environ.get("B2T_COMM_SMS", "0")))   # (only useful with B2T_DP_BUCKETS=bucketed: SMs the backward tail leaves to the overlapping collectives)
        eng.forward(features, day_indicies, training=True, smooth_mode=1 if ta['smooth_data'] else 0,
                    smooth_std=float(ta['smooth_kernel_std']), smooth_size=int(ta['smooth_kernel_size']), cut=cut,
                    white_noise_std=float(ta['white_noise_std']), offset_noise_std=float(ta['constant_offset_std']), seed=seed,
                    want_logits=False)
        ps, st = self.args['model']['patch_size'], self.args['model']['patch_stride']
        adjusted_lens = (((n_time_steps - cut) - ps) / st + 1).to(torch.int32)
        loss_vec = eng.ctc_loss(labels, adjusted_lens, phone_seq_lens, grad_scale=1.0 / (B * self.world_size),
                                max_target_len=batch.get('max_phone_seq_len'))
        eng.backward()
        if self.world_size > 1:
            eng.all_reduce_grads()                  # the step's gradient all-reduce (flat gradients + day-touched flags), bucket by bucket behind backward
        self.optimizer.step()
        self.learning_rate_scheduler.step()
        loss = loss_vec.mean()
        return loss, self.optimizer.last_stats

    def train(self):
        '''
        Train the model
        '''
        self.model.train()
        train_losses, val_losses, val_PERs, val_results = [], [], [], []
        val_steps_since_improvement = 0
        save_best_checkpoint = _get(self.args, 'save_best_checkpoint', True)
        early_stopping = _get(self.args, 'early_stopping', True)
        early_stopping_val_steps = self.args['early_stopping_val_steps']
        train_start_time = time.time()
        i = -1
        for i, batch in enumerate(_DevicePrefetcher(self.train_loader, self.device)):
            self.model.train()
            start_time = time.time()
            loss, stats = self._train_step(batch)
            if not torch.isfinite(stats[0]).item() and self.args['grad_norm_clip_value'] > 0:
                raise RuntimeError("The total norm of the gradients is non-finite, so it cannot be clipped (error_if_nonfinite)")
            loss_value = loss.item()                 # the reference synchronises here every step as well (rnn_trainer.py:562)
            train_step_duration = time.time() - start_time
            train_losses.append(loss_value)
            if i % self.args['batches_per_train_log'] == 0:
                self.logger.info(f'Train batch {i}: loss: {loss_value:.2f} grad norm: {stats[0].item():.2f} time: {train_step_duration:.3f}')
            if i % self.args['batches_per_val_step'] == 0 or i == (self.args['num_training_batches'] - 1):
                self.logger.info(f"Running test after training batch: {i}")
                start_time = time.time()
                val_metrics = self.validation(loader=self.val_loader, return_logits=self.args['save_val_logits'], return_data=self.args['save_val_data'])
                val_step_duration = time.time() - start_time
                self.logger.info(f'Val batch {i}: PER (avg): {val_metrics["avg_PER"]:.4f} CTC Loss (avg): {val_metrics["avg_loss"]:.4f} time: {val_step_duration:.3f}')
                if self.args['log_individual_day_val_PER']:
                    for day in val_metrics['day_PERs'].keys():
                        d = val_metrics['day_PERs'][day]
                        if d['total_seq_length'] > 0:
                            self.logger.info(f"{self.args['dataset']['sessions'][day]} val PER: {d['total_edit_distance'] / d['total_seq_length']:0.4f}")
                val_PERs.append(val_metrics['avg_PER'])
                val_losses.append(val_metrics['avg_loss'])
                val_results.append(val_metrics)
                new_best = False
                if val_metrics['avg_PER'] < self.best_val_PER:
                    self.logger.info(f"New best test PER {self.best_val_PER:.4f} --> {val_metrics['avg_PER']:.4f}")
                    self.best_val_PER, self.best_val_loss, new_best = val_metrics['avg_PER'], val_metrics['avg_loss'], True
                elif val_metrics['avg_PER'] == self.best_val_PER and (val_metrics['avg_loss'] < self.best_val_loss):
                    self.logger.info(f"New best test loss {self.best_val_loss:.4f} --> {val_metrics['avg_loss']:.4f}")
                    self.best_val_loss, new_best = val_metrics['avg_loss'], True
                if new_best:
                    if save_best_checkpoint:
                        self.logger.info("Checkpointing model")
                        self.save_model_checkpoint(f'{self.args["checkpoint_dir"]}/best_checkpoint', self.best_val_PER, self.best_val_loss)
                    if self.args['save_val_metrics'] and self.is_main and os.path.isdir(self.args['checkpoint_dir']):
                        with open(f'{self.args["checkpoint_dir"]}/val_metrics.pkl', 'wb') as f:
                            pickle.dump(val_metrics, f)
                    val_steps_since_improvement = 0
                else:
                    val_steps_since_improvement += 1
                if self.args['save_all_val_steps']:
                    self.save_model_checkpoint(f'{self.args["checkpoint_dir"]}/checkpoint_batch_{i}', val_metrics['avg_PER'], val_metrics['avg_loss'])
                if early_stopping and (val_steps_since_improvement >= early_stopping_val_steps):
                    self.logger.info(f'Overall validation PER has not improved in {early_stopping_val_steps} validation steps. Stopping training early at batch: {i}')
                    break
        training_duration = time.time() - train_start_time
        self.logger.info(f'Best avg val PER achieved: {self.best_val_PER:.5f}')
        self.logger.info(f'Total training time: {(training_duration / 60):.2f} minutes')
        if self.args['save_final_model'] and val_PERs:
            self.save_model_checkpoint(f'{self.args["checkpoint_dir"]}/final_checkpoint_batch_{i}', val_PERs[-1], val_losses[-1])
        return {'train_losses': train_losses, 'val_losses': val_losses, 'val_PERs': val_PERs, 'val_metrics': val_results}

    def validation(self, loader, return_logits=False, return_data=False):
        '''
        Calculate metrics on the validation dataset (rnn_trainer.py:653-770).  Greedy decode and edit distance are
        integer kernels on the device; only their results come back to the host.
        '''
        self.model.eval()
        ta = self.transform_args
        metrics = {}
        if return_logits:
            metrics['logits'], metrics['n_time_steps'] = [], []
        if return_data:
            metrics['input_features'] = []
        for k in ('decoded_seqs', 'true_seq', 'phone_seq_lens', 'transcription', 'losses', 'block_nums', 'trial_nums', 'day_indicies'):
            metrics[k] = []
        total_edit_distance, total_seq_length = 0, 0
        probs = _get(self.args['dataset'], 'dataset_probability_val', None)
        n_sessions = len(self.args['dataset']['sessions'])
        day_per = {d: {'total_edit_distance': 0, 'total_seq_length': 0} for d in range(n_sessions) if probs is None or probs[d] == 1}
        ps, st = self.args['model']['patch_size'], self.args['model']['patch_stride']
        for i, batch in enumerate(loader):
            features = batch['input_features'].to(self.device)
            labels = batch['seq_class_ids'].to(self.device)
            n_time_steps = batch['n_time_steps'].to(self.device)
            phone_seq_lens = batch['phone_seq_lens'].to(self.device)
            day_indicies = batch['day_indicies'].to(self.device)
            day = int(batch['day_indicies'][0].item())
            if probs is not None and probs[day] == 0:
                if _get(self.args, 'log_val_skip_logs', False):
                    self.logger.info(f"Skipping validation on day {day}")
                continue
            with torch.no_grad():
                B, T, _ = features.shape
                eng = self.model.engine(B, T, training=False)
                logits, _ = eng.forward(features, day_indicies, training=False, smooth_mode=1 if ta['smooth_data'] else 0,
                                        smooth_std=float(ta['smooth_kernel_std']), smooth_size=int(ta['smooth_kernel_size']),
                                        want_logits=return_logits)
                adjusted_lens = ((n_time_steps - ps) / st + 1).to(torch.int32)
                max_len = int(batch['phone_seq_lens'].max()) if not batch['phone_seq_lens'].is_cuda else None
                loss = eng.ctc_loss(labels, adjusted_lens, phone_seq_lens, grad_scale=1.0, want_grad=False, max_target_len=max_len).mean()
                dec, dlen, ed = eng.greedy_edit(labels, adjusted_lens, phone_seq_lens, max_target_len=max_len)
            metrics['losses'].append(loss.cpu().detach().numpy())
            dec_h, dlen_h, ed_h = dec.cpu().numpy(), dlen.cpu().numpy(), ed.cpu().numpy()
            decoded_seqs = [dec_h[b, :dlen_h[b]] for b in range(dec_h.shape[0])]
            batch_edit_distance = int(ed_h.sum())
            day_per[day]['total_edit_distance'] += batch_edit_distance
            day_per[day]['total_seq_length'] += torch.sum(phone_seq_lens).item()
            total_edit_distance += batch_edit_distance
            total_seq_length += torch.sum(phone_seq_lens)
            if return_logits:
                metrics['logits'].append(logits.cpu().float().numpy())
                metrics['n_time_steps'].append(adjusted_lens.cpu().numpy())
            if return_data:
                metrics['input_features'].append(batch['input_features'].cpu().numpy())
            metrics['decoded_seqs'].append(decoded_seqs)
            metrics['true_seq'].append(batch['seq_class_ids'].cpu().numpy())
            metrics['phone_seq_lens'].append(batch['phone_seq_lens'].cpu().numpy())
            metrics['transcription'].append(batch['transcriptions'].cpu().numpy())
            metrics['losses'].append(loss.detach().item())
            metrics['block_nums'].append(batch['block_nums'].numpy())
            metrics['trial_nums'].append(batch['trial_nums'].numpy())
            metrics['day_indicies'].append(batch['day_indicies'].cpu().numpy())
        avg_PER = total_edit_distance / total_seq_length
        metrics['day_PERs'] = day_per
        metrics['avg_PER'] = avg_PER.item()
        metrics['avg_loss'] = np.mean(metrics['losses'])
        return metrics
