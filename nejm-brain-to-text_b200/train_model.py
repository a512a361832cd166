"""Entry point mirroring ``model_training/train_model.py:1-6``: load rnn_args.yaml, build the trainer, train.

    python -m nejm_brain_to_text_b200.train_model [rnn_args.yaml]            (after b2t_pkg.load())
    torchrun --nproc-per-node 8 ... train_model.py rnn_args.yaml             (data parallel, one rank per GPU)
"""
import sys


def main(path="rnn_args.yaml"):
    try:
        from omegaconf import OmegaConf
        args = OmegaConf.load(path)
    except ImportError:
        import yaml
        with open(path) as f:
            args = yaml.safe_load(f)
    from .rnn_trainer import BrainToTextDecoder_Trainer
    trainer = BrainToTextDecoder_Trainer(args)
    metrics = trainer.train()
    return metrics


if __name__ == "__main__":
    main(*sys.argv[1:2])
