#!/usr/bin/env python3
"""Drop-in for ``python -m score_based_channels.test_score`` (reference
``src/score_based_channels/test_score.py``): same flags, same ``results.pt`` keys and array shapes.

    python -m score_based_channels_b200.test_score [--gpu 0] [--train CDL-C] [--test CDL-C]
                                                   [--spacing 0.5] [--pilot_alpha 0.6]

Differences in execution only: the 17 SNR points are stacked on the batch axis and the whole
2311 x 3-step trajectory of every (SNR, channel) pair runs in one launch of the fused kernel; under
torchrun the (SNR x channel) axis is sharded over the GPUs with one all-gather of the NMSE log.
Extra, optional flags (not in the reference): --ckpt, --out_dir, --levels, --num_channels, --precision,
--seed, --no_plot."""
import argparse
import copy
import itertools
import os
import sys

import numpy as np
import torch

from . import entry_common as ec
from .loaders import Channels


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpu', type=int, default=0)
    parser.add_argument('--train', type=str, default='CDL-C')
    parser.add_argument('--test', type=str, default='CDL-C')
    parser.add_argument('--save_channels', type=int, default=0)
    parser.add_argument('--spacing', nargs='+', type=float, default=[0.5])
    parser.add_argument('--pilot_alpha', nargs='+', type=float, default=[0.6])
    # extras
    parser.add_argument('--ckpt', type=str, default=None, help='checkpoint path (default ./models/score/<train>/final_model.pt)')
    parser.add_argument('--out_dir', type=str, default=None)
    parser.add_argument('--levels', type=int, default=None, help='run only the first N sigma levels (debug)')
    parser.add_argument('--num_channels', type=int, default=100)
    parser.add_argument('--precision', type=str, default=None, choices=[None, 'auto', 'tf32x3', 'tf32', 'fp16x2'])
    parser.add_argument('--seed', type=int, default=None, help='seed torch / numpy / the sampler RNG (reference: unseeded)')
    parser.add_argument('--no_plot', action='store_true')
    args = parser.parse_args(argv)

    dev, rank, ws = ec.pick_device(args.gpu)
    if args.seed is not None:
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)
    sampler_seed = ec.common_seed(args.seed, dev, ws)

    # Target file (test_score.py:32-36)
    target_dir = './models/score/%s' % args.train
    target_file = args.ckpt or os.path.join(target_dir, 'final_model.pt')
    contents = ec.load_checkpoint(target_file, args.train)
    config = contents['config']

    # Default hyper-parameters for pilot_alpha = 0.6, all SNR points (test_score.py:38-54)
    alpha_step, beta_noise = 3e-11, 0.01
    config.sampling.steps_each = 3

    diffuser = ec.build_model(config, contents['model_state'], dev, args.precision)

    train_seed, val_seed = 1234, 4321
    config.data.channel = args.train
    dataset = Channels(train_seed, config, norm=config.data.norm_channels)

    snr_range = np.arange(-10, 32.5, 2.5)
    spacing_range = np.asarray(args.spacing)
    pilot_alpha_range = np.asarray(args.pilot_alpha)
    noise_range = 10 ** (-snr_range / 10.) * config.data.image_size[1]
    num_channels = args.num_channels
    num_levels = int(config.model.num_classes) if args.levels is None else int(args.levels)
    steps_each = int(config.sampling.steps_each)

    nmse_log = np.zeros((len(spacing_range), len(pilot_alpha_range), len(snr_range),
                         int(num_levels * steps_each), num_channels))
    result_dir = args.out_dir or './results/score/train-%s_test-%s' % (args.train, args.test)
    if rank == 0:
        os.makedirs(result_dir, exist_ok=True)

    meta_params = itertools.product(spacing_range, pilot_alpha_range)
    for meta_idx, (spacing, pilot_alpha) in enumerate(meta_params):
        spacing_idx, pilot_alpha_idx = np.unravel_index(meta_idx, (len(spacing_range), len(pilot_alpha_range)))
        val_config = copy.deepcopy(config)
        val_config.data.channel = args.test
        val_config.data.spacing_list = [spacing]
        val_config.data.num_pilots = int(np.floor(config.data.image_size[1] * pilot_alpha))
        val_dataset = Channels(val_seed, val_config, norm=[dataset.mean, dataset.std], allow_other_seed=False)
        print('There are %d validation channels' % len(val_dataset))
        if len(val_dataset) < num_channels:   # the reference's DataLoader(batch_size=num_channels, drop_last=True) yields nothing
            raise ValueError('validation set holds %d channels, fewer than --num_channels %d: averages over the '
                             'missing columns would be diluted' % (len(val_dataset), num_channels))
        n = num_channels
        items = [val_dataset[i] for i in range(n)]
        val_P = torch.from_numpy(np.stack([it['P'] for it in items])).to(dev)
        val_P = torch.conj(torch.transpose(val_P, -1, -2)).contiguous()      # Hermitian pilots (:110)
        val_H_herm = torch.from_numpy(np.stack([it['H_herm'] for it in items])).to(dev)
        val_H = (val_H_herm[:, 0] + 1j * val_H_herm[:, 1]).contiguous()
        init_val_H = torch.randn_like(val_H)
        if ws > 1:   # every rank must hold the same inputs: rank 0's draw is broadcast
            for t in (val_P, val_H, init_val_H):
                torch.distributed.broadcast(torch.view_as_real(t), src=0)
        # measurement noise from a dedicated generator: the same draw whatever the number of ranks
        gen = torch.Generator(device=dev)
        gen.manual_seed(sampler_seed + 17 * meta_idx)
        nm = ec.ald_over_snr(diffuser, val_P, val_H, init_val_H, noise_range, alpha_step, beta_noise,
                             float(val_config.model.sigma_end), num_levels, steps_each,
                             seed=sampler_seed + meta_idx, generator=gen)
        nmse_log[spacing_idx, pilot_alpha_idx, :, :, :n] = nm

    # Use average estimation error to select best number of steps (test_score.py:173-175)
    avg_nmse = np.mean(nmse_log, axis=-1)
    best_nmse = np.min(avg_nmse, axis=-1)

    if rank == 0:
        def plot():
            from matplotlib import pyplot as plt
            plt.rcParams['font.size'] = 14
            plt.figure(figsize=(10, 10))
            for alpha_idx, local_alpha in enumerate(pilot_alpha_range):
                plt.plot(snr_range, 10 * np.log10(best_nmse[0, alpha_idx]), linewidth=4, label='Alpha=%.2f' % local_alpha)
            plt.grid(); plt.legend()
            plt.title('Score-based channel estimation')
            plt.xlabel('SNR [dB]'); plt.ylabel('NMSE [dB]')
            plt.tight_layout()
            plt.savefig(os.path.join(result_dir, 'results.png'), dpi=300, bbox_inches='tight')
            plt.close()
        if not args.no_plot:
            ec.maybe_plot(plot)
        save_dict = {'nmse_log': nmse_log, 'avg_nmse': avg_nmse, 'best_nmse': best_nmse,
                     'spacing_range': spacing_range, 'pilot_alpha_range': pilot_alpha_range,
                     'snr_range': snr_range, 'val_config': val_config}
        torch.save(save_dict, os.path.join(result_dir, 'results.pt'))
        print('SNR [dB]      :', ' '.join('%6.1f' % s for s in snr_range))
        print('best NMSE [dB]:', ' '.join('%6.2f' % v for v in 10 * np.log10(best_nmse[0, 0])))
    return {'nmse_log': nmse_log, 'avg_nmse': avg_nmse, 'best_nmse': best_nmse, 'snr_range': snr_range}


if __name__ == '__main__':
    main()
    ec.finalize()
