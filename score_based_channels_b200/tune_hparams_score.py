#!/usr/bin/env python3
"""Drop-in for ``python -m score_based_channels.tune_hparams_score`` (reference
``src/score_based_channels/tune_hparams_score.py``): same flags, same ``<channel>-hyperparameters.pt``
keys and array shapes (``nmse_log [n_alpha, n_beta, n_snr, steps, 100]``).

As in the reference the grid is alpha_step x beta_noise only; the best number of steps "N" is the
arg-min over the per-step NMSE log (tune_hparams_score.py:151-152).  Every grid cell re-draws its
pilots and initial estimate (tune_hparams_score.py:76-97).  All cells x SNR points x channels are
independent trajectories: they are sharded over the available GPUs."""
import argparse
import copy
import itertools
import os

import numpy as np
import torch

from . import entry_common as ec
from .loaders import Channels


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpu', type=int, default=0)
    parser.add_argument('--channel', type=str, default='CDL-C')
    parser.add_argument('--spacing', type=float, default=0.5)
    parser.add_argument('--alpha_step_range', nargs='+', type=float, default=[3e-11, 6e-11, 1e-10, 3e-10])
    parser.add_argument('--beta_noise_range', nargs='+', type=float, default=[0.1, 0.01, 0.001])
    parser.add_argument('--pilot_alpha', type=float, default=0.6)
    # extras
    parser.add_argument('--ckpt', type=str, default=None)
    parser.add_argument('--out_dir', type=str, default=None)
    parser.add_argument('--levels', type=int, default=None)
    parser.add_argument('--num_channels', type=int, default=100)
    parser.add_argument('--precision', type=str, default=None, choices=[None, 'auto', 'tf32x3', 'tf32', 'fp16x2'])
    parser.add_argument('--seed', type=int, default=None)
    parser.add_argument('--no_plot', action='store_true')
    args = parser.parse_args(argv)

    dev, rank, ws = ec.pick_device(args.gpu)
    if args.seed is not None:
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)
    sampler_seed = ec.common_seed(args.seed, dev, ws)

    target_dir = './models/score/%s' % args.channel
    target_file = args.ckpt or os.path.join(target_dir, 'final_model.pt')
    contents = ec.load_checkpoint(target_file, args.channel)
    config = contents['config']
    diffuser = ec.build_model(config, contents['model_state'], dev, args.precision)

    train_seed, val_seed = 1234, 4321
    config.data.channel = args.channel
    dataset = Channels(train_seed, config, norm=config.data.norm_channels)

    snr_range = np.arange(-10, 32.5, 2.5)
    alpha_step_range = np.asarray(args.alpha_step_range)
    beta_noise_range = np.asarray(args.beta_noise_range)
    noise_range = 10 ** (-snr_range / 10.) * config.data.image_size[1]
    num_levels = int(config.model.num_classes) if args.levels is None else int(args.levels)
    steps_each = int(config.sampling.steps_each)
    nch = args.num_channels

    nmse_log = np.zeros((len(alpha_step_range), len(beta_noise_range), len(snr_range),
                         int(num_levels * steps_each), nch))
    result_dir = args.out_dir or './results/score'
    if rank == 0:
        os.makedirs(result_dir, exist_ok=True)

    meta_params = itertools.product(alpha_step_range, beta_noise_range)
    for meta_idx, (alpha_step, beta_noise) in enumerate(meta_params):
        alpha_idx, beta_idx = np.unravel_index(meta_idx, (len(alpha_step_range), len(beta_noise_range)))
        val_config = copy.deepcopy(config)
        val_config.data.channel = args.channel
        val_config.data.spacing_list = [args.spacing]
        val_config.data.num_pilots = int(np.floor(config.data.image_size[1] * args.pilot_alpha))
        val_dataset = Channels(val_seed, val_config, norm=[dataset.mean, dataset.std], allow_other_seed=False)
        print('There are %d validation channels' % len(val_dataset))
        if len(val_dataset) < nch:   # the reference's DataLoader(batch_size=num_channels, drop_last=True) yields nothing
            raise ValueError('validation set holds %d channels, fewer than --num_channels %d: averages over the '
                             'missing columns would be diluted' % (len(val_dataset), nch))
        n = nch
        items = [val_dataset[i] for i in range(n)]
        val_P = torch.from_numpy(np.stack([it['P'] for it in items])).to(dev)
        val_P = torch.conj(torch.transpose(val_P, -1, -2)).contiguous()
        val_H_herm = torch.from_numpy(np.stack([it['H_herm'] for it in items])).to(dev)
        val_H = (val_H_herm[:, 0] + 1j * val_H_herm[:, 1]).contiguous()
        init_val_H = torch.randn_like(val_H)
        if ws > 1:
            for t in (val_P, val_H, init_val_H):
                torch.distributed.broadcast(torch.view_as_real(t), src=0)
        # measurement noise from a dedicated generator: the same draw whatever the number of ranks
        gen = torch.Generator(device=dev)
        gen.manual_seed(sampler_seed + 17 * meta_idx)
        nm = ec.ald_over_snr(diffuser, val_P, val_H, init_val_H, noise_range, float(alpha_step), float(beta_noise),
                             float(val_config.model.sigma_end), num_levels, steps_each,
                             seed=sampler_seed + meta_idx, id_base=meta_idx * len(snr_range) * n, generator=gen)
        nmse_log[alpha_idx, beta_idx, :, :, :n] = nm

    # Average estimation error and best stopping point (tune_hparams_score.py:150-152)
    avg_nmse = np.mean(nmse_log, axis=-1)
    best_nmse = np.min(avg_nmse, axis=-1)
    best_alpha_snr, best_beta_snr = [], []
    for snr_idx in range(len(snr_range)):
        local_nmse = best_nmse[..., snr_idx].flatten()
        best_idx = np.argmin(local_nmse)
        best_alpha_idx, best_beta_idx = np.unravel_index(best_idx, (len(alpha_step_range), len(beta_noise_range)))
        best_alpha_snr.append(alpha_step_range[best_alpha_idx])
        best_beta_snr.append(beta_noise_range[best_beta_idx])

    if rank == 0:
        def plot():
            from matplotlib import pyplot as plt
            plt.rcParams['font.size'] = 14
            plt.figure(figsize=(10, 10))
            for alpha_idx, local_alpha in enumerate(alpha_step_range):
                for beta_idx, local_beta in enumerate(beta_noise_range):
                    plt.plot(snr_range, 10 * np.log10(best_nmse[alpha_idx, beta_idx]), linewidth=4,
                             label='Alpha=%.2e, Beta=%.2e' % (local_alpha, local_beta))
            plt.grid(); plt.legend()
            plt.title('Score-based hyperparameter search')
            plt.xlabel('SNR [dB]'); plt.ylabel('NMSE [dB]')
            plt.tight_layout()
            plt.savefig(os.path.join(result_dir, '%s-hyperparameters.png' % args.channel), dpi=300, bbox_inches='tight')
            plt.close()
        if not args.no_plot:
            ec.maybe_plot(plot)
        torch.save({'nmse_log': nmse_log, 'avg_nmse': avg_nmse, 'best_nmse': best_nmse,
                    'best_alpha_snr': best_alpha_snr, 'best_beta_snr': best_beta_snr, 'snr_range': snr_range,
                    'alpha_step_range': alpha_step_range, 'beta_noise_range': beta_noise_range,
                    'config': config, 'args': args},
                   os.path.join(result_dir, '%s-hyperparameters.pt' % args.channel))
        print('best alpha per SNR:', best_alpha_snr)
        print('best beta  per SNR:', best_beta_snr)
    return {'nmse_log': nmse_log, 'avg_nmse': avg_nmse, 'best_nmse': best_nmse,
            'best_alpha_snr': best_alpha_snr, 'best_beta_snr': best_beta_snr}


if __name__ == '__main__':
    main()
    ec.finalize()
