#!/usr/bin/env python3
# -*- coding: utf-8 -*-
"""Drop-in for ``python -m score_based_channels.test_mmse`` (reference
``src/score_based_channels/test_mmse.py``): the approximate-MMSE estimator of Fig. 5c -- ``mmse_avg`` independent
posterior samples per channel drawn with the annealed-Langevin sampler (per-SNR step size / noise boost / early
stop, optional start point and data-consistency boost) and averaged.

Same flags as the reference (``--gpu --model --channel --start_point --spacing --pilot_alpha --steps_each
--dc_boost --num_classes``; ``--normalize_grad`` is parsed and unused there, test_mmse.py:24) and the same result-file
keys (``test_mmse.py:278-293``: spacing_range, pilot_alpha_range, args, config, snr_range, val_config, oracle_log,
oracle_H, saved_H) plus ``mmse_nmse`` = NMSE of the averaged estimate.  The reference script hard-codes private
checkpoint / hyper-parameter files (``:43-61,121-122``) that are not shipped: ``--ckpt`` and ``--hyper`` name them
here, and without ``--hyper`` the test_score defaults (step 3e-11, noise 0.01, no early stop) are used for every
SNR.  The reference re-appends the per-SNR batches to lists it has already tensorised (``:178-192``); the evident
intent -- one batch of ``kept_samples x mmse_avg`` trajectories per SNR point -- is what runs here.

The whole per-SNR batch is ONE launch of the fused kernel (``sampler.ald_run``): ``dc_boost`` and the early stop
(``stop_step``) are kernel arguments, so nothing returns to the host inside the loop.  Under torchrun the
(channel x posterior-sample) axis is sharded over the GPUs."""
from __future__ import annotations

import argparse
import copy
import itertools
import os

import numpy as np
import torch

from . import dist as sdist
from . import entry_common as ec
from . import sampler
from .loaders import Channels


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpu', type=int, default=1)
    parser.add_argument('--model', type=str, default='CDL-C')
    parser.add_argument('--channel', type=str, default='CDL-C')
    parser.add_argument('--start_point', type=str, default='Noise', choices=['Noise', 'Adjoint', 'LS'])
    parser.add_argument('--spacing', nargs='+', type=float, default=[0.5])
    parser.add_argument('--pilot_alpha', nargs='+', type=float, default=[0.6])
    parser.add_argument('--steps_each', type=int, default=3)
    parser.add_argument('--normalize_grad', type=bool, default=False)
    parser.add_argument('--dc_boost', type=float, default=1)
    parser.add_argument('--num_classes', type=int, default=2311)
    # additions (the reference hard-codes these)
    parser.add_argument('--ckpt', type=str, default=None)
    parser.add_argument('--hyper', type=str, default=None, help="'our_hyperparams_<model>.pt' of test_mmse.py:121")
    parser.add_argument('--out_dir', type=str, default=None)
    parser.add_argument('--kept_samples', type=int, default=100)
    parser.add_argument('--mmse_avg', type=int, default=50)
    parser.add_argument('--snr_range', nargs='+', type=float, default=None)
    parser.add_argument('--levels', type=int, default=None, help='run only the first N sigma levels (debug)')
    parser.add_argument('--precision', type=str, default=None, choices=[None, 'auto', 'tf32x3', 'tf32', 'fp16x2'])
    parser.add_argument('--seed', type=int, default=None)
    args = parser.parse_args(argv)

    dev, rank, ws = ec.pick_device(args.gpu if torch.cuda.device_count() > args.gpu else 0)
    if args.seed is not None:
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)
    sampler_seed = ec.common_seed(args.seed, dev, ws)

    contents = ec.load_checkpoint(args.ckpt or './models/score/%s/final_model.pt' % args.model, args.model)
    config = contents['config']
    config.sampling.sigma = 0.
    config.purpose = 'train'
    # "More sigmas" (test_mmse.py:69-74): the schedule is re-derived for num_classes levels
    config.model.num_classes = args.num_classes
    config.model.sigma_rate = (config.model.sigma_end / config.model.sigma_begin) ** (1 / (config.model.num_classes - 1))
    state = {k: v for k, v in contents['model_state'].items()}
    if args.num_classes != int(state['sigmas'].numel()):
        state.pop('sigmas')          # "Load weights WITHOUT SIGMAS" (:79): keep the re-derived schedule
    diffuser = ec.build_model(config, state, dev, args.precision) if 'sigmas' in state else \
        _build_without_sigmas(config, state, dev, args.precision)

    train_seed, val_seed = 1234, 4321
    config.data.channel = args.model
    config.data.array = 'ULA'
    dataset = Channels(train_seed, config, norm=config.data.norm_channels)

    config.sampling.steps_each = args.steps_each
    num_levels = int(config.model.num_classes) if args.levels is None else int(args.levels)
    total_steps = num_levels * args.steps_each
    snr_range = np.arange(-30, 17.5, 2.5) if args.snr_range is None else np.asarray(args.snr_range, dtype=float)
    spacing_range = np.asarray(args.spacing)
    pilot_alpha_range = np.asarray(args.pilot_alpha)
    noise_range = 10 ** (-snr_range / 10.)                                  # test_mmse.py:100 (no Nt factor here)
    kept, navg = args.kept_samples, args.mmse_avg
    Nt, Nr = int(config.data.image_size[1]), int(config.data.image_size[0])

    # per-(pilot_alpha, SNR) hyper-parameters (test_mmse.py:120-127)
    shape = (len(pilot_alpha_range), len(snr_range))
    if args.hyper:
        hp = torch.load(args.hyper, map_location='cpu', weights_only=False)
        best_step, best_noise, best_stop = (np.asarray(hp[k]) for k in ('best_step_idx', 'best_noise_idx', 'best_stop_idx'))
    else:
        best_step, best_noise = np.full(shape, 3e-11), np.full(shape, 0.01)
        best_stop = np.full(shape, total_steps - 1, dtype=np.int64)

    oracle_log = np.full((len(spacing_range), len(pilot_alpha_range), len(snr_range), total_steps, kept, navg), np.nan)
    saved_H = np.zeros((len(spacing_range), len(pilot_alpha_range), len(snr_range), kept, navg, Nt, Nr), dtype=np.complex64)
    mmse_nmse = np.zeros((len(spacing_range), len(pilot_alpha_range), len(snr_range), kept))
    result_dir = args.out_dir or 'TWC_rebuttal_MMSE_aug6_seed%d' % val_seed
    if rank == 0:
        os.makedirs(result_dir, exist_ok=True)

    oracle_H, val_config = None, None
    for meta_idx, (spacing, pilot_alpha) in enumerate(itertools.product(spacing_range, pilot_alpha_range)):
        spacing_idx, pilot_alpha_idx = np.unravel_index(meta_idx, (len(spacing_range), len(pilot_alpha_range)))
        val_config = copy.deepcopy(config)
        val_config.purpose = 'val'
        val_config.data.channel = args.channel
        val_config.data.spacing_list = [spacing]
        val_config.data.num_pilots = int(np.floor(config.data.num_pilots * pilot_alpha))
        val_dataset = Channels(val_seed, val_config, norm=[dataset.mean, dataset.std], allow_other_seed=False)
        print('There are %d validation channels!' % len(val_dataset))
        n = min(kept, len(val_dataset))
        items = [val_dataset[i] for i in range(n)]
        val_P = torch.from_numpy(np.stack([it['P'] for it in items])).to(dev)
        val_P = torch.conj(torch.transpose(val_P, -1, -2)).contiguous()      # test_mmse.py:159
        val_H_herm = torch.from_numpy(np.stack([it['H_herm'] for it in items])).to(dev)
        val_H = (val_H_herm[:, 0] + 1j * val_H_herm[:, 1]).contiguous()
        if ws > 1:
            for t in (val_P, val_H):
                torch.distributed.broadcast(torch.view_as_real(t), src=0)
        oracle_H = val_H.cpu().numpy()

        for snr_idx, local_noise in enumerate(noise_range):
            step_size = float(best_step[pilot_alpha_idx, snr_idx])
            noise_boost = float(best_noise[pilot_alpha_idx, snr_idx])
            target_stop = int(best_stop[pilot_alpha_idx, snr_idx])
            local_Y = torch.matmul(val_P, val_H)
            local_Y = (local_Y + float(np.sqrt(local_noise)) * torch.randn_like(local_Y)).contiguous()
            if ws > 1:   # rank 0's measurement noise is THE draw: broadcast before anything is derived from it
                torch.distributed.broadcast(torch.view_as_real(local_Y), src=0)
            # every channel is replicated mmse_avg times: same pilots / measurements, independent chains (:180-187)
            gP = val_P.repeat_interleave(navg, dim=0)
            gY = local_Y.repeat_interleave(navg, dim=0)
            gH = val_H.repeat_interleave(navg, dim=0)
            if args.start_point == 'Noise':
                current = torch.randn_like(gH)
            elif args.start_point == 'Adjoint':
                current = torch.matmul(torch.conj(torch.transpose(gP, -1, -2)), gY)
            else:   # 'LS' (:199-201): minimum-norm least squares of P x = y on the host
                current = torch.linalg.lstsq(gP.cpu(), gY.cpu(), driver='gelsd').solution.to(dev)
            current = current.contiguous()     # lstsq returns a column-major solution
            if ws > 1:
                torch.distributed.broadcast(torch.view_as_real(current), src=0)
            total = n * navg
            lo, hi = sdist.shard_range(total, rank, ws)
            ids = torch.arange(lo, hi, dtype=torch.int64, device=dev) + (meta_idx * len(snr_range) + snr_idx) * total
            X, nlog = sampler.ald_run(diffuser, gP[lo:hi], gY[lo:hi], current[lo:hi], gH[lo:hi],
                                      noise_var=float(local_noise), alpha_step=step_size, beta=noise_boost,
                                      sigma_end=float(val_config.model.sigma_end), level_begin=0, level_end=num_levels,
                                      steps_each=args.steps_each, seed=sampler_seed, sample_ids=ids,
                                      dc_boost=float(args.dc_boost), stop_step=min(target_stop, total_steps - 1))
            nlog = sdist.gather_columns(nlog, total)
            X = torch.view_as_complex(sdist.gather_rows(torch.view_as_real(X).contiguous(), total))
            if target_stop < total_steps - 1:
                print('Early stopping at step %d!' % target_stop)
            oracle_log[spacing_idx, pilot_alpha_idx, snr_idx, :, :n] = \
                nlog.double().cpu().numpy().reshape(total_steps, n, navg)
            Xs = X.view(n, navg, Nt, Nr)
            saved_H[spacing_idx, pilot_alpha_idx, snr_idx, :n] = Xs.cpu().numpy()
            est = Xs.mean(dim=1)                                              # the approximate-MMSE estimate
            mmse_nmse[spacing_idx, pilot_alpha_idx, snr_idx, :n] = (
                torch.sum(torch.abs(est - val_H) ** 2, dim=(-1, -2)) / torch.sum(torch.abs(val_H) ** 2, dim=(-1, -2))
            ).double().cpu().numpy()

    if rank == 0:
        save_dict = {'spacing_range': spacing_range, 'pilot_alpha_range': pilot_alpha_range, 'args': args,
                     'config': config, 'snr_range': snr_range, 'val_config': val_config, 'oracle_log': oracle_log,
                     'oracle_H': oracle_H, 'saved_H': saved_H, 'mmse_nmse': mmse_nmse}
        torch.save(save_dict, os.path.join(result_dir, 'model_%s_channel_%s.pt' % (args.model, args.channel)))
        print('SNR [dB]            :', ' '.join('%6.1f' % s for s in snr_range))
        print('approx-MMSE NMSE [dB]:', ' '.join('%6.2f' % v for v in 10 * np.log10(mmse_nmse[0, 0].mean(-1))))
    return {'oracle_log': oracle_log, 'saved_H': saved_H, 'mmse_nmse': mmse_nmse, 'snr_range': snr_range}


def _build_without_sigmas(config, state, dev, precision):
    """strict load of everything but the `sigmas` buffer, which keeps the schedule re-derived from the config."""
    from .ncsnv2 import NCSNv2Deepest
    diffuser = NCSNv2Deepest(config, precision=precision).to(dev)
    own = diffuser.state_dict()
    state = dict(state)
    state['sigmas'] = own['sigmas']
    diffuser.load_state_dict(state)
    diffuser.eval()
    return diffuser


if __name__ == '__main__':
    main()
    ec.finalize()
