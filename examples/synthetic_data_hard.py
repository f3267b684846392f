#!/usr/bin/env python
"""The reference's "hard" synthetic experiment (test/synthetic_data_hard_test.py, BASELINE.json configs[1]) driven through
this package: N = 100 points, D = 60 outputs made of four groups of 15 GP draws whose ARD relevances use different
pairs of the five true inputs, Q = 10 latent dimensions, M = 50 inducing points, truncation T = 20, Adam lr = 0.01.

    python examples/synthetic_data_hard.py [--iters 2500] [--mode d|t] [--cuda-graph] [--out results.npz]

Needs a B200 (the product path has no CPU fallback).  The result file has the keys the reference's analysis scripts
read (src/utils/constants.py:38-73).  The data recipe follows the description of the reference script
(:21-30 shape constants, :54-64 kernels, :92-105 sampling, :116-118 standardisation); it is regenerated here with
numpy instead of a TensorFlow session, so the draws differ from the reference's even at the same seed.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ARD_MASKS = ([1.0, 1.0, 0.0, 0.0, 0.0], [1.0, 0.0, 1.0, 0.0, 0.0], [0.0, 1.0, 0.0, 1.0, 0.0], [1.0, 0.0, 0.0, 1.0, 0.0])


def make_data(num_samples=100, dims_per_group=15, input_dim=5, seed=10):
    """[N, 4 * dims_per_group] standardised outputs and the group label of every column."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((num_samples, input_dim))
    columns, labels = [], []
    for g, mask in enumerate(ARD_MASKS):
        xs = x * np.sqrt(np.asarray(mask))
        sq = (xs ** 2).sum(1)
        k = np.exp(-0.5 * (sq[:, None] + sq[None, :] - 2.0 * xs @ xs.T))
        noise_var = (0.1 + 0.2 * rng.standard_normal()) ** 2
        chol = np.linalg.cholesky(k + (noise_var + 1e-8) * np.eye(num_samples))
        columns.append(chol @ rng.standard_normal((num_samples, dims_per_group)))
        labels += [g] * dims_per_group
    y = np.concatenate(columns, axis=1)
    y = (y - y.mean(0)) / y.std(0)
    return y, np.asarray(labels)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2500)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--mode", choices=("d", "t"), default="d", help="d: dp_gp_lvm (the script's model); t: dp_gp_lvm_t")
    ap.add_argument("--cuda-graph", action="store_true", help="replay one captured CUDA graph per iteration")
    ap.add_argument("--out", default="synthetic_hard_results.npz")
    args = ap.parse_args()

    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    from dp_gp_lvm_b200.train import save_results, train

    y, labels = make_data()
    np.random.seed(10)                                  # the D-mode factory draws its initial point from numpy's global RNG
    if args.mode == "d":
        model = dp_gp_lvm(y_train=y, num_latent_dims=10, num_inducing_points=50, truncation_level=20, mask_size=1)
    else:
        model = dp_gp_lvm_t(y_train=y, num_latent_dims=10, num_inducing_points=50, truncation_level=20, mask_size=1, seed=10)
    seconds, history = train(model, learning_rate=args.lr, train_iter=args.iters, name="DP-GP-LVM", use_cuda_graph=args.cuda_graph)
    save_results(model, args.out, y, train_opt_time=seconds)

    # which truncation component every output dimension ended up in, against the generating groups
    assign = model.assignments.detach().cpu().numpy().argmax(1)
    print("iterations/s: %.1f" % (args.iters / seconds))
    for g in range(len(ARD_MASKS)):
        comps, counts = np.unique(assign[labels == g], return_counts=True)
        print("group %d -> components %s" % (g, dict(zip(comps.tolist(), counts.tolist()))))
    print("saved", args.out)


if __name__ == "__main__":
    main()
