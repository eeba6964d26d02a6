"""``python -m idelucs_b200`` — the reference CLI surface (idelucs/__main__.py:274-318: same
flags and defaults) driving the B200-native hot path.  Orchestration only: voters x epochs,
majority vote over voters (k-means on one-hot votes, utils.py:582-602) or HDBSCAN on the latent space when
--n_clusters=0 (the ``hdbscan`` package if importable, else scikit-learn's implementation), metrics and TSV output."""
import argparse
import os
import sys
import time

import numpy as np


def run(args):
    import torch  # noqa: F401
    from . import models
    from .utils import SummaryFasta, cluster_acc, label_features
    start = time.time()
    use_hdbscan = False
    if args["n_clusters"] == 0:   # idelucs/__main__.py:75-83: 200 output units, clusters from HDBSCAN on the latent space
        args["n_clusters"], use_hdbscan = 200, True
    model = models.IID_model(args)
    model.names, model.lengths, model.GT, model.cluster_dis = SummaryFasta(model.sequence_file, model.GT_file)
    print(model.cluster_dis)
    print(f"No. Sequences: \t {len(model.lengths):,}")
    print(f"Min. Length: \t {np.min(model.lengths):,}")
    print(f"Max. Length: \t {np.max(model.lengths):,}")
    print(f"Avg. Length: \t {round(float(np.mean(model.lengths)), 2):,}")
    model.build_dataloader()
    predictions, latent, probabilities = [], None, None
    min_loss, model_latent = np.inf, None
    for voter in range(args["n_voters"]):
        sys.stdout.write(f"\r........... Training Model ({voter + 1}/{args['n_voters']})................")
        sys.stdout.flush()
        model.net.apply(models.weights_init)
        model.epoch = 0
        voter_min = np.inf
        for _ in range(args["n_epochs"]):
            voter_min = min(voter_min, model.contrastive_training_epoch())
        y_pred, probabilities, latent = model.predict()
        if voter_min < min_loss:
            min_loss, model_latent = voter_min, latent
        if not use_hdbscan:   # relabel in order of first appearance (__main__.py:131-141)
            y_pred = y_pred.astype(np.int32)
            first = {}
            y_pred = np.array([first.setdefault(int(c), len(first)) for c in y_pred], dtype=np.int32)
            predictions.append(y_pred)
    if use_hdbscan:   # __main__.py:149-156 (the last voter's latent space, like the reference)
        msz = len(model.names) // 100 + 1
        try:
            import hdbscan
            clusterer = hdbscan.HDBSCAN(min_cluster_size=msz, gen_min_span_tree=True, prediction_data=True)
        except ImportError:   # same algorithm from scikit-learn (>= 1.3) when the hdbscan package is not installed
            from sklearn.cluster import HDBSCAN
            clusterer = HDBSCAN(min_cluster_size=max(msz, 2))
        clusterer.fit(latent)
        y_pred = clusterer.labels_ + 1
        probabilities = clusterer.probabilities_
        args["n_clusters"] = int(np.max(y_pred)) + 1
    else:
        y_pred, probabilities = label_features(np.array(predictions), args["n_clusters"])
    out_dir = os.path.join(os.getcwd(), "Results", os.path.basename(args["sequence_file"]).split(".")[0],
                           time.strftime("%b_%d_%H-%M-%S"))
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "assignments.tsv"), "w") as fh:
        fh.write("sequence_id\tassignment\tconfidence_score\n")
        for name, c, pr in zip(model.names, y_pred, probabilities):
            fh.write(f"{name}\t{int(c)}\t{float(pr):.6f}\n")
    print("\n")
    if model.GT is not None:
        from sklearn import metrics
        labels = {c: i for i, c in enumerate(sorted(set(model.GT)))}
        y_true = np.array([labels[c] for c in model.GT])
        _, acc = cluster_acc(y_true, np.asarray(y_pred))
        print(f"ACC \t {acc:.5f}")
        print(f"ARI \t {metrics.adjusted_rand_score(y_true, y_pred):.5f}")
        print(f"NMI \t {metrics.adjusted_mutual_info_score(y_true, y_pred):.5f}")
    print(f"Elapsed \t {time.time() - start:.1f} s   results in {out_dir}")
    return y_pred


def main(argv=None):
    parser = argparse.ArgumentParser(prog="idelucs")
    parser.add_argument("--sequence_file", action="store", type=str)
    parser.add_argument("--n_clusters", action="store", type=int, default=0,
                        help="Expected or maximum number of clusters (0: fine-grained clusters via HDBSCAN)")
    parser.add_argument("--n_epochs", action="store", type=int, default=100)
    parser.add_argument("--n_mimics", action="store", type=int, default=3, help="data augmentations per sequence")
    parser.add_argument("--batch_sz", action="store", type=int, default=256)
    parser.add_argument("--GT_file", action="store", type=str, default=None)
    parser.add_argument("--k", action="store", type=int, default=6, help="k-mer length")
    parser.add_argument("--optimizer", action="store", type=str, default="RMSprop")
    parser.add_argument("--scheduler", action="store", type=str, default="None")
    parser.add_argument("--weight", action="store", type=float, default=0.25, help="weight of the IIC term")
    parser.add_argument("--lambda", action="store", type=float, default=2.8, help="cluster balance")
    parser.add_argument("--lr", action="store", type=float, default=1e-3)
    parser.add_argument("--n_voters", action="store", type=int, default=5)
    parser.add_argument("--model_size", action="store", type=str, default="linear")
    parser.add_argument("--plot", action="store", type=bool, default=False)
    args = vars(parser.parse_args(argv))
    print("\nTraining Parameters:")
    for key in args:
        print(f"{key} \t -> {args[key]}")
    if not args["sequence_file"]:
        parser.error("--sequence_file is required")
    run(args)


if __name__ == "__main__":
    main()
