"""``python -m idelucs_b200`` — the reference CLI surface (idelucs/__main__.py:274-318: same
flags and defaults) driving the B200-native hot path.  Orchestration only: voters x epochs,
majority vote over voters (k-means on one-hot votes, utils.py:582-602) or HDBSCAN when
--n_clusters=0 and the ``hdbscan`` package is importable, metrics and TSV output."""
import argparse
import os
import sys
import time

import numpy as np


def label_features(predictions, n_clusters):
    """vote ensemble of idelucs/utils.py:582-602 (k-means over centred one-hot votes)"""
    from sklearn.cluster import KMeans
    n_v, n = predictions.shape
    feats = np.zeros((n, n_v * n_clusters))
    for v in range(n_v):
        feats[np.arange(n), v * n_clusters + predictions[v]] = 1.0
    feats -= feats.sum(axis=0) / n
    return KMeans(n_clusters=n_clusters, init="k-means++", n_init=10).fit_predict(feats)


def run(args):
    import torch  # noqa: F401
    from . import models
    from .utils import SummaryFasta, cluster_acc
    start = time.time()
    use_hdbscan = False
    if args["n_clusters"] == 0:
        try:
            import hdbscan
        except ImportError:
            raise SystemExit("--n_clusters=0 needs the 'hdbscan' package (not installed); pass --n_clusters > 0")
        args["n_clusters"], use_hdbscan = 200, True
    model = models.IID_model(args)
    model.names, model.lengths, model.GT, model.cluster_dis = SummaryFasta(model.sequence_file, model.GT_file)
    print(model.cluster_dis)
    print(f"No. Sequences: \t {len(model.lengths):,}")
    print(f"Min. Length: \t {np.min(model.lengths):,}")
    print(f"Max. Length: \t {np.max(model.lengths):,}")
    print(f"Avg. Length: \t {round(float(np.mean(model.lengths)), 2):,}")
    model.build_dataloader()
    predictions, latent = [], None
    for voter in range(args["n_voters"]):
        sys.stdout.write(f"\r........... Training Model ({voter + 1}/{args['n_voters']})................")
        sys.stdout.flush()
        model.net.apply(models.weights_init)
        model.epoch = 0
        for _ in range(args["n_epochs"]):
            model.contrastive_training_epoch()
        y_pred, _, latent = model.predict()
        predictions.append(y_pred.astype(np.int64))
    if use_hdbscan:
        y_pred = hdbscan.HDBSCAN(min_cluster_size=len(model.names) // 100 + 1).fit_predict(latent)
    elif len(predictions) > 1:
        y_pred = label_features(np.stack(predictions), args["n_clusters"])
    else:
        y_pred = predictions[0]
    out_dir = os.path.join(os.getcwd(), "Results", os.path.basename(args["sequence_file"]).split(".")[0],
                           time.strftime("%b_%d_%H-%M-%S"))
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "assignments.tsv"), "w") as fh:
        fh.write("sequence_id\tassignment\n")
        for name, c in zip(model.names, y_pred):
            fh.write(f"{name}\t{int(c)}\n")
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
