"""TEST INFRASTRUCTURE ONLY (container-only: needs /root/reference).

End-to-end accuracy distribution of the LIVE reference (idelucs/models.py:46-172 driven the way
idelucs/__main__.py:100-147 and idelucs/cluster.py:32-52 drive it), CPU, one voter per seed:

    for seed in seeds:  torch / numpy / random seeded -> IID_model(args) -> build_dataloader()
                        -> n_epochs x contrastive_training_epoch() -> predict() -> ACC / ARI vs GT

plus one 5-voter ensemble per dataset (label_features, idelucs/utils.py:582-602) for the published
rows Example/ALL_RESULTS.tsv:3 and :19.  Writes tests/golden/e2e_reference.json, which the GPU
test tests/test_gpu_e2e_stats.py compares the B200 path's distribution with (Mann-Whitney U).

    python oracle/gen_e2e_golden.py [--seeds 20] [--out tests/golden/e2e_reference.json]
"""
import argparse
import gzip
import json
import os
import random
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

DATASETS = {
    # name: (n_clusters, batch_sz)  — BASELINE.md §3 / Example/ALL_RESULTS.tsv:3, 19
    "Influenza-A": (5, 512),
    "Actinopterygii": (3, 256),
}


def base_args(fasta, gt, n_clusters, batch_sz):
    return {"sequence_file": fasta, "GT_file": gt, "n_clusters": n_clusters, "k": 6, "model_size": "linear",
            "n_mimics": 3, "batch_sz": batch_sz, "lambda": 2.8, "lr": 1e-3, "weight": 0.25, "scheduler": None,
            "optimizer": "RMSprop", "n_epochs": 35, "n_voters": 1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "e2e_reference.json"))
    ap.add_argument("--voters", type=int, default=5)
    a = ap.parse_args()
    import ref_live
    ref = ref_live.load()
    import torch
    from sklearn.metrics import adjusted_rand_score
    from idelucs import models as rmodels   # the reference's
    from idelucs import utils as rutils

    tmp = tempfile.mkdtemp(prefix="idl_e2e_")
    res = {"how": "live reference on CPU, k=6, n_mimics=3, 35 epochs, lambda=2.8, w=0.25, RMSprop lr=1e-3, one voter per seed; "
                  "seeds set on torch / numpy / random before IID_model() (oracle/gen_e2e_golden.py)", "datasets": {}}
    for name, (C, B) in DATASETS.items():
        fasta = os.path.join(tmp, name + ".fas")
        with gzip.open(os.path.join(ROOT, "tests", "golden", name + ".fas.gz"), "rb") as fi, open(fasta, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        gt = os.path.join(ROOT, "tests", "golden", name + "_GT.tsv")
        names, lengths, GT, dis = rutils.SummaryFasta(fasta, gt)
        uniq = sorted(set(GT))
        y_true = np.array([uniq.index(g) for g in GT])
        accs, aris, secs = [], [], []
        for seed in range(a.seeds):
            torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
            t0 = time.time()
            args = base_args(fasta, gt, C, B)
            m = rmodels.IID_model(args)
            m.names, m.lengths, m.GT, m.cluster_dis = names, lengths, GT, dis
            m.build_dataloader()
            for _ in range(args["n_epochs"]):
                m.contrastive_training_epoch()
            y_pred, _, _ = m.predict()
            _, acc = rutils.cluster_acc(y_true, np.asarray(y_pred))
            accs.append(float(acc)); aris.append(float(adjusted_rand_score(y_true, y_pred))); secs.append(time.time() - t0)
            print(name, "seed", seed, "acc %.4f ari %.4f (%.1f s)" % (accs[-1], aris[-1], secs[-1]), flush=True)
        # one ensemble run (voters share the dataloader, weights re-initialised per voter — __main__.py:100-147)
        torch.manual_seed(1000); np.random.seed(1000); random.seed(1000)
        args = base_args(fasta, gt, C, B)
        m = rmodels.IID_model(args)
        m.names, m.lengths, m.GT, m.cluster_dis = names, lengths, GT, dis
        m.build_dataloader()
        preds = []
        for v in range(a.voters):
            m.net.apply(rmodels.weights_init); m.epoch = 0
            for _ in range(args["n_epochs"]):
                m.contrastive_training_epoch()
            yp = np.asarray(m.predict()[0]).astype(np.int32)
            d, cnt = {}, 0
            for i in range(yp.shape[0]):       # __main__.py:131-140 relabelling in order of first appearance
                if yp[i] not in d:
                    d[yp[i]] = cnt; cnt += 1
                yp[i] = d[yp[i]]
            preds.append(yp)
        y_ens, _ = rutils.label_features(np.array(preds), C)
        _, acc_e = rutils.cluster_acc(y_true, np.asarray(y_ens))
        res["datasets"][name] = {"n_clusters": C, "batch_sz": B, "acc": accs, "ari": aris, "seconds_per_run": float(np.mean(secs)),
                                 "ensemble5_acc": float(acc_e), "ensemble5_ari": float(adjusted_rand_score(y_true, y_ens))}
        print(name, "ensemble acc %.4f" % acc_e, flush=True)
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
