"""Where the wall time of one CLI run goes (development aid): interpreter / CUDA start-up, ingest, 5 voters x 35 epochs, predict, vote.
usage: python tools/cli_phases.py file.fas"""
import os, sys, time
t0 = time.time()
sys.path.insert(0, os.getcwd())
import numpy as np, torch
t1 = time.time()
from idelucs_b200 import models
from idelucs_b200.utils import SummaryFasta, label_features
t2 = time.time()
args = dict(sequence_file=sys.argv[1], n_clusters=5, n_epochs=35, n_mimics=3, batch_sz=512, GT_file=None, k=6, optimizer="RMSprop", scheduler="None",
            weight=0.25, lr=1e-3, n_voters=5, model_size="linear", plot=False)
args["lambda"] = 2.8
m = models.IID_model(args)
m.names, m.lengths, m.GT, m.cluster_dis = SummaryFasta(m.sequence_file, m.GT_file)
torch.cuda.synchronize(); t3 = time.time()
m.build_dataloader()
torch.cuda.synchronize(); t4 = time.time()
tt = tp = 0.0
preds = []
for v in range(5):
    m.net.apply(models.weights_init); m.epoch = 0
    a = time.time()
    for _ in range(35):
        m.contrastive_training_epoch()
    torch.cuda.synchronize(); b = time.time()
    y, p, lat = m.predict()
    c = time.time()
    tt += b - a; tp += c - b
    preds.append(y.astype(np.int32))
t5 = time.time()
label_features(np.array(preds), 5)
t6 = time.time()
print("import torch %.2f | import package %.2f | model + SummaryFasta %.2f | build_dataloader %.2f | train 5 x 35 epochs %.2f | predict x5 %.2f | vote %.2f | total %.2f"
      % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, tt, tp, t6 - t5, t6 - t0))
print("steps per epoch", len(m.dataloader))
