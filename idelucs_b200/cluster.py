"""sklearn-style wrapper with the reference's signature (idelucs/cluster.py:9-52)."""
import sys

from . import models
from .utils import SummaryFasta


class iDeLUCS_cluster(object):
    def __init__(self, sequence_file, n_clusters=4, n_epochs=500, n_mimics=3, batch_sz=512, k=4, weight=0.25, n_voters=1):
        self.args = dict(sequence_file=sequence_file, n_clusters=n_clusters, n_epochs=n_epochs, n_mimics=n_mimics,
                         batch_sz=batch_sz, GT_file=None, k=k, optimizer="RMSprop", weight=weight, n_voters=n_voters,
                         lr=1e-3, model_size="linear", scheduler=None)
        self.args["lambda"] = 2.8

    def fit_predict(self, kmers=None):
        """Trains n_voters re-initialised models and returns the LAST voter's (y_pred, latent),
        like the reference (cluster.py:40-52; optimiser state is not reset between voters)."""
        model = models.IID_model(self.args)
        model.names, model.lengths, model.GT, model.cluster_dis = SummaryFasta(model.sequence_file, model.GT_file)
        model.build_dataloader()
        y_pred = latent = None
        for voter in range(self.args["n_voters"]):
            sys.stdout.write(f"\r........... Training Model ({voter + 1}/{self.args['n_voters']})................")
            sys.stdout.flush()
            model.net.apply(models.weights_init)
            model.epoch = 0
            for _ in range(self.args["n_epochs"]):
                model.contrastive_training_epoch()
            y_pred, _, latent = model.predict()
        return y_pred, latent
